// K5s — the frame-parallel spectral kernels SPECIALISED for one plan (compiled by NVRTC, see spectral_spec_source() in
// spec.cpp).  The generated translation unit is
//     #define QG_SPEC_ONLY
//     #include "interp.cu"                 // exec(): the SAME op semantics as every other kernel
//     #include "spectral_pod.h"
//     namespace qg { constexpr Instr kSpCode[] = {...}; constexpr SpSegment kSpSegs[] = {...}; ... }   + SP_* counts
//     #include "spectral_kernel.cuh"
// What the generic kernels of spectral.cu pay per decoded mini-tape instruction (dispatch, operand decoding, a shared-memory
// round trip per edge) is gone: a thread carries one sample / one bin through its mini-tape in registers, every transform
// size is a compile-time constant.  Same arithmetic in the same order: results are bit-identical to spectral.cu's, which are
// bit-identical to the time-vector kernel's (tests/test_gpu_spectral.py).
#pragma once

namespace qg {

constexpr int SP_PS = SP_P + SP_NS;

struct SpLane {
  float x[SP_PS + (SP_NSLOTS > 2 ? SP_NSLOTS : 2)];
  const float* tables;
  __device__ __forceinline__ float& at(int i) { return x[i]; }
  __device__ __forceinline__ float& out(int base, int k) { return x[base + k]; }
  __device__ __forceinline__ float& in(int f) { return x[f]; }
  __device__ __forceinline__ float& sc(int i) { return x[i]; }
  __device__ __forceinline__ float& ring(uint32_t, uint32_t) const { return const_cast<SpLane*>(this)->x[0]; }   // never reached: mini-tapes hold no ring ops
  __device__ __forceinline__ int first() { return 0; }
  __device__ __forceinline__ bool more(int k) const { return k < 1; }
  __device__ __forceinline__ int next(int k) { return k + 1; }
  __device__ __forceinline__ uint32_t sample() const { return 0u; }
  __device__ __forceinline__ uint32_t count() const { return 1u; }
};
__device__ __forceinline__ uint32_t ring_len(const SpLane&, uint32_t) { return 1u; }
__device__ __forceinline__ void reset_range(const SpLane&, uint32_t) {}
__device__ __forceinline__ void lane_fft(const SpLane&, uint32_t, uint32_t, int, const float*, bool) {}

#include "spectral_fft.cuh"

// mini-tape [i, hi) for ONE column whose time is t (the instruction's own offset is in `pad`)
template <int i, int hi>
__device__ __forceinline__ void sp_run(SpLane& L, const SpArgs& a, int v, long t) {
  if constexpr (i < hi) {
    constexpr Instr I = kSpCode[i];
    constexpr long d = (long)(int32_t)I.pad;
    if constexpr (I.op == OP_NOISE) {
      L.x[I.out] = d_noise(__float_as_uint(L.x[I.s]) + (uint32_t)(t + d) + 1u);
    } else if constexpr (I.op == OP_WAVE) {
      constexpr uint32_t len = I.aux2;
      uint32_t m;
      if constexpr ((len & (len - 1u)) == 0u) m = (__float_as_uint(L.x[I.s]) + (uint32_t)(t + d)) & (len - 1u);
      else { long mm = ((long)__float_as_uint(L.x[I.s]) + t + d) % (long)len; m = (uint32_t)(mm < 0 ? mm + (long)len : mm); }
      L.x[I.out] = __ldg(L.tables + I.aux + m);
    } else if constexpr (I.op == OP_IMPULSE) {
      L.x[I.out] = (__float_as_uint(L.x[I.s]) == 0u && t + d == 0) ? 1.0f : 0.0f;
    } else if constexpr (I.op == OP_DELAY || I.op == OP_TICK) {
      L.x[I.out] = t + d >= 0 ? L.x[I.in[0]] : 0.0f;
    } else if constexpr (I.op == OP_STREAM_IN) {
      L.x[I.out] = a.y[((size_t)I.aux * a.V + v) * (size_t)a.ring + ((uint32_t)t & ((uint32_t)a.ring - 1u))];
    } else {
      int pc = 0;
      exec(I, L, pc);
    }
    sp_run<i + 1, hi>(L, a, v, t);
  }
}

__device__ __forceinline__ void sp_load_scalars(SpLane& L, const SpArgs& a, int v) {
  L.tables = a.tables;
#pragma unroll
  for (int p = 0; p < SP_P; p++) L.x[p] = a.params[(size_t)p * a.Vp + v];
#pragma unroll
  for (int s = 0; s < SP_NS; s++) L.x[SP_P + s] = a.state_init[(size_t)s * a.Vp + v];
#pragma unroll
  for (int k = 0; k < (SP_NSLOTS > 2 ? SP_NSLOTS : 2); k++) L.x[SP_PS + k] = 0.0f;
}

// one frame of segment SEG whose output starts at time tb (see spectral.cu for the timing of the reference)
template <int SEG>
__device__ __forceinline__ void sp_frame(const SpArgs& a, int v, long tb, int tid, int nth) {
  constexpr SpSegment sg = kSpSegs[SEG];
  constexpr int LG = sg.lg, N = 1 << LG, half = N >> 1, sh = 32 - LG;
  const long tau = tb - N;
  float* yre = sg.y_re >= 0 ? a.y + ((size_t)sg.y_re * a.V + v) * (size_t)a.ring : nullptr;
  float* yim = sg.y_im >= 0 ? a.y + ((size_t)sg.y_im * a.V + v) * (size_t)a.ring : nullptr;
  const long lo = a.t0 - tb, hi = a.t0 + a.T - tb;
  if (hi <= 0 || lo >= N) return;
  const int i_lo = lo > 0 ? (int)lo : 0, i_hi = hi < N ? (int)hi : N;
  const uint32_t ymask = (uint32_t)a.ring - 1u, y0 = (uint32_t)tb;
  if (tau + N <= 0) {
    for (int i = i_lo + tid; i < i_hi; i += nth) {
      if (yre) yre[(y0 + (uint32_t)i) & ymask] = 0.0f;
      if (yim) yim[(y0 + (uint32_t)i) & ymask] = 0.0f;
    }
    return;
  }
  constexpr int f_off = 0, g_off = CPAD(SP_NMAX) + 1;   // float2 offsets into shared memory
  float2* f = QG_SMEM_C + f_off;
  float2* g = QG_SMEM_C + g_off;
  // twiddles stay in global memory behind the read-only cache: staging the table in shared memory was measured 1.3 % SLOWER
  // (99.95 vs 98.73 ms on configs[3]) — the loads hit L1 and the copy costs more than it saves
  const float2* tw = reinterpret_cast<const float2*>(a.tables + sg.tw);
  SpLane L;
  sp_load_scalars(L, a, v);
  // ---- the frame's N input samples, bit-reversed into the transform buffer
  for (int m = tid; m < N; m += nth) {
    const long t = tau - N + m;
    float xv = 0.0f;
    if (t >= 0) {
      sp_run<sg.pre_lo, sg.pre_hi>(L, a, v, t);
      xv = L.x[sg.pre_x];
    }
    f[CPAD(__brev((uint32_t)m) >> sh)] = make_float2(xv, 0.0f);
  }
  __syncthreads();
  sp_fft_n<LG, false>(f_off, tw, 0, 1.0f, tid, nth);
  // ---- the bin chain (bins 0 .. N/2 and their mirror when the chain's behaviour under conjugation is known)
  constexpr bool sym = sg.sym_re != 0;
  constexpr int i_end = sym ? half + 1 : N;
  for (int i = tid; i < i_end; i += nth) {
    float2 z;
    if (i <= half) z = f[CPAD(i)];
    else { z = f[CPAD(N - i)]; z.y = -z.y; }
    L.x[SP_PS] = z.x; L.x[SP_PS + 1] = z.y;
    sp_run<sg.ch_lo, sg.ch_hi>(L, a, v, tau + i);
    const float2 r = make_float2(L.x[sg.in_re_x], L.x[sg.in_im_x]);
    g[CPAD(__brev((uint32_t)i) >> sh)] = tau + i >= 0 ? r : make_float2(0.0f, 0.0f);
    if (sym && i > 0 && i < half) {
      const int im = N - i;
      g[CPAD(__brev((uint32_t)im) >> sh)] = tau + im >= 0 ? make_float2(r.x * (float)sg.sym_re, r.y * (float)sg.sym_im) : make_float2(0.0f, 0.0f);
    }
  }
  __syncthreads();
  sp_fft_n<LG, false>(g_off, tw, 0, -1.0f, tid, nth);
  const float sc = 1.0f / (float)N;
  for (int i = i_lo + tid; i < i_hi; i += nth) {
    const float2 z = g[CPAD(i)];
    if (yre) yre[(y0 + (uint32_t)i) & ymask] = z.x * sc;
    if (yim) yim[(y0 + (uint32_t)i) & ymask] = z.y * sc;
  }
}
template <int S>
__device__ __forceinline__ void sp_frame_of(int seg, const SpArgs& a, int v, long c, int frame, int tid, int nth) {
  if constexpr (S < SP_NSEG) {
    if (seg == S) {
      constexpr int N = 1 << kSpSegs[S].lg;
      sp_frame<S>(a, v, c * (long)SP_C + ((N - kSpSegs[S].start) & (N - 1)) + (long)frame * N, tid, nth);
    } else {
      sp_frame_of<S + 1>(seg, a, v, c, frame, tid, nth);
    }
  }
}

extern "C" __global__ void __launch_bounds__(256, SP_MIN_BLOCKS) k_sp_frames(SpArgs a, long c0, int n_rounds) {
  const int per_voice = SP_NITEMS * n_rounds;
  const int v = (int)(blockIdx.x / (unsigned)per_voice), w = (int)(blockIdx.x % (unsigned)per_voice);
  const SpItem it = kSpItems[w % SP_NITEMS];
  sp_frame_of<0>(it.seg, a, v, c0 + w / SP_NITEMS, it.frame, threadIdx.x, blockDim.x);
}

// the post-graph: one thread = one output sample at a time, SP_POST_BLOCK consecutive samples of one voice per CTA
extern "C" __global__ void __launch_bounds__(256) k_sp_post(SpArgs a, long t_lo, long t_hi) {
  const int nblk = (int)((t_hi - t_lo + SP_POST_BLOCK - 1) / SP_POST_BLOCK);
  const int v = (int)(blockIdx.x / (unsigned)nblk), b = (int)(blockIdx.x % (unsigned)nblk);
  const long tbase = t_lo + (long)b * SP_POST_BLOCK;
  const int n = (int)(t_hi - tbase < SP_POST_BLOCK ? t_hi - tbase : SP_POST_BLOCK);
  SpLane L;
  sp_load_scalars(L, a, v);
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const long t = tbase + j;
    sp_run<SP_POST_LO, SP_POST_HI>(L, a, v, t);
    const size_t tt = (size_t)(t - a.t0);
#pragma unroll
    for (int c = 0; c < SP_NOUT; c++) {
      const size_t o = a.frame_major ? (tt * a.V + v) * SP_NOUT + c : ((size_t)v * SP_NOUT + c) * a.T + tt;
      a.out[o] = L.x[kSpOutX[c]];
    }
  }
}

}  // namespace qg
