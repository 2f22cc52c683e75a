// K1s — lane kernel SPECIALISED for one tape (compiled by NVRTC at bank-build time, see spec.cpp).
//
// The generated translation unit is
//     #define QG_SPEC_ONLY
//     #include "interp.cu"                  // lane types, exec() and its helpers — the SAME op semantics as every interpreter
//     namespace qg { constexpr Instr kTape[] = {...}; constexpr uint16_t kOutX[] = {...}; }   + QG_SPEC_* counts
//     #include "spec_kernel.cuh"
// Every instruction is a compile-time constant, so exec()'s dispatch folds to the one case it needs, operand indices are
// literals, and the lane's X array (parameters, state, temporaries) has only constant subscripts: it lives in registers.
// What the interpreters pay per decoded instruction (dispatch, operand decoding, a shared-memory round trip per edge)
// is gone; what is left is the ops' own arithmetic — the reference's `Net::tick` without the virtual calls and edge copies
// (/root/reference/src/process.rs:1347-1351).  Uniform (feed-forward / feedback) tapes only: nested-net control flow keeps
// the SIMT-stack interpreter.
#pragma once

namespace qg {

constexpr int SPEC_P = QG_SPEC_P, SPEC_NS = QG_SPEC_NS, SPEC_NT = QG_SPEC_NT, SPEC_NIN = QG_SPEC_NIN, SPEC_NOUT = QG_SPEC_NOUT;
constexpr int SPEC_N = QG_SPEC_N, SPEC_NX = SPEC_P + SPEC_NS + SPEC_NT;

struct RegLane {
  float x[SPEC_NX > 0 ? SPEC_NX : 1];
  int v, Vp;
  float* rings;
  const Ring* ring_tab;
  const float* tables;
  __device__ __forceinline__ float& at(int i) { return x[i]; }
  __device__ __forceinline__ float& out(int base, int k) { return x[base + k]; }
  __device__ __forceinline__ float& in(int f) { return x[f]; }
  __device__ __forceinline__ float& sc(int i) { return x[i]; }
  __device__ __forceinline__ float& ring(uint32_t r, uint32_t pos) const {
    return rings[(size_t)(ring_tab[r].offset + pos) * (size_t)Vp + (size_t)v];
  }
  __device__ __forceinline__ int first() { return 0; }
  __device__ __forceinline__ bool more(int k) const { return k < 1; }
  __device__ __forceinline__ int next(int k) { return k + 1; }
  __device__ __forceinline__ uint32_t sample() const { return 0u; }
  __device__ __forceinline__ uint32_t count() const { return 1u; }
};
__device__ __forceinline__ uint32_t ring_len(const RegLane& L, uint32_t r) { return L.ring_tab[r].length; }
// nested-net resets and per-lane FFTs never reach a specialised kernel (spec.cpp refuses those tapes); exec() still has to
// compile its dead cases
__device__ __forceinline__ void reset_range(const RegLane&, uint32_t) {}
__device__ __forceinline__ void lane_fft(const RegLane&, uint32_t, uint32_t, int, const float*, bool) {}

template <int i>
__device__ __forceinline__ void spec_run(RegLane& L, int& pc) {
  if constexpr (i < SPEC_N) {
    constexpr Instr I = kTape[i];
    exec(I, L, pc);
    spec_run<i + 1>(L, pc);
  }
}

#if defined(QG_SPEC_PREFETCH)
// EXPERIMENTAL (QG_SPEC_PREFETCH=1 in the environment at qg_bank_set_path time; not yet run on hardware): delay lines
// without a dependent HBM load per sample.  A delay op's ring position is a counter, and the reads of 8 consecutive samples
// do not depend on the writes of those samples when the ring is at least 8 long (the rule of K1b's whole-block access):
// per block of 8 samples the kernel issues all ring reads of every such delay op up front (independent loads), runs the 8
// samples out of registers, then stores the 8 writes.
constexpr int SPEC_BT = 8;
__host__ __device__ constexpr bool spec_is_pref(int i) { return kTape[i].op == OP_DELAY && kRingLen[kTape[i].aux] >= (uint32_t)SPEC_BT; }
__host__ __device__ constexpr int spec_pref_before(int upto) { int c = 0; for (int k = 0; k < upto; k++) c += spec_is_pref(k) ? 1 : 0; return c; }
constexpr int SPEC_ND = spec_pref_before(SPEC_N);

struct DelayBlock {
  float rd[SPEC_ND > 0 ? SPEC_ND : 1][SPEC_BT];
  float wr[SPEC_ND > 0 ? SPEC_ND : 1][SPEC_BT];
};

// j is a template parameter so that every rd / wr subscript is a literal (registers)
template <int i, int j>
__device__ __forceinline__ void spec_run_blk(RegLane& L, DelayBlock& D, int& pc) {
  if constexpr (i < SPEC_N) {
    constexpr Instr I = kTape[i];
    if constexpr (spec_is_pref(i)) {
      constexpr int d = spec_pref_before(i);
      D.wr[d][j] = L.x[I.in[0]];
      L.x[I.out] = D.rd[d][j];
    } else {
      exec(I, L, pc);
    }
    spec_run_blk<i + 1, j>(L, D, pc);
  }
}
template <int i>
__device__ __forceinline__ void spec_ring_reads(RegLane& L, DelayBlock& D, int n) {
  if constexpr (i < SPEC_N) {
    if constexpr (spec_is_pref(i)) {
      constexpr Instr I = kTape[i];
      constexpr int d = spec_pref_before(i);
      constexpr uint32_t len = kRingLen[I.aux];
      const uint32_t idx = __float_as_uint(L.x[I.s]);
      float* const rb = L.rings + (size_t)L.ring_tab[I.aux].offset * (size_t)L.Vp + (size_t)L.v;
#pragma unroll
      for (int j = 0; j < SPEC_BT; j++)
        if (j < n) D.rd[d][j] = rb[(size_t)ring_wrap(idx + (uint32_t)j, len) * (size_t)L.Vp];
    }
    spec_ring_reads<i + 1>(L, D, n);
  }
}
template <int i>
__device__ __forceinline__ void spec_ring_writes(RegLane& L, DelayBlock& D, int n) {
  if constexpr (i < SPEC_N) {
    if constexpr (spec_is_pref(i)) {
      constexpr Instr I = kTape[i];
      constexpr int d = spec_pref_before(i);
      constexpr uint32_t len = kRingLen[I.aux];
      const uint32_t idx = __float_as_uint(L.x[I.s]);
      float* const rb = L.rings + (size_t)L.ring_tab[I.aux].offset * (size_t)L.Vp + (size_t)L.v;
#pragma unroll
      for (int j = 0; j < SPEC_BT; j++)
        if (j < n) rb[(size_t)ring_wrap(idx + (uint32_t)j, len) * (size_t)L.Vp] = D.wr[d][j];
      L.x[I.s] = __uint_as_float(ring_wrap(idx + (uint32_t)n, len));
    }
    spec_ring_writes<i + 1>(L, D, n);
  }
}
#endif

extern "C" __global__ void __launch_bounds__(128) k_spec(InterpArgs a) {
  const int nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  float* tiles = QG_SMEM_F;                                  // [n_out][nwarps][32][33]
  const int v = blockIdx.x * nt + tid;                       // padded voice index, always < Vp
  RegLane L;
  L.v = v; L.Vp = a.Vp; L.rings = a.rings; L.ring_tab = a.ring_tab; L.tables = a.tables;
#pragma unroll
  for (int p = 0; p < SPEC_P; p++) L.x[p] = a.params[(size_t)p * a.Vp + v];
#pragma unroll
  for (int s = 0; s < SPEC_NS; s++) L.x[SPEC_P + s] = a.state[(size_t)s * a.Vp + v];
#pragma unroll
  for (int k = 0; k < SPEC_NT; k++) L.x[SPEC_P + SPEC_NS + k] = 0.0f;

  const int warp_v0 = blockIdx.x * nt + warp * 32;
#if defined(QG_SPEC_PREFETCH)
  DelayBlock D;
#endif
  for (long t = 0; t < a.T; t++) {
#pragma unroll
    for (int c = 0; c < SPEC_NIN; c++) {
      size_t idx = a.in_frame_major ? ((size_t)t * a.V + v) * SPEC_NIN + c : ((size_t)v * SPEC_NIN + c) * a.T + t;
      L.x[SPEC_P + SPEC_NS + c] = v < a.V ? a.in[idx] : 0.0f;
    }
    int pc = 0;
#if defined(QG_SPEC_PREFETCH)
    {
      const int j = (int)(t & (SPEC_BT - 1));
      const int n = (a.T - (t - j)) < SPEC_BT ? (int)(a.T - (t - j)) : SPEC_BT;      // samples in this block
      if (j == 0) spec_ring_reads<0>(L, D, n);
      switch (j) {
        case 0: spec_run_blk<0, 0>(L, D, pc); break;
        case 1: spec_run_blk<0, 1>(L, D, pc); break;
        case 2: spec_run_blk<0, 2>(L, D, pc); break;
        case 3: spec_run_blk<0, 3>(L, D, pc); break;
        case 4: spec_run_blk<0, 4>(L, D, pc); break;
        case 5: spec_run_blk<0, 5>(L, D, pc); break;
        case 6: spec_run_blk<0, 6>(L, D, pc); break;
        default: spec_run_blk<0, 7>(L, D, pc); break;
      }
      if (j == n - 1) spec_ring_writes<0>(L, D, n);
    }
#else
    spec_run<0>(L, pc);
#endif
    // ---- outputs: same staging and the same left-to-right group mix as k_interp
    if (a.out_frame_major) {
      if (v < a.V) {
#pragma unroll
        for (int c = 0; c < SPEC_NOUT; c++) a.out[((size_t)t * a.V + v) * SPEC_NOUT + c] = L.x[kOutX[c]];
      }
    } else {
      const int tt = (int)(t & 31);
#pragma unroll
      for (int c = 0; c < SPEC_NOUT; c++) tiles[(((size_t)c * nwarps + warp) * 32 + lane) * 33 + tt] = L.x[kOutX[c]];
      if (tt == 31 || t == a.T - 1) {
        __syncwarp();
        const long t_base = t - tt;
        const int ncols = tt + 1;
        for (int c = 0; c < SPEC_NOUT; c++) {
          const float* tile = tiles + ((size_t)c * nwarps + warp) * 32 * 33;
          if (a.group <= 1) {
            for (int r = 0; r < 32; r++) {
              int vv = warp_v0 + r;
              if (vv < a.V && lane < ncols) a.out[((size_t)vv * SPEC_NOUT + c) * a.T + t_base + lane] = tile[r * 33 + lane];
            }
          } else {
            const int G = a.group;
            const float inv = 1.0f / (float)G;
            for (int g0 = 0; g0 < 32; g0 += G) {
              int gi = (warp_v0 + g0) / G;
              if (warp_v0 + g0 + G <= a.V && lane < ncols) {
                float acc = tile[g0 * 33 + lane];
                for (int r = 1; r < G; r++) acc += tile[(g0 + r) * 33 + lane];
                a.out[((size_t)gi * SPEC_NOUT + c) * a.T + t_base + lane] = acc * inv;
              }
            }
          }
        }
        __syncwarp();
      }
    }
  }
#pragma unroll
  for (int s = 0; s < SPEC_NS; s++) a.state[(size_t)s * a.Vp + v] = L.x[SPEC_P + s];
}

}  // namespace qg
