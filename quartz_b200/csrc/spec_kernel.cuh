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
  for (long t = 0; t < a.T; t++) {
#pragma unroll
    for (int c = 0; c < SPEC_NIN; c++) {
      size_t idx = a.in_frame_major ? ((size_t)t * a.V + v) * SPEC_NIN + c : ((size_t)v * SPEC_NIN + c) * a.T + t;
      L.x[SPEC_P + SPEC_NS + c] = v < a.V ? a.in[idx] : 0.0f;
    }
    int pc = 0;
    spec_run<0>(L, pc);
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
