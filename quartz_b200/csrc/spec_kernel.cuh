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

// Delay lines without a dependent HBM load per sample.  A delay op's ring position is a counter, and the reads of 8
// consecutive samples do not depend on the writes of those samples when the ring is at least 8 long (the rule of K1b's
// whole-block access): per block of 8 samples the kernel issues all ring reads of every such delay op up front (8
// independent 128-byte lines per warp and op), runs the 8 samples out of registers, then stores the 8 writes.
// Measured without it (round 1): one dependent load per sample made configs[4]'s delay archetype 6x SLOWER than K1b.
#if !defined(QG_SPEC_BT_DELAY)
#define QG_SPEC_BT_DELAY 8   // samples per block when a delay line is prefetched (measured on configs[4] d: 4 / 8 / 16 in profiles/README.md)
#endif
__host__ __device__ constexpr bool spec_is_pref(int i) { return kTape[i].op == OP_DELAY && kRingLen[kTape[i].aux] >= (uint32_t)QG_SPEC_BT_DELAY; }
__host__ __device__ constexpr int spec_pref_before(int upto) { int c = 0; for (int k = 0; k < upto; k++) c += spec_is_pref(k) ? 1 : 0; return c; }
constexpr int SPEC_ND = spec_pref_before(SPEC_N);
// samples per block: 8 when a delay line is prefetched, else 1 (the tape is instantiated once per sample of a block, and
// NVRTC's compile time grows with it)
constexpr int SPEC_BT = SPEC_ND > 0 ? QG_SPEC_BT_DELAY : 1;

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

struct SpecOut {
  int tile0;        // float index (into the CTA's shared memory) of this lane's row in channel 0's tile: ((warp * 32) + lane) * 33
  int chan_stride;  // floats between the tiles of consecutive output channels: nwarps * 32 * 33
  int warp_tile0;   // float index of row 0 of this warp's tile for channel 0
  int lane, warp_v0;
};

// one sample (index j of its block): inputs, the tape, outputs parked in the warp's 32 x 33 tile (or stored, frame-major).
// `tt` (the sample's column in the tile) is a 32-bit value and the tile is addressed by index: the per-sample code of a
// one-sample feedback voice is 16 instructions, and the loop around it used to cost twice that in 64-bit counters and
// re-materialised shared-memory addresses (profiles/README.md, round 2).  The tape is instantiated ONCE per sample of a
// block (NVRTC's compile time grows with every copy): the layout test and the tail test `j < n` stay run-time tests.
// frame-major output [T][V][outputs] (the interleave of audio.rs:113-117; single graphs and small banks): out of line, the
// voice-major loop pays one uniform test for it
__device__ __noinline__ void spec_store_fm(const InterpArgs& a, const SpecOut& o, int v, long t, int tt) {
  if (v >= a.V) return;
  for (int c = 0; c < SPEC_NOUT; c++) a.out[((size_t)t * a.V + v) * SPEC_NOUT + c] = QG_SMEM_F[o.tile0 + c * o.chan_stride + tt];
}
template <int j>
__device__ __forceinline__ void spec_sample(const InterpArgs& a, RegLane& L, DelayBlock& D, const SpecOut& o, long t, int tt, bool FM) {
  const int v = L.v;
#pragma unroll
  for (int c = 0; c < SPEC_NIN; c++) {
    size_t idx = a.in_frame_major ? ((size_t)t * a.V + v) * SPEC_NIN + c : ((size_t)v * SPEC_NIN + c) * a.T + t;
    L.x[SPEC_P + SPEC_NS + c] = v < a.V ? a.in[idx] : 0.0f;
  }
  int pc = 0;
  spec_run_blk<0, j>(L, D, pc);
  // voice-major (the bulk layout): park the outputs in the tile; frame-major stores them from the same slots, out of line
#pragma unroll
  for (int c = 0; c < SPEC_NOUT; c++) QG_SMEM_F[o.tile0 + c * o.chan_stride + tt] = L.x[kOutX[c]];
  if (FM) spec_store_fm(a, o, v, t, tt);
}
// voice-major outputs: the tile holds the `ncols` samples that start at t_base; same staging and the same left-to-right group
// mix as k_interp
__device__ __noinline__ void spec_flush(const InterpArgs& a, const SpecOut& o, long t_base, int ncols) {
  __syncwarp();
  for (int c = 0; c < SPEC_NOUT; c++) {
    const int tile = o.warp_tile0 + c * o.chan_stride;
    if (a.group <= 1) {
      for (int r = 0; r < 32; r++) {
        int vv = o.warp_v0 + r;
        if (vv < a.V && o.lane < ncols) a.out[((size_t)vv * SPEC_NOUT + c) * a.T + t_base + o.lane] = QG_SMEM_F[tile + r * 33 + o.lane];
      }
    } else {
      const int G = a.group;
      const float inv = 1.0f / (float)G;
      for (int g0 = 0; g0 < 32; g0 += G) {
        int gi = (o.warp_v0 + g0) / G;
        if (o.warp_v0 + g0 + G <= a.V && o.lane < ncols) {
          float acc = QG_SMEM_F[tile + g0 * 33 + o.lane];
          for (int r = 1; r < G; r++) acc += QG_SMEM_F[tile + (g0 + r) * 33 + o.lane];
          a.out[((size_t)gi * SPEC_NOUT + c) * a.T + t_base + o.lane] = acc * inv;
        }
      }
    }
  }
  __syncwarp();
}
// a block of SPEC_BT samples whose first sample is t0 = tile column tt0
template <int j>
__device__ __forceinline__ void spec_block(const InterpArgs& a, RegLane& L, DelayBlock& D, const SpecOut& o, long t0, int tt0, int n, bool FM) {
  if constexpr (j < SPEC_BT) {
    if (SPEC_BT == 1 || j < n) spec_sample<j>(a, L, D, o, t0 + j, tt0 + j, FM);      // one-sample blocks are always full
    spec_block<j + 1>(a, L, D, o, t0, tt0, n, FM);
  }
}

__device__ __forceinline__ void spec_render(const InterpArgs& a, RegLane& L, const SpecOut& o) {
  const bool FM = a.out_frame_major != 0;
  DelayBlock D;
  // tiles of 32 samples (one flush each); inside a tile, blocks of SPEC_BT samples (a divisor of 32) with 32-bit counters
  for (long tb = 0; tb < a.T; tb += 32) {
    const int nb = a.T - tb < 32 ? (int)(a.T - tb) : 32;
    for (int j0 = 0; j0 < nb; j0 += SPEC_BT) {
      const int n = nb - j0 < SPEC_BT ? nb - j0 : SPEC_BT;
      spec_ring_reads<0>(L, D, n);
      spec_block<0>(a, L, D, o, tb + j0, j0, n, FM);
      spec_ring_writes<0>(L, D, n);
    }
    if (!FM) spec_flush(a, o, tb, nb);
  }
}

#if !defined(QG_SPEC_MINB)
#define QG_SPEC_MINB 8   // 64 registers: measured on configs[4] (a 192 -> 189, d 83 -> 67 ms; 10 blocks spill the delay kernel: 88 ms)
#endif
extern "C" __global__ void __launch_bounds__(128, QG_SPEC_MINB) k_spec(InterpArgs a) {
  const int nt = blockDim.x, tid = threadIdx.x;
  SpecOut o;
  o.lane = tid & 31;
  const int warp = tid >> 5, nwarps = nt >> 5;
  o.chan_stride = nwarps * 32 * 33;                          // tiles: [n_out][nwarps][32][33]
  o.warp_tile0 = warp * 32 * 33;
  o.tile0 = o.warp_tile0 + o.lane * 33;
  o.warp_v0 = blockIdx.x * nt + warp * 32;
  const int v = blockIdx.x * nt + tid;                       // padded voice index, always < Vp
  RegLane L;
  L.v = v; L.Vp = a.Vp; L.rings = a.rings; L.ring_tab = a.ring_tab; L.tables = a.tables;
#pragma unroll
  for (int p = 0; p < SPEC_P; p++) L.x[p] = a.params[(size_t)p * a.Vp + v];
#pragma unroll
  for (int s = 0; s < SPEC_NS; s++) L.x[SPEC_P + s] = a.state[(size_t)s * a.Vp + v];
#pragma unroll
  for (int k = 0; k < SPEC_NT; k++) L.x[SPEC_P + SPEC_NS + k] = 0.0f;
  spec_render(a, L, o);
#pragma unroll
  for (int s = 0; s < SPEC_NS; s++) a.state[(size_t)s * a.Vp + v] = L.x[SPEC_P + s];
}

}  // namespace qg
