// Hand-fused kernels for hot tape shapes (see fused.h).
//
// K2  k_noise_svf_scan — `white() >> <fixed linear filter>` banks (BASELINE configs[1]): ONE WARP PER VOICE, time-parallel.
//     A filter with fixed coefficients is LTI:  s' = A s + B x,  y = C s + D x  with at most two scanned state words
//     (SVF: ic1, ic2; direct-form biquad: y1, y2; one-poles: y1 — the filter family is a template parameter).  A block of
//     32*K consecutive samples (K = 32) is split over the 32 lanes (K contiguous samples each = one 128-byte output row):
//       1. every lane runs its K samples from zero filter state (noise is counter-based, so lane j just starts its
//          counter at base + j*K and recomputes the two inputs before its chunk) and parks the K zero-state outputs in a
//          shared tile laid out as the TMA box (SWIZZLE_128B, conflict-free float4 accesses);
//       2. the per-lane end states are combined with a 5-step warp-shuffle scan of the affine maps
//          s -> A^K s + c_j  (Kogge-Stone, matrices A^K, A^2K, ... A^16K precomputed per voice in f64);
//       3. every lane adds the homogeneous response (C A^i) s_start to its K outputs IN PLACE (rows C A^i precomputed);
//       4. the 32*K outputs (one contiguous 4 KB run of the voice's row) leave with ONE TMA tensor store
//          (cp.async.bulk.tensor.2d -> UTMASTG).
//     Small banks use S time segments per voice: a state-only pre-pass computes each segment's zero-state end
//     state, the segment start states are chained (f64 square-and-multiply of A^len), then every segment renders.
//     Arithmetic uses explicit FMAs and a re-associated recurrence: parity is the f32 audio tolerance
//     (<= 1e-4 abs, <= -90 dBFS), demonstrated at full length and on extreme parameters in tests/test_gpu_fused.py;
//     noise samples and the persisted counter are bit-exact.  Ill-conditioned direct-form biquads are refused by the
//     bank (capi.cu: biquads_well_conditioned) and run in the reference's operation order instead.
#include "fused.h"

#include <cuda.h>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_math.cuh"
#include "tape.h"

namespace qg {

namespace {

// K = samples per lane per block, B = 32*K samples per warp block.  K = 32: one lane = one 128-byte row of the tile,
// scan and loop overhead amortised over 1,024 samples.
constexpr int K2_K = 32;

__device__ __forceinline__ void bulk_store_block(float* gdst, const float* ssrc, unsigned bytes) {
  // one elected lane: shared -> global bulk async copy
  unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
// TMA tensor store of one K-row x 128-byte box (SWIZZLE_128B layout in shared memory) -> SASS UTMASTG
__device__ __forceinline__ void tma_store_box(const CUtensorMap* tmap, const void* ssrc, int row0) {
  unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(tmap), "r"(0), "r"(row0), "r"(s)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

struct SvfC { float a1, a2, a3, m0, m1, m2; };

// zero-state / any-state SVF tick with explicit FMAs; returns the mode output.
// LP: the lowpass mix (m0, m1, m2) = (0, 0, 1) is known at plan time, the output is v2 itself.
template <bool LP = false>
__device__ __forceinline__ float svf_fma(float x, float& ic1, float& ic2, const SvfC& c) {
  float v3 = x - ic2;
  float v1 = __fmaf_rn(c.a2, v3, c.a1 * ic1);
  float v2 = __fmaf_rn(c.a3, v3, __fmaf_rn(c.a2, ic1, ic2));
  ic1 = __fmaf_rn(2.0f, v1, -ic1);
  ic2 = __fmaf_rn(2.0f, v2, -ic2);
  if (LP) return v2;
  return __fmaf_rn(c.m2, v2, __fmaf_rn(c.m1, v1, c.m0 * x));
}

// same tick fed with the un-scaled noise integer hf (x = hf * 2^-31, an exact power-of-two scaling): the scaling rides on
// the FMAs that consume x.  Lowpass needs no v1: ic1' = 2 v1 - ic1 = (2 a1 - 1) ic1 + 2 a2 v3 with c11 = 2 a1 - 1 and
// c12 = 2 a2 both exact in f32 (a1 in [0.5, 1)) -> 6 FP ops per sample instead of 7.
struct SvfK { float c11, c12, m0s; };
template <bool LP>
__device__ __forceinline__ float svf_fma_noise(float hf, float& ic1, float& ic2, const SvfC& c, const SvfK& k) {
  const float sc = 1.0f / 2147483648.0f;
  float v3 = __fmaf_rn(hf, sc, -ic2);
  if (LP) {
    float v2 = __fmaf_rn(c.a3, v3, __fmaf_rn(c.a2, ic1, ic2));
    ic1 = __fmaf_rn(k.c12, v3, k.c11 * ic1);
    ic2 = __fmaf_rn(2.0f, v2, -ic2);
    return v2;
  }
  float v1 = __fmaf_rn(c.a2, v3, c.a1 * ic1);
  float v2 = __fmaf_rn(c.a3, v3, __fmaf_rn(c.a2, ic1, ic2));
  ic1 = __fmaf_rn(2.0f, v1, -ic1);
  ic2 = __fmaf_rn(2.0f, v2, -ic2);
  return __fmaf_rn(c.m2, v2, __fmaf_rn(c.m1, v1, k.m0s * hf));
}
// (measured: replacing the hash's SHF with IMAD.HI to unload the half-rate ALU pipe is slower — 13.4 vs 11.8 ms)
__device__ __forceinline__ float noise_int(uint32_t counter) { return (float)(int32_t)d_hash32x(counter); }
__device__ __forceinline__ uint32_t __umad_opaque(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
// d_hash32x with the caller supplying hi = x >> 16
__device__ __forceinline__ float noise_int_hi(uint32_t x, uint32_t hi) {
  x = (x ^ hi) * 0x21f0aaadU;
  x = (x ^ (x >> 15)) * 0x735a2d97U;
  return (float)(int32_t)(x ^ (x >> 15));
}

// ---- filter families K2 serves (template parameter F).  Every one is a linear recurrence with at most two state words
// that are scanned; the previous INPUT samples some of them use (biquad x1, x2; one-poles x1) are not state for the scan:
// noise is counter-based, so a lane recomputes the two samples before its chunk.
//   F_SVF / F_SVF_LP  Simper SVF, state (ic1, ic2); LP = lowpass mix known at plan time (6 FP ops per tick)
//   F_BIQUAD          direct form I (FunDSP Biquad): y0 = b0 x0 + b1 x1 + b2 x2 - a1 y1 - a2 y2, scanned state (y1, y2)
//   F_ONEPOLE         lowpole / highpole / dcblock / allpole (kind 0..3, as OP_ONEPOLE), scanned state (y1)
enum : int { F_SVF = 0, F_SVF_LP = 1, F_BIQUAD = 2, F_ONEPOLE = 3 };
struct FiltC {
  SvfC c;          // SVF: a1 a2 a3 m0 m1 m2 ; BIQUAD: a1 a2 (c.a1, c.a2), b0 b1 b2 scaled by 2^-31 (c.a3, c.m0, c.m1) ; ONEPOLE: coeff (c.a1)
  SvfK k;          // SVF only
  int kind;        // ONEPOLE kind
};
// one tick fed with the un-scaled noise integer hf (x = hf * 2^-31); (s1, s2) scanned state, (h1, h2) previous hf values
template <int F>
__device__ __forceinline__ float filt_tick(float hf, float& s1, float& s2, float& h1, float& h2, const FiltC& f) {
  const float sc = 1.0f / 2147483648.0f;
  if (F == F_SVF) return svf_fma_noise<false>(hf, s1, s2, f.c, f.k);
  if (F == F_SVF_LP) return svf_fma_noise<true>(hf, s1, s2, f.c, f.k);
  if (F == F_BIQUAD) {
    const float fb = __fmaf_rn(f.c.a1, s1, f.c.a2 * s2);
    const float y0 = __fmaf_rn(f.c.a3, hf, __fmaf_rn(f.c.m0, h1, __fmaf_rn(f.c.m1, h2, -fb)));
    h2 = h1; h1 = hf; s2 = s1; s1 = y0;
    return y0;
  }
  const float c = f.c.a1;
  float y;
  switch (f.kind) {
    case 0: y = __fmaf_rn(c, s1, ((1.0f - c) * sc) * hf); break;                    // (1 - c) x + c y1
    case 1: y = c * __fmaf_rn(sc, hf - h1, s1); break;                              // c (y1 + x - x1)
    case 2: y = __fmaf_rn(c, s1, sc * (hf - h1)); break;                            // x - x1 + c y1
    default: y = __fmaf_rn(c, __fmaf_rn(sc, hf, -s1), sc * h1); break;              // c (x - y1) + x1
  }
  h1 = hf; s1 = y;
  return y;
}
// state transition matrix A and the row (c1, c2) with y_h(i) = (c1, c2) A^i s_start for the homogeneous output response
template <int F>
__device__ __forceinline__ void filt_matrices(const FiltC& f, double& A11, double& A12, double& A21, double& A22, double& c1, double& c2) {
  if (F == F_SVF || F == F_SVF_LP) {
    const double a1 = f.c.a1, a2 = f.c.a2, a3 = f.c.a3;
    A11 = 2 * a1 - 1; A12 = -2 * a2; A21 = 2 * a2; A22 = 1 - 2 * a3;
    // y_h = m1*v1_h + m2*v2_h with v1_h = a1 s1 - a2 s2, v2_h = a2 s1 + (1 - a3) s2
    c1 = (double)f.c.m1 * a1 + (double)f.c.m2 * a2; c2 = -(double)f.c.m1 * a2 + (double)f.c.m2 * (1 - a3);
  } else if (F == F_BIQUAD) {
    A11 = -(double)f.c.a1; A12 = -(double)f.c.a2; A21 = 1; A22 = 0;
    c1 = A11; c2 = A12;                                                             // y0 is the first component of A s
  } else {
    A11 = f.kind == 3 ? -(double)f.c.a1 : (double)f.c.a1; A12 = 0; A21 = 0; A22 = 0;
    c1 = A11; c2 = 0;
  }
}
template <int F>
__device__ __forceinline__ FiltC filt_load(const float* __restrict__ params, int Vp, int v, int p0, int kind) {
  FiltC f;
  f.kind = kind;
  const float sc = 1.0f / 2147483648.0f;
#define QG_P(i) params[(size_t)(p0 + (i)) * Vp + v]
  if (F == F_SVF || F == F_SVF_LP) {
    f.c.a1 = QG_P(0); f.c.a2 = QG_P(1); f.c.a3 = QG_P(2); f.c.m0 = QG_P(3); f.c.m1 = QG_P(4); f.c.m2 = QG_P(5);
    f.k.c11 = 2.0f * f.c.a1 - 1.0f; f.k.c12 = 2.0f * f.c.a2; f.k.m0s = f.c.m0 * sc;
  } else if (F == F_BIQUAD) {
    f.c.a1 = QG_P(0); f.c.a2 = QG_P(1); f.c.a3 = QG_P(2) * sc; f.c.m0 = QG_P(3) * sc; f.c.m1 = QG_P(4) * sc; f.c.m2 = 0.0f;
    f.k.c11 = f.k.c12 = f.k.m0s = 0.0f;
  } else {
    f.c.a1 = QG_P(0); f.c.a2 = f.c.a3 = f.c.m0 = f.c.m1 = f.c.m2 = 0.0f;
    f.k.c11 = f.k.c12 = f.k.m0s = 0.0f;
  }
#undef QG_P
  return f;
}
// where the scanned state words and the input history live in the bank's state table (relative to the op's first state word)
template <int F> struct FiltState {
  static constexpr int s1 = F == F_BIQUAD ? 2 : (F == F_ONEPOLE ? 1 : 0);   // first scanned word
  static constexpr int n_s = F == F_ONEPOLE ? 1 : 2;                          // scanned words
  static constexpr int n_h = F == F_BIQUAD ? 2 : (F == F_ONEPOLE ? 1 : 0);    // history words (x1[, x2]) at offset 0
};

// Per-warp shared memory.  `tile` holds the block's 32*K samples as float4 chunks: the zero-state outputs are parked in
// it, corrected IN PLACE and shipped from it, so a warp needs 128*K bytes (+ constants) and 28 warps fit on an SM.
//   STORE == 2: tile is laid out as the TMA box (K rows x 128 B, SWIZZLE_128B): sample s of the block sits in row s/32,
//               16-byte chunk ((s%32)/4) ^ (row & 7) — conflict-free for the per-lane float4 accesses;
//   STORE <= 1: tile is [K/4][32 lanes] float4 (conflict-free); STORE == 1 stages the corrected block linearly in `lin`
//               for the bulk copy (bank conflicts on that path — it only serves rows that are not 128-byte aligned).
template <int K, int STORE>
struct __align__(1024) WarpSmem {
  float4 tile[K / 4 * 32];
  float4 mp[5];          // A^(K*2^i) as (a, b, c, d)
  float4 rc[K / 2];      // homogeneous response rows C*A^i: (r1_i, r2_i, r1_{i+1}, r2_{i+1})
  float lin[STORE == 1 ? 32 * K : 4];
};

// MODE 0: write samples; MODE 1: state-only pre-pass (zero start state, no output) for segment chaining
// STORE (decided on the host from the output alignment): 0 plain scalar stores; 1 linear bulk async copy of the block
// (rows 16-byte aligned); 2 TMA tensor store from the 128-byte-swizzled tile (T % 32 == 0).
template <int MODE, int F, int STORE, int K>
__global__ void __launch_bounds__(128, 7) k_noise_svf_scan(const float* __restrict__ params, float* __restrict__ state, int Vp,
                                                           int V, long T, int S, long seg_len, int p_svf, int s_noise,
                                                           int s_svf, int kind, float* __restrict__ out,
                                                           float* __restrict__ seg_state, const __grid_constant__ CUtensorMap tmap) {
  constexpr int B = 32 * K;
  using FS = FiltState<F>;
  __shared__ WarpSmem<K, MODE == 0 ? STORE : 0> sm[4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto& W = sm[warp];
  const long w = (long)blockIdx.x * 4 + warp;        // warp id = voice * S + segment
  const int v = (int)(w / S), seg = (int)(w % S);
  if (v >= V) return;
  const FiltC fc = filt_load<F>(params, Vp, v, p_svf, kind);
  // state-space form of the tick:  s' = A s + B x ; y = C s + D x
  double A11, A12, A21, A22, c1, c2;
  filt_matrices<F>(fc, A11, A12, A21, A22, c1, c2);
  // per-voice constants in f64, rounded once, kept in shared memory (read as broadcasts):
  //   rc[i] = C * A^i  (the homogeneous output response i samples after a block start state), i < K
  //   mp[i] = A^(K*2^i), i = 0..4  (scan matrices)
  if (lane == 0) {
    double m11 = 1, m12 = 0, m21 = 0, m22 = 1;
    float* rcf = reinterpret_cast<float*>(W.rc);
    for (int i = 0; i < K; i++) {
      rcf[2 * i] = (float)(c1 * m11 + c2 * m21);
      rcf[2 * i + 1] = (float)(c1 * m12 + c2 * m22);
      double n11 = A11 * m11 + A12 * m21, n12 = A11 * m12 + A12 * m22, n21 = A21 * m11 + A22 * m21, n22 = A21 * m12 + A22 * m22;
      m11 = n11; m12 = n12; m21 = n21; m22 = n22;
    }
    for (int i = 0; i < 5; i++) {
      W.mp[i] = make_float4((float)m11, (float)m12, (float)m21, (float)m22);
      double n11 = m11 * m11 + m12 * m21, n12 = m11 * m12 + m12 * m22, n21 = m21 * m11 + m22 * m21, n22 = m21 * m12 + m22 * m22;
      m11 = n11; m12 = n12; m21 = n21; m22 = n22;
    }
  }
  __syncwarp();
  const long t_begin = (long)seg * seg_len;
  const long t_end = (t_begin + seg_len < T) ? t_begin + seg_len : T;
  const uint32_t counter0 = __float_as_uint(state[(size_t)s_noise * Vp + v]);
  float S1, S2;   // start state of the current block (warp-uniform)
  if (MODE == 1 || seg > 0) {
    if (MODE == 1) { S1 = 0.0f; S2 = 0.0f; }
    else { S1 = seg_state[((size_t)v * (S + 1) + seg) * 2]; S2 = seg_state[((size_t)v * (S + 1) + seg) * 2 + 1]; }
  } else { S1 = state[(size_t)(s_svf + FS::s1) * Vp + v]; S2 = FS::n_s > 1 ? state[(size_t)(s_svf + FS::s1 + 1) * Vp + v] : 0.0f; }
  // input history of the voice's first sample of this launch comes from the persisted state (x = hf * 2^-31 exactly)
  const float xp1 = FS::n_h >= 1 ? state[(size_t)s_svf * Vp + v] * 2147483648.0f : 0.0f;
  const float xp2 = FS::n_h >= 2 ? state[(size_t)(s_svf + 1) * Vp + v] * 2147483648.0f : 0.0f;
  // hf of the two samples before absolute sample index ta (counter of sample ta is counter0 + ta + 1)
  auto history = [&](long ta, float& h1, float& h2) {
    if (FS::n_h == 0) { h1 = 0.0f; h2 = 0.0f; return; }
    h1 = ta >= 1 ? noise_int(counter0 + (uint32_t)ta) : xp1;
    h2 = FS::n_h < 2 ? 0.0f : (ta >= 2 ? noise_int(counter0 + (uint32_t)ta - 1u) : (ta == 1 ? xp1 : xp2));
  };
  float* orow = MODE == 0 ? out + (size_t)v * T : nullptr;
  constexpr bool can_bulk = MODE == 0 && STORE >= 1;
  constexpr bool swz = MODE == 0 && STORE == 2;
  // float4 slot of this lane's chunk i4 (samples lane*K + 4*i4 .. +3 of the block) inside the tile
  auto slot = [&](int i4) -> int {
    if (!swz) return i4 * 32 + lane;
    const int s4 = lane * (K / 4) + i4, row = s4 >> 3, ch = s4 & 7;
    return row * 8 + (ch ^ (row & 7));
  };

  const uint32_t one = S > 0 ? 1u : 0u;   // always 1; not a compile-time constant
  long t = t_begin;
  for (; t + B <= t_end; t += B) {
    if (swz) {   // the previous block's tensor store must have finished reading the tile
      if (lane == 0) bulk_wait_read_0();
      __syncwarp();
    }
    // ---- 1. zero-state run of this lane's K samples, outputs parked in the tile.  The hash's first step x ^ (x >> 16)
    // needs x >> 16, which is constant over the chunk unless its counters cross a multiple of 65,536 (one block in 64):
    // the common case hoists it (one SHF less per sample), the crossing block takes the plain path (warp-uniform branch)
    float e1 = 0.0f, e2 = 0.0f;
    {
      float h1, h2;
      history(t + (long)lane * K, h1, h2);
      const uint32_t base = counter0 + (uint32_t)t + (uint32_t)(lane * K) + 1u;
      const bool cross = (base & 0xffffu) > (0xffffu - (uint32_t)(K - 1));
      if (!__any_sync(0xffffffffu, cross)) {
        const uint32_t hi = base >> 16;
        // counter + j as IMAD (one * j + base, `one` opaque to ptxas): the add moves from the half-rate ALU pipe, which the
        // hash's shifts and xors already load, to the FMA pipe
#define QG_CTR(j) __umad_opaque(one, (uint32_t)(j), base)
#pragma unroll
        for (int i4 = 0; i4 < K / 4; i4++) {
          float4 y;
          y.x = filt_tick<F>(noise_int_hi(QG_CTR(4 * i4 + 0), hi), e1, e2, h1, h2, fc);
          y.y = filt_tick<F>(noise_int_hi(QG_CTR(4 * i4 + 1), hi), e1, e2, h1, h2, fc);
          y.z = filt_tick<F>(noise_int_hi(QG_CTR(4 * i4 + 2), hi), e1, e2, h1, h2, fc);
          y.w = filt_tick<F>(noise_int_hi(QG_CTR(4 * i4 + 3), hi), e1, e2, h1, h2, fc);
          if (MODE == 0) W.tile[slot(i4)] = y;
        }
#undef QG_CTR
      } else {
#pragma unroll 2
        for (int i4 = 0; i4 < K / 4; i4++) {
          float4 y;
          y.x = filt_tick<F>(noise_int(base + (uint32_t)(4 * i4 + 0)), e1, e2, h1, h2, fc);
          y.y = filt_tick<F>(noise_int(base + (uint32_t)(4 * i4 + 1)), e1, e2, h1, h2, fc);
          y.z = filt_tick<F>(noise_int(base + (uint32_t)(4 * i4 + 2)), e1, e2, h1, h2, fc);
          y.w = filt_tick<F>(noise_int(base + (uint32_t)(4 * i4 + 3)), e1, e2, h1, h2, fc);
          if (MODE == 0) W.tile[slot(i4)] = y;
        }
      }
    }
    // ---- 2. Kogge-Stone scan of the affine maps  e <- A^(K d) e(lane - d) + e ; lane 0 inherits the block start state
    {
      const float4 m = W.mp[0];
      const float s1 = lane == 0 ? S1 : 0.0f, s2 = lane == 0 ? S2 : 0.0f;
      e1 = __fmaf_rn(m.x, s1, __fmaf_rn(m.y, s2, e1));
      e2 = __fmaf_rn(m.z, s1, __fmaf_rn(m.w, s2, e2));
    }
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const int d = 1 << i;
      const float4 m = W.mp[i];
      float r1 = __shfl_up_sync(0xffffffffu, e1, d), r2 = __shfl_up_sync(0xffffffffu, e2, d);
      r1 = lane >= d ? r1 : 0.0f;
      r2 = lane >= d ? r2 : 0.0f;
      e1 = __fmaf_rn(m.x, r1, __fmaf_rn(m.y, r2, e1));
      e2 = __fmaf_rn(m.z, r1, __fmaf_rn(m.w, r2, e2));
    }
    float h1 = __shfl_up_sync(0xffffffffu, e1, 1), h2 = __shfl_up_sync(0xffffffffu, e2, 1);
    if (lane == 0) { h1 = S1; h2 = S2; }
    S1 = __shfl_sync(0xffffffffu, e1, 31);
    S2 = __shfl_sync(0xffffffffu, e2, 31);
    if (MODE == 0) {
      // ---- 3. homogeneous correction, in place
      if (STORE == 1) {
        if (lane == 0) bulk_wait_read_0();          // the previous block's copy has finished reading W.lin
        __syncwarp();
      }
#pragma unroll
      for (int i4 = 0; i4 < K / 4; i4++) {
        float4 y = W.tile[slot(i4)];
        const float4 ra = W.rc[2 * i4], rb = W.rc[2 * i4 + 1];
        y.x = __fmaf_rn(ra.x, h1, __fmaf_rn(ra.y, h2, y.x));
        y.y = __fmaf_rn(ra.z, h1, __fmaf_rn(ra.w, h2, y.y));
        y.z = __fmaf_rn(rb.x, h1, __fmaf_rn(rb.y, h2, y.z));
        y.w = __fmaf_rn(rb.z, h1, __fmaf_rn(rb.w, h2, y.w));
        if (swz) W.tile[slot(i4)] = y;
        else if (STORE == 1) reinterpret_cast<float4*>(&W.lin[lane * K])[i4] = y;
        else { float* g = orow + t + lane * K + 4 * i4; g[0] = y.x; g[1] = y.y; g[2] = y.z; g[3] = y.w; }
      }
      // ---- 4. one bulk async copy per block (UTMASTG / UBLKCP)
      if (can_bulk) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (swz) tma_store_box(&tmap, &W.tile[0], (int)(((long)v * T + t) >> 5));
          else bulk_store_block(orow + t, &W.lin[0], B * 4);
        }
      }
    }
  }
  // ---- tail (< B samples): sequential on every lane (redundant), lane 0 stores
  {
    float s1 = S1, s2 = S2, h1, h2;
    history(t, h1, h2);
    for (long tt = t; tt < t_end; tt++) {
      float yv = filt_tick<F>(noise_int(counter0 + (uint32_t)tt + 1u), s1, s2, h1, h2, fc);
      if (MODE == 0 && lane == 0) orow[tt] = yv;
    }
    S1 = s1; S2 = s2;
  }
  if (MODE == 1) {
    if (lane == 0) { seg_state[((size_t)v * (S + 1) + seg) * 2] = S1; seg_state[((size_t)v * (S + 1) + seg) * 2 + 1] = S2; }
  } else {
    if (can_bulk && lane == 0) bulk_wait_all();
    if (seg == S - 1 && lane == 0) {
      if (S == 1) {
        state[(size_t)(s_svf + FS::s1) * Vp + v] = S1;
        if (FS::n_s > 1) state[(size_t)(s_svf + FS::s1 + 1) * Vp + v] = S2;
        if (FS::n_h >= 1) {   // the last inputs of this launch become the persisted history
          float h1, h2;
          history(T, h1, h2);
          state[(size_t)s_svf * Vp + v] = h1 * (1.0f / 2147483648.0f);
          if (FS::n_h >= 2) state[(size_t)(s_svf + 1) * Vp + v] = h2 * (1.0f / 2147483648.0f);
        }
        state[(size_t)s_noise * Vp + v] = __uint_as_float(counter0 + (uint32_t)T);
      } else {
        // other segments of this voice may not have read the persisted state yet: publish through slot S,
        // k_finalize_segments copies it into the bank state after the render kernel
        seg_state[((size_t)v * (S + 1) + S) * 2] = S1;
        seg_state[((size_t)v * (S + 1) + S) * 2 + 1] = S2;
      }
    }
  }
}

template <int F>
__global__ void k_finalize_segments(float* __restrict__ state, int Vp, int V, long T, int S, int s_noise, int s_svf,
                                    const float* __restrict__ seg_state) {
  using FS = FiltState<F>;
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const uint32_t counter0 = __float_as_uint(state[(size_t)s_noise * Vp + v]);
  if (FS::n_h >= 1) {   // last inputs of the launch -> persisted history (x1[, x2]); T >= 2 whenever time is segmented
    state[(size_t)s_svf * Vp + v] = d_noise(counter0 + (uint32_t)T);
    if (FS::n_h >= 2) state[(size_t)(s_svf + 1) * Vp + v] = d_noise(counter0 + (uint32_t)T - 1u);
  }
  state[(size_t)(s_svf + FS::s1) * Vp + v] = seg_state[((size_t)v * (S + 1) + S) * 2];
  if (FS::n_s > 1) state[(size_t)(s_svf + FS::s1 + 1) * Vp + v] = seg_state[((size_t)v * (S + 1) + S) * 2 + 1];
  state[(size_t)s_noise * Vp + v] = __uint_as_float(counter0 + (uint32_t)T);
}

// Chain the segment start states: start_0 = persisted state, start_{g+1} = A^len_g start_g + zs_end_g.
// One thread per voice; A^len by square-and-multiply in f64.
template <int F>
__global__ void k_chain_segments(const float* __restrict__ params, const float* __restrict__ state, int Vp, int V, long T, int S,
                                 long seg_len, int p_svf, int s_svf, int kind, float* __restrict__ seg_state) {
  using FS = FiltState<F>;
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const FiltC fc = filt_load<F>(params, Vp, v, p_svf, kind);
  double A11, A12, A21, A22, c1, c2;
  filt_matrices<F>(fc, A11, A12, A21, A22, c1, c2);
  double s1 = state[(size_t)(s_svf + FS::s1) * Vp + v], s2 = FS::n_s > 1 ? state[(size_t)(s_svf + FS::s1 + 1) * Vp + v] : 0.0;
  for (int g = 0; g < S; g++) {
    float z1 = seg_state[((size_t)v * (S + 1) + g) * 2], z2 = seg_state[((size_t)v * (S + 1) + g) * 2 + 1];
    seg_state[((size_t)v * (S + 1) + g) * 2] = (float)s1;
    seg_state[((size_t)v * (S + 1) + g) * 2 + 1] = (float)s2;
    long len = (long)(g + 1) * seg_len <= T ? seg_len : T - (long)g * seg_len;
    if (len < 0) len = 0;
    double p11 = 1, p12 = 0, p21 = 0, p22 = 1, b11 = A11, b12 = A12, b21 = A21, b22 = A22;
    for (long e = len; e > 0; e >>= 1) {
      if (e & 1) {
        double n11 = b11 * p11 + b12 * p21, n12 = b11 * p12 + b12 * p22, n21 = b21 * p11 + b22 * p21, n22 = b21 * p12 + b22 * p22;
        p11 = n11; p12 = n12; p21 = n21; p22 = n22;
      }
      double n11 = b11 * b11 + b12 * b21, n12 = b11 * b12 + b12 * b22, n21 = b21 * b11 + b22 * b21, n22 = b21 * b12 + b22 * b22;
      b11 = n11; b12 = n12; b21 = n21; b22 = n22;
    }
    double n1 = p11 * s1 + p12 * s2 + (double)z1, n2 = p21 * s1 + p22 * s2 + (double)z2;
    s1 = n1; s2 = n2;
  }
}

// K1f  k_polysynth — `sine(f) >> <fixed SVF>` * ar(a, ak, r, rk), optional group mix (BASELINE configs[2]).
//      One lane per voice, time sequential, every per-voice quantity in registers.
//      * phase and envelope time accumulate with the reference's exact operation order (no FMA): a 1-ulp slip
//        would drift over 480,000 samples;
//      * lfo() control points are jittered per voice (hash-seeded), so a naive `if (t >= t1)` makes ~28 % of all
//        warp-steps execute the expensive segment update for a single lane.  Instead the NEXT control point is
//        computed ahead of time at warp-uniform window boundaries (window < shortest possible segment), and the
//        per-sample crossing only rotates registers;
//      * SVF / envelope interpolation use FMA and a per-segment reciprocal: f32 audio tolerance, not bits;
//      * group mix (K6): a 32x33 shared tile per warp, lane j adds column j over the group's voices left to right.
// lfo control functions with constant parameters: xd(t) = exp(-t p), xD, ar (functions.rs:505-507, 517-540, 547-555)
struct EnvC { float c[4]; int shape; };
__device__ __forceinline__ float env_eval(float tt, const EnvC& e) { return d_env_eval(e.shape, 0, tt, e.c, e.c); }

constexpr int PS_THREADS = 64;   // 2 warps per block: 1024 blocks for 65,536 voices spread evenly (6.9 per SM)
// OSC 0: sine(f) (SFU); OSC 1: band-limited wavetable oscillator saw / square / triangle / soft_saw (f fixed per voice: the
// table is chosen once, a sample costs the phase step + a 4-point interpolated read)
template <bool LP, int OSC>
__global__ void __launch_bounds__(PS_THREADS) k_polysynth(const float* __restrict__ params, float* __restrict__ state, int Vp, int V,
                                                   long T, int G, int look_tiles, int p_f, int p_sd, int p_svf, int p_env,
                                                   int s_ph, int s_svf, int s_env, int env_shape, const float* __restrict__ wt_hdr,
                                                   float* __restrict__ out) {
  __shared__ float tile[PS_THREADS / 32][32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v = blockIdx.x * PS_THREADS + threadIdx.x;   // < Vp (padded voices run on copies of the last voice)
  const int warp_v0 = blockIdx.x * PS_THREADS + warp * 32;
#define PRM(i) params[(size_t)(i) * Vp + v]
#define ST(i) state[(size_t)(i) * Vp + v]
  const float inc = PRM(p_f) * PRM(p_sd);                // input[0] * sample_duration
  const SvfC c = {PRM(p_svf), PRM(p_svf + 1), PRM(p_svf + 2), PRM(p_svf + 3), PRM(p_svf + 4), PRM(p_svf + 5)};
  const EnvC e = {{PRM(p_env), PRM(p_env + 1), PRM(p_env + 2), PRM(p_env + 3)}, env_shape};
  const float esd = PRM(p_env + 4);
  // wavetable oscillator: the frequency is a per-voice constant, so is the table
  uint32_t wt_hint = OSC == 1 ? __float_as_uint(ST(s_ph + 1)) : 0u, wt_len = 2;
  const float* wt_tb = OSC == 1 ? d_wavetable_select(wt_hdr, PRM(p_f), wt_hint, wt_len) : nullptr;
  const float wt_flen = (float)wt_len;
  float phase = ST(s_ph), ic1 = ST(s_svf), ic2 = ST(s_svf + 1);
  float et = ST(s_env), t0 = ST(s_env + 1), t1 = ST(s_env + 2), v0 = ST(s_env + 3), v1 = ST(s_env + 4);
  uint64_t th = (uint64_t)__float_as_uint(ST(s_env + 5)) | ((uint64_t)__float_as_uint(ST(s_env + 6)) << 32);
  uint32_t first = __float_as_uint(ST(s_env + 7));
  // prologue: bring the envelope "inside a segment" exactly like the per-sample code would on its first tick
  if (et >= t1) {
    if (first) { v1 = env_eval(0.0f, e); first = 0u; }
    t0 = t1; v0 = v1;
    t1 = t0 + d_lerp(0.75f, 1.25f, d_rnd1(th)) * 0.002f;
    v1 = env_eval(t1, e);
    th += 1;
  }
  // The interpolated envelope lerp(v0, v1, (t - t0) / (t1 - t0)) is advanced incrementally (env += slope per sample,
  // anchored once per control point): <= ~200 f32 additions per anchor.
  // Two segments are live per lane: the CURRENT one (envC, dC, ends at t1) and the NEXT one (envN, dN, control point
  // (nt1, nv1) computed ahead of time).  A look-ahead window (look_tiles * 32 samples) is shorter than the shortest lfo
  // segment, so a lane crosses at most one control point per window: inside the window a sample only SELECTS between
  // the two running values (et >= t1), and the register rotation + the next look-ahead (hash, powf) happen at the
  // warp-uniform window boundary.
  float inv = 1.0f / (t1 - t0);
  float envC = __fmaf_rn(v1 - v0, (et - t0) * inv, v0);
  float dC = (v1 - v0) * inv * esd;
  float nt1 = 0.0f, nv1 = 0.0f, envN = 0.0f, dN = 0.0f;
  bool crossed = false;              // the last tested sample of the window was past t1
  uint32_t ncross = 0;               // control points consumed in this launch (added to the 64-bit hash counter at the end)
  float* my_tile = &tile[warp][lane][0];

  // lowpass: ic1' = (2 a1 - 1) ic1 + 2 a2 v3 (both constants exact in f32), no v1 needed: 6 FP ops per tick
  const float c11 = 2.0f * c.a1 - 1.0f, c12 = 2.0f * c.a2;
  const bool small_inc = __all_sync(0xffffffffu, inc >= 0.0f && inc < 1.0f);   // warp-uniform
  auto sample = [&](int i, auto small_inc_t) {
    // ---- sine (FunDSP Sine::tick): output from the phase before the increment.  The phase recurrence is exact;
    // sin(2 pi p) uses the SFU (MUFU.SIN works on the fractional revolution, so p in [0, 1) needs no folding; the two
    // roundings of p * 2pi * (1 / 2pi) cost < 5e-7 rad)
    const float p = phase;
    phase = p + inc;
    // p in [0, 1): for 0 <= inc < 1 the sum is in [0, 2) and `phase - floor(phase)` is exactly a conditional `- 1`
    // (keeps FRND off the quarter-rate XU pipe that MUFU.SIN already uses); other increments take the general form
    if (decltype(small_inc_t)::value) { if (phase >= 1.0f) phase -= 1.0f; }
    else phase -= floorf(phase);
    // sine reads the phase BEFORE the step, the wavetable oscillators the phase AFTER it (FunDSP Sine / WaveSynth)
    const float x = OSC == 0 ? __sinf(p * QG_TAU) : d_wavetable_interp_fma(wt_tb, wt_len, wt_flen, phase);
    // ---- SVF
    float y;
    if (LP) {
      const float v3 = x - ic2;
      y = __fmaf_rn(c.a3, v3, __fmaf_rn(c.a2, ic1, ic2));
      ic1 = __fmaf_rn(c12, v3, c11 * ic1);
      ic2 = __fmaf_rn(2.0f, y, -ic2);
    } else {
      y = svf_fma<false>(x, ic1, ic2, c);
    }
    // ---- envelope (lfo)
    crossed = et >= t1;
    et += esd;
    my_tile[i] = y * (crossed ? envN : envC);
    envC += dC;
    envN += dN;
  };
  auto rotate = [&]() {              // the lane went past t1 in the window that just ended
    t0 = t1; v0 = v1; t1 = nt1; v1 = nv1;
    envC = envN; dC = dN;
    ncross += 1u;
    crossed = false;
  };

  long tile_idx = 0;
  bool have_next = false;
  const float sc = 1.0f / (float)(G > 0 ? G : 1);
  const int group0 = G > 1 ? warp_v0 / G : 0;            // first output row of this warp (32 % G == 0)
  for (long tb = 0; tb < T; tb += 32, tile_idx++) {
    if ((tile_idx % look_tiles) == 0) {                    // warp-uniform window boundary
      if (crossed) { rotate(); have_next = false; }
      if (!have_next) {                                    // look one control point ahead
        nt1 = t1 + d_lerp(0.75f, 1.25f, d_rnd1(th + (uint64_t)ncross)) * 0.002f;
        nv1 = env_eval(nt1, e);
        const float ninv = 1.0f / (nt1 - t1);
        envN = __fmaf_rn(nv1 - v1, (et - t1) * ninv, v1);  // the next segment's line, evaluated at the current time
        dN = (nv1 - v1) * ninv * esd;
        have_next = true;
      }
    }
    const int n = (T - tb) < 32 ? (int)(T - tb) : 32;
    if (n == 32 && small_inc) {
#pragma unroll 8
      for (int i = 0; i < 32; i++) sample(i, std::true_type{});
    } else if (n == 32) {
#pragma unroll 8
      for (int i = 0; i < 32; i++) sample(i, std::false_type{});
    } else {
      for (int i = 0; i < n; i++) sample(i, std::false_type{});
    }
    __syncwarp();
    if (lane < n) {
      if (G <= 1) {
        for (int r = 0; r < 32; r++)
          if (warp_v0 + r < V) out[(size_t)(warp_v0 + r) * T + tb + lane] = tile[warp][r][lane];
      } else {
        for (int g0 = 0, gi = group0; g0 < 32; g0 += G, gi++) {
          if (warp_v0 + g0 + G <= V) {
            float acc = tile[warp][g0][lane];
            for (int r = 1; r < G; r++) acc += tile[warp][g0 + r][lane];
            out[(size_t)gi * T + tb + lane] = acc * sc;
          }
        }
      }
    }
    __syncwarp();
  }
  if (crossed) rotate();
  th += (uint64_t)ncross;
  if (v < V) {
    ST(s_ph) = phase; ST(s_svf) = ic1; ST(s_svf + 1) = ic2;
    if (OSC == 1) ST(s_ph + 1) = __uint_as_float(wt_hint);
    ST(s_env) = et; ST(s_env + 1) = t0; ST(s_env + 2) = t1; ST(s_env + 3) = v0; ST(s_env + 4) = v1;
    ST(s_env + 5) = __uint_as_float((uint32_t)th); ST(s_env + 6) = __uint_as_float((uint32_t)(th >> 32));
    ST(s_env + 7) = __uint_as_float(first);
  }
#undef PRM
#undef ST
}

}  // namespace

FusedPlan plan_fused(const Tape& t) {
  FusedPlan pl;
  const auto& c = t.code;
  // white() >> <fixed linear filter>: K2
  if (t.h.n_inputs == 0 && t.h.n_outputs == 1 && c.size() == 2 && c[0].op == OP_NOISE &&
      (c[1].op == OP_SVF || c[1].op == OP_BIQUAD || c[1].op == OP_ONEPOLE) && c[1].in[0] == c[0].out && t.out_x.size() == 1 &&
      t.out_x[0] == c[1].out) {
    pl.id = FUSED_NOISE_SVF;
    pl.p[0] = c[1].p;                                   // first coefficient (X index == parameter index)
    pl.s[0] = c[0].s - (int)t.h.n_params;               // noise counter
    pl.s[1] = c[1].s - (int)t.h.n_params;               // first state word of the filter
    if (c[1].op == OP_SVF) {
      // the mix (m0, m1, m2) is a function of the filter mode alone for lowpass: identical for every voice
      pl.p[1] = (t.params[c[1].p + 3] == 0.0f && t.params[c[1].p + 4] == 0.0f && t.params[c[1].p + 5] == 1.0f) ? 1 : 0;
    } else if (c[1].op == OP_BIQUAD) {
      pl.p[1] = 2;
    } else {
      pl.p[1] = 3;
      pl.p[2] = c[1].n;                                 // one-pole kind
    }
    return pl;
  }
  // sine(f) >> svf(fixed)  *  ar(a,ak,r,rk)
  const int P = (int)t.h.n_params;
  // <sine | saw | square | triangle | soft_saw>(f) >> svf(fixed)  *  <xd(p) | xD(d,k) | ar(a,ak,r,rk)>   (constant parameters)
  if (t.h.n_inputs == 0 && t.h.n_outputs == 1 && c.size() == 4 && (c[0].op == OP_SINE || c[0].op == OP_WAVETABLE) && c[0].in[0] < P &&
      c[1].op == OP_SVF && c[1].in[0] == c[0].out && c[2].op == OP_ENVELOPE && (c[2].n >> 8) == 0 && (c[2].n & 0xff) <= 2 &&
      c[3].op == OP_MUL && c[3].in[0] == c[1].out && c[3].in[1] == c[2].out && t.out_x.size() == 1 && t.out_x[0] == c[3].out) {
    pl.id = FUSED_SINE_SVF_ENV;
    pl.p[0] = c[0].in[0]; pl.p[1] = c[0].p; pl.p[2] = c[1].p; pl.p[3] = c[2].p;
    pl.s[0] = c[0].s - P; pl.s[1] = c[1].s - P; pl.s[2] = c[2].s - P;
    pl.p[4] = (t.params[c[1].p + 3] == 0.0f && t.params[c[1].p + 4] == 0.0f && t.params[c[1].p + 5] == 1.0f) ? 1 : 0;
    pl.p[5] = c[0].op == OP_WAVETABLE ? 1 : 0;          // oscillator kind
    pl.p[6] = (int)c[0].aux;                            // wavetable: table-set header offset
    pl.p[7] = c[2].n & 0xff;                            // envelope shape
  }
  return pl;
}

const char* fused_name(int id) {
  return id == FUSED_NOISE_SVF ? "k_noise_svf_scan" : id == FUSED_SINE_SVF_ENV ? "k_polysynth" : "none";
}

// 2-D view of a voice-major f32 buffer as rows of 32 samples (128 B); box = K rows, shared-memory layout SWIZZLE_128B.
// cuTensorMapEncodeTiled is fetched through the runtime's driver entry point (no link-time dependency on libcuda).
static bool encode_rows32(CUtensorMap* tm, float* base, size_t n_rows, int box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)p;
  });
  if (!fn) return false;
  cuuint64_t dims[2] = {32, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launch_fused(const FusedPlan& pl, const FusedArgs& a, cudaStream_t stream, int* launches) {
  if (pl.id == FUSED_NOISE_SVF) {
    if (a.group != 1) return cudaErrorNotSupported;
    // segments per voice: enough warps to fill 148 SMs x ~24 warps when the bank is small
    constexpr int K = K2_K;
    constexpr int B = 32 * K;
    int S = 1;
    const long target_warps = 148L * 24;
    if (a.V < target_warps) {
      S = (int)((target_warps + a.V - 1) / a.V);
      long max_s = a.T / (8L * B);
      if (S > max_s) S = (int)(max_s < 1 ? 1 : max_s);
    }
    long seg_len = ((a.T + S - 1) / S + B - 1) / B * B;
    S = (int)((a.T + seg_len - 1) / seg_len);
    if (S < 1) S = 1;
    long warps = (long)a.V * S;
    unsigned blocks = (unsigned)((warps + 3) / 4);
    const int filt = pl.p[1], kind = filt == 3 ? pl.p[2] : 0;
    float* seg = nullptr;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
#define QG_K2(MODEV, FV, SV) k_noise_svf_scan<MODEV, FV, SV, K><<<blocks, 128, 0, stream>>>(a.params, a.state, a.Vp, a.V, a.T, S, seg_len, pl.p[0], pl.s[0], pl.s[1], kind, MODEV == 0 ? a.out : nullptr, seg, tmap)
#define QG_K2_F(MODEV, SV) do { if (filt == 1) QG_K2(MODEV, F_SVF_LP, SV); else if (filt == 0) QG_K2(MODEV, F_SVF, SV); else if (filt == 2) QG_K2(MODEV, F_BIQUAD, SV); else QG_K2(MODEV, F_ONEPOLE, SV); } while (0)
    if (S > 1) {
      size_t need = (size_t)a.V * (S + 1) * 2 * sizeof(float);
      if (need > *a.scratch_bytes) {
        if (*a.scratch) cudaFreeAsync(*a.scratch, stream);   // bank buffers live in the stream-ordered pool (capi.cu: dev_malloc)
        *a.scratch = nullptr; *a.scratch_bytes = 0;
        cudaError_t e = cudaMallocAsync((void**)a.scratch, need, stream);
        if (e != cudaSuccess) return e;
        *a.scratch_bytes = need;
      }
      seg = *a.scratch;
      QG_K2_F(1, 0);
      const unsigned cb = (a.V + 127) / 128;
      if (filt == 2) k_chain_segments<F_BIQUAD><<<cb, 128, 0, stream>>>(a.params, a.state, a.Vp, a.V, a.T, S, seg_len, pl.p[0], pl.s[1], kind, seg);
      else if (filt == 3) k_chain_segments<F_ONEPOLE><<<cb, 128, 0, stream>>>(a.params, a.state, a.Vp, a.V, a.T, S, seg_len, pl.p[0], pl.s[1], kind, seg);
      else k_chain_segments<F_SVF><<<cb, 128, 0, stream>>>(a.params, a.state, a.Vp, a.V, a.T, S, seg_len, pl.p[0], pl.s[1], kind, seg);
      if (launches) *launches += 2;
    }
    // output path: TMA tensor store needs 32-sample rows on 128-byte boundaries; linear bulk copies need 16-byte rows
    int store = ((((size_t)(uintptr_t)a.out) & 15) == 0) && ((a.T & 3) == 0) ? 1 : 0;
    if (store == 1 && ((((size_t)(uintptr_t)a.out) & 127) == 0) && (a.T % 32) == 0 && ((size_t)a.V * (size_t)a.T / 32) < 0x7fffffffull &&
        encode_rows32(&tmap, a.out, (size_t)a.V * (size_t)a.T / 32, K))
      store = 2;
    // rows that are 16-byte but not 128-byte aligned could use the linear bulk copy (STORE == 1, kept in the kernel template);
    // it is not instantiated: the case is rare and the scalar-store variant serves it
    if (store == 2) QG_K2_F(0, 2); else QG_K2_F(0, 0);
#undef QG_K2_F
#undef QG_K2
    if (launches) *launches += 1;
    if (S > 1) {
      const unsigned cb = (a.V + 127) / 128;
      if (filt == 2) k_finalize_segments<F_BIQUAD><<<cb, 128, 0, stream>>>(a.state, a.Vp, a.V, a.T, S, pl.s[0], pl.s[1], seg);
      else if (filt == 3) k_finalize_segments<F_ONEPOLE><<<cb, 128, 0, stream>>>(a.state, a.Vp, a.V, a.T, S, pl.s[0], pl.s[1], seg);
      else k_finalize_segments<F_SVF><<<cb, 128, 0, stream>>>(a.state, a.Vp, a.V, a.T, S, pl.s[0], pl.s[1], seg);
      if (launches) *launches += 1;
    }
    return cudaGetLastError();
  }
  if (pl.id == FUSED_SINE_SVF_ENV) {
    if (a.group < 1 || a.group > 32 || (32 % a.group) != 0) return cudaErrorNotSupported;
    // the envelope look-ahead runs every `look_tiles` tiles of 32 samples and must be shorter than the shortest
    // lfo segment (0.75 * 2 ms)
    const double min_seg = 0.0015 * (double)a.sample_rate - 2.0;
    int look_tiles = min_seg >= 64.0 ? 2 : (min_seg >= 32.0 ? 1 : 0);
    if (look_tiles == 0) return cudaErrorNotSupported;
#define QG_PS(LPV, OSCV) k_polysynth<LPV, OSCV><<<a.Vp / PS_THREADS, PS_THREADS, 0, stream>>>(a.params, a.state, a.Vp, a.V, a.T, a.group, look_tiles, pl.p[0], pl.p[1], pl.p[2], pl.p[3], pl.s[0], pl.s[1], pl.s[2], pl.p[7], a.tables ? a.tables + pl.p[6] : nullptr, a.out)
    if (pl.p[5] && !a.tables) return cudaErrorNotSupported;
    if (launches) *launches += 1;
    // sine voices: two voices per lane in packed f32x2 arithmetic (fused_poly.cu); the wavetable oscillators (a 4-point
    // interpolated table read per sample) keep one voice per lane
    if (!pl.p[5] && look_tiles == 2) return launch_polysynth_x2(pl, a, stream);
    if (pl.p[5]) { if (pl.p[4]) QG_PS(true, 1); else QG_PS(false, 1); }
    else { if (pl.p[4]) QG_PS(true, 0); else QG_PS(false, 0); }   // sine below 44.1 kHz: 32-sample envelope windows
#undef QG_PS
    return cudaGetLastError();
  }
  return cudaErrorNotSupported;
}

}  // namespace qg
