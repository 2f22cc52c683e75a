// K1f  k_polysynth_x2 — `sine(f) >> <fixed SVF> * <xd | xD | ar>(constants)` with optional group mix (BASELINE configs[2]),
// TWO VOICES PER LANE in packed f32x2 arithmetic (sm_100a FFMA2 / FADD2 / FMUL2: one issue slot, two FMAs per lane).
//
// What the reference does per voice and sample (FunDSP Sine::tick, Svf::tick, lfo envelope; functions.rs:547-555 for the
// `ar` shape, process.rs:1756 for the sum of voices): phase += f / sr (wrapped, f32), sin(2 pi phase), 12-flop SVF tick,
// piecewise-linear envelope through jittered ~2 ms control points, multiply, add into the mix.  Here:
//   * a warp owns 64 consecutive voices, lane l the pair (2l, 2l+1): every recurrence (phase, SVF, envelope lines) is one
//     packed instruction per pair; phase keeps the reference's exact f32 operation order (packed add / conditional -1 are
//     RN per component), because a 1-ulp slip per sample drifts audibly over 480,000 samples;
//   * the envelope: inside a window (64 samples, shorter than the shortest lfo segment) at most one control point is
//     crossed, so the piecewise-linear interpolant equals min (concave kink) or max (convex kink) of the CURRENT and the
//     NEXT segment's lines.  max is turned into min by rendering the voice NEGATED (sine argument, SVF state and both lines
//     times -1: the SVF is linear and IEEE arithmetic is sign-symmetric, so this is exact) — a sample costs two packed
//     line steps and one FMNMX per voice; no time accumulation, no compare/select per sample.  The exact f32 time
//     recurrence t += 1/sr runs in closed form per window (constant rounded step inside a binade, checked) and decides,
//     like the reference's `t >= t1`, when a control point is consumed;
//   * control points (64-bit hash -> jitter, pow shapes through ex2(k lg2 x): |error| < 2e-7) are computed one segment
//     ahead at the warp-uniform window boundary;
//   * the pair's two products are added in the lane, parked as float4 (4 consecutive samples, conflict-free STS.128) in a
//     [lane][sample] tile, and the group sum is 16 LDS.128 + packed adds per 4 output samples, stored as 16-byte words.
// Arithmetic is FMA-contracted: parity is the f32 audio tolerance (<= 1e-4 abs, <= -90 dBFS), tests/test_gpu_fused.py.
#include "fused.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "dev_math.cuh"
#include "tape.h"

namespace qg {

namespace {

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float& comp(float2& v, int j) { return j ? v.y : v.x; }
__device__ __forceinline__ float compc(const float2& v, int j) { return j ? v.y : v.x; }
// FunDSP Sine: phase -= floor(phase).  For a phase in [0, 2) (0 <= increment < 1) that is a conditional -1: w = p - 1 is
// exact when p >= 1 and negative otherwise, and as UNSIGNED integers bits(w) < bits(p) exactly when w >= 0 — one packed add
// and an integer min per voice (ALU pipe).  (FSET.BF, the float-valued compare, measured at ~4 cycles per warp instruction
// on B200: it made the wrap the most expensive part of the recurrence.)
__device__ __forceinline__ float2 wrap01(float2 p) {
  const float2 w = __fadd2_rn(p, make_float2(-1.0f, -1.0f));
  return make_float2(__uint_as_float(min(__float_as_uint(p.x), __float_as_uint(w.x))),
                     __uint_as_float(min(__float_as_uint(p.y), __float_as_uint(w.y))));
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// lfo control functions with constant parameters: xd, xD, ar (functions.rs:505-507, 517-540, 547-555).  Every shape ends
// in one ex2(k lg2 x): relative error ~ |k log2 x| * 2e-7 — the control-point values are in [0, 1] and the audio tolerance
// is 1e-4.  k == 1 (the `ar` attack of configs[2]) stays exact.  Selects only: `shape` is warp-uniform.
__device__ __forceinline__ float env_point(int shape, float tt, float c0, float c1, float c2, float c3) {
  const bool s0 = shape == 0, s1 = shape == 1, att = tt < c0;
  const float x = s1 ? (c0 - tt) * rcp_approx(c0) : (att ? tt * rcp_approx(c0) : (c2 - (tt - c0)) * rcp_approx(c2));
  const float k = (s1 || att) ? c1 : c3;
  const bool live = s0 || (s1 ? att : tt < c0 + c2);
  float r = exp2f(s0 ? -tt * c0 * 1.4426950408889634f : k * __log2f(x));
  r = (!s0 && k == 1.0f) ? x : r;
  r = (!s0 && k == 0.0f) ? 1.0f : r;
  return live ? r : 0.0f;
}
// FunDSP rnd1: (hash >> 11) as f64 / 2^53, rounded to f32 — one RN rounding of the 53-bit integer, then an exact scaling
__device__ __forceinline__ float rnd1_f32(uint64_t x) { return __ull2float_rn(d_hash64a(x) >> 11) * (1.0f / 9007199254740992.0f); }

constexpr int PX_PITCH = 68;   // floats per tile row: 64 samples + 4 (rows stay 16-byte aligned, STS.128 / LDS.128 conflict-free)

template <bool LP, int G>
__global__ void __launch_bounds__(32, 16) k_polysynth_x2(const float* __restrict__ params, float* __restrict__ state, int Vp, int V,
                                                    long T, int win, int vec_ok, int p_f, int p_sd, int p_svf, int p_env, int s_ph,
                                                    int s_svf, int s_env, int env_shape, float* __restrict__ out, int S, long T1,
                                                    int W, float* __restrict__ state_out) {
  constexpr int ROWS = G == 1 ? 64 : 32;   // tile rows: one per voice, or one per lane (the pair is summed in the lane)
  constexpr int RPO = G == 1 ? 1 : G / 2;  // tile rows per output row
  constexpr int NOUT = ROWS / RPO;         // output rows of one warp
  __shared__ __align__(16) float tile[ROWS * PX_PITCH];
  const int lane = threadIdx.x;
  // S == 2: two warps per block of 64 voices split the render in time.  Role 0 renders [0, T1).  Role 1 walks the exact
  // recurrences (phase, envelope time, control points) over [0, T1 - W) without rendering, lets the SVF settle from zero
  // state over the W samples before T1 (W >= the filter's decay to 1e-9, fused_settle()), then renders [T1, T) and owns
  // the final state.  More warps per scheduler hide the latency that 2 voices per lane concentrate in one warp.
  const int NB = gridDim.x / S;
  const int role = blockIdx.x / NB;
  const int v0 = (blockIdx.x - role * NB) * 64;   // first voice of the warp (Vp is a multiple of 128: padded voices copy the last one)
  const int va = v0 + 2 * lane;
#define PRM2(i) (*reinterpret_cast<const float2*>(&params[(size_t)(i) * Vp + va]))
#define ST2(i) (*reinterpret_cast<float2*>(&state[(size_t)(i) * Vp + va]))
  const float2 inc = __fmul2_rn(PRM2(p_f), PRM2(p_sd));          // input[0] * sample_duration
  const float2 a1 = PRM2(p_svf), a2 = PRM2(p_svf + 1), a3 = PRM2(p_svf + 2);
  const float2 m0 = PRM2(p_svf + 3), m1 = PRM2(p_svf + 4), m2 = PRM2(p_svf + 5);
  // lowpass: ic1' = (2 a1 - 1) ic1 + 2 a2 v3 (both constants exact in f32), no v1 needed: 6 packed ops per tick
  const float2 c11 = f2(2.0f * a1.x - 1.0f, 2.0f * a1.y - 1.0f), c12 = f2(2.0f * a2.x, 2.0f * a2.y);
  const float2 two = f2(2.0f, 2.0f);
  const float2 ec0 = PRM2(p_env), ec1 = PRM2(p_env + 1), ec2 = PRM2(p_env + 2), ec3 = PRM2(p_env + 3), esd = PRM2(p_env + 4);
  float2 ph = ST2(s_ph), ic1 = ST2(s_svf), ic2 = ST2(s_svf + 1);
  float2 et = ST2(s_env), t0 = ST2(s_env + 1), t1 = ST2(s_env + 2), ev0 = ST2(s_env + 3), ev1 = ST2(s_env + 4);
  const float2 thl = ST2(s_env + 5), thh = ST2(s_env + 6), firstw = ST2(s_env + 7);
  uint64_t th[2] = {(uint64_t)__float_as_uint(thl.x) | ((uint64_t)__float_as_uint(thh.x) << 32),
                    (uint64_t)__float_as_uint(thl.y) | ((uint64_t)__float_as_uint(thh.y) << 32)};
  uint32_t first[2] = {__float_as_uint(firstw.x), __float_as_uint(firstw.y)};
  float2 nt1 = f2(0.0f, 0.0f), nv1 = nt1, invC, invN = nt1;
  uint32_t ncross[2] = {0u, 0u};
  bool crossed[2] = {false, false};
  // prologue: bring each envelope "inside a segment" exactly like the per-sample code would on its first tick
#pragma unroll
  for (int j = 0; j < 2; j++) {
    if (comp(et, j) >= comp(t1, j)) {
      if (first[j]) { comp(ev1, j) = env_point(env_shape, 0.0f, compc(ec0, j), compc(ec1, j), compc(ec2, j), compc(ec3, j)); first[j] = 0u; }
      comp(t0, j) = comp(t1, j); comp(ev0, j) = comp(ev1, j);
      comp(t1, j) = comp(t0, j) + d_lerp(0.75f, 1.25f, rnd1_f32(th[j])) * 0.002f;
      comp(ev1, j) = env_point(env_shape, comp(t1, j), compc(ec0, j), compc(ec1, j), compc(ec2, j), compc(ec3, j));
      th[j] += 1;
    }
    comp(invC, j) = rcp_approx(comp(t1, j) - comp(t0, j));
  }
  // the voice is rendered as sgn * voice (see header): sine argument scale, SVF state and envelope lines carry the sign
  float2 sgn = f2(1.0f, 1.0f), stau = f2(QG_TAU, QG_TAU);
  float2 sC, sN, dC, dN;   // sgn * (current / next segment's line at the running sample), and their per-sample steps
  const bool small_inc = __all_sync(0xffffffffu, inc.x >= 0.0f && inc.x < 1.0f && inc.y >= 0.0f && inc.y < 1.0f);   // warp-uniform

  // one sample of the pair: returns sgn * svf_output (ys) and sgn * envelope (m): ys * m = output * envelope
  auto sample = [&](float2& ys, float2& m, auto small_t) {
    const float2 p = ph;
    ph = __fadd2_rn(p, inc);
    if (decltype(small_t)::value) ph = wrap01(ph);
    else ph = f2(ph.x - floorf(ph.x), ph.y - floorf(ph.y));
    // sine reads the phase BEFORE the step; MUFU.SIN works on the fractional revolution, so p in [0, 1) needs no folding
    const float2 arg = __fmul2_rn(p, stau);
    const float2 x = f2(__sinf(arg.x), __sinf(arg.y));
    const float2 v3 = __fadd2_rn(x, neg2(ic2));
    if (LP) {
      ys = __ffma2_rn(a3, v3, __ffma2_rn(a2, ic1, ic2));
      ic1 = __ffma2_rn(c12, v3, __fmul2_rn(c11, ic1));
      ic2 = __ffma2_rn(two, ys, neg2(ic2));
    } else {
      const float2 v1 = __ffma2_rn(a2, v3, __fmul2_rn(a1, ic1));
      const float2 v2 = __ffma2_rn(a3, v3, __ffma2_rn(a2, ic1, ic2));
      ic1 = __ffma2_rn(two, v1, neg2(ic1));
      ic2 = __ffma2_rn(two, v2, neg2(ic2));
      ys = __ffma2_rn(m2, v2, __ffma2_rn(m1, v1, __fmul2_rn(m0, x)));
    }
    m = f2(fminf(sC.x, sN.x), fminf(sC.y, sN.y));
    sC = __fadd2_rn(sC, dC);
    sN = __fadd2_rn(sN, dN);
  };

  float4* tile4 = reinterpret_cast<float4*>(tile);
  const int chunks = win >> 2;             // 16-byte chunks per tile row
  const int chunk_shift = win == 64 ? 4 : 3;
  const float gscale = 1.0f / (float)G;

  // ---- window boundary (warp-uniform point), for the window of n samples that starts at the current time: consume a
  // crossed control point, look one ahead, anchor both lines at the window's first sample, advance the exact time
  // recurrence over the window.  Straight-line code (selects, no branches) so that the two voices' chains — 64-bit hash,
  // MUFU sequences — interleave with each other and with the group sums of the window that just ended; n == 0 is harmless.
  auto boundary = [&](int n) {
    bool slow_any = false;
    float e_first[2], dt_j[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const bool cr = crossed[j];
      comp(t0, j) = cr ? comp(t1, j) : comp(t0, j);
      comp(ev0, j) = cr ? comp(ev1, j) : comp(ev0, j);
      comp(t1, j) = cr ? comp(nt1, j) : comp(t1, j);
      comp(ev1, j) = cr ? comp(nv1, j) : comp(ev1, j);
      comp(invC, j) = cr ? comp(invN, j) : comp(invC, j);
      ncross[j] += cr ? 1u : 0u;
      // the control point after t1 (recomputed every window: the same value for as long as the segment stands)
      const float nt = comp(t1, j) + d_lerp(0.75f, 1.25f, rnd1_f32(th[j] + (uint64_t)ncross[j])) * 0.002f;
      comp(nt1, j) = nt;
      comp(nv1, j) = env_point(env_shape, nt, compc(ec0, j), compc(ec1, j), compc(ec2, j), compc(ec3, j));
      comp(invN, j) = rcp_approx(nt - comp(t1, j));
      // exact t += sd, n times: inside one binade the rounded step is a constant (two equal steps imply it stays constant:
      // a tie can only alternate on the first step), so the window's times are e + i * d in closed form
      const float e = comp(et, j), sd = compc(esd, j);
      const float e1 = e + sd, e2 = e1 + sd, d = e1 - e;
      const float e_end = __fmaf_rn((float)n, d, e), e_last = __fmaf_rn((float)(n - 1), d, e);
      const bool fast = (e2 - e1) == d && (__float_as_uint(e) >> 23) == (__float_as_uint(e_end) >> 23) && e > 0.0f;
      slow_any = slow_any || !fast;
      e_first[j] = e;
      dt_j[j] = d;
      comp(et, j) = e_end;
      crossed[j] = n > 0 && e_last >= comp(t1, j);
    }
    if (__any_sync(0xffffffffu, slow_any)) {     // start of a render, binade crossings (a handful of windows per render): step
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const float e = e_first[j], sd = compc(esd, j);
        float w = e, e_last = e;
        for (int i = 0; i < n; i++) { e_last = w; w += sd; }
        comp(et, j) = w;
        dt_j[j] = n > 0 ? (w - e) * rcp_approx((float)n) : sd;
        crossed[j] = n > 0 && e_last >= comp(t1, j);
      }
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const float e = e_first[j];
      // lines through (t0, v0)-(t1, v1) and (t1, v1)-(nt1, nv1), evaluated at the window's first sample, stepped by dt
      const float gC = (comp(ev1, j) - comp(ev0, j)) * comp(invC, j), gN = (comp(nv1, j) - comp(ev1, j)) * comp(invN, j);
      const float lc = __fmaf_rn(gC, e - comp(t0, j), comp(ev0, j)), ln = __fmaf_rn(gN, e - comp(t1, j), comp(ev1, j));
      // concave kink (slope decreases): min of the two lines; convex: max = -min of the negated lines
      const float s = gN <= gC ? 1.0f : -1.0f;
      const float flip = s * comp(sgn, j);       // -1: the voice changes sign for this window
      comp(sgn, j) = s;
      comp(stau, j) *= flip;
      comp(ic1, j) *= flip;
      comp(ic2, j) *= flip;
      comp(sC, j) = s * lc; comp(sN, j) = s * ln;
      comp(dC, j) = s * gC * dt_j[j]; comp(dN, j) = s * gN * dt_j[j];
    }
  };

  const long t_begin = role == 0 ? 0 : T1 - W;              // multiples of win
  const long t_end = (S == 2 && role == 0) ? T1 : T;
  const long t_emit = role == 0 ? 0 : T1;
  if (role == 1 && t_begin > 0) {
    // ---- pre-pass over [0, t_begin): state only.  Phase: the reference's f32 recurrence, sample by sample (it has no
    // closed form: the rounding depends on the binade the running sum is in).
    if (small_inc) {
#pragma unroll 8
      for (long i = 0; i < t_begin; i++) {
        ph = __fadd2_rn(ph, inc);
        ph = wrap01(ph);
      }
    } else {
      for (long i = 0; i < t_begin; i++) {
        ph = __fadd2_rn(ph, inc);
        ph = f2(ph.x - floorf(ph.x), ph.y - floorf(ph.y));
      }
    }
    // Envelope time: exact closed form inside a binade (as in boundary()), in strides that grow while it holds.  Control
    // points: the reference consumes one at the first sample whose time reaches it, so after the pre-pass every control
    // point <= the time of the last pre-pass sample is gone.
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const float sd = compc(esd, j);
      float e = comp(et, j), e_last = e;
      long left = t_begin;
      long stride = 64;
      while (left > 0) {
        const long k = left < stride ? left : stride;
        const float e1 = e + sd, e2 = e1 + sd, d = e1 - e;
        const float e_end = __fmaf_rn((float)k, d, e);
        if ((e2 - e1) == d && (__float_as_uint(e) >> 23) == (__float_as_uint(e_end) >> 23) && e > 0.0f) {
          e_last = __fmaf_rn((float)(k - 1), d, e);
          e = e_end;
          left -= k;
          if (stride < 65536) stride *= 2;
        } else if (k <= 8) {
          for (long i = 0; i < k; i++) { e_last = e; e += sd; }
          left -= k;
        } else {
          stride = k / 8;
        }
      }
      comp(et, j) = e;
      bool moved = false;
      while (e_last >= comp(t1, j)) {
        comp(t0, j) = comp(t1, j);
        comp(t1, j) = comp(t0, j) + d_lerp(0.75f, 1.25f, rnd1_f32(th[j] + (uint64_t)ncross[j])) * 0.002f;
        ncross[j] += 1u;
        moved = true;
      }
      if (moved) {
        comp(ev0, j) = env_point(env_shape, comp(t0, j), compc(ec0, j), compc(ec1, j), compc(ec2, j), compc(ec3, j));
        comp(ev1, j) = env_point(env_shape, comp(t1, j), compc(ec0, j), compc(ec1, j), compc(ec2, j), compc(ec3, j));
        comp(invC, j) = rcp_approx(comp(t1, j) - comp(t0, j));
      }
    }
  }
  if (role == 1) { ic1 = f2(0.0f, 0.0f); ic2 = ic1; }      // the SVF settles over [T1 - W, T1)
  {
    const long left = t_end - t_begin;
    boundary(left < (long)win ? (int)left : win);
  }
  for (long tb = t_begin; tb < t_end; tb += win) {
    const int n = (t_end - tb) < (long)win ? (int)(t_end - tb) : win;
    const bool emit = tb >= t_emit;
    // ---- the window's samples
    auto full_window = [&](auto small_t) {
#pragma unroll 2
      for (int q = 0; q < chunks; q++) {
        float za[4], zb[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float2 ys, m;
          sample(ys, m, small_t);
          if (G == 1) { za[i] = ys.x * m.x; zb[i] = ys.y * m.y; }
          else za[i] = __fmaf_rn(ys.y, m.y, ys.x * m.x);
        }
        tile4[lane * (PX_PITCH / 4) + q] = make_float4(za[0], za[1], za[2], za[3]);
        if (G == 1) tile4[(32 + lane) * (PX_PITCH / 4) + q] = make_float4(zb[0], zb[1], zb[2], zb[3]);
      }
    };
    if (n == win && small_inc) full_window(std::true_type{});
    else if (n == win) full_window(std::false_type{});
    else {
      for (int i = 0; i < n; i++) {
        float2 ys, m;
        sample(ys, m, std::false_type{});
        if (G == 1) { tile[lane * PX_PITCH + i] = ys.x * m.x; tile[(32 + lane) * PX_PITCH + i] = ys.y * m.y; }
        else tile[lane * PX_PITCH + i] = __fmaf_rn(ys.y, m.y, ys.x * m.x);
      }
    }
    __syncwarp();
    // ---- the next window's boundary work shares a basic block with this window's group sums
    {
      const long left = t_end - tb - n;
      boundary(left < (long)win ? (int)left : win);
    }
    // ---- group sums (pairwise tree over the group's tile rows) and 16-byte stores
    for (int it = lane; emit && it < NOUT * chunks; it += 32) {
      const int orow = it >> chunk_shift, ch = it & (chunks - 1);
      float2 lo[RPO], hi[RPO];
#pragma unroll
      for (int k = 0; k < RPO; k++) {
        const float4 r = tile4[(orow * RPO + k) * (PX_PITCH / 4) + ch];
        lo[k] = f2(r.x, r.y); hi[k] = f2(r.z, r.w);
      }
#pragma unroll
      for (int w = 1; w < RPO; w *= 2) {
#pragma unroll
        for (int k = 0; k + w < RPO; k += 2 * w) { lo[k] = __fadd2_rn(lo[k], lo[k + w]); hi[k] = __fadd2_rn(hi[k], hi[k + w]); }
      }
      float4 acc = make_float4(lo[0].x, lo[0].y, hi[0].x, hi[0].y);
      long orow_g;       // output row
      bool valid;
      if (G == 1) {      // tile row r < 32: voice 2r of the warp; r >= 32: voice 2(r - 32) + 1
        const int vw = orow < 32 ? 2 * orow : 2 * (orow - 32) + 1;
        orow_g = (long)v0 + vw;
        valid = v0 + vw < V;
      } else {
        acc.x *= gscale; acc.y *= gscale; acc.z *= gscale; acc.w *= gscale;
        orow_g = (long)(v0 / G) + orow;
        valid = v0 + (orow + 1) * G <= V;
      }
      if (valid) {
        float* o = out + (size_t)orow_g * T + tb + 4 * ch;
        if (vec_ok && 4 * ch + 3 < n) *reinterpret_cast<float4*>(o) = acc;
        else {
          if (4 * ch + 0 < n) o[0] = acc.x;
          if (4 * ch + 1 < n) o[1] = acc.y;
          if (4 * ch + 2 < n) o[2] = acc.z;
          if (4 * ch + 3 < n) o[3] = acc.w;
        }
      }
    }
    __syncwarp();
  }
  // ---- persist (same state words as the interpreters' OP_SINE / OP_SVF / OP_ENVELOPE)
#pragma unroll
  for (int j = 0; j < 2; j++) {
    if (crossed[j]) {
      comp(t0, j) = comp(t1, j); comp(ev0, j) = comp(ev1, j); comp(t1, j) = comp(nt1, j); comp(ev1, j) = comp(nv1, j);
      ncross[j] += 1u;
    }
    th[j] += (uint64_t)ncross[j];
    if (comp(sgn, j) < 0.0f) { comp(ic1, j) = -comp(ic1, j); comp(ic2, j) = -comp(ic2, j); }
  }
#undef ST2
#define ST2(i) (*reinterpret_cast<float2*>(&state_out[(size_t)(i) * Vp + va]))
  if (va < V && role == S - 1) {     // va + 1 may be the first padded voice: its words are padding too (Vp > V), writing them is harmless
    ST2(s_ph) = ph; ST2(s_svf) = ic1; ST2(s_svf + 1) = ic2;
    ST2(s_env) = et; ST2(s_env + 1) = t0; ST2(s_env + 2) = t1; ST2(s_env + 3) = ev0; ST2(s_env + 4) = ev1;
    ST2(s_env + 5) = f2(__uint_as_float((uint32_t)th[0]), __uint_as_float((uint32_t)th[1]));
    ST2(s_env + 6) = f2(__uint_as_float((uint32_t)(th[0] >> 32)), __uint_as_float((uint32_t)(th[1] >> 32)));
    ST2(s_env + 7) = f2(__uint_as_float(first[0]), __uint_as_float(first[1]));
  }
#undef PRM2
#undef ST2
}

}  // namespace

// p[] / s[] as filled by plan_fused() for FUSED_SINE_SVF_ENV; serves the sine oscillator (the wavetable oscillators keep
// the one-voice-per-lane kernel in fused.cu)
cudaError_t launch_polysynth_x2(const FusedPlan& pl, const FusedArgs& a, int win, cudaStream_t stream) {
  const int vec_ok = ((((size_t)(uintptr_t)a.out) & 15) == 0 && (a.T & 3) == 0) ? 1 : 0;
  const int NB = a.Vp / 64;
  // time segments (see the kernel): worthwhile while the bank leaves schedulers short of warps, possible when the filters
  // forget their state within a small part of the render
  int S = 1, W = 0;
  long T1 = a.T;
  const char* force = getenv("QG_POLY_SEGMENTS");
  const bool allow = !(force && force[0] == '1');
  if (allow && NB <= 148 * 12 && a.settle > 0) {
    W = (a.settle + win - 1) / win * win;
    if (a.T >= 16L * W && a.T >= 32768) {
      S = 2;
      // role 1 spends ~0.12 of a rendered sample's cost on a pre-pass sample: T1 = (T1 - W) * 0.12 + (T - T1 + W)
      T1 = (long)(((double)a.T + 0.88 * (double)W) / 1.88) / win * win;
      if (T1 <= W || T1 >= a.T) { S = 1; T1 = a.T; W = 0; }
    }
  }
  float* state_out = a.state;
  const int rows = std::max(pl.s[0] + 1, std::max(pl.s[1] + 2, pl.s[2] + 8));
  if (S == 2) {   // role 1 finishes while role 0 warps of other blocks may not have read their start state yet
    const size_t need = (size_t)rows * a.Vp * sizeof(float);
    if (need > *a.scratch_bytes) {
      if (*a.scratch) cudaFree(*a.scratch);
      *a.scratch = nullptr; *a.scratch_bytes = 0;
      cudaError_t e = cudaMalloc((void**)a.scratch, need);
      if (e != cudaSuccess) return e;
      *a.scratch_bytes = need;
    }
    state_out = *a.scratch;
  }
  const unsigned blocks = (unsigned)(NB * S);
#define QG_PX(LPV, GV) k_polysynth_x2<LPV, GV><<<blocks, 32, 0, stream>>>(a.params, a.state, a.Vp, a.V, a.T, win, vec_ok, pl.p[0], pl.p[1], pl.p[2], pl.p[3], pl.s[0], pl.s[1], pl.s[2], pl.p[7], a.out, S, T1, W, state_out)
#define QG_PX_G(LPV)                                     \
  switch (a.group) {                                     \
    case 1: QG_PX(LPV, 1); break;                        \
    case 2: QG_PX(LPV, 2); break;                        \
    case 4: QG_PX(LPV, 4); break;                        \
    case 8: QG_PX(LPV, 8); break;                        \
    case 16: QG_PX(LPV, 16); break;                      \
    case 32: QG_PX(LPV, 32); break;                      \
    default: return cudaErrorNotSupported;               \
  }
  if (pl.p[4]) { QG_PX_G(true) } else { QG_PX_G(false) }
#undef QG_PX_G
#undef QG_PX
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || S == 1) return e;
  // the final state (phase; SVF ic1, ic2; the envelope's 8 words) moves from the scratch table into the bank's
  const size_t row = (size_t)a.Vp * sizeof(float);
  e = cudaMemcpyAsync(a.state + (size_t)pl.s[0] * a.Vp, state_out + (size_t)pl.s[0] * a.Vp, row, cudaMemcpyDeviceToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(a.state + (size_t)pl.s[1] * a.Vp, state_out + (size_t)pl.s[1] * a.Vp, 2 * row, cudaMemcpyDeviceToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(a.state + (size_t)pl.s[2] * a.Vp, state_out + (size_t)pl.s[2] * a.Vp, 8 * row, cudaMemcpyDeviceToDevice, stream);
  return e;
}

}  // namespace qg
