// K1f  k_polysynth_x2 — `sine(f) >> <fixed SVF> * <xd | xD | ar>(constants)` with optional group mix (BASELINE configs[2]),
// TWO VOICES PER LANE in packed f32x2 arithmetic (sm_100a FFMA2 / FADD2 / FMUL2).
//
// What the reference does per voice and sample (FunDSP Sine::tick, Svf::tick, lfo envelope; functions.rs:547-555 for the
// `ar` shape, process.rs:1756 for the sum of voices): phase += f / sr (wrapped, f32), sin(2 pi phase), 12-flop SVF tick,
// piecewise-linear envelope through jittered ~2 ms control points, multiply, add into the mix.  Here:
//   * a warp owns 64 consecutive voices, lane l the pair (2l, 2l+1): every recurrence (phase, SVF, envelope lines) is one
//     packed instruction per pair; phase keeps the reference's exact f32 operation order (packed add / conditional -1 are
//     RN per component), because a 1-ulp slip per sample drifts audibly over 480,000 samples;
//   * the envelope: inside a window of 64 samples (shorter than the shortest lfo segment) at most one control point is
//     crossed, so the piecewise-linear interpolant equals min (concave kink) or max (convex kink) of the CURRENT and the
//     NEXT segment's lines.  max is turned into min by rendering the voice NEGATED (sine argument, SVF state and both lines
//     times -1: the SVF is linear and IEEE arithmetic is sign-symmetric, so this is exact) — a sample costs two packed
//     line steps and one FMNMX per voice; no time accumulation, no compare/select per sample.  The exact f32 time
//     recurrence t += 1/sr runs in closed form per window (constant rounded step inside a binade, checked) and decides,
//     like the reference's `t >= t1`, when a control point is consumed;
//   * control points (64-bit hash -> jitter, pow shapes through ex2(k lg2 x): |error| < 2e-7) are computed one segment
//     ahead at the warp-uniform window boundary;
//   * the pair's two products are added in the lane, parked as float4 (4 consecutive samples, conflict-free STS.128) in a
//     [lane][sample] tile, and the group sum is 16 LDS.128 + packed adds per 4 output samples, stored as 16-byte words;
//   * SOFTWARE PIPELINE: measured on B200 (profiles/r02_ubench_pipes.txt) a packed op occupies the dispatch port for two
//     cycles, so the 64-sample loop is dispatch-bound (33 cycles per pair and sample) while the window boundary (hash,
//     MUFU chains) and the group sums are latency-bound.  One straight-line block per window therefore holds the group sums
//     of the PREVIOUS window (double-buffered tile), the boundary work of the NEXT window (results applied after the
//     samples) and the fully unrolled samples of THIS window: the scheduler fills the latency holes with sample work.
// Arithmetic is FMA-contracted: parity is the f32 audio tolerance (<= 1e-4 abs, <= -90 dBFS), tests/test_gpu_fused.py.
#include "fused.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "dev_math.cuh"
#include "tape.h"

namespace qg {

namespace {

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float& comp(float2& v, int j) { return j ? v.y : v.x; }
__device__ __forceinline__ float compc(const float2& v, int j) { return j ? v.y : v.x; }
// FunDSP Sine: phase -= floor(phase).  For a phase in [0, 2) (0 <= increment < 1) that is exactly a conditional -1:
// 1.0f / 0.0f from FSET.BF, one packed subtract.  (An unsigned-integer min of bits(p) and bits(p - 1) does the same on the
// ALU pipe; it measured 5 % slower in this kernel.)
__device__ __forceinline__ float ge_one(float x) {
  float r;
  asm("set.ge.f32.f32 %0, %1, 0f3F800000;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float2 wrap01(float2 p) { return __fadd2_rn(p, make_float2(-ge_one(p.x), -ge_one(p.y))); }
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// lfo control functions with constant parameters: xd, xD, ar (functions.rs:505-507, 517-540, 547-555).  Every shape ends
// in one ex2(k lg2 x): relative error ~ |k log2 x| * 2e-7 — the control-point values are in [0, 1] and the audio tolerance
// is 1e-4.  k == 1 (the `ar` attack of configs[2]) stays exact.  Selects only: `shape` is warp-uniform.
__device__ __forceinline__ float env_point(int shape, float tt, float c0, float c1, float c2, float c3, float rc0, float rc2) {
  const bool s0 = shape == 0, s1 = shape == 1, att = tt < c0;
  const float xa = (s1 ? c0 - tt : tt) * rc0, xr = (c2 - (tt - c0)) * rc2;     // rc0 = 1 / c0, rc2 = 1 / c2: per-voice constants
  const float x = (s1 || att) ? xa : xr;
  const float k = (s1 || att) ? c1 : c3;
  const bool live = s0 || (s1 ? att : tt < c0 + c2);
  float r = exp2f(s0 ? -tt * c0 * 1.4426950408889634f : k * __log2f(x));
  r = (!s0 && k == 1.0f) ? x : r;
  r = (!s0 && k == 0.0f) ? 1.0f : r;
  return live ? r : 0.0f;
}
// FunDSP rnd1: (hash >> 11) as f64 / 2^53, rounded to f32 — one RN rounding of the 53-bit integer, then an exact scaling
__device__ __forceinline__ float rnd1_f32(uint64_t x) { return __ull2float_rn(d_hash64a(x) >> 11) * (1.0f / 9007199254740992.0f); }

constexpr int PX_WIN = 64;     // samples per window: shorter than the shortest lfo segment (0.75 * 2 ms) from 44.1 kHz up
constexpr int PX_PITCH = 68;   // floats per tile row: 64 samples + 4 (rows stay 16-byte aligned, STS.128 / LDS.128 conflict-free)

template <bool LP, int G>
__global__ void __launch_bounds__(32) k_polysynth_x2(const float* __restrict__ params, float* __restrict__ state, int Vp, int V,
                                                    long T, int vec_ok, int p_f, int p_sd, int p_svf, int p_env, int s_ph,
                                                    int s_svf, int s_env, int env_shape, float* __restrict__ out) {
  constexpr int ROWS = G == 1 ? 64 : 32;   // tile rows: one per voice, or one per lane (the pair is summed in the lane)
  constexpr int RPO = G == 1 ? 1 : G / 2;  // tile rows per output row
  constexpr int NOUT = ROWS / RPO;         // output rows of one warp
  constexpr int CH = PX_WIN / 4;           // 16-byte chunks per tile row
  __shared__ __align__(16) float tile_mem[2][ROWS * PX_PITCH];
  const int lane = threadIdx.x;
  const int v0 = blockIdx.x * 64;          // first voice of the warp (Vp is a multiple of 128: padded voices copy the last one)
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
  const float2 rc0 = f2(rcp_approx(ec0.x), rcp_approx(ec0.y)), rc2 = f2(rcp_approx(ec2.x), rcp_approx(ec2.y));
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
      if (first[j]) { comp(ev1, j) = env_point(env_shape, 0.0f, compc(ec0, j), compc(ec1, j), compc(ec2, j), compc(ec3, j), compc(rc0, j), compc(rc2, j)); first[j] = 0u; }
      comp(t0, j) = comp(t1, j); comp(ev0, j) = comp(ev1, j);
      comp(t1, j) = comp(t0, j) + d_lerp(0.75f, 1.25f, rnd1_f32(th[j])) * 0.002f;
      comp(ev1, j) = env_point(env_shape, comp(t1, j), compc(ec0, j), compc(ec1, j), compc(ec2, j), compc(ec3, j), compc(rc0, j), compc(rc2, j));
      th[j] += 1;
    }
    comp(invC, j) = rcp_approx(comp(t1, j) - comp(t0, j));
  }
  // the voice is rendered as sgn * voice (see header): sine argument scale, SVF state and envelope lines carry the sign
  float2 sgn = f2(1.0f, 1.0f), stau = f2(QG_TAU, QG_TAU);
  float2 sC = f2(0.0f, 0.0f), sN = sC, dC = sC, dN = sC;   // sgn * (current / next segment's line at the running sample), per-sample steps
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

  // ---- window boundary (warp-uniform point) for the window of n samples that starts at the current envelope time:
  // consume a crossed control point, look one ahead, advance the exact time recurrence over the window, and work out the
  // sign and both lines at the window's first sample.  Nothing here is read by sample(): the results (struct Next) are
  // applied after the samples of the window before.  Selects, no branches; n == 0 is harmless.
  struct Next { float sC[2], sN[2], dC[2], dN[2], s[2], e_first[2]; bool slow; };
  auto lines_of = [&](int j, float e, float dt, Next& nx) {
    // lines through (t0, v0)-(t1, v1) and (t1, v1)-(nt1, nv1), evaluated at the window's first sample, stepped by dt
    const float gC = (comp(ev1, j) - comp(ev0, j)) * comp(invC, j), gN = (comp(nv1, j) - comp(ev1, j)) * comp(invN, j);
    const float lc = __fmaf_rn(gC, e - comp(t0, j), comp(ev0, j)), ln = __fmaf_rn(gN, e - comp(t1, j), comp(ev1, j));
    // concave kink (slope decreases): min of the two lines; convex: max = -min of the negated lines
    const float s = gN <= gC ? 1.0f : -1.0f;
    nx.s[j] = s;
    nx.sC[j] = s * lc; nx.sN[j] = s * ln;
    nx.dC[j] = s * gC * dt; nx.dN[j] = s * gN * dt;
  };
  auto prepare = [&](int n, Next& nx) {
    nx.slow = false;
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
      comp(nv1, j) = env_point(env_shape, nt, compc(ec0, j), compc(ec1, j), compc(ec2, j), compc(ec3, j), compc(rc0, j), compc(rc2, j));
      comp(invN, j) = rcp_approx(nt - comp(t1, j));
      // exact t += sd, n times: inside one binade the rounded step is a constant (two equal steps imply it stays constant:
      // a tie can only alternate on the first step), so the window's times are e + i * d in closed form
      const float e = comp(et, j), sd = compc(esd, j);
      const float e1 = e + sd, e2 = e1 + sd, d = e1 - e;
      const float e_end = __fmaf_rn((float)n, d, e), e_last = __fmaf_rn((float)(n - 1), d, e);
      const bool fast = (e2 - e1) == d && (__float_as_uint(e) >> 23) == (__float_as_uint(e_end) >> 23) && e > 0.0f;
      nx.slow = nx.slow || !fast;
      nx.e_first[j] = e;
      comp(et, j) = e_end;
      crossed[j] = n > 0 && e_last >= comp(t1, j);
      lines_of(j, e, d, nx);
    }
  };
  // start of a render and binade crossings (a handful of windows per render): step the time recurrence
  auto fixup = [&](int n, Next& nx) {
    if (!__any_sync(0xffffffffu, nx.slow)) return;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const float e = nx.e_first[j], sd = compc(esd, j);
      float w = e, e_last = e;
      for (int i = 0; i < n; i++) { e_last = w; w += sd; }
      comp(et, j) = w;
      crossed[j] = n > 0 && e_last >= comp(t1, j);
      lines_of(j, e, n > 0 ? (w - e) * rcp_approx((float)n) : sd, nx);
    }
  };
  auto apply = [&](const Next& nx) {
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const float flip = nx.s[j] * comp(sgn, j);       // -1: the voice changes sign for the coming window
      comp(sgn, j) = nx.s[j];
      comp(stau, j) *= flip;
      comp(ic1, j) *= flip;
      comp(ic2, j) *= flip;
      comp(sC, j) = nx.sC[j]; comp(sN, j) = nx.sN[j]; comp(dC, j) = nx.dC[j]; comp(dN, j) = nx.dN[j];
    }
  };
  // ---- group sums of a finished window (tile buffer `tb4`, first sample t_w, n_w samples; pairwise tree over the group's
  // tile rows) and 16-byte stores.  `live` == false (no previous window yet) only disables the stores.
  const float gscale = 1.0f / (float)G;
  auto reduce = [&](const float4* tb4, long t_w, int n_w, bool live, auto fast_t) {
    constexpr bool FAST = decltype(fast_t)::value;       // full window, 16-byte stores: straight-line, predicated stores only
#pragma unroll(FAST && NOUT * CH <= 128 ? (NOUT * CH + 31) / 32 : 1)
    for (int it = lane; it < NOUT * CH; it += 32) {
      const int orow = it / CH, ch = it % CH;
      float2 lo[RPO], hi[RPO];
#pragma unroll
      for (int k = 0; k < RPO; k++) {
        const float4 r = tb4[(orow * RPO + k) * (PX_PITCH / 4) + ch];
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
      float* o = out + (size_t)orow_g * T + t_w + 4 * ch;
      if (FAST) {
        if (valid && live) *reinterpret_cast<float4*>(o) = acc;
      } else if (valid && live) {
        if (vec_ok && 4 * ch + 3 < n_w) *reinterpret_cast<float4*>(o) = acc;
        else {
          if (4 * ch + 0 < n_w) o[0] = acc.x;
          if (4 * ch + 1 < n_w) o[1] = acc.y;
          if (4 * ch + 2 < n_w) o[2] = acc.z;
          if (4 * ch + 3 < n_w) o[3] = acc.w;
        }
      }
    }
  };

  Next nx;
  {
    const int n0 = T < (long)PX_WIN ? (int)T : PX_WIN;
    prepare(n0, nx);
    fixup(n0, nx);
    apply(nx);
  }
  int buf = 0;
  int n = 0;
  long tb = 0;
  for (; tb < T; tb += PX_WIN) {
    n = (T - tb) < (long)PX_WIN ? (int)(T - tb) : PX_WIN;
    const long left = T - tb - n;
    const int n_next = left < (long)PX_WIN ? (int)left : PX_WIN;
    float* tile = tile_mem[buf];
    float4* tile4 = reinterpret_cast<float4*>(tile);
    const float4* prev4 = reinterpret_cast<const float4*>(tile_mem[buf ^ 1]);
    if (n == PX_WIN && small_inc && vec_ok) {
      // one straight-line block: previous window's group sums, next window's boundary, this window's 64 samples
      reduce(prev4, tb - PX_WIN, PX_WIN, tb > 0, std::true_type{});
      prepare(n_next, nx);
#pragma unroll
      for (int q = 0; q < CH; q++) {
        float za[4], zb[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float2 ys, m;
          sample(ys, m, std::true_type{});
          if (G == 1) { za[i] = ys.x * m.x; zb[i] = ys.y * m.y; }
          else za[i] = __fmaf_rn(ys.y, m.y, ys.x * m.x);
        }
        tile4[lane * (PX_PITCH / 4) + q] = make_float4(za[0], za[1], za[2], za[3]);
        if (G == 1) tile4[(32 + lane) * (PX_PITCH / 4) + q] = make_float4(zb[0], zb[1], zb[2], zb[3]);
      }
    } else {
      reduce(prev4, tb - PX_WIN, PX_WIN, tb > 0, std::false_type{});
      prepare(n_next, nx);
      for (int i = 0; i < n; i++) {
        float2 ys, m;
        sample(ys, m, std::false_type{});
        if (G == 1) { tile[lane * PX_PITCH + i] = ys.x * m.x; tile[(32 + lane) * PX_PITCH + i] = ys.y * m.y; }
        else tile[lane * PX_PITCH + i] = __fmaf_rn(ys.y, m.y, ys.x * m.x);
      }
    }
    fixup(n_next, nx);
    apply(nx);
    __syncwarp();
    buf ^= 1;
  }
  if (T > 0) reduce(reinterpret_cast<const float4*>(tile_mem[buf ^ 1]), tb - PX_WIN, n, true, std::false_type{});
  // ---- persist (same state words as the interpreters' OP_SINE / OP_SVF / OP_ENVELOPE); the last prepare(0) already
  // consumed a control point crossed in the final window
#pragma unroll
  for (int j = 0; j < 2; j++) {
    th[j] += (uint64_t)ncross[j];
    if (comp(sgn, j) < 0.0f) { comp(ic1, j) = -comp(ic1, j); comp(ic2, j) = -comp(ic2, j); }
  }
  if (va < V) {     // va + 1 may be the first padded voice: its words are padding too (Vp > V), writing them is harmless
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

// p[] / s[] as filled by plan_fused() for FUSED_SINE_SVF_ENV; serves the sine oscillator at sample rates whose shortest lfo
// segment holds a 64-sample window (the wavetable oscillators and lower rates keep the one-voice-per-lane kernel in fused.cu)
cudaError_t launch_polysynth_x2(const FusedPlan& pl, const FusedArgs& a, cudaStream_t stream) {
  const int vec_ok = ((((size_t)(uintptr_t)a.out) & 15) == 0 && (a.T & 3) == 0) ? 1 : 0;
  const unsigned blocks = (unsigned)(a.Vp / 64);
#define QG_PX(LPV, GV) k_polysynth_x2<LPV, GV><<<blocks, 32, 0, stream>>>(a.params, a.state, a.Vp, a.V, a.T, vec_ok, pl.p[0], pl.p[1], pl.p[2], pl.p[3], pl.s[0], pl.s[1], pl.s[2], pl.p[7], a.out)
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
  return cudaGetLastError();
}

}  // namespace qg
