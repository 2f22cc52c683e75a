// Device-side scalar semantics.  The reference is Rust: casts saturate, f32::min/max ignore NaN, signum(±0)=±1,
// `%` is fmod, rustc never contracts a*b+c.  The whole extension is compiled with -fmad=false (no implicit FMA),
// default IEEE div/sqrt and no FTZ so that trigger/threshold paths (`!= 0.`, `>=`, quantize, ramp wrap) decide
// exactly like the CPU; transcendental functions come from libdevice and agree to ~2 ulp (DESIGN.md tolerance).
#pragma once
#if defined(__CUDACC_RTC__)
#include "rtc_compat.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "coefs.h"

namespace qg {

// saturating, NaN -> 0 (cvt.rzi.u64.f32 alone maps NaN to 0x8000000000000000)
__device__ __forceinline__ uint64_t d_as_usize(float x) { return x != x ? 0ull : __float2ull_rz(x); }
__device__ __forceinline__ int32_t d_as_i32(float x) { return __float2int_rz(x); }       // saturating, NaN -> 0
__device__ __forceinline__ bool d_is_normal(float x) {
  float a = fabsf(x);
  return a >= 1.17549435e-38f && a <= 3.402823466e+38f;
}
__device__ __forceinline__ float d_signum(float x) { return (x != x) ? x : copysignf(1.0f, x); }
__device__ __forceinline__ float d_rem_euclid(float a, float b) {
  float r = fmodf(a, b);
  return r < 0.0f ? r + fabsf(b) : r;
}
__device__ __forceinline__ float d_clamp(float x, float lo, float hi) {
  float r = x;
  if (r < lo) r = lo;
  if (r > hi) r = hi;
  return r;
}
__device__ __forceinline__ float d_fract(float x) { return x - truncf(x); }

__device__ __forceinline__ uint64_t d_atto(uint64_t state, uint64_t data) {
  uint64_t r = (state << 5) | (state >> 59);
  return (r ^ data) * 0x517cc1b727220a95ULL;
}
__device__ __forceinline__ uint64_t d_hash64a(uint64_t x) {
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}
__device__ __forceinline__ uint64_t d_hash64b(uint64_t x) {
  x = (x ^ (x >> 32)) * 0xd6e8feb86659fd93ULL;
  x = (x ^ (x >> 32)) * 0xd6e8feb86659fd93ULL;
  return x ^ (x >> 32);
}
__device__ __forceinline__ uint32_t d_hash32x(uint32_t x) {
  x = (x ^ (x >> 16)) * 0x21f0aaadU;
  x = (x ^ (x >> 15)) * 0x735a2d97U;
  return x ^ (x >> 15);
}
__device__ __forceinline__ float d_rnd1(uint64_t x) { return (float)((double)(d_hash64a(x) >> 11) / 9007199254740992.0); }
__device__ __forceinline__ float d_rnd2(uint64_t x) { return (float)((double)(d_hash64b(x) >> 11) / 9007199254740992.0); }
// white(): one sample from the counter
__device__ __forceinline__ float d_noise(uint32_t counter) {
  return (float)(int32_t)d_hash32x(counter) * (1.0f / 2147483648.0f);
}

// exp2 through f64: pitch maps (semitone_ratio) feed phase accumulators, where a 1-ulp frequency difference against the
// CPU libm (which rounds correctly) would integrate into an audible phase drift; the f64 result rounds to the same f32
__device__ __forceinline__ float d_exp2_cr(float x) { return (float)exp2((double)x); }
// same reasoning for the rest of the exp/log family (xerp, db_amp, pow ... routinely compute oscillator frequencies)
__device__ __forceinline__ float d_exp_cr(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float d_log_cr(float x) { return (float)log((double)x); }
__device__ __forceinline__ float d_log2_cr(float x) { return (float)log2((double)x); }
__device__ __forceinline__ float d_log10_cr(float x) { return (float)log10((double)x); }
__device__ __forceinline__ float d_pow_cr(float x, float y) { return (float)pow((double)x, (double)y); }
__device__ __forceinline__ float d_lerp(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float d_delerp(float a, float b, float x) { return (x - a) / (b - a); }
__device__ __forceinline__ float d_xerp(float a, float b, float t) { return d_exp_cr(d_lerp(d_log_cr(a), d_log_cr(b), t)); }
__device__ __forceinline__ float d_dexerp(float a, float b, float x) { return d_log_cr(x / a) / d_log_cr(b / a); }
__device__ __forceinline__ float d_exp10(float x) { return d_exp_cr(x * 2.30258509299404568402f); }
__device__ __forceinline__ float d_spline(float y0, float y1, float y2, float y3, float t) {
  return y1 + t / 2.0f * (y2 - y0 + t * (2.0f * y0 - 5.0f * y1 + 4.0f * y2 - y3 + t * (3.0f * (y1 - y2) + y3 - y0)));
}
__device__ __forceinline__ float d_smooth5(float x) { return ((6.0f * x - 15.0f) * x + 10.0f) * x * x * x; }
__device__ __forceinline__ float d_a_weight(float f) {
  const float c0 = 12194.0f * 12194.0f, c1 = 20.6f * 20.6f, c2 = 107.7f * 107.7f, c3 = 737.9f * 737.9f;
  const float c4 = 1.2589254f;
  float f2 = f * f;
  return c4 * c0 * f2 * f2 / ((f2 + c1) * sqrtf((f2 + c2) * (f2 + c3)) * (f2 + c0));
}

// Band-limited wavetable oscillators (FunDSP WaveSynth + Wavetable::read/at, restated; table blob built by lower.cpp
// make_wave): header = [n][n x (limit, offset(bits), length(bits))][samples...].
// Table for |f|: search from `hint`, which is only a starting point.
__device__ __forceinline__ const float* d_wavetable_select(const float* hdr, float f, uint32_t& hint, uint32_t& len) {
  const uint32_t nt = (uint32_t)hdr[0];
  const float af = fabsf(f);
  while (hint + 1 < nt && af >= hdr[1 + 3 * hint]) hint++;
  while (hint > 0 && af < hdr[1 + 3 * (hint - 1)]) hint--;
  len = __float_as_uint(hdr[3 + 3 * hint]);
  return hdr + __float_as_uint(hdr[2 + 3 * hint]);
}
// read a table (length a power of two) at phase ph in [0, 1) with the 4-point optimal interpolator
__device__ __forceinline__ float d_wavetable_interp(const float* tb, uint32_t len, float ph) {
  const uint32_t mask = len - 1;
  float pp = (float)len * ph;
  uint32_t i1 = (uint32_t)pp;
  float w = pp - (float)i1;
  uint32_t i0 = (i1 + len - 1) & mask;
  i1 &= mask;
  float a0 = __ldg(tb + i0), a1 = __ldg(tb + i1), a2 = __ldg(tb + ((i1 + 1) & mask)), a3 = __ldg(tb + ((i1 + 2) & mask));   // read-only path
  float z = w - 0.5f, even1 = a2 + a1, odd1 = a2 - a1, even2 = a3 + a0, odd2 = a3 - a0;
  float c0 = even1 * 0.46567255120778489f + even2 * 0.03432729708429672f;
  float c1 = odd1 * 0.53743830753560162f + odd2 * 0.15429462557307461f;
  float c2 = even1 * -0.25194210134021744f + even2 * 0.25194744935939062f;
  float c3 = odd1 * -0.46896069955075126f + odd2 * 0.15578800670302476f;
  float c4 = even1 * 0.00986988334359864f + even2 * -0.00989340017126506f;
  return (((c4 * z + c3) * z + c2) * z + c1) * z + c0;
}
// same read with contracted arithmetic (14 FMA-pipe ops instead of 27) for the fused kernels, which are held to the f32 audio
// tolerance rather than to the interpreters' operation order
__device__ __forceinline__ float d_wavetable_interp_fma(const float* __restrict__ tb, uint32_t len, float flen, float ph) {
  const uint32_t mask = len - 1;
  const float pp = flen * ph;
  uint32_t i1 = (uint32_t)pp;
  const float z = (pp - (float)i1) - 0.5f;
  const uint32_t i0 = (i1 + len - 1) & mask;
  i1 &= mask;
  const float a0 = __ldg(tb + i0), a1 = __ldg(tb + i1), a2 = __ldg(tb + ((i1 + 1) & mask)), a3 = __ldg(tb + ((i1 + 2) & mask));
  const float even1 = a2 + a1, odd1 = a2 - a1, even2 = a3 + a0, odd2 = a3 - a0;
  const float c0 = __fmaf_rn(even1, 0.46567255120778489f, even2 * 0.03432729708429672f);
  const float c1 = __fmaf_rn(odd1, 0.53743830753560162f, odd2 * 0.15429462557307461f);
  const float c2 = __fmaf_rn(even1, -0.25194210134021744f, even2 * 0.25194744935939062f);
  const float c3 = __fmaf_rn(odd1, -0.46896069955075126f, odd2 * 0.15578800670302476f);
  const float c4 = __fmaf_rn(even1, 0.00986988334359864f, even2 * -0.00989340017126506f);
  return __fmaf_rn(__fmaf_rn(__fmaf_rn(__fmaf_rn(c4, z, c3), z, c2), z, c1), z, c0);
}
__device__ __forceinline__ float d_wavetable_read(const float* hdr, float f, float ph, uint32_t& hint) {
  uint32_t len;
  const float* tb = d_wavetable_select(hdr, f, hint, len);
  return d_wavetable_interp(tb, len, ph);
}

// Simper SVF tick (FunDSP Svf): updates (ic1, ic2), returns m0*v0 + m1*v1 + m2*v2
__device__ __forceinline__ float d_svf_tick(float v0, float& ic1, float& ic2, float a1, float a2, float a3, float m0,
                                            float m1, float m2) {
  float v3 = v0 - ic2;
  float v1 = a1 * ic1 + a2 * v3;
  float v2 = ic2 + a2 * ic1 + a3 * v3;
  ic1 = 2.0f * v1 - ic1;
  ic2 = 2.0f * v2 - ic2;
  return m0 * v0 + m1 * v1 + m2 * v2;
}

// lfo()/lfo_in() control functions written in the reference (functions.rs:505-507, 517-540, 547-576, 811)
__device__ __forceinline__ float d_env_eval(int shape, int nin, float tt, const float* c, const float* in) {
  switch (shape) {
    case 0: { float p = nin ? in[0] : c[0]; return expf(-tt * p); }
    case 1: {
      float d = nin >= 1 ? in[0] : c[0];
      float k = nin == 2 ? in[1] : (nin == 1 ? c[0] : c[1]);
      return tt < d ? powf((d - tt) / d, k) : 0.0f;
    }
    case 2: {
      float a, ak, r, rk;
      if (nin == 0) { a = c[0]; ak = c[1]; r = c[2]; rk = c[3]; }
      else if (nin == 2) { a = in[0]; r = in[1]; ak = c[0]; rk = c[1]; }
      else { a = in[0]; ak = in[1]; r = in[2]; rk = in[3]; }
      if (tt < a) return powf(tt / a, ak);
      if (tt < a + r) return powf((r - (tt - a)) / r, rk);
      return 0.0f;
    }
    default: return tt;
  }
}

}  // namespace qg
