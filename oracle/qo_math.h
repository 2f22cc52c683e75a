// TEST INFRASTRUCTURE — CPU oracle for the quartz audio-graph hot path.  Not part of the product:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// Scalar helpers that restate (a) Rust `f32` semantics the reference relies on and (b) the few
// FunDSP 0.18.2 math functions reachable from /root/reference/src/functions.rs.  FunDSP's source is NOT
// in /root/reference (crates.io dependency, Cargo.lock:1844-1847): everything marked [U] below is a
// restatement of the published algorithm from memory and is PARITY-UNPINNED; everything marked [P] is
// pinned by in-tree reference code or by the golden vectors in tests/golden/quartz_assets.json.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace qo {

// ---- Rust cast semantics: `f32 as usize/i32/u64/i64` saturate, NaN -> 0  [P: nodes.rs:29,86,145,729]
static inline uint64_t as_usize(float x) {
  if (!(x > 0.0f)) return 0;                         // NaN, negatives, -0
  if (x >= 18446744073709551616.0f) return UINT64_MAX;
  return (uint64_t)x;
}
static inline int32_t as_i32(float x) {
  if (x != x) return 0;
  if (x >= 2147483648.0f) return INT32_MAX;
  if (x <= -2147483648.0f) return INT32_MIN;
  return (int32_t)x;
}
static inline int64_t as_i64(float x) {
  if (x != x) return 0;
  if (x >= 9223372036854775808.0f) return INT64_MAX;
  if (x <= -9223372036854775808.0f) return INT64_MIN;
  return (int64_t)x;
}
static inline bool is_normal(float x) { return std::fpclassify(x) == FP_NORMAL; }
// f32::signum: 1.0 for +0/positive/+inf, -1.0 for -0/negative/-inf, NaN for NaN  [P: functions.rs:1068]
static inline float signum(float x) { return x != x ? x : (std::signbit(x) ? -1.0f : 1.0f); }
// f32::min / f32::max ignore a NaN operand  [P: functions.rs:899-909]
static inline float rmin(float a, float b) { return std::fmin(a, b); }
static inline float rmax(float a, float b) { return std::fmax(a, b); }
// f32::rem_euclid  [P: functions.rs:923]
static inline float rem_euclid(float a, float b) {
  float r = std::fmod(a, b);
  return r < 0.0f ? r + std::fabs(b) : r;
}
// f32::clamp(lo, hi)  (NaN stays NaN)
static inline float rclamp(float x, float lo, float hi) {
  float r = x;
  if (r < lo) r = lo;
  if (r > hi) r = hi;
  return r;
}
static inline float fract(float x) { return x - std::trunc(x); }   // f32::fract

// ---- hashes  [U]
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
// AttoHash::hash (FxHasher step) [U]
static inline uint64_t atto(uint64_t state, uint64_t data) {
  return (rotl64(state, 5) ^ data) * 0x517cc1b727220a95ULL;
}
static inline uint64_t hash64a(uint64_t x) {   // SplitMix64 finaliser [U]
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}
static inline uint64_t hash64b(uint64_t x) {   // degski 64-bit [U]
  x = (x ^ (x >> 32)) * 0xd6e8feb86659fd93ULL;
  x = (x ^ (x >> 32)) * 0xd6e8feb86659fd93ULL;
  return x ^ (x >> 32);
}
static inline uint32_t hash32x(uint32_t x) {   // 2-round 32-bit mixer used by white() [U]
  x = (x ^ (x >> 16)) * 0x21f0aaadU;
  x = (x ^ (x >> 15)) * 0x735a2d97U;
  return x ^ (x >> 15);
}
static inline double rnd1(uint64_t x) { return (double)(hash64a(x) >> 11) / 9007199254740992.0; }
static inline double rnd2(uint64_t x) { return (double)(hash64b(x) >> 11) / 9007199254740992.0; }

// ---- FunDSP math [U unless noted]
static inline float lerp(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline float lerp11(float a, float b, float t) { return lerp(a, b, t * 0.5f + 0.5f); }
static inline float delerp(float a, float b, float x) { return (x - a) / (b - a); }
static inline float delerp11(float a, float b, float x) { return (x - a) / (b - a) * 2.0f - 1.0f; }
static inline float xerp(float a, float b, float t) { return std::exp(lerp(std::log(a), std::log(b), t)); }
static inline float xerp11(float a, float b, float t) { return xerp(a, b, t * 0.5f + 0.5f); }
static inline float dexerp(float a, float b, float x) { return std::log(x / a) / std::log(b / a); }
static inline float dexerp11(float a, float b, float x) { return dexerp(a, b, x) * 2.0f - 1.0f; }
static inline float exp10f_(float x) { return std::exp(x * 2.30258509299404568402f); }
static inline float db_amp(float db) { return exp10f_(db / 20.0f); }
static inline float amp_db(float a) { return std::log10(a) * 20.0f; }
static inline float a_weight(float f) {
  const float c0 = 12194.0f * 12194.0f, c1 = 20.6f * 20.6f, c2 = 107.7f * 107.7f, c3 = 737.9f * 737.9f;
  const float c4 = 1.2589254f;
  float f2 = f * f;
  return c4 * c0 * f2 * f2 / ((f2 + c1) * std::sqrt((f2 + c2) * (f2 + c3)) * (f2 + c0));
}
static inline float spline(float y0, float y1, float y2, float y3, float t) {   // Catmull-Rom
  return y1 + t / 2.0f * (y2 - y0 + t * (2.0f * y0 - 5.0f * y1 + 4.0f * y2 - y3 + t * (3.0f * (y1 - y2) + y3 - y0)));
}
static inline float softsign(float x) { return x / (1.0f + std::fabs(x)); }
static inline float smooth3(float x) { return (3.0f - 2.0f * x) * x * x; }
static inline float smooth5(float x) { return ((6.0f * x - 15.0f) * x + 10.0f) * x * x * x; }
static inline float smooth7(float x) {
  float x2 = x * x;
  return x2 * x2 * (35.0f - 84.0f * x + (70.0f - 20.0f * x) * x2);
}
static inline float smooth9(float x) {
  float x2 = x * x;
  return ((((70.0f * x - 315.0f) * x + 540.0f) * x - 420.0f) * x + 126.0f) * x2 * x2 * x;
}
static inline float uparc(float x) { return 1.0f - std::sqrt(rmax(0.0f, 1.0f - x * x)); }
static inline float downarc(float x) { return std::sqrt(rmax(0.0f, (2.0f - x) * x)); }
static inline float sine_ease(float x) { return (1.0f - std::cos(x * 3.14159265358979323846f)) * 0.5f; }
static const float TAU_F = 6.28318530717958647692f;
static const float PI_F = 3.14159265358979323846f;
static inline float sin_hz(float hz, float t) { return std::sin(t * hz * TAU_F); }
static inline float cos_hz(float hz, float t) { return std::cos(t * hz * TAU_F); }
static inline float sqr_hz(float hz, float t) {
  float x = t * hz;
  x = x - std::floor(x);
  return x < 0.5f ? 1.0f : -1.0f;
}
static inline float tri_hz(float hz, float t) {
  float x = t * hz;
  x = x - std::floor(x);
  return std::fabs(x - 0.5f) * 4.0f - 1.0f;
}
// [P] pinned bit-exact by assets/wip (tests/golden): exp2(x / 12)
static inline float semitone_ratio(float x) { return std::exp2(x / 12.0f); }
static inline float dissonance(float f0, float f1) {
  float q = std::fabs(f0 - f1) / (0.021f * rmin(f0, f1) + 19.0f);
  return 5.531753f * (std::exp(-0.84f * q) - std::exp(-1.38f * q));
}

}  // namespace qo
