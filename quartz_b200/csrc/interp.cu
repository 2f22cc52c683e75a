// The general paths: every kernel here evaluates an arbitrary lowered op tape through the same exec().
//
// K1  k_interp<uniform | divergent>  one lane = one voice, the whole render loop runs inside the kernel, sample by sample;
//     tapes with nested nets (kr / select / seq / reset ...) run a SIMT-stack emulation.  Replaces the reference's
//     per-sample virtual dispatch (`for _ in 0..len { net.tick(&[], &mut s) }`, /root/reference/src/process.rs:1347-1351
//     -> FunDSP Net::tick -> Box<dyn AudioUnit>::tick per vertex).
// K1b k_interp_blk<8>  same layout, every instruction decoded once per block of 8 samples (feed-forward tapes).
// K3/K4 + time-parallel path  k_interp_tv  one CTA per voice, threads = the samples of a hop: stateless ops on all threads,
//     block-level linear-recurrence scans for fixed LTI filters, cooperative shared-memory FFTs for rfft / ifft, exact
//     phase recurrences stepped by one thread.
// Data layout (DESIGN.md "HBM layout"): every per-voice table is [index][voice] so the 32 lanes of a warp read
// one 128-byte line; per-lane working values X live in shared memory as X[index][thread] (bank = thread, so
// every access is conflict-free); outputs are staged through per-warp shared tiles and written as full sectors
// (voice-major) or, for group mixes, summed left-to-right over the voices of a group (K6).
#if defined(__CUDACC_RTC__)
#include "rtc_compat.h"
#else
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>
#endif

#include "dev_math.cuh"
#include "kernels.h"
#include "tape.h"

namespace qg {

// All interpreter kernels carve their dynamic shared memory out of this one symbol.  Addressing it by INDEX (instead of
// through float* members) lets ptxas emit LDS/STS with plain offsets; generic pointers into shared memory were being
// re-materialised (S2UR CgaCtaId + ULEA) at every use.
extern __shared__ __align__(16) unsigned char qg_smem[];
#define QG_SMEM_F (reinterpret_cast<float*>(qg_smem))

// Execution context of one voice in lane mode (one thread = one voice).
struct Lane {
  float* x;        // shared: X[i] at x[i * nt]
  int nt;
  int v;           // voice index (padded space)
  int Vp;
  float* rings;
  const Ring* ring_tab;
  const float* tables;
  const float* state_init;
  const uint8_t* state_keep;
  const ResetRange* resets;
  int P;
  __device__ __forceinline__ float& at(int i) const { return x[i * nt]; }
  // operand access by role (exec() uses these): k-th output of an instruction, an input operand, a parameter/state scalar
  __device__ __forceinline__ float& out(int base, int k) const { return x[(base + k) * nt]; }
  __device__ __forceinline__ float& in(int f) const { return x[f * nt]; }
  __device__ __forceinline__ float& sc(int i) const { return x[i * nt]; }
  __device__ __forceinline__ float& ring(uint32_t r, uint32_t pos) const {
    return rings[(size_t)(ring_tab[r].offset + pos) * (size_t)Vp + (size_t)v];
  }
  // one sample per exec() call
  __device__ __forceinline__ int first() { return 0; }
  __device__ __forceinline__ bool more(int k) const { return k < 1; }
  __device__ __forceinline__ int next(int k) { return k + 1; }
  __device__ __forceinline__ uint32_t sample() const { return 0u; }
  __device__ __forceinline__ uint32_t count() const { return 1u; }
};
// Block lane (k_interp_blk): one thread = one voice, every instruction is applied to a block of up to BT consecutive
// samples before the next one is decoded (the tape is feed-forward, so an op only needs its own state and its operands'
// values for the same sample).  Parameters and state stay scalars; temporary i becomes BT values at
// x[(PS + (i - PS) * BT + j) * nt].
template <int BT>
struct BlockLane : Lane {
  int PS;          // P + NS: first temporary index
  int j, n;        // current sample of the block, samples in this block
  // k_interp_blk rewrites the temporary indices of the staged tape to block layout (i' = PS + (i - PS) * BT, see blk_index),
  // so an access is one add and one multiply; outputs are always temporaries, parameters/state always scalars, only input
  // operands can be either
  __device__ __forceinline__ float& out(int base, int k) const { return x[(base + k * BT + j) * nt]; }
  __device__ __forceinline__ float& in(int f) const { return x[(f + (f >= PS ? j : 0)) * nt]; }
  __device__ __forceinline__ float& sc(int i) const { return x[i * nt]; }
  __device__ __forceinline__ float& tr(int f, int jj) const { return x[(f + (f >= PS ? jj : 0)) * nt]; }   // translated index, sample jj
  __device__ __forceinline__ int first() { j = 0; return 0; }
  __device__ __forceinline__ bool more(int k) const { return k < n; }
  __device__ __forceinline__ int next(int k) { j = k + 1; return k + 1; }
  __device__ __forceinline__ uint32_t sample() const { return (uint32_t)j; }
  __device__ __forceinline__ uint32_t count() const { return (uint32_t)n; }
};
template <int BT>
__device__ __forceinline__ uint16_t blk_index(uint16_t i, int PS) {
  return (i == 0xffffu || (int)i < PS) ? i : (uint16_t)(PS + ((int)i - PS) * BT);
}
// Execution context of one THREAD of the time-vector kernel (one CTA = one voice, threads = the samples of a hop):
// parameters/state are per-voice scalars, temporaries are arrays of H samples.  Only stateless ops go through exec();
// the thread applies each decoded instruction to its samples j = j0, j0 + stride, ... < n.
struct TvSample {
  int ps_off, tmp_off;   // float offsets into the CTA's shared memory: [P + NS] scalars, temporaries [index][H]
  int PS, H;
  const float* tables;
  int j0, stride, n, j;
  __device__ __forceinline__ float& at(int i) const { return QG_SMEM_F[i < PS ? ps_off + i : tmp_off + (i - PS) * H + j]; }
  __device__ __forceinline__ float& out(int base, int k) const { return QG_SMEM_F[tmp_off + (base + k - PS) * H + j]; }
  __device__ __forceinline__ float& in(int f) const { return QG_SMEM_F[f < PS ? ps_off + f : tmp_off + (f - PS) * H + j]; }
  __device__ __forceinline__ float& sc(int i) const { return QG_SMEM_F[ps_off + i]; }
  __device__ __forceinline__ float& ring(uint32_t, uint32_t) const { return QG_SMEM_F[ps_off]; }   // never used by stateless ops
  __device__ __forceinline__ int first() { j = j0; return j0; }
  __device__ __forceinline__ bool more(int k) const { return k < n; }
  __device__ __forceinline__ int next(int k) { j = k + stride; return k + stride; }
  __device__ __forceinline__ uint32_t sample() const { return 0u; }
  __device__ __forceinline__ uint32_t count() const { return 1u; }
};

// Operand cursor of one instruction.  The generic form resolves an operand index at every access (Lane, TvSample and any
// access outside a per-sample loop).
template <class LaneT>
struct Operands {
  LaneT& L; const Instr& I; int k;
  __device__ __forceinline__ Operands(LaneT& l, const Instr& i, bool start) : L(l), I(i), k(0) { if (start) k = L.first(); }
  __device__ __forceinline__ bool more() const { return L.more(k); }
  __device__ __forceinline__ void next() { k = L.next(k); }
  __device__ __forceinline__ float& out(int q) const { return L.out((int)I.out, q); }
  __device__ __forceinline__ float& in(int q) const { return L.in((int)I.in[q]); }
};
// Block lanes walk their BT samples with one pointer per operand: the index arithmetic (is the operand a temporary or a
// scalar, times the lane stride) is done once per instruction instead of once per access (it was 7 of every ~25 issued
// instructions of an `add`), and pointers of operands a case does not touch are dead code.
template <int BT>
struct BlockOperands {
  BlockLane<BT>& L; const Instr& I;
  float* po; float* pi[5]; int si[5]; int k;
  __device__ __forceinline__ BlockOperands(BlockLane<BT>& l, const Instr& i, bool) : L(l), I(i), k(0) {
    L.j = 0;
    po = L.x + (int)I.out * L.nt;
#pragma unroll
    for (int q = 0; q < 5; q++) { const int f = (int)I.in[q]; pi[q] = L.x + f * L.nt; si[q] = f >= L.PS ? L.nt : 0; }
  }
  __device__ __forceinline__ bool more() const { return k < L.n; }
  __device__ __forceinline__ void next() {
    k++; L.j = k; po += L.nt;
#pragma unroll
    for (int q = 0; q < 5; q++) pi[q] += si[q];
  }
  __device__ __forceinline__ float& out(int q) const { return po[q * BT * L.nt]; }
  __device__ __forceinline__ float& in(int q) const { return *pi[q]; }
};
template <class LaneT> struct OperandsOf { typedef Operands<LaneT> type; };
template <int BT> struct OperandsOf<BlockLane<BT>> { typedef BlockOperands<BT> type; };

#define X(i) L.at((int)(i))
// applies a case body to every sample of the lane's block (exactly once for Lane / TvSample); inside it `q_` is the
// per-sample cursor, outside it the function-scope `q_` of exec() resolves operands by index
#define QG_EACH for (typename OperandsOf<LaneT>::type q_(L, I, true); q_.more(); q_.next())
#define XU(i) __float_as_uint(X(i))
#define SETU(i, u) X(i) = __uint_as_float(u)
// operand access by role inside exec(): k-th output, k-th input operand, parameter/state scalar
#define XO(k) q_.out((int)(k))
#define XI(k) q_.in(k)
#define XS(i) L.sc((int)(i))
#define XSU(i) __float_as_uint(XS(i))
#define SETSU(i, u) XS(i) = __uint_as_float(u)

template <class LaneT>
__device__ __forceinline__ float& ring_at(const LaneT& L, uint32_t ring, uint32_t pos) { return L.ring(ring, pos); }
__device__ __forceinline__ uint32_t ring_len(const Lane& L, uint32_t r) { return L.ring_tab[r].length; }
__device__ __forceinline__ uint32_t ring_wrap(uint32_t pos, uint32_t len) { return pos >= len ? pos - len : pos; }
__device__ __forceinline__ uint32_t ring_len(const TvSample&, uint32_t) { return 1u; }

static __device__ __noinline__ void reset_range(const TvSample&, uint32_t) {}
static __device__ __noinline__ void reset_range(const Lane& L, uint32_t id) {
  ResetRange r = L.resets[id];
  for (int s = r.s_lo; s < r.s_hi; s++)
    if (!L.state_keep[s - L.P]) X(s) = L.state_init[(size_t)(s - L.P) * L.Vp + L.v];   // a nested seq() keeps its event list
  for (int g = r.ring_lo; g < r.ring_hi; g++) {
    uint32_t len = L.ring_tab[g].length;
    for (uint32_t k = 0; k < len; k++) ring_at(L, g, k) = 0.0f;
  }
}

// in-place radix-2 FFT over two of the lane's HBM rings (re, im); inverse scales by 1/N
static __device__ __noinline__ void lane_fft(const TvSample&, uint32_t, uint32_t, int, const float*, bool) {}
static __device__ __noinline__ void lane_fft(const Lane& L, uint32_t rre, uint32_t rim, int lg, const float* tw, bool inverse) {
  uint32_t N = 1u << lg;
  for (uint32_t i = 1, j = 0; i < N; i++) {
    uint32_t bit = N >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      float a = ring_at(L, rre, i), b = ring_at(L, rim, i);
      ring_at(L, rre, i) = ring_at(L, rre, j); ring_at(L, rim, i) = ring_at(L, rim, j);
      ring_at(L, rre, j) = a; ring_at(L, rim, j) = b;
    }
  }
  for (uint32_t len = 2; len <= N; len <<= 1) {
    uint32_t half = len >> 1, step = N / len;
    for (uint32_t k = 0; k < half; k++) {
      float wr = tw[2 * k * step], wi = tw[2 * k * step + 1];
      if (inverse) wi = -wi;
      for (uint32_t i = k; i < N; i += len) {
        float ur = ring_at(L, rre, i), ui = ring_at(L, rim, i);
        float vr = ring_at(L, rre, i + half), vi = ring_at(L, rim, i + half);
        float tr = vr * wr - vi * wi, ti = vr * wi + vi * wr;
        ring_at(L, rre, i) = ur + tr; ring_at(L, rim, i) = ui + ti;
        ring_at(L, rre, i + half) = ur - tr; ring_at(L, rim, i + half) = ui - ti;
      }
    }
  }
  if (inverse) {
    float s = 1.0f / (float)N;
    for (uint32_t i = 0; i < N; i++) { ring_at(L, rre, i) *= s; ring_at(L, rim, i) *= s; }
  }
}

// Execute one instruction for this lane.  `pc` is only touched by control-flow ops.
template <class LaneT>
__device__ __forceinline__ void exec(const Instr& I, LaneT& L, int& pc) {
  Operands<LaneT> q_(L, I, false);   // operand access outside a per-sample loop
  switch (I.op) {
    case OP_NOP: break;
    case OP_MOV: QG_EACH { XO(0) = XI(0); } break;
    case OP_ZERO: QG_EACH { XO(0) = 0.0f; } break;
    case OP_LD_STATE: QG_EACH { XO(0) = XS(I.s); } break;
    case OP_ST_STATE: QG_EACH { XS(I.s) = XI(0); } break;
    case OP_ADD: QG_EACH { XO(0) = XI(0) + XI(1); } break;
    case OP_SUB: QG_EACH { XO(0) = XI(0) - XI(1); } break;
    case OP_MUL: QG_EACH { XO(0) = XI(0) * XI(1); } break;
    case OP_GT: QG_EACH { XO(0) = XI(0) > XI(1) ? 1.0f : 0.0f; } break;
    case OP_LT: QG_EACH { XO(0) = XI(0) < XI(1) ? 1.0f : 0.0f; } break;
    case OP_EQ: QG_EACH { XO(0) = XI(0) == XI(1) ? 1.0f : 0.0f; } break;
    case OP_NE: QG_EACH { XO(0) = XI(0) != XI(1) ? 1.0f : 0.0f; } break;
    case OP_GE: QG_EACH { XO(0) = XI(0) >= XI(1) ? 1.0f : 0.0f; } break;
    case OP_LE: QG_EACH { XO(0) = XI(0) <= XI(1) ? 1.0f : 0.0f; } break;
    case OP_MIN: QG_EACH { XO(0) = fminf(XI(0), XI(1)); } break;
    case OP_MAX: QG_EACH { XO(0) = fmaxf(XI(0), XI(1)); } break;
    case OP_POW: QG_EACH { XO(0) = d_pow_cr(XI(0), XI(1)); } break;
    case OP_REM: QG_EACH { XO(0) = d_rem_euclid(XI(0), XI(1)); } break;
    case OP_LOG: QG_EACH { XO(0) = d_log_cr(XI(0)) / d_log_cr(XI(1)); } break;
    case OP_BITAND: QG_EACH { XO(0) = (float)(d_as_i32(XI(0)) & d_as_i32(XI(1))); } break;
    case OP_BITOR: QG_EACH { XO(0) = (float)(d_as_i32(XI(0)) | d_as_i32(XI(1))); } break;
    case OP_BITXOR: QG_EACH { XO(0) = (float)(d_as_i32(XI(0)) ^ d_as_i32(XI(1))); } break;
    case OP_SHL: QG_EACH { XO(0) = (float)(int32_t)((uint32_t)d_as_i32(XI(0)) << (uint32_t)(d_as_usize(XI(1)) & 31)); } break;
    case OP_SHR: QG_EACH { XO(0) = (float)(d_as_i32(XI(0)) >> (uint32_t)(d_as_usize(XI(1)) & 31)); } break;
    case OP_HYPOT: QG_EACH { XO(0) = hypotf(XI(0), XI(1)); } break;
    case OP_ATAN2: QG_EACH { XO(0) = atan2f(XI(0), XI(1)); } break;
    case OP_DISSONANCE: QG_EACH {
      float f0 = XI(0), f1 = XI(1);
      float q = fabsf(f0 - f1) / (0.021f * fminf(f0, f1) + 19.0f);
      XO(0) = 5.531753f * (expf(-0.84f * q) - expf(-1.38f * q));
    } break;
    case OP_SIN_HZ: QG_EACH { XO(0) = sinf(XI(1) * XI(0) * QG_TAU); } break;
    case OP_COS_HZ: QG_EACH { XO(0) = cosf(XI(1) * XI(0) * QG_TAU); } break;
    case OP_SQR_HZ: QG_EACH { float x = XI(1) * XI(0); x = x - floorf(x); XO(0) = x < 0.5f ? 1.0f : -1.0f; } break;
    case OP_TRI_HZ: QG_EACH { float x = XI(1) * XI(0); x = x - floorf(x); XO(0) = fabsf(x - 0.5f) * 4.0f - 1.0f; } break;
    case OP_PDHALF_BI: QG_EACH {   // functions.rs:677-688
      float x = XI(0), mid = d_clamp(XI(1), -1.0f, 1.0f);
      if (x < mid) { float ls = mid != -1.0f ? 1.0f / (mid + 1.0f) : 0.0f; XO(0) = ls * x; }
      else { float rs = mid != 1.0f ? 1.0f / (1.0f - mid) : 0.0f; XO(0) = rs * (x - mid) + 0.5f; }
    } break;
    case OP_PDHALF_UNI: QG_EACH {   // functions.rs:689-706
      float x = XI(0), m = XI(1);
      float mid = m >= 1.0f ? 1.0f : (m <= -1.0f ? 0.0f : (m + 1.0f) / 2.0f);
      if (x < mid) { float ls = mid != 0.0f ? 0.5f / mid : 0.0f; XO(0) = ls * x; }
      else { float rs = mid != 1.0f ? 0.5f / (1.0f - mid) : 0.0f; XO(0) = rs * (x - mid) + 0.5f; }
    } break;
    case OP_LERP: QG_EACH { XO(0) = d_lerp(XI(0), XI(1), XI(2)); } break;
    case OP_LERP11: QG_EACH { XO(0) = d_lerp(XI(0), XI(1), XI(2) * 0.5f + 0.5f); } break;
    case OP_DELERP: QG_EACH { XO(0) = d_delerp(XI(0), XI(1), XI(2)); } break;
    case OP_DELERP11: QG_EACH { XO(0) = d_delerp(XI(0), XI(1), XI(2)) * 2.0f - 1.0f; } break;
    case OP_XERP: QG_EACH { XO(0) = d_xerp(XI(0), XI(1), XI(2)); } break;
    case OP_XERP11: QG_EACH { XO(0) = d_xerp(XI(0), XI(1), XI(2) * 0.5f + 0.5f); } break;
    case OP_DEXERP: QG_EACH { XO(0) = d_dexerp(XI(0), XI(1), XI(2)); } break;
    case OP_DEXERP11: QG_EACH { XO(0) = d_dexerp(XI(0), XI(1), XI(2)) * 2.0f - 1.0f; } break;
    case OP_SPLINE: QG_EACH { XO(0) = d_spline(XI(0), XI(1), XI(2), XI(3), XI(4)); } break;
    case OP_ABS: QG_EACH { XO(0) = fabsf(XI(0)); } break;
    case OP_SIGNUM: QG_EACH { XO(0) = d_signum(XI(0)); } break;
    case OP_FLOOR: QG_EACH { XO(0) = floorf(XI(0)); } break;
    case OP_FRACT: QG_EACH { XO(0) = d_fract(XI(0)); } break;
    case OP_CEIL: QG_EACH { XO(0) = ceilf(XI(0)); } break;
    case OP_ROUND: QG_EACH { XO(0) = roundf(XI(0)); } break;
    case OP_SQRT: QG_EACH { XO(0) = sqrtf(XI(0)); } break;
    case OP_EXP: QG_EACH { XO(0) = d_exp_cr(XI(0)); } break;
    case OP_EXP2: QG_EACH { XO(0) = d_exp2_cr(XI(0)); } break;
    case OP_EXP10: QG_EACH { XO(0) = d_exp10(XI(0)); } break;
    case OP_LN_1P_FN: QG_EACH { XO(0) = log1pf(XI(0)); } break;
    case OP_EXP_M1_FN: QG_EACH { XO(0) = expm1f(XI(0)); } break;
    case OP_LN: QG_EACH { XO(0) = d_log_cr(XI(0)); } break;
    case OP_LOG2: QG_EACH { XO(0) = d_log2_cr(XI(0)); } break;
    case OP_LOG10: QG_EACH { XO(0) = d_log10_cr(XI(0)); } break;
    case OP_SIN: QG_EACH { XO(0) = sinf(XI(0)); } break;
    case OP_COS: QG_EACH { XO(0) = cosf(XI(0)); } break;
    case OP_TAN: QG_EACH { XO(0) = tanf(XI(0)); } break;
    case OP_ASIN: QG_EACH { XO(0) = asinf(XI(0)); } break;
    case OP_ACOS: QG_EACH { XO(0) = acosf(XI(0)); } break;
    case OP_ATAN: QG_EACH { XO(0) = atanf(XI(0)); } break;
    case OP_SINH: QG_EACH { XO(0) = sinhf(XI(0)); } break;
    case OP_COSH: QG_EACH { XO(0) = coshf(XI(0)); } break;
    case OP_TANH: QG_EACH { XO(0) = tanhf(XI(0)); } break;
    case OP_ASINH: QG_EACH { XO(0) = asinhf(XI(0)); } break;
    case OP_ACOSH: QG_EACH { XO(0) = acoshf(XI(0)); } break;
    case OP_ATANH: QG_EACH { XO(0) = atanhf(XI(0)); } break;
    case OP_SQUARED: QG_EACH { float x = XI(0); XO(0) = x * x; } break;
    case OP_CUBED: QG_EACH { float x = XI(0); XO(0) = x * x * x; } break;
    case OP_DB_AMP: QG_EACH { XO(0) = d_exp10(XI(0) / 20.0f); } break;
    case OP_AMP_DB: QG_EACH { XO(0) = d_log10_cr(XI(0)) * 20.0f; } break;
    case OP_A_WEIGHT: QG_EACH { XO(0) = d_a_weight(XI(0)); } break;
    case OP_SOFTSIGN: QG_EACH { float x = XI(0); XO(0) = x / (1.0f + fabsf(x)); } break;
    case OP_SMOOTH3: QG_EACH { float x = XI(0); XO(0) = (3.0f - 2.0f * x) * x * x; } break;
    case OP_SMOOTH5: QG_EACH { XO(0) = d_smooth5(XI(0)); } break;
    case OP_SMOOTH7: QG_EACH { float x = XI(0), x2 = x * x; XO(0) = x2 * x2 * (35.0f - 84.0f * x + (70.0f - 20.0f * x) * x2); } break;
    case OP_SMOOTH9: QG_EACH {
      float x = XI(0), x2 = x * x;
      XO(0) = ((((70.0f * x - 315.0f) * x + 540.0f) * x - 420.0f) * x + 126.0f) * x2 * x2 * x;
    } break;
    case OP_UPARC: QG_EACH { float x = XI(0); XO(0) = 1.0f - sqrtf(fmaxf(0.0f, 1.0f - x * x)); } break;
    case OP_DOWNARC: QG_EACH { float x = XI(0); XO(0) = sqrtf(fmaxf(0.0f, (2.0f - x) * x)); } break;
    case OP_SINE_EASE: QG_EACH { XO(0) = (1.0f - cosf(XI(0) * QG_PI)) * 0.5f; } break;
    case OP_SEMITONE_RATIO: QG_EACH { XO(0) = d_exp2_cr(XI(0) / 12.0f); } break;
    case OP_RND1: QG_EACH { XO(0) = d_rnd1(d_as_usize(XI(0))); } break;
    case OP_RND2: QG_EACH { XO(0) = d_rnd2(d_as_usize(XI(0))); } break;
    case OP_DEG: QG_EACH { XO(0) = XI(0) * 57.2957795130823208767981548141051703f; } break;
    case OP_RAD: QG_EACH { XO(0) = XI(0) * (QG_PI / 180.0f); } break;
    case OP_RECIP: QG_EACH { XO(0) = 1.0f / XI(0); } break;
    case OP_NORMAL: QG_EACH { float x = XI(0); XO(0) = d_is_normal(x) ? x : 0.0f; } break;
    case OP_CLIP: QG_EACH { XO(0) = d_clamp(XI(0), XS(I.p), XS(I.p + 1)); } break;
    case OP_WRAP2: QG_EACH { float p0 = XS(I.p), r = XS(I.p + 1); XO(0) = fmodf(fmodf(XI(0) - p0, r) + r, r) + p0; } break;
    case OP_WRAP1: QG_EACH { float x0 = XS(I.p), x = XI(0); XO(0) = x - x0 * floorf(x / x0); } break;
    case OP_MIRROR: QG_EACH {   // functions.rs:1167-1180
      float p0 = XS(I.p), p1 = XS(I.p + 1), r = XS(I.p + 2), x = XI(0);
      float n = d_is_normal(x) ? x : 0.0f, res;
      if (n >= p0 && n <= p1) res = n;
      else {
        float distance = fminf(n - p1, p0 - n);
        float folds = floorf(distance / r);
        if ((n > p1 && fmodf(folds, 2.0f) == 0.0f) || (n < p0 && fmodf(folds, 2.0f) != 0.0f)) res = p0 + (distance - folds * r);
        else res = p1 - (distance - folds * r);
      }
      XO(0) = res;
    } break;
    case OP_POL: QG_EACH { float a = XI(0), b = XI(1); XO(0) = hypotf(a, b); XO(1) = atan2f(b, a); } break;
    case OP_CAR: QG_EACH { float a = XI(0), b = XI(1), sn, cs; sincosf(b, &sn, &cs); XO(0) = a * cs; XO(1) = a * sn; } break;
    case OP_DIVN: QG_EACH { XO(0) = XI(0) / (float)I.n; } break;
    case OP_JOIN: QG_EACH {   // left-to-right sum of n <= 5 operands, divided by aux when aux != 0 (join(n): mean of n)
      float s = XI(0);
      if (I.n > 1) s += XI(1);
      if (I.n > 2) s += XI(2);
      if (I.n > 3) s += XI(3);
      if (I.n > 4) s += XI(4);
      XO(0) = I.aux ? s / (float)I.aux : s;
    } break;
    case OP_PAN: QG_EACH { float x = XI(0); XO(0) = XS(I.p) * x; XO(1) = XS(I.p + 1) * x; } break;
    case OP_PAN_VAR: QG_EACH {
      float x = XI(0), pan = XI(1);
      if (pan != XS(I.s)) { float l, r; pan_weights(pan, &l, &r); XS(I.s) = pan; XS(I.s + 1) = l; XS(I.s + 2) = r; }
      XO(0) = XS(I.s + 1) * x; XO(1) = XS(I.s + 2) * x;
    } break;
    case OP_ROTATE: QG_EACH {
      float a = XI(0), b = XI(1), c = XS(I.p), s = XS(I.p + 1);
      XO(0) = c * a - s * b; XO(1) = s * a + c * b;
    } break;
    // ---------------------------------------------------------------- sources
    case OP_SINE: QG_EACH {
      float ph = XS(I.s);
      float np = ph + XI(0) * XS(I.p);
      np -= floorf(np);
      XS(I.s) = np;
      XO(0) = sinf(ph * QG_TAU);
    } break;
    case OP_NOISE: QG_EACH { uint32_t c = XSU(I.s) + 1u; SETSU(I.s, c); XO(0) = d_noise(c); } break;
    case OP_IMPULSE: QG_EACH { uint32_t f = XSU(I.s); XO(0) = f ? 0.0f : 1.0f; SETSU(I.s, 1u); } break;
    case OP_RAMP: QG_EACH {   // nodes.rs:476-483
      float val = XS(I.s);
      XO(0) = val;
      val += XI(0) / XS(I.p);
      if (val >= 1.0f) val -= 1.0f;
      XS(I.s) = val;
    } break;
    case OP_WAVETABLE: QG_EACH {   // FunDSP WaveSynth + Wavetable::read/at (restated, see lower.cpp make_wave)
      float f = XI(0);
      float ph = XS(I.s) + f * XS(I.p);
      ph -= floorf(ph);
      XS(I.s) = ph;
      uint32_t hint = XSU(I.s + 1);
      XO(0) = d_wavetable_read(L.tables + I.aux, f, ph, hint);
      SETSU(I.s + 1, hint);
    } break;
    case OP_WAVE: QG_EACH {
      uint32_t i = XSU(I.s);
      XO(0) = __ldg(L.tables + I.aux + i);
      i += 1;
      if (i >= I.aux2) i = 0;
      SETSU(I.s, i);
    } break;
    // ---------------------------------------------------------------- filters
    case OP_SVF: QG_EACH {
      float ic1 = XS(I.s), ic2 = XS(I.s + 1);
      XO(0) = d_svf_tick(XI(0), ic1, ic2, XS(I.p), XS(I.p + 1), XS(I.p + 2), XS(I.p + 3), XS(I.p + 4), XS(I.p + 5));
      XS(I.s) = ic1; XS(I.s + 1) = ic2;
    } break;
    case OP_SVF_VAR: QG_EACH {
      int mode = I.n & 0xff, nvar = I.n >> 8;
      float hz = nvar >= 1 ? XI(1) : XS(I.p), q = nvar >= 2 ? XI(2) : XS(I.p + 1), g = nvar >= 3 ? XI(3) : XS(I.p + 2);
      if (hz != XS(I.s + 8) || q != XS(I.s + 9) || g != XS(I.s + 10)) {
        float c[6];
        svf_coefs(mode, hz, q, g, XS(I.p + 3), c);
        for (int k = 0; k < 6; k++) XS(I.s + 2 + k) = c[k];
        XS(I.s + 8) = hz; XS(I.s + 9) = q; XS(I.s + 10) = g;
      }
      float ic1 = XS(I.s), ic2 = XS(I.s + 1);
      XO(0) = d_svf_tick(XI(0), ic1, ic2, XS(I.s + 2), XS(I.s + 3), XS(I.s + 4), XS(I.s + 5), XS(I.s + 6), XS(I.s + 7));
      XS(I.s) = ic1; XS(I.s + 1) = ic2;
    } break;
    case OP_BIQUAD: QG_EACH {
      float x0 = XI(0), x1 = XS(I.s), x2 = XS(I.s + 1), y1 = XS(I.s + 2), y2 = XS(I.s + 3);
      float y0 = XS(I.p + 2) * x0 + XS(I.p + 3) * x1 + XS(I.p + 4) * x2 - XS(I.p) * y1 - XS(I.p + 1) * y2;
      XS(I.s) = x0; XS(I.s + 1) = x1; XS(I.s + 2) = y0; XS(I.s + 3) = y1;
      XO(0) = y0;
    } break;
    case OP_BIQUAD_VAR: QG_EACH {
      int kind = I.n & 0xff, nvar = I.n >> 8;
      float c0 = XI(1), c1 = nvar >= 2 ? XI(2) : XS(I.s + 10);
      if (c0 != XS(I.s + 9) || c1 != XS(I.s + 10)) {
        float c[5];
        biquad_coefs(kind, c0, c1, XS(I.p), c);
        for (int k = 0; k < 5; k++) XS(I.s + 4 + k) = c[k];
        XS(I.s + 9) = c0; XS(I.s + 10) = c1;
      }
      float x0 = XI(0), x1 = XS(I.s), x2 = XS(I.s + 1), y1 = XS(I.s + 2), y2 = XS(I.s + 3);
      float y0 = XS(I.s + 6) * x0 + XS(I.s + 7) * x1 + XS(I.s + 8) * x2 - XS(I.s + 4) * y1 - XS(I.s + 5) * y2;
      XS(I.s) = x0; XS(I.s + 1) = x1; XS(I.s + 2) = y0; XS(I.s + 3) = y1;
      XO(0) = y0;
    } break;
    case OP_ONEPOLE: case OP_ONEPOLE_VAR: QG_EACH {
      float coeff;
      if (I.op == OP_ONEPOLE) coeff = XS(I.p);
      else {
        float p = XI(1);
        if (p != XS(I.s + 3)) { XS(I.s + 3) = p; XS(I.s + 2) = onepole_coef(I.n, p, XS(I.p)); }
        coeff = XS(I.s + 2);
      }
      float x = XI(0), x1 = XS(I.s), y1 = XS(I.s + 1), y;
      switch (I.n) {
        case 0: y = (1.0f - coeff) * x + coeff * y1; break;
        case 1: y = coeff * (y1 + x - x1); break;
        case 2: y = x - x1 + coeff * y1; break;
        default: y = coeff * (x - y1) + x1; break;
      }
      XS(I.s) = x; XS(I.s + 1) = y;
      XO(0) = y;
    } break;
    case OP_PINKPASS: QG_EACH {
      float w = XI(0);
      float b0 = 0.99886f * XS(I.s) + w * 0.0555179f;
      float b1 = 0.99332f * XS(I.s + 1) + w * 0.0750759f;
      float b2 = 0.96900f * XS(I.s + 2) + w * 0.1538520f;
      float b3 = 0.86650f * XS(I.s + 3) + w * 0.3104856f;
      float b4 = 0.55000f * XS(I.s + 4) + w * 0.5329522f;
      float b5 = -0.7616f * XS(I.s + 5) - w * 0.0168980f;
      float pink = b0 + b1 + b2 + b3 + b4 + b5 + XS(I.s + 6) + w * 0.5362f;
      XS(I.s) = b0; XS(I.s + 1) = b1; XS(I.s + 2) = b2; XS(I.s + 3) = b3; XS(I.s + 4) = b4; XS(I.s + 5) = b5;
      XS(I.s + 6) = w * 0.115926f;
      XO(0) = pink * 0.11f;
    } break;
    case OP_FIR: QG_EACH {
      int n = I.n;
      for (int k = n - 1; k > 0; k--) XS(I.s + k) = XS(I.s + k - 1);
      XS(I.s) = XI(0);
      float acc = 0.0f;
      for (int k = 0; k < n; k++) acc += XS(I.p + k) * XS(I.s + k);
      XO(0) = acc;
    } break;
    // ---------------------------------------------------------------- delays
    case OP_TICK: QG_EACH { float v = XS(I.s); XS(I.s) = XI(0); XO(0) = v; } break;
    case OP_DELAY: QG_EACH {
      uint32_t i = XSU(I.s), len = ring_len(L, I.aux);
      float& slot = ring_at(L, I.aux, i);
      float o = slot;
      slot = XI(0);
      i = i + 1 == len ? 0 : i + 1;
      SETSU(I.s, i);
      XO(0) = o;
    } break;
    case OP_TAP: QG_EACH {
      uint32_t idx = XSU(I.s), len = ring_len(L, I.aux), mask = len - 1;
      ring_at(L, I.aux, idx) = XI(0);
      float tap = d_clamp(XI(1), XS(I.p), XS(I.p + 1)) * XS(I.p + 2);
      if (tap != tap) tap = 0.0f;
      uint32_t fl = (uint32_t)d_as_usize(tap);
      float d = tap - (float)fl;
      uint32_t i1 = (idx + len - fl) & mask;
      if (I.n) {
        uint32_t i0 = (i1 + 1) & mask, i2 = (i1 + len - 1) & mask, i3 = (i1 + len - 2) & mask;
        XO(0) = d_spline(ring_at(L, I.aux, i0), ring_at(L, I.aux, i1), ring_at(L, I.aux, i2), ring_at(L, I.aux, i3), d);
      } else {
        uint32_t i2 = (i1 + len - 1) & mask;
        XO(0) = d_lerp(ring_at(L, I.aux, i1), ring_at(L, I.aux, i2), d);
      }
      SETSU(I.s, (idx + 1) & mask);
    } break;
    case OP_SAMP_DELAY: QG_EACH {   // nodes.rs:726-731: push_front, pop_back, then index from the front
      uint32_t head = XSU(I.s), len = ring_len(L, I.aux);
      head = head == 0 ? len - 1 : head - 1;
      ring_at(L, I.aux, head) = XI(0);
      SETSU(I.s, head);
      uint64_t k = d_as_usize(XI(1));
      float o = 0.0f;
      if (k < (uint64_t)len) { uint32_t pos = head + (uint32_t)k; if (pos >= len) pos -= len; o = ring_at(L, I.aux, pos); }
      XO(0) = o;
    } break;
    case OP_ENVELOPE: QG_EACH {
      int shape = I.n & 0xff, nin = I.n >> 8;
      float t = XS(I.s), t0 = XS(I.s + 1), t1 = XS(I.s + 2), v0 = XS(I.s + 3), v1 = XS(I.s + 4);
      if (t >= t1) {
        float c[4] = {XS(I.p), XS(I.p + 1), XS(I.p + 2), XS(I.p + 3)};
        float in[4];
        in[0] = nin > 0 ? XI(0) : 0.0f; in[1] = nin > 1 ? XI(1) : 0.0f;
        in[2] = nin > 2 ? XI(2) : 0.0f; in[3] = nin > 3 ? XI(3) : 0.0f;
        if (XSU(I.s + 7)) { v1 = d_env_eval(shape, nin, 0.0f, c, in); SETSU(I.s + 7, 0u); }
        uint64_t th = (uint64_t)XSU(I.s + 5) | ((uint64_t)XSU(I.s + 6) << 32);
        t0 = t1;
        v0 = v1;
        float next = d_lerp(0.75f, 1.25f, d_rnd1(th)) * 0.002f;
        t1 = t0 + next;
        v1 = d_env_eval(shape, nin, t1, c, in);
        th += 1;
        SETSU(I.s + 5, (uint32_t)th); SETSU(I.s + 6, (uint32_t)(th >> 32));
        XS(I.s + 1) = t0; XS(I.s + 2) = t1; XS(I.s + 3) = v0; XS(I.s + 4) = v1;
      }
      float u = d_delerp(t0, t1, t);
      XS(I.s) = t + XS(I.p + 4);
      XO(0) = d_lerp(v0, v1, u);
    } break;
    case OP_DECLICK: QG_EACH {
      float t = XS(I.s), dur = XS(I.p), x = XI(0);
      if (t < dur) { XO(0) = x * d_smooth5(t / dur); XS(I.s) = t + XS(I.p + 1); }
      else XO(0) = x;
    } break;
    // ---------------------------------------------------------------- in-tree stateful nodes
    case OP_SHIFT_REG: QG_EACH {   // nodes.rs:173-185
      if (XI(1) != 0.0f) {
        for (int k = 7; k > 0; k--) XS(I.s + k) = XS(I.s + k - 1);
        XS(I.s) = XI(0);
      }
      for (int k = 0; k < 8; k++) XO(k) = XS(I.s + k);
    } break;
    case OP_SNH: QG_EACH {   // nodes.rs:811-816
      if (XI(1) != 0.0f) XS(I.s) = XI(0);
      XO(0) = XS(I.s);
    } break;
    case OP_QUANTIZE: {   // nodes.rs:213-228
      // short step tables (the scenes use 8 entries) are read once per instruction; a NaN pad never wins `d < dist`
      float tb[8];
      const bool small = I.aux2 <= 8u;
#pragma unroll
      for (uint32_t k = 0; k < 8; k++) tb[k] = (small && k < I.aux2) ? __ldg(L.tables + I.aux + k) : __int_as_float(0x7fc00000);   // read-only for the kernel's lifetime: hoistable out of the sample loop
      QG_EACH {
        float n = XI(0), range = XS(I.p);
        float wrapped = n - range * floorf(n / range);
        float nearest = 0.0f, dist = FLT_MAX;
        if (small) {
#pragma unroll
          for (uint32_t k = 0; k < 8; k++) {
            float d = fabsf(wrapped - tb[k]);
            if (d < dist) { nearest = tb[k]; dist = d; }
          }
        } else {
          for (uint32_t k = 0; k < I.aux2; k++) {
            float v = __ldg(L.tables + I.aux + k);
            float d = fabsf(wrapped - v);
            if (d < dist) { nearest = v; dist = d; }
          }
        }
        XO(0) = n + nearest - wrapped;
      }
    } break;
    case OP_ARR_GET: QG_EACH {   // nodes.rs:143-149
      uint64_t k = d_as_usize(XI(0));
      XO(0) = k < (uint64_t)I.aux2 ? __ldg(L.tables + I.aux + (uint32_t)k) : 0.0f;
    } break;
    // ---------------------------------------------------------------- control flow
    case OP_KR_BEGIN: {   // nodes.rs:272-275
      uint32_t c = XSU(I.s);
      if (c == 0) SETSU(I.s, I.aux2);
      else pc = (int)I.aux;
      break;
    }
    case OP_KR_END: SETSU(I.s, XSU(I.s) - 1u); break;
    case OP_RESET_EVERY: {   // nodes.rs:353-358
      uint32_t c = XSU(I.s);
      if (c >= I.aux) { { const LaneT Lc = L; reset_range(Lc, I.aux2); } c = 0; }
      SETSU(I.s, c + 1u);
      break;
    }
    case OP_RESET_IF: if (XI(0) != 0.0f) { const LaneT Lc = L; reset_range(Lc, I.aux2); } break;
    case OP_RESET_V: {   // nodes.rs:435-440
      uint32_t c = XSU(I.s);
      uint64_t lim = d_as_usize(roundf(XI(0) * XS(I.p)));
      if ((uint64_t)c >= lim) { { const LaneT Lc = L; reset_range(Lc, I.aux2); } c = 0; }
      SETSU(I.s, c + 1u);
      break;
    }
    case OP_JNE_IDX: if (d_as_usize(XI(0)) != (uint64_t)I.aux2) pc = (int)I.aux; break;
    case OP_SEQ_TRIG: {   // nodes.rs:76-91
      const int nk = I.n;
      if (XI(0) != 0.0f) {
        const uint64_t k = d_as_usize(XI(1));
        if (k < (uint64_t)nk) {
          const int b = I.s + 1 + 5 * (int)k;
          { const LaneT Lc = L; reset_range(Lc, I.aux2 + (uint32_t)k); }
          const uint64_t dl = d_as_usize(roundf(XI(2) * XS(I.p))), du = d_as_usize(roundf(XI(3) * XS(I.p)));
          const uint32_t stamp = XSU(I.s) + 1u;
          SETSU(I.s, stamp);
          SETU(b, 1u);
          SETU(b + 1, dl > 0xffffffffull ? 0xffffffffu : (uint32_t)dl);
          SETU(b + 2, du > 0xffffffffull ? 0xffffffffu : (uint32_t)du);
          SETU(b + 3, stamp);
        }
      }
      for (int k = 0; k < nk; k++) {   // events.retain(|x| x.2 != 0)
        const int b = I.s + 1 + 5 * k;
        if (XU(b) && XU(b + 2) == 0u) SETU(b, 0u);
        SETU(b + 4, 0u);
      }
      break;
    }
    case OP_SEQ_GATE: {   // nodes.rs:95-105
      const int b = I.s + 1 + 5 * (int)I.n;
      bool tick = false;
      if (XU(b)) {
        if (XU(b + 1) == 0u) { SETU(b + 2, XU(b + 2) - 1u); SETU(b + 4, 1u); tick = true; }
        else SETU(b + 1, XU(b + 1) - 1u);
      }
      if (!tick) pc = (int)I.aux;
      break;
    }
    case OP_SEQ_END: {   // out += buffer, in event-list (= push) order
      const int nk = I.n;
      float acc = 0.0f;
      uint32_t done = 0u;
      for (int r = 0; r < nk; r++) {
        int best = -1;
        uint32_t bs = 0xffffffffu;
        for (int k = 0; k < nk; k++) {
          const int b = I.s + 1 + 5 * k;
          const uint32_t stamp = XU(b + 3);
          if (XU(b + 4) && stamp > done && stamp <= bs) { bs = stamp; best = k; }
        }
        if (best < 0) break;
        acc += X(I.in[0] + best);
        done = bs;
      }
      XO(0) = acc;
      break;
    }
    // ---------------------------------------------------------------- feedback
    // block lanes run FB_READ for every sample of the block before FB_WRITE stores any (ring length >= block length is
    // checked by the launcher), so both address the ring at idx + j and the index advances once per block
    case OP_FB_READ: QG_EACH { XO(0) = XI(0) + ring_at(L, I.aux, ring_wrap(XSU(I.s) + L.sample(), ring_len(L, I.aux))); } break;
    case OP_FB_WRITE: {
      const uint32_t len = ring_len(L, I.aux), i = XSU(I.s);
      QG_EACH { ring_at(L, I.aux, ring_wrap(i + L.sample(), len)) = XI(0); }
      if (I.n) SETSU(I.s, ring_wrap(i + L.count(), len));
      break;
    }
    case OP_FB1_READ: QG_EACH { XO(0) = XI(0) + XS(I.s); } break;
    case OP_FB1_WRITE: QG_EACH { XS(I.s) = XI(0); } break;
    // ---------------------------------------------------------------- spectral nodes (per-lane path)
    case OP_RFFT: QG_EACH {   // nodes.rs:625-642
      uint32_t N = 1u << I.n, i = XSU(I.s), nx = i + 1 == N ? 0 : i + 1;
      SETSU(I.s, nx);
      if (i == 0) {
        for (uint32_t k = 0; k < N; k++) { ring_at(L, I.aux + 1, k) = ring_at(L, I.aux, k); ring_at(L, I.aux + 2, k) = 0.0f; }
        { const LaneT Lc = L; lane_fft(Lc, I.aux + 1, I.aux + 2, I.n, L.tables + I.aux2, false); }
      }
      ring_at(L, I.aux, i) = XI(0);
      if (i <= N / 2) { XO(0) = ring_at(L, I.aux + 1, i); XO(1) = ring_at(L, I.aux + 2, i); }
      else { XO(0) = ring_at(L, I.aux + 1, N - i); XO(1) = -ring_at(L, I.aux + 2, N - i); }
    } break;
    case OP_IFFT: QG_EACH {   // nodes.rs:681-693
      uint32_t N = 1u << I.n, i = XSU(I.s), nx = i + 1 == N ? 0 : i + 1;
      SETSU(I.s, nx);
      if (i == 0) {
        for (uint32_t k = 0; k < N; k++) { ring_at(L, I.aux + 2, k) = ring_at(L, I.aux, k); ring_at(L, I.aux + 3, k) = ring_at(L, I.aux + 1, k); }
        { const LaneT Lc = L; lane_fft(Lc, I.aux + 2, I.aux + 3, I.n, L.tables + I.aux2, true); }
      }
      ring_at(L, I.aux, i) = XI(0); ring_at(L, I.aux + 1, i) = XI(1);
      XO(0) = ring_at(L, I.aux + 2, i); XO(1) = ring_at(L, I.aux + 3, i);
    } break;
    default: break;
  }
}

// spec.cpp compiles this file up to here (QG_SPEC_ONLY) together with a kernel specialised for one tape
#if !defined(QG_SPEC_ONLY)
// ------------------------------------------------------------------------------------------------ kernels
template <bool DIVERGENT>
__global__ void __launch_bounds__(128) k_interp(InterpArgs a) {
  const int nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  Instr* code = reinterpret_cast<Instr*>(qg_smem);
  float* xs = reinterpret_cast<float*>(qg_smem + (size_t)a.n_instr * sizeof(Instr));
  const int nx = a.P + a.NS + a.NT;
  float* tiles = xs + (size_t)nx * nt;                       // [n_out][nwarps][32][33] when a.tile
  {   // stage the tape
    const uint4* src = reinterpret_cast<const uint4*>(a.code);
    uint4* dst = reinterpret_cast<uint4*>(code);
    for (int i = tid; i < a.n_instr * 2; i += nt) dst[i] = src[i];
  }
  const int v = blockIdx.x * nt + tid;                       // padded voice index, always < Vp
  Lane L;
  L.x = xs + tid; L.nt = nt; L.v = v; L.Vp = a.Vp; L.rings = a.rings; L.ring_tab = a.ring_tab; L.tables = a.tables;
  L.state_init = a.state_init; L.state_keep = a.state_keep; L.resets = a.resets; L.P = a.P;
  for (int p = 0; p < a.P; p++) X(p) = a.params[(size_t)p * a.Vp + v];
  for (int s = 0; s < a.NS; s++) X(a.P + s) = a.state[(size_t)s * a.Vp + v];
  for (int k = 0; k < a.NT; k++) X(a.P + a.NS + k) = 0.0f;
  __syncthreads();

  const int warp_v0 = blockIdx.x * nt + warp * 32;
  const int in_base = a.P + a.NS;
  for (long t = 0; t < a.T; t++) {
    for (int c = 0; c < a.n_in; c++) {
      size_t idx = a.in_frame_major ? ((size_t)t * a.V + v) * a.n_in + c : ((size_t)v * a.n_in + c) * a.T + t;
      X(in_base + c) = v < a.V ? a.in[idx] : 0.0f;
    }
    if (!DIVERGENT) {
      int pc = 0;
      for (int i = 0; i < a.n_instr; i++) { const Instr I = code[i]; exec(I, L, pc); }
    } else {
      // SIMT-stack emulation: always run the lowest pending instruction; lanes that jumped ahead wait there
      int pc = 0;
      for (;;) {
        int m = __reduce_min_sync(0xffffffffu, pc);
        if (m >= a.n_instr) break;
        if (pc == m) { pc = m + 1; const Instr I = code[m]; exec(I, L, pc); }
      }
    }
    // ---- outputs
    if (a.out_frame_major) {
      if (v < a.V)
        for (int c = 0; c < a.n_out; c++) a.out[((size_t)t * a.V + v) * a.n_out + c] = X(a.out_x[c]);
    } else {
      const int tt = (int)(t & 31);
      for (int c = 0; c < a.n_out; c++) tiles[(((size_t)c * nwarps + warp) * 32 + lane) * 33 + tt] = X(a.out_x[c]);
      if (tt == 31 || t == a.T - 1) {
        __syncwarp();
        const long t_base = t - tt;
        const int ncols = tt + 1;
        for (int c = 0; c < a.n_out; c++) {
          const float* tile = tiles + ((size_t)c * nwarps + warp) * 32 * 33;
          if (a.group <= 1) {
            for (int r = 0; r < 32; r++) {
              int vv = warp_v0 + r;
              if (vv < a.V && lane < ncols) a.out[((size_t)vv * a.n_out + c) * a.T + t_base + lane] = tile[r * 33 + lane];
            }
          } else {
            // K6 group mix: voices of a group are summed left to right, then scaled by 1/G (`(v0+v1+..) >> mul(1/G)`)
            const int G = a.group;
            const float inv = 1.0f / (float)G;
            for (int g0 = 0; g0 < 32; g0 += G) {
              int gi = (warp_v0 + g0) / G;
              if (warp_v0 + g0 + G <= a.V && lane < ncols) {
                float acc = tile[g0 * 33 + lane];
                for (int r = 1; r < G; r++) acc += tile[(g0 + r) * 33 + lane];
                a.out[((size_t)gi * a.n_out + c) * a.T + t_base + lane] = acc * inv;
              }
            }
          }
        }
        __syncwarp();
      }
    }
  }
  for (int s = 0; s < a.NS; s++) a.state[(size_t)s * a.Vp + v] = X(a.P + s);
}

// K1b — block-mode lane interpreter for feed-forward (uniform) tapes: one lane = one voice, but every instruction is decoded
// ONCE per block of BT samples and applied to the whole block (FunDSP's `process()` does the same with 64-sample blocks),
// so the decode / dispatch cost that dominates k_interp is amortised and HBM delay lines are read BT lines at a time
// (BT independent loads in flight per warp instead of one dependent load per sample).
template <int BT>
__global__ void __launch_bounds__(128) k_interp_blk(InterpArgs a) {
  const int nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  Instr* code = reinterpret_cast<Instr*>(qg_smem);
  float* xs = reinterpret_cast<float*>(qg_smem + (size_t)a.n_instr * sizeof(Instr));
  const int PS = a.P + a.NS;
  const int nx = PS + a.NT * BT;
  // output staging: a [32 voices][BT + 1] tile per warp and output — one block of samples wide, so that the kernel's shared
  // memory is dominated by X and more warps fit (the lane kernels are latency-bound); rows leave as BT*4-byte segments
  constexpr int TW = BT + 1;
  float* tiles = xs + (size_t)nx * nt;                       // [n_out][nwarps][32][TW]
  for (int i = tid; i < a.n_instr; i += nt) {   // stage the tape with its temporary indices rewritten to block layout
    Instr I = a.code[i];
    I.out = blk_index<BT>(I.out, PS);
    for (int k = 0; k < 5; k++) I.in[k] = blk_index<BT>(I.in[k], PS);
    code[i] = I;
  }
  const int v = blockIdx.x * nt + tid;
  BlockLane<BT> L;
  L.x = xs + tid; L.nt = nt; L.v = v; L.Vp = a.Vp; L.rings = a.rings; L.ring_tab = a.ring_tab; L.tables = a.tables;
  L.state_init = a.state_init; L.state_keep = a.state_keep; L.resets = a.resets; L.P = a.P; L.PS = PS; L.j = 0; L.n = BT;
  for (int p = 0; p < a.P; p++) L.x[p * nt] = a.params[(size_t)p * a.Vp + v];
  for (int s = 0; s < a.NS; s++) L.x[(a.P + s) * nt] = a.state[(size_t)s * a.Vp + v];
  for (int k = 0; k < a.NT * BT; k++) L.x[(PS + k) * nt] = 0.0f;
  __syncthreads();

  const int warp_v0 = blockIdx.x * nt + warp * 32;
  for (long t0 = 0; t0 < a.T; t0 += BT) {
    const int n = (a.T - t0) < BT ? (int)(a.T - t0) : BT;
    L.n = n;
    for (int c = 0; c < a.n_in; c++)
      for (int j = 0; j < n; j++) {
        const long t = t0 + j;
        size_t idx = a.in_frame_major ? ((size_t)t * a.V + v) * a.n_in + c : ((size_t)v * a.n_in + c) * a.T + t;
        L.tr(PS + c * BT, j) = v < a.V ? a.in[idx] : 0.0f;
      }
    for (int i = 0; i < a.n_instr; i++) {
      const Instr I = code[i];
      if (I.op == OP_DELAY) {
        // whole-block delay line access: all reads first (independent loads), then the writes
        const uint2 rg2 = __ldg(reinterpret_cast<const uint2*>(L.ring_tab + I.aux));   // {offset, length} in one load
        Ring rg; rg.offset = rg2.x; rg.length = rg2.y;
        const uint32_t len = rg.length, idx = __float_as_uint(L.x[I.s * nt]);
        if (len >= (uint32_t)BT) {
          float* const rb = L.rings + (size_t)rg.offset * (size_t)L.Vp + (size_t)v;   // this voice's column of the ring
          const uint32_t Vp = (uint32_t)L.Vp;
          float o[BT];
#pragma unroll
          for (int j = 0; j < BT; j++) if (j < n) o[j] = rb[(size_t)ring_wrap(idx + (uint32_t)j, len) * Vp];
          const float* src = L.x + (int)I.in[0] * nt;
          const int sstep = (int)I.in[0] >= PS ? nt : 0;
#pragma unroll
          for (int j = 0; j < BT; j++) if (j < n) rb[(size_t)ring_wrap(idx + (uint32_t)j, len) * Vp] = src[j * sstep];
          float* dst = L.x + (int)I.out * nt;
#pragma unroll
          for (int j = 0; j < BT; j++) if (j < n) dst[j * nt] = o[j];
          L.x[I.s * nt] = __uint_as_float(ring_wrap(idx + (uint32_t)n, len));
          continue;
        }
      }
      int pc = 0;
      exec(I, L, pc);
    }
    // ---- outputs
    if (a.out_frame_major) {
      if (v < a.V)
        for (int j = 0; j < n; j++)
          for (int c = 0; c < a.n_out; c++) a.out[((size_t)(t0 + j) * a.V + v) * a.n_out + c] = L.tr(blk_index<BT>(a.out_x[c], PS), j);
    } else {
      for (int c = 0; c < a.n_out; c++) {
        const int ox = blk_index<BT>(a.out_x[c], PS);
        float* trow = tiles + (((size_t)c * nwarps + warp) * 32 + lane) * TW;
        const float* src = L.x + ox * nt;
        const int sstep = ox >= PS ? nt : 0;
        for (int j = 0; j < n; j++) trow[j] = src[j * sstep];
      }
      __syncwarp();
      for (int c = 0; c < a.n_out; c++) {
        const float* tile = tiles + ((size_t)c * nwarps + warp) * 32 * TW;
        if (a.group <= 1) {
          // lane -> (voice row, sample): 32 / BT rows per pass, each row a contiguous run of BT samples
          constexpr int RP = 32 / BT;
          const int jj = lane % BT, rr = lane / BT;
          for (int r0 = 0; r0 < 32; r0 += RP) {
            const int r = r0 + rr, vv = warp_v0 + r;
            if (vv < a.V && jj < n) a.out[((size_t)vv * a.n_out + c) * a.T + t0 + jj] = tile[r * TW + jj];
          }
        } else {
          // K6 group mix: voices of a group are summed left to right, then scaled by 1/G; lane -> (group, sample)
          const int G = a.group, per_pass = 32 / BT;         // groups handled per pass
          const float inv = 1.0f / (float)G;
          const int jj = lane % BT;
          for (int gp = 0; gp * G < 32; gp += per_pass) {
            const int g0 = (gp + lane / BT) * G;
            if (g0 < 32 && warp_v0 + g0 + G <= a.V && jj < n) {
              float acc = tile[g0 * TW + jj];
              for (int r = 1; r < G; r++) acc += tile[(g0 + r) * TW + jj];
              a.out[((size_t)((warp_v0 + g0) / G) * a.n_out + c) * a.T + t0 + jj] = acc * inv;
            }
          }
        }
      }
      __syncwarp();
    }
  }
  for (int s = 0; s < a.NS; s++) a.state[(size_t)s * a.Vp + v] = L.x[(a.P + s) * nt];
}

// state_init[s][v] = default word, then hash-seeded words (phase / noise seed / envelope hash), salted per voice
__global__ void k_init_state(float* state_init, const uint32_t* defaults, int NS, int Vp, const HashInit* hi, int n_hi,
                             const uint64_t* salts) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= Vp) return;
  for (int s = 0; s < NS; s++) state_init[(size_t)s * Vp + v] = __uint_as_float(defaults[s]);
  uint64_t salt = salts ? salts[v] : 0ull;
  for (int k = 0; k < n_hi; k++) {
    uint64_t h = hi[k].hash;
    if (salt) h = d_atto(h, salt);
    uint32_t w;
    switch (hi[k].kind) {
      case INIT_SINE_PHASE: w = __float_as_uint(d_rnd1(h)); break;
      case INIT_NOISE_SEED: w = (uint32_t)h; break;
      case INIT_HASH_LO: w = (uint32_t)h; break;
      default: w = (uint32_t)(h >> 32); break;
    }
    state_init[(size_t)hi[k].state * Vp + v] = __uint_as_float(w);
  }
}

__global__ void k_reset_state(float* state, const float* __restrict__ state_init, const uint8_t* __restrict__ keep, int NS, int Vp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)NS * Vp && !keep[i / Vp]) state[i] = state_init[i];
}

__global__ void k_broadcast_params(float* params, const float* tmpl, int P, int Vp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)P * Vp) params[i] = tmpl[i / Vp];
}

// Stream path (src/audio.rs:85-118): frames pulled by the device callback are sanitised (non-normal -> 0), clamped
// to [-1, 1] and interleaved L R.  src = voice-major rows [n_ch][n] of ONE graph; mono graphs get a silent right channel
// (process.rs:1897 `net | dc(0.)`), any other arity plays silence (process.rs:1901).
__global__ void k_stereo_frames(const float* __restrict__ src, int n_ch, long n, float* __restrict__ frames) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float l = 0.0f, r = 0.0f;
  if (n_ch == 1) l = src[t];
  else if (n_ch == 2) { l = src[t]; r = src[n + t]; }
  l = d_is_normal(l) ? d_clamp(l, -1.0f, 1.0f) : 0.0f;
  r = d_is_normal(r) ? d_clamp(r, -1.0f, 1.0f) : 0.0f;
  reinterpret_cast<float2*>(frames)[t] = make_float2(l, r);
}
// the same frames in the device sample types cpal offers besides f32 (audio.rs:115-116 `T::from_sample`, dasp_sample's
// conversions [external, restated]): i16 = (s * 32768) as i16 (Rust's float -> int cast truncates and saturates),
// u16 = that value in offset binary
__global__ void k_stereo_frames_i16(const float* __restrict__ src, int n_ch, long n, int offset_binary, uint32_t* __restrict__ frames) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float l = 0.0f, r = 0.0f;
  if (n_ch == 1) l = src[t];
  else if (n_ch == 2) { l = src[t]; r = src[n + t]; }
  l = d_is_normal(l) ? d_clamp(l, -1.0f, 1.0f) : 0.0f;
  r = d_is_normal(r) ? d_clamp(r, -1.0f, 1.0f) : 0.0f;
  int li = max(-32768, min(32767, d_as_i32(l * 32768.0f))), ri = max(-32768, min(32767, d_as_i32(r * 32768.0f)));
  if (offset_binary) { li += 32768; ri += 32768; }
  frames[t] = ((uint32_t)li & 0xffffu) | ((uint32_t)ri << 16);          // little-endian L then R
}

// full mix of the rows of a [R][T] buffer into one [T] row: rows are added in index order (per output sample)
__global__ void k_mix_rows(const float* rows, int R, long T, float scale, float* out) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float acc = 0.0f;
  for (int r = 0; r < R; r++) acc += rows[(size_t)r * T + t];
  out[t] = acc * scale;
}

// ------------------------------------------------------------------------------------------------ K3/K4 + time-vector mode
// k_interp_tv: ONE CTA PER VOICE, threads = the samples of a hop (H <= 512).  For banks with few voices and long
// feed-forward graphs (the spectral patches: BASELINE configs[3]) lane-per-voice leaves the GPU empty; here every op
// of the tape is applied to a whole hop at once.  Stateless ops reuse exec(); counter-like sources jump; delay lines
// are block-copied; rfft/ifft (nodes.rs:601-700) become cooperative shared-memory radix-2 transforms (K3/K4) that run
// exactly when the node's counter wraps — hops are aligned to the frame grid by construction (H | N, H | start).

// K3/K4 — cooperative FFT of N = 1 << lg points held in shared memory (re, im), input in bit-reversed order.
// The arithmetic is the radix-2 decimation-in-time butterfly network of the per-lane path (same twiddle table, same
// operation order per butterfly, inverse scaled by 1/N), but a thread carries 2^R points through R consecutive stages in
// registers (R = 3: radix-8 passes), so an N = 2048 transform takes 4 passes / barriers instead of 11 and a quarter of the
// shared-memory traffic.  Results are bit-identical to the radix-2 schedule.  Buffers are padded by one float per 32 to
// spread the strided accesses of the first passes over the banks.
#define FPAD(i) ((i) + ((i) >> 5))
template <int R>
__device__ __forceinline__ void tv_fft_pass(float* fr, float* fi, int lg, int s, const float* __restrict__ tw, bool inverse, int tid,
                                            int nth) {
  const uint32_t N = 1u << lg, h = 1u << s;
  for (uint32_t g = tid; g < (N >> R); g += nth) {
    const uint32_t k = g & (h - 1), base = ((g >> s) << (s + R)) | k;
    float xr[1 << R], xi[1 << R];
#pragma unroll
    for (int m = 0; m < (1 << R); m++) { const uint32_t idx = base + (uint32_t)m * h; xr[m] = fr[FPAD(idx)]; xi[m] = fi[FPAD(idx)]; }
#pragma unroll
    for (int q = 0; q < R; q++) {            // stage s + q: partners differ in bit q of m
      const int hq = 1 << q;
#pragma unroll
      for (int m = 0; m < (1 << R); m++) {
        if (m & hq) continue;
        const uint32_t kq = k + (uint32_t)(m & (hq - 1)) * h;      // index within the butterfly group of this stage
        const uint32_t ti = kq << (lg - 1 - (s + q));              // k * (N / len)
        const float wr = __ldg(tw + 2 * ti);
        float wi = __ldg(tw + 2 * ti + 1);
        if (inverse) wi = -wi;
        const float ur = xr[m], ui = xi[m], vr = xr[m + hq], vi = xi[m + hq];
        const float tr = vr * wr - vi * wi, tim = vr * wi + vi * wr;
        xr[m] = ur + tr; xi[m] = ui + tim;
        xr[m + hq] = ur - tr; xi[m + hq] = ui - tim;
      }
    }
#pragma unroll
    for (int m = 0; m < (1 << R); m++) { const uint32_t idx = base + (uint32_t)m * h; fr[FPAD(idx)] = xr[m]; fi[FPAD(idx)] = xi[m]; }
  }
}
__device__ void tv_fft(float* fr, float* fi, int lg, const float* tw, bool inverse, int tid, int nth) {
  const uint32_t N = 1u << lg;
  for (int s = 0; s < lg;) {
    const int r = lg - s >= 3 ? 3 : lg - s;
    if (r == 3) tv_fft_pass<3>(fr, fi, lg, s, tw, inverse, tid, nth);
    else if (r == 2) tv_fft_pass<2>(fr, fi, lg, s, tw, inverse, tid, nth);
    else tv_fft_pass<1>(fr, fi, lg, s, tw, inverse, tid, nth);
    s += r;
    __syncthreads();
  }
  if (inverse) {
    const float sc = 1.0f / (float)N;
    for (uint32_t k = tid; k < N; k += nth) { fr[FPAD(k)] *= sc; fi[FPAD(k)] *= sc; }
    __syncthreads();
  }
}

// one out-of-line copy of the generic per-sample code for the time-vector kernel's one-thread paths (keeps the hot loop's
// instruction footprint small)
__device__ __noinline__ void exec_one_thread(const Instr& I, TvSample& L) {
  int dummy = 0;
  exec(I, L, dummy);
}

constexpr int TV_SEGCAP = 72;   // lfo segments a 512-sample hop can cross when a segment is at least 8 samples long
// (idx + j) mod len for idx < len: one conditional subtract in the common case (ring at least one hop long)
__device__ __forceinline__ uint32_t tv_wrap(uint32_t p, uint32_t len) {
  if (p >= len) { p -= len; if (p >= len) p %= len; }
  return p;
}

__global__ void __launch_bounds__(256, 4) k_interp_tv(TvArgs a) {
  const int tid = threadIdx.x, nth = blockDim.x, v = blockIdx.x;
  Instr* code = reinterpret_cast<Instr*>(qg_smem);
  // region offsets (floats) come precomputed as kernel parameters (tv_layout)
#define H a.H
#define PS a.PS
#define ps_off a.ps_off
#define tmp_off a.tmp_off
#define oldv_off a.oldv_off
#define fr_off a.fr_off
#define fi_off a.fi_off
#define lti_off a.lti_off
#define scan_off a.scan_off
#define segi_off a.segi_off
#define segt_off a.segt_off
#define ps (QG_SMEM_F + ps_off)
#define tmp (QG_SMEM_F + tmp_off)
#define oldv (QG_SMEM_F + oldv_off)
#define fr (QG_SMEM_F + fr_off)
#define fi (QG_SMEM_F + fi_off)
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.code);
    uint4* dst = reinterpret_cast<uint4*>(code);
    for (int i = tid; i < a.n_instr * 2; i += nth) dst[i] = src[i];
  }
  for (int p = tid; p < a.P; p += nth) ps[p] = a.params[(size_t)p * a.Vp + v];
  for (int s = tid; s < a.NS; s += nth) ps[a.P + s] = a.state[(size_t)s * a.Vp + v];
  for (int k = tid; k < a.NT * H; k += nth) tmp[k] = 0.0f;
  __syncthreads();
  // ---- scan matrices of the fixed-coefficient LTI filters (one thread per filter, f64, rounded once):
  // a thread advances CPT consecutive samples, so the per-thread map is s -> M s + e with M = A^CPT
  const int CPT = (H + nth - 1) / nth;                    // samples per thread in the LTI scans (<= 4: H <= 512, >= 128 threads)
  for (int pc = tid; pc < a.n_instr; pc += nth) {
    const Instr I = code[pc];
    if (!op_is_lti(I.op)) continue;
    double A11, A12, A21, A22;
    if (I.op == OP_SVF) {
      const double a1 = ps[I.p], a2 = ps[I.p + 1], a3 = ps[I.p + 2];
      A11 = 2 * a1 - 1; A12 = -2 * a2; A21 = 2 * a2; A22 = 1 - 2 * a3;
    } else if (I.op == OP_BIQUAD) {
      A11 = -(double)ps[I.p]; A12 = -(double)ps[I.p + 1]; A21 = 1; A22 = 0;
    } else {
      const double c = ps[I.p];
      A11 = I.n == 3 ? -c : c; A12 = 0; A21 = 0; A22 = 0;
    }
    double m11 = 1, m12 = 0, m21 = 0, m22 = 1;
    for (int k = 0; k < CPT; k++) {
      const double n11 = A11 * m11 + A12 * m21, n12 = A11 * m12 + A12 * m22, n21 = A21 * m11 + A22 * m21, n22 = A21 * m12 + A22 * m22;
      m11 = n11; m12 = n12; m21 = n21; m22 = n22;
    }
    float4* tab = reinterpret_cast<float4*>(QG_SMEM_F + lti_off + (int)I.aux * TV_LTI_FLOATS);
    double p11 = 1, p12 = 0, p21 = 0, p22 = 1;             // M^l, l = 0..32
    for (int l = 0; l <= 32; l++) {
      if (l < 32) tab[5 + l] = make_float4((float)p11, (float)p12, (float)p21, (float)p22);
      else tab[37] = make_float4((float)p11, (float)p12, (float)p21, (float)p22);
      if (l == 1) tab[0] = tab[5 + l];
      if (l == 2) tab[1] = tab[5 + l];
      if (l == 4) tab[2] = tab[5 + l];
      if (l == 8) tab[3] = tab[5 + l];
      if (l == 16) tab[4] = tab[5 + l];
      const double n11 = m11 * p11 + m12 * p21, n12 = m11 * p12 + m12 * p22, n21 = m21 * p11 + m22 * p21, n22 = m21 * p12 + m22 * p22;
      p11 = n11; p12 = n12; p21 = n21; p22 = n22;
    }
  }
  float* rg = a.rings + (size_t)v * a.ring_floats;        // voice-major rings in this mode
#define RING(r) (rg + a.ring_tab[r].offset)
#define TMP(i) (tmp + ((int)(i) - PS) * H)                 // temporaries only
#define SRC(i, j) ((int)(i) < PS ? ps[(int)(i)] : tmp[((int)(i) - PS) * H + (j)])
  __syncthreads();
  // a previous call may have stopped in the middle of a hop: the first pass only completes that hop, so that
  // transform frames stay aligned with hop boundaries (all rfft/ifft counters are congruent modulo H)
  int first = 0;
  if (a.align_s >= 0) {
    const uint32_t ph = __float_as_uint(ps[a.align_s]) % (uint32_t)H;
    first = ph ? H - (int)ph : 0;
  }
  for (long t0 = 0; t0 < a.T;) {
    int n = (int)(a.T - t0 < H ? a.T - t0 : H);
    if (t0 == 0 && first > 0 && first < n) n = first;
    for (int c = 0; c < a.n_in; c++)
      for (int j = tid; j < n; j += nth)
        tmp[(size_t)c * H + j] = a.frame_major ? a.in[((size_t)(t0 + j) * a.V + v) * a.n_in + c] : a.in[((size_t)v * a.n_in + c) * a.T + t0 + j];
    __syncthreads();
    for (int pc = 0; pc < a.n_instr; pc++) {
      const Instr I = code[pc];
      if (I.pad != 0) {   // stateless op: lower() stored the end of its run in `pad` (0 for stateful ops)
        // a run of stateless ops is applied sample by sample without intermediate barriers
        const int pe = (int)I.pad;   // end of the run, precomputed by lower()
        // op-outer / sample-inner: every instruction is decoded once per thread and applied to all of the thread's samples;
        // a thread only ever touches its own sample columns, so the run needs no barrier between ops
        TvSample L{ps_off, tmp_off, PS, H, a.tables, tid, nth, n, tid};
        int dummy = 0;
        for (int q = pc; q < pe; q++) { const Instr Iq = code[q]; exec(Iq, L, dummy); }
        pc = pe - 1;
      } else {
        switch (I.op == OP_BIQUAD && !a.biquad_scan ? (uint16_t)OP_COUNT_ : I.op) {   // OP_COUNT_ -> the generic one-thread path
          case OP_NOISE: {
            const uint32_t c = __float_as_uint(ps[I.s]);
            __syncthreads();
            for (int j = tid; j < n; j += nth) TMP(I.out)[j] = d_noise(c + (uint32_t)j + 1u);
            if (tid == 0) ps[I.s] = __uint_as_float(c + (uint32_t)n);
            break;
          }
          case OP_WAVE: {
            const uint32_t idx = __float_as_uint(ps[I.s]);
            __syncthreads();
            for (int j = tid; j < n; j += nth) TMP(I.out)[j] = a.tables[I.aux + tv_wrap(idx + (uint32_t)j, I.aux2)];
            if (tid == 0) ps[I.s] = __uint_as_float(tv_wrap(idx + (uint32_t)n, I.aux2));
            break;
          }
          case OP_IMPULSE: {
            const uint32_t f = __float_as_uint(ps[I.s]);
            __syncthreads();
            for (int j = tid; j < n; j += nth) TMP(I.out)[j] = (f == 0u && j == 0) ? 1.0f : 0.0f;
            if (tid == 0) ps[I.s] = __uint_as_float(1u);
            break;
          }
          case OP_SVF: case OP_BIQUAD: case OP_ONEPOLE: {
            // Block-level parallel linear-recurrence scan over the hop.  Thread t owns samples [t*CPT, t*CPT + cnt):
            //   1. zero-state run of its samples (real inputs, zero filter state) -> e_t;
            //   2. scan of the affine maps s -> M s + e_t: Kogge-Stone inside a warp (M^(2^i)), warp totals chained by
            //      one thread (M^32), start state of lane l = M^l P_warp + E_(l-1);
            //   3. the samples are re-run from the true start state with the op's own arithmetic (same code as exec()).
            const float4* tab = reinterpret_cast<const float4*>(QG_SMEM_F + lti_off + (int)I.aux * TV_LTI_FLOATS);
            float* wtot = QG_SMEM_F + scan_off;             // [nwarps][2] warp totals, then [nwarps][2] warp start states
            float* wsta = wtot + 2 * (nth >> 5);
            const int j0 = tid * CPT;
            const int cnt = j0 >= n ? 0 : (n - j0 < CPT ? n - j0 : CPT);
            const int lane = tid & 31, wid = tid >> 5;
            // inputs and the two previous inputs (history comes from the hop or from the persisted state)
            float xin[4] = {0.0f, 0.0f, 0.0f, 0.0f}, xm1 = 0.0f, xm2 = 0.0f;
            for (int k = 0; k < cnt; k++) xin[k] = SRC(I.in[0], j0 + k);
            if (I.op == OP_BIQUAD) {
              xm1 = j0 >= 1 ? SRC(I.in[0], j0 - 1) : ps[I.s];
              xm2 = j0 >= 2 ? SRC(I.in[0], j0 - 2) : (j0 == 1 ? ps[I.s] : ps[I.s + 1]);
            } else if (I.op == OP_ONEPOLE) {
              xm1 = j0 >= 1 ? SRC(I.in[0], j0 - 1) : ps[I.s];
            }
            float S1, S2;                                   // hop start state
            if (I.op == OP_SVF) { S1 = ps[I.s]; S2 = ps[I.s + 1]; }
            else if (I.op == OP_BIQUAD) { S1 = ps[I.s + 2]; S2 = ps[I.s + 3]; }
            else { S1 = ps[I.s + 1]; S2 = 0.0f; }
            float c0 = ps[I.p], c1 = 0.0f, c2 = 0.0f, c3 = 0.0f, c4 = 0.0f, c5 = 0.0f;
            if (I.op != OP_ONEPOLE) { c1 = ps[I.p + 1]; c2 = ps[I.p + 2]; c3 = ps[I.p + 3]; c4 = ps[I.p + 4]; }
            if (I.op == OP_SVF) c5 = ps[I.p + 5];
            const int kind = I.n;
            // one tick with the op's own arithmetic; (s1, s2) is the scanned state, (h1, h2) the input history
            auto tick = [&](float x, float& s1, float& s2, float& h1, float& h2) -> float {
              if (I.op == OP_SVF) return d_svf_tick(x, s1, s2, c0, c1, c2, c3, c4, c5);
              if (I.op == OP_BIQUAD) {
                const float y0 = c2 * x + c3 * h1 + c4 * h2 - c0 * s1 - c1 * s2;
                h2 = h1; h1 = x; s2 = s1; s1 = y0;
                return y0;
              }
              float y;
              switch (kind) {
                case 0: y = (1.0f - c0) * x + c0 * s1; break;
                case 1: y = c0 * (s1 + x - h1); break;
                case 2: y = x - h1 + c0 * s1; break;
                default: y = c0 * (x - s1) + h1; break;
              }
              h1 = x; s1 = y;
              return y;
            };
            float e1 = 0.0f, e2 = 0.0f;
            { float h1 = xm1, h2 = xm2; for (int k = 0; k < cnt; k++) tick(xin[k], e1, e2, h1, h2); }
#pragma unroll
            for (int i = 0; i < 5; i++) {
              const int d = 1 << i;
              const float4 m = tab[i];
              float r1 = __shfl_up_sync(0xffffffffu, e1, d), r2 = __shfl_up_sync(0xffffffffu, e2, d);
              if (lane >= d) { e1 += m.x * r1 + m.y * r2; e2 += m.z * r1 + m.w * r2; }
            }
            float p1 = __shfl_up_sync(0xffffffffu, e1, 1), p2 = __shfl_up_sync(0xffffffffu, e2, 1);
            if (lane == 0) { p1 = 0.0f; p2 = 0.0f; }
            if (lane == 31) { wtot[2 * wid] = e1; wtot[2 * wid + 1] = e2; }
            __syncthreads();                                 // also: every thread has read its inputs and the old state
            if (tid == 0) {
              const float4 m32 = tab[37];
              float q1 = S1, q2 = S2;
              for (int w = 0; w < (nth >> 5); w++) {
                wsta[2 * w] = q1; wsta[2 * w + 1] = q2;
                const float t1 = m32.x * q1 + m32.y * q2 + wtot[2 * w], t2 = m32.z * q1 + m32.w * q2 + wtot[2 * w + 1];
                q1 = t1; q2 = t2;
              }
            }
            __syncthreads();
            {
              const float4 ml = tab[5 + lane];
              const float q1 = wsta[2 * wid], q2 = wsta[2 * wid + 1];
              float s1 = ml.x * q1 + ml.y * q2 + p1, s2 = ml.z * q1 + ml.w * q2 + p2;
              float h1 = xm1, h2 = xm2;
              for (int k = 0; k < cnt; k++) TMP(I.out)[j0 + k] = tick(xin[k], s1, s2, h1, h2);
              if (cnt > 0 && j0 + cnt == n) {              // owner of the hop's last sample persists the state
                if (I.op == OP_SVF) { ps[I.s] = s1; ps[I.s + 1] = s2; }
                else if (I.op == OP_BIQUAD) { ps[I.s] = h1; ps[I.s + 1] = h2; ps[I.s + 2] = s1; ps[I.s + 3] = s2; }
                else { ps[I.s] = h1; ps[I.s + 1] = s1; }
              }
            }
            break;
          }
          case OP_SINE: {   // exec(): out = sin(phase * TAU) from the phase BEFORE the increment; phase += f * (1/sr), wrapped
            // A run of consecutive, mutually independent oscillators (chords, additive banks: `sine(f0) + sine(f1) + ...`)
            // is stepped concurrently: one warp leader per oscillator runs that oscillator's exact f32 phase recurrence.
            int pe = pc + 1;
            while (pe < a.n_instr && code[pe].op == OP_SINE) {
              bool dep = false;
              for (int q = pc; q < pe; q++) dep = dep || code[pe].in[0] == code[q].out;
              if (dep) break;
              pe++;
            }
            __syncthreads();
            for (int q = pc; q < pe; q++) {                  // increments, in parallel
              const Instr Q = code[q];
              const float isr = ps[Q.p];
              for (int j = tid; j < n; j += nth) TMP(Q.out)[j] = SRC(Q.in[0], j) * isr;
            }
            __syncthreads();
            if ((tid & 31) == 0) {
              for (int q = pc + (tid >> 5); q < pe; q += nth >> 5) {
                const Instr Q = code[q];
                float ph = ps[Q.s];
                float* o = TMP(Q.out);
                const float inc0 = o[0];
                if ((int)Q.in[0] < PS && inc0 >= 0.0f && inc0 < 1.0f && ph >= 0.0f && ph < 1.0f) {
                  // constant frequency (`sine(440)`: what quartz renders most): the increment is not reloaded, and for a phase
                  // in [0, 2) `ph -= floor(ph)` IS a conditional `- 1` (exact either way) — the dependent chain per sample is
                  // FADD, FSETP, predicated FADD instead of LDS, FADD, FRND, FADD
#pragma unroll 4
                  for (int j = 0; j < n; j++) { o[j] = ph; ph += inc0; if (ph >= 1.0f) ph -= 1.0f; }
                } else {
                  for (int j = 0; j < n; j++) { const float inc = o[j]; o[j] = ph; ph += inc; ph -= floorf(ph); }
                }
                ps[Q.s] = ph;
              }
            }
            __syncthreads();
            for (int q = pc; q < pe; q++) {
              float* o = TMP(code[q].out);
              for (int j = tid; j < n; j += nth) o[j] = sinf(o[j] * QG_TAU);
            }
            pc = pe - 1;
            break;
          }
          case OP_RAMP: {   // nodes.rs:476-483: out = val; val += f / sr; if (val >= 1) val -= 1
            const float sr = ps[I.p];
            __syncthreads();
            for (int j = tid; j < n; j += nth) TMP(I.out)[j] = SRC(I.in[0], j) / sr;
            __syncthreads();
            if (tid == 0) {
              float val = ps[I.s];
              float* o = TMP(I.out);
              if ((int)I.in[0] < PS) {                      // constant frequency (`dc(f) >> ramp()`): the increment stays in a register
                const float inc0 = o[0];
#pragma unroll 4
                for (int j = 0; j < n; j++) { o[j] = val; val += inc0; if (val >= 1.0f) val -= 1.0f; }
              } else {
                for (int j = 0; j < n; j++) { const float inc = o[j]; o[j] = val; val += inc; if (val >= 1.0f) val -= 1.0f; }
              }
              ps[I.s] = val;
            }
            break;
          }
          case OP_WAVETABLE: {   // phase recurrence on one thread (phase AFTER the increment is what is read), lookups on all
            const float isr = ps[I.p];
            __syncthreads();
            if (tid == 0) {
              float ph = ps[I.s];
              for (int j = 0; j < n; j++) { ph += SRC(I.in[0], j) * isr; ph -= floorf(ph); oldv[j] = ph; }
              ps[I.s] = ph;
            }
            __syncthreads();
            const uint32_t hint0 = __float_as_uint(ps[I.s + 1]);
            __syncthreads();
            for (int j = tid; j < n; j += nth) {
              uint32_t hint = hint0;
              TMP(I.out)[j] = d_wavetable_read(a.tables + I.aux, SRC(I.in[0], j), oldv[j], hint);
              if (j == n - 1) ps[I.s + 1] = __uint_as_float(hint);
            }
            break;
          }
          case OP_ENVELOPE: {   // lfo()/lfo_in(): control points every ~2 ms, linear interpolation in between
            const float dt = ps[I.p + 4];
            if (!(0.0015f / dt >= 9.0f)) {                   // segments shorter than 8 samples: no room in the segment table
              if (tid == 0) { TvSample L{ps_off, tmp_off, PS, H, a.tables, 0, 1, n, 0}; exec_one_thread(I, L); }
              break;
            }
            float* segi = QG_SMEM_F + segi_off;
            float4* segt = reinterpret_cast<float4*>(QG_SMEM_F + segt_off);
            __syncthreads();
            if (tid == 0) {                                  // time accumulation + control-point updates, exactly like exec()
              const int shape = I.n & 0xff, nin = I.n >> 8;
              float t = ps[I.s], t0 = ps[I.s + 1], t1 = ps[I.s + 2], v0 = ps[I.s + 3], v1 = ps[I.s + 4];
              int ns = 0;
              segt[0] = make_float4(t0, t1, v0, v1);
              for (int j = 0; j < n; j++) {
                if (t >= t1) {
                  float c[4] = {ps[I.p], ps[I.p + 1], ps[I.p + 2], ps[I.p + 3]};
                  float in[4];
                  in[0] = nin > 0 ? SRC(I.in[0], j) : 0.0f; in[1] = nin > 1 ? SRC(I.in[1], j) : 0.0f;
                  in[2] = nin > 2 ? SRC(I.in[2], j) : 0.0f; in[3] = nin > 3 ? SRC(I.in[3], j) : 0.0f;
                  if (__float_as_uint(ps[I.s + 7])) { v1 = d_env_eval(shape, nin, 0.0f, c, in); ps[I.s + 7] = __uint_as_float(0u); }
                  uint64_t th = (uint64_t)__float_as_uint(ps[I.s + 5]) | ((uint64_t)__float_as_uint(ps[I.s + 6]) << 32);
                  t0 = t1; v0 = v1;
                  t1 = t0 + d_lerp(0.75f, 1.25f, d_rnd1(th)) * 0.002f;
                  v1 = d_env_eval(shape, nin, t1, c, in);
                  th += 1;
                  ps[I.s + 5] = __uint_as_float((uint32_t)th); ps[I.s + 6] = __uint_as_float((uint32_t)(th >> 32));
                  ns = ns + 1 < TV_SEGCAP ? ns + 1 : ns;
                  segt[ns] = make_float4(t0, t1, v0, v1);
                }
                oldv[j] = t;
                segi[j] = __int_as_float(ns);
                t += dt;
              }
              ps[I.s] = t; ps[I.s + 1] = t0; ps[I.s + 2] = t1; ps[I.s + 3] = v0; ps[I.s + 4] = v1;
            }
            __syncthreads();
            for (int j = tid; j < n; j += nth) {
              const float4 sg = segt[__float_as_int(segi[j])];
              TMP(I.out)[j] = d_lerp(sg.z, sg.w, d_delerp(sg.x, sg.y, oldv[j]));
            }
            break;
          }
          case OP_TICK: {
            const float prev = ps[I.s];
            __syncthreads();
            for (int j = tid; j < n; j += nth) TMP(I.out)[j] = j == 0 ? prev : SRC(I.in[0], j - 1);
            if (tid == 0) ps[I.s] = SRC(I.in[0], n - 1);
            break;
          }
          case OP_DELAY: {
            const uint32_t Lr = a.ring_tab[I.aux].length, idx = __float_as_uint(ps[I.s]);
            float* r = RING(I.aux);
            __syncthreads();
            for (int j = tid; j < n; j += nth)
              TMP(I.out)[j] = (uint32_t)j < Lr ? r[tv_wrap(idx + (uint32_t)j, Lr)] : SRC(I.in[0], j - (int)Lr);
            __syncthreads();
            const int jlo = n > (int)Lr ? n - (int)Lr : 0;
            for (int j = jlo + tid; j < n; j += nth) r[tv_wrap(idx + (uint32_t)j, Lr)] = SRC(I.in[0], j);
            if (tid == 0) ps[I.s] = __uint_as_float(tv_wrap(idx + (uint32_t)n, Lr));
            break;
          }
          case OP_TAP: {   // write-then-read per sample; `oldv` keeps what the hop overwrites (ring length >= H)
            const uint32_t Lr = a.ring_tab[I.aux].length, mask = Lr - 1, idx = __float_as_uint(ps[I.s]);
            float* r = RING(I.aux);
            __syncthreads();
            for (int j = tid; j < n; j += nth) oldv[j] = r[(idx + (uint32_t)j) & mask];
            __syncthreads();
            for (int j = tid; j < n; j += nth) r[(idx + (uint32_t)j) & mask] = SRC(I.in[0], j);
            __syncthreads();
            const float mn = ps[I.p], mx = ps[I.p + 1], sr = ps[I.p + 2];
            for (int j = tid; j < n; j += nth) {
              float tap = d_clamp(SRC(I.in[1], j), mn, mx) * sr;
              if (tap != tap) tap = 0.0f;
              const uint32_t fl = (uint32_t)d_as_usize(tap);
              const float d = tap - (float)fl;
              const uint32_t i1 = (idx + (uint32_t)j + Lr - (fl & mask)) & mask;
              auto rd = [&](uint32_t p) -> float {
                const uint32_t jj = (p - idx) & mask;
                return (jj < (uint32_t)n && jj > (uint32_t)j) ? oldv[jj] : r[p];
              };
              if (I.n) {
                const uint32_t i0 = (i1 + 1) & mask, i2 = (i1 + Lr - 1) & mask, i3 = (i1 + Lr - 2) & mask;
                TMP(I.out)[j] = d_spline(rd(i0), rd(i1), rd(i2), rd(i3), d);
              } else {
                const uint32_t i2 = (i1 + Lr - 1) & mask;
                TMP(I.out)[j] = d_lerp(rd(i1), rd(i2), d);
              }
            }
            if (tid == 0) ps[I.s] = __uint_as_float((idx + (uint32_t)n) & mask);
            break;
          }
          case OP_RFFT: {   // nodes.rs:625-642
            const int lg = I.n;
            const uint32_t N = 1u << lg, i0 = __float_as_uint(ps[I.s]);
            float *rin = RING(I.aux), *rre = RING(I.aux + 1), *rim = RING(I.aux + 2);
            __syncthreads();
            if (i0 == 0) {   // K3: the frame is complete -> transform it before this hop's samples are stored
              for (uint32_t k = tid; k < N; k += nth) { const uint32_t rv = __brev(k) >> (32 - lg); fr[FPAD(rv)] = rin[k]; fi[FPAD(rv)] = 0.0f; }
              __syncthreads();
              tv_fft(fr, fi, lg, a.tables + I.aux2, false, tid, nth);
              for (uint32_t k = tid; k < N; k += nth) { rre[k] = fr[FPAD(k)]; rim[k] = fi[FPAD(k)]; }
              __syncthreads();
            }
            for (int j = tid; j < n; j += nth) {
              const uint32_t i = i0 + (uint32_t)j;
              rin[i] = SRC(I.in[0], j);
              if (i <= N / 2) { TMP(I.out)[j] = rre[i]; TMP(I.out + 1)[j] = rim[i]; }
              else { TMP(I.out)[j] = rre[N - i]; TMP(I.out + 1)[j] = -rim[N - i]; }
            }
            if (tid == 0) ps[I.s] = __uint_as_float((i0 + (uint32_t)n) & (N - 1));
            break;
          }
          case OP_IFFT: {   // nodes.rs:681-693
            const int lg = I.n;
            const uint32_t N = 1u << lg, i0 = __float_as_uint(ps[I.s]);
            float *ire = RING(I.aux), *iim = RING(I.aux + 1), *ore = RING(I.aux + 2), *oim = RING(I.aux + 3);
            __syncthreads();
            if (i0 == 0) {   // K4: full complex inverse transform of the collected bins
              for (uint32_t k = tid; k < N; k += nth) { const uint32_t rv = __brev(k) >> (32 - lg); fr[FPAD(rv)] = ire[k]; fi[FPAD(rv)] = iim[k]; }
              __syncthreads();
              tv_fft(fr, fi, lg, a.tables + I.aux2, true, tid, nth);
              for (uint32_t k = tid; k < N; k += nth) { ore[k] = fr[FPAD(k)]; oim[k] = fi[FPAD(k)]; }
              __syncthreads();
            }
            for (int j = tid; j < n; j += nth) {
              const uint32_t i = i0 + (uint32_t)j;
              ire[i] = SRC(I.in[0], j); iim[i] = SRC(I.in[1], j);
              TMP(I.out)[j] = ore[i]; TMP(I.out + 1)[j] = oim[i];
            }
            if (tid == 0) ps[I.s] = __uint_as_float((i0 + (uint32_t)n) & (N - 1));
            break;
          }
          default: {
            // scalar-state op without a block form (variable filters, envelopes, sample-and-hold, wavetable phase ...):
            // one thread steps it through the hop with the generic per-sample code; plan_tv() admits only ring-free ops here
            __syncthreads();
            if (tid == 0) {
              TvSample L{ps_off, tmp_off, PS, H, a.tables, 0, 1, n, 0};
              exec_one_thread(I, L);
            }
            break;
          }
        }
      }
      __syncthreads();
    }
    for (int c = 0; c < a.n_out; c++) {
      const int ox = a.out_x[c];
      for (int j = tid; j < n; j += nth) {
        const size_t o = a.frame_major ? ((size_t)(t0 + j) * a.V + v) * a.n_out + c : ((size_t)v * a.n_out + c) * a.T + t0 + j;
        a.out[o] = SRC(ox, j);
      }
    }
    __syncthreads();
    t0 += n;
  }
  for (int s = tid; s < a.NS; s += nth) a.state[(size_t)s * a.Vp + v] = ps[a.P + s];
#undef RING
#undef TMP
#undef SRC
#undef ps
#undef tmp
#undef oldv
#undef fr
#undef fi
#undef H
#undef PS
#undef ps_off
#undef tmp_off
#undef oldv_off
#undef fr_off
#undef fi_off
#undef lti_off
#undef scan_off
#undef segi_off
#undef segt_off
}

// shared-memory layout of k_interp_tv, in floats: tape | scalars | temporaries | tap scratch | transform buffers | LTI scan
// matrices | scan scratch (2 x 2 x 8 warps) | envelope segment indices | envelope segment table
static size_t tv_layout(TvArgs& a) {
  a.PS = a.P + a.NS;
  a.ps_off = a.n_instr * (int)(sizeof(Instr) / 4);
  a.tmp_off = a.ps_off + ((a.PS + 3) & ~3);
  a.oldv_off = a.tmp_off + a.NT * a.H;
  a.fr_off = a.oldv_off + a.H;
  a.fi_off = a.fr_off + (int)FPAD(a.fft_n);
  a.lti_off = (a.fi_off + (int)FPAD(a.fft_n) + 3) & ~3;
  a.scan_off = a.lti_off + a.n_lti * TV_LTI_FLOATS;
  a.segi_off = a.scan_off + 32;
  a.segt_off = a.segi_off + a.H;
  return (size_t)(a.segt_off + TV_SEGCAP * 4) * sizeof(float);
}
size_t tv_smem_bytes(const TvArgs& a_in) {
  TvArgs a = a_in;
  return tv_layout(a);
}

cudaError_t launch_interp_tv(const TvArgs& a_in, cudaStream_t stream, int* launches) {
  TvArgs a = a_in;
  size_t smem = tv_layout(a);
  cudaError_t e = cudaFuncSetAttribute(k_interp_tv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // 256 threads: 2 samples per thread and hop.  (128 threads = 4 samples per thread halves the per-op decode work but was
  // measured 35 % slower on configs[3]: the kernel is latency-bound, it wants the warps.)
  k_interp_tv<<<a.V, 256, smem, stream>>>(a);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

// FP32 pipe roofline probe: 8 independent FFMA chains per thread, no memory traffic.  bench.py divides the flops by the
// measured launch time to get the FP32 (non-tensor) peak the compute-bound workloads are compared with.
__global__ void __launch_bounds__(256) k_fp32_peak(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
      x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
    }
  }
  float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456f) out[0] = s;   // keeps the chains alive without a store in the common case
}
cudaError_t launch_fp32_peak(float* out, int blocks, int iters, cudaStream_t stream) {
  k_fp32_peak<<<blocks, 256, 0, stream>>>(out, iters, 0.999f, 0.001f);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ launchers
static size_t interp_smem(const InterpArgs& a, int nt, bool tile) {
  size_t b = (size_t)a.n_instr * sizeof(Instr) + (size_t)(a.P + a.NS + a.NT) * nt * sizeof(float);
  if (tile) b += (size_t)a.n_out * (nt / 32) * 32 * 33 * sizeof(float);
  return b;
}

static size_t blk_smem(const InterpArgs& a, int nt, bool tile, int bt) {
  size_t b = (size_t)a.n_instr * sizeof(Instr) + (size_t)(a.P + a.NS + a.NT * bt) * nt * sizeof(float);
  if (tile) b += (size_t)a.n_out * (nt / 32) * 32 * (bt + 1) * sizeof(float);
  return b;
}
constexpr int BLK_BT = 8;
int interp_block_len() { return BLK_BT; }

cudaError_t launch_interp(const InterpArgs& a_in, bool divergent, bool block_ok, cudaStream_t stream, int* launches) {
  InterpArgs a = a_in;
  bool tile = !a.out_frame_major && a.n_out > 0;
  cudaError_t e;
  if (!divergent && block_ok && a.T >= 4 * BLK_BT) {
    // block mode: BT samples per decoded instruction; pick the block width that still leaves >= 8 warps per SM
    int nt = 128;
    const size_t per_sm = 220 * 1024;
    while (nt > 32 && (blk_smem(a, nt, tile, BLK_BT) > per_sm / 4 || (a.Vp / nt) < 148)) nt >>= 1;
    size_t smem = blk_smem(a, nt, tile, BLK_BT);
    if (smem <= per_sm / 2) {
      e = cudaFuncSetAttribute(k_interp_blk<BLK_BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      k_interp_blk<BLK_BT><<<a.Vp / nt, nt, smem, stream>>>(a);
      if (launches) *launches += 1;
      return cudaGetLastError();
    }
  }
  int nt = 128;
  const size_t limit = 200 * 1024;
  while (nt > 32 && interp_smem(a, nt, tile) > limit) nt >>= 1;
  if (interp_smem(a, nt, tile) > limit) return cudaErrorInvalidConfiguration;
  // few voices: smaller blocks spread the voices over more SMs
  while (nt > 32 && (a.Vp / nt) < 148) nt >>= 1;
  size_t smem = interp_smem(a, nt, tile);
  int blocks = a.Vp / nt;
  if (divergent) {
    e = cudaFuncSetAttribute(k_interp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_interp<true><<<blocks, nt, smem, stream>>>(a);
  } else {
    e = cudaFuncSetAttribute(k_interp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_interp<false><<<blocks, nt, smem, stream>>>(a);
  }
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_init_state(float* state_init, const uint32_t* defaults, int NS, int Vp, const HashInit* hi, int n_hi,
                              const uint64_t* salts, cudaStream_t stream) {
  k_init_state<<<(Vp + 127) / 128, 128, 0, stream>>>(state_init, defaults, NS, Vp, hi, n_hi, salts);
  return cudaGetLastError();
}
cudaError_t launch_reset_state(float* state, const float* state_init, const uint8_t* keep, int NS, int Vp, cudaStream_t stream) {
  size_t n = (size_t)NS * Vp;
  if (n == 0) return cudaSuccess;
  k_reset_state<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(state, state_init, keep, NS, Vp);
  return cudaGetLastError();
}
cudaError_t launch_broadcast_params(float* params, const float* tmpl, int P, int Vp, cudaStream_t stream) {
  size_t n = (size_t)P * Vp;
  if (n == 0) return cudaSuccess;
  k_broadcast_params<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(params, tmpl, P, Vp);
  return cudaGetLastError();
}
cudaError_t launch_stereo_frames(const float* src, int n_ch, long n, float* frames, cudaStream_t stream) {
  k_stereo_frames<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, n_ch, n, frames);
  return cudaGetLastError();
}
cudaError_t launch_stereo_frames_i16(const float* src, int n_ch, long n, int offset_binary, void* frames, cudaStream_t stream) {
  k_stereo_frames_i16<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, n_ch, n, offset_binary, (uint32_t*)frames);
  return cudaGetLastError();
}
cudaError_t launch_mix_rows(const float* rows, int R, long T, float scale, float* out, cudaStream_t stream) {
  k_mix_rows<<<(unsigned)((T + 255) / 256), 256, 0, stream>>>(rows, R, T, scale, out);
  return cudaGetLastError();
}

#endif  // !QG_SPEC_ONLY

}  // namespace qg
