// Op tape: the flat, topologically ordered program one voice executes per sample.
//
// A quartz audio graph (FunDSP `Net` built by /root/reference/src/functions.rs:111 `str_to_net` and the
// connective arms of /root/reference/src/process.rs:1669-1876) is lowered once on the host into this POD
// layout and evaluated by the CUDA kernels in interp.cu / fused.cu.  Per voice ("lane") the kernels keep
// ONE float array X, addressed by the 16-bit indices carried in the instructions, split in three regions:
//   X[0, P)            parameters  — derived coefficients, loaded once per launch from the bank's [P][V] table
//   X[P, P+NS)         state       — recurrences/counters (u32 counters bit-cast), persisted in [NS][V]
//   X[P+NS, P+NS+T)    temporaries — edge values of the current sample (net inputs first)
// Long delay lines live in HBM rings laid out [pos][voice] so a warp touches one 128-byte line per access.
#pragma once
#if defined(__CUDACC_RTC__)
#include "rtc_compat.h"
#else
#include <stdint.h>
#endif

namespace qg {

enum Op : uint16_t {
  OP_NOP = 0,
  // ---- data movement
  OP_MOV,        // T[out] = X[a]
  OP_ZERO,       // T[out] = 0
  OP_LD_STATE,   // T[out] = S[s]
  OP_ST_STATE,   // S[s] = X[a]
  // ---- arithmetic on two operands (operands may be temporaries or parameters)
  OP_ADD, OP_SUB, OP_MUL,
  OP_GT, OP_LT, OP_EQ, OP_NE, OP_GE, OP_LE,
  OP_MIN, OP_MAX, OP_POW, OP_REM, OP_LOG,
  OP_BITAND, OP_BITOR, OP_BITXOR, OP_SHL, OP_SHR,
  OP_HYPOT, OP_ATAN2, OP_DISSONANCE, OP_SIN_HZ, OP_COS_HZ, OP_SQR_HZ, OP_TRI_HZ,
  OP_PDHALF_BI, OP_PDHALF_UNI,
  // ---- three operands
  OP_LERP, OP_LERP11, OP_DELERP, OP_DELERP11, OP_XERP, OP_XERP11, OP_DEXERP, OP_DEXERP11,
  OP_SPLINE,     // five operands
  // ---- unary maps
  OP_ABS, OP_SIGNUM, OP_FLOOR, OP_FRACT, OP_CEIL, OP_ROUND, OP_SQRT, OP_EXP, OP_EXP2, OP_EXP10,
  OP_LN_1P_FN,   // computes ln(1+x)  (reached by the op string "exp_m1": functions.rs:1077)
  OP_EXP_M1_FN,  // computes exp(x)-1 (reached by the op string "ln_1p":  functions.rs:1078)
  OP_LN, OP_LOG2, OP_LOG10, OP_SIN, OP_COS, OP_TAN, OP_ASIN, OP_ACOS, OP_ATAN, OP_SINH, OP_COSH, OP_TANH,
  OP_ASINH, OP_ACOSH, OP_ATANH, OP_SQUARED, OP_CUBED, OP_DB_AMP, OP_AMP_DB, OP_A_WEIGHT, OP_SOFTSIGN,
  OP_SMOOTH3, OP_SMOOTH5, OP_SMOOTH7, OP_SMOOTH9, OP_UPARC, OP_DOWNARC, OP_SINE_EASE, OP_SEMITONE_RATIO,
  OP_RND1, OP_RND2, OP_DEG, OP_RAD, OP_RECIP, OP_NORMAL,
  OP_CLIP,       // clamp(X[a], P[p], P[p+1])
  OP_WRAP2,      // P[p]=p0, P[p+1]=r      (functions.rs:1149-1155)
  OP_WRAP1,      // P[p]=x                 (functions.rs:1156-1160)
  OP_MIRROR,     // P[p]=p0, p1, r         (functions.rs:1163-1181)
  OP_POL, OP_CAR,   // 2 -> 2 (out, out+1)
  OP_JOIN,       // T[out] = X[in0] + ... + X[in(n-1)] left to right (n <= 5), divided by aux when aux != 0 (join(n): mean)
  OP_DIVN,       // T[out] = X[a] / n
  OP_PAN,        // fixed pan: P[p]=l, P[p+1]=r ; 1 -> 2
  OP_PAN_VAR,    // (x, pan) -> 2 ; S[s]=cached pan, S[s+1]=l, S[s+2]=r
  OP_ROTATE,     // P[p]=c, P[p+1]=s ; 2 -> 2
  // ---- sources
  OP_SINE,       // S[s]=phase; P[p]=1/sr ; in a = frequency
  OP_NOISE,      // S[s]=u32 counter
  OP_IMPULSE,    // S[s]=fired flag
  OP_RAMP,       // S[s]=val; P[p]=sr ; in a = frequency      (nodes.rs:459-492)
  OP_WAVE,       // S[s]=u32 index; table aux, length aux2
  OP_WAVETABLE,  // band-limited saw/square/triangle/soft_saw: in a = frequency; P[p]=1/sr; S[s]=phase, S[s+1]=u32 table hint;
                 // aux = table-set header in the tables region: n, then n x (limit, offset, length), see lower.cpp
  // ---- filters
  OP_SVF,        // fixed coefficients P[p..p+5] = a1 a2 a3 m0 m1 m2 ; S[s]=ic1, S[s+1]=ic2
  OP_SVF_VAR,    // n = (mode) | (nvar << 8); inputs a=x, b.. = hz,q,gain ; P[p..p+3]=hz q gain sr defaults ;
                 // S[s..s+10] = ic1 ic2 a1 a2 a3 m0 m1 m2 hz q gain (cached)
  OP_BIQUAD,     // P[p..p+4] = a1 a2 b0 b1 b2 ; S[s..s+3] = x1 x2 y1 y2
  OP_BIQUAD_VAR, // n = kind | (nvar << 8) ; P[p..p+2] = p0 p1 sr ; S[s..s+10] = x1 x2 y1 y2 a1 a2 b0 b1 b2 p0 p1
  OP_ONEPOLE,    // n = kind (0 lowpole 1 highpole 2 dcblock 3 allpole) ; P[p]=coeff ; S[s]=x1, S[s+1]=y1
  OP_ONEPOLE_VAR,// n = kind ; P[p]=param default, P[p+1]=sr ; S[s..s+3] = x1 y1 coeff param
  OP_PINKPASS,   // S[s..s+6]
  OP_FIR,        // n taps ; P[p..p+n) weights ; S[s..s+n) history
  // ---- delays
  OP_TICK,       // S[s]
  OP_DELAY,      // ring aux ; S[s]=u32 index
  OP_TAP,        // n = cubic(1)/linear(0) ; ring aux ; P[p]=min, max, sr ; S[s]=u32 index
  OP_SAMP_DELAY, // ring aux (length = max) ; S[s]=u32 head   (nodes.rs:707-738)
  // ---- envelopes (lfo / lfo_in): n = shape | (n_inputs << 8) ; P[p..p+3] constants, P[p+4] = 1/sr ;
  //      S[s..s+7] = t t0 t1 v0 v1 thash_lo thash_hi first
  OP_ENVELOPE,
  OP_DECLICK,    // P[p]=dur, P[p+1]=1/sr ; S[s]=t
  // ---- in-tree stateful nodes
  OP_SHIFT_REG,  // (x, trig) -> 8 ; S[s..s+7]            (nodes.rs:157-190)
  OP_SNH,        // (x, trig) -> 1 ; S[s]                 (nodes.rs:795-821)
  OP_QUANTIZE,   // table aux length n ; P[p]=range       (nodes.rs:196-229)
  OP_ARR_GET,    // table aux length n                    (nodes.rs:127-150)
  // ---- control flow (nested nets); targets are absolute instruction indices in aux
  OP_KR_BEGIN,   // S[s]=u32 count ; n=period ; if count != 0 jump aux, else count = n      (nodes.rs:271-278)
  OP_KR_END,     // S[s] count -= 1
  OP_RESET_EVERY,// S[s]=u32 count ; aux=n samples ; reset range aux2 if count >= n ; count = (..)+1  (nodes.rs:351-360)
  OP_RESET_IF,   // if X[a] != 0 reset range aux2                                              (nodes.rs:393-400)
  OP_RESET_V,    // S[s]=u32 count ; P[p]=sr ; if count >= round(X[a]*sr) reset ; count += 1   (nodes.rs:433-442)
  OP_JNE_IDX,    // if (X[a] as usize) != n jump aux                                           (nodes.rs:27-33)
  OP_SEQ_TRIG,   // Seq trigger handling; see lower.cpp                                        (nodes.rs:74-92)
  OP_SEQ_GATE,   // per-net gate; jump aux when the net's event is not playing                 (nodes.rs:95-105)
  OP_SEQ_END,
  // ---- feedback (FunDSP FeedbackUnit): ring aux, S[s]=u32 index
  OP_FB_READ,    // T[out] = X[a] + ring[idx]
  OP_FB_WRITE,   // ring[idx] = X[a] ; n = 1 on the last channel -> advance idx
  // one-sample feedback (the default `feedback()` of process.rs:1500-1507): the delay line is ONE value, kept in the
  // state region (shared memory during a launch) instead of an HBM ring.  Sample-by-sample kernels only.
  OP_FB1_READ,   // T[out] = X[a] + S[s]
  OP_FB1_WRITE,  // S[s] = X[a]
  // ---- spectral nodes, per-lane path (nodes.rs:601-700): rings aux.. hold the node's buffers, aux2 = twiddle
  //      table offset, n = log2(N), S[s] = u32 count.  RFFT rings: in[N], re[N], im[N]; IFFT: in_re in_im out_re out_im
  OP_RFFT, OP_IFFT,
  // ---- streams (stage-split spectral path): read/write HBM signal arrays [stream][voice][time]
  OP_STREAM_IN, OP_STREAM_OUT,
  OP_COUNT_
};

#if defined(__CUDACC__)
#define QG_TAPE_HD __host__ __device__
#else
#define QG_TAPE_HD
#endif
// ops that read only the current sample's operands (no per-voice memory): they can be applied to any sample of a hop
// fixed-coefficient linear recurrences: the time-vector kernel evaluates them with a block-level scan over the hop
QG_TAPE_HD inline bool op_is_lti(uint16_t op) { return op == OP_SVF || op == OP_BIQUAD || op == OP_ONEPOLE; }
// phase accumulators: exact f32 sequential recurrences; the time-vector kernel runs the recurrence on one thread and the
// per-sample function (sin, table lookup) on all of them
QG_TAPE_HD inline bool op_is_phasor(uint16_t op) { return op == OP_SINE || op == OP_RAMP; }
QG_TAPE_HD inline bool op_is_stateless(uint16_t op) {
  return op == OP_NOP || op == OP_MOV || op == OP_ZERO || (op >= OP_ADD && op <= OP_PAN) || op == OP_ROTATE ||
         op == OP_QUANTIZE || op == OP_ARR_GET;
}

struct Instr {
  uint16_t op;
  uint16_t out;      // first output temporary (multi-output ops write out, out+1, ...)
  uint16_t in[5];    // operand indices into X (a parameter or a temporary)
  uint16_t p;        // first parameter index (into X)
  uint16_t s;        // first state index (into X)
  uint16_t n;        // small immediate (tap count, mode, period, ...)
  uint32_t aux;      // table offset / ring id / jump target / sample count
  uint32_t aux2;     // reset range id
  uint32_t pad;      // stateless instructions: end (exclusive) of the run of stateless instructions they belong to
};
static_assert(sizeof(Instr) == 32, "Instr must stay 32 bytes");

struct Ring {        // HBM delay line: element (pos, voice) at ring_base[(offset + pos) * V + voice]
  uint32_t offset;   // in units of V floats
  uint32_t length;
};
struct ResetRange {  // what "reset the inner net" means for a nested net
  uint16_t s_lo, s_hi;       // X indices (state region) restored from the bank's init table
  uint16_t ring_lo, ring_hi; // rings cleared
};
// How a state word is initialised per voice.
enum InitKind : uint32_t {
  INIT_SINE_PHASE = 1,   // rnd1(hash) as f32
  INIT_NOISE_SEED = 2,   // hash as u32
  INIT_HASH_LO = 3,      // low 32 bits of hash
  INIT_HASH_HI = 4,      // high 32 bits of hash
};
struct HashInit {
  uint32_t state;      // state index relative to the state region
  uint32_t kind;
  uint64_t hash;       // structural hash from ping(); salted per voice: hash' = salt ? atto(hash, salt) : hash
};

struct TapeHeader {
  uint32_t magic;      // 'QGTP'
  uint32_t version;
  uint32_t n_instr;
  uint32_t n_params;   // P
  uint32_t n_state;    // NS
  uint32_t n_temps;    // T
  uint32_t n_inputs;   // net inputs  (temporaries 0..n_inputs-1 are filled from the input buffer)
  uint32_t n_outputs;  // net outputs
  uint32_t n_rings;
  uint32_t ring_floats;// sum of ring lengths (floats per voice)
  uint32_t n_resets;
  uint32_t n_hash_init;
  uint32_t table_floats;
  uint32_t flags;      // TAPE_DIVERGENT when control-flow instructions are present
  uint32_t n_raw;      // raw (op-string) parameters of the graph, in order of appearance
  uint32_t fused_id;   // FUSED_* when the tape matches a hand-fused kernel shape, else 0
  float sample_rate;
  uint32_t n_lti;      // fixed-coefficient LTI filters (OP_SVF / OP_BIQUAD / OP_ONEPOLE): their `aux` is an index < n_lti
  uint32_t reserved[6];
};
static const uint32_t TAPE_MAGIC = 0x50544751u;
static const uint32_t TAPE_VERSION = 1;
static const uint32_t TAPE_DIVERGENT = 1u;
// floats of scan-matrix table per LTI filter in the time-vector kernel: M^(2^i) for i < 5, M^l for l < 32, M^32
static const int TV_LTI_FLOATS = (5 + 32 + 1) * 4;

}  // namespace qg
