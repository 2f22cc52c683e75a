// Host container of a lowered tape + the parameter derivation needed to batch voices into a bank.
#pragma once
#include <string>
#include <vector>

#include "graph.h"
#include "tape.h"

namespace qg {

// How a block of device parameters is derived from a block of raw (op-string) parameters.
enum DeriveKind : uint32_t {
  D_COPY = 1,      // P[i] = raw[i]            (n values)
  D_SVF,           // raw = hz, q[, gain] -> a1 a2 a3 m0 m1 m2   (mode, sr)
  D_SVF_DEFAULTS,  // variable-input SVF: P = hz q gain sr defaults, raw = the fixed trailing parameters
  D_BIQUAD,        // kind 1/2 fixed: raw -> a1 a2 b0 b1 b2
  D_ONEPOLE,       // raw[0] -> coeff (kind, sr)
  D_WRAP2,         // raw p0,p1 -> min, (max-min)
  D_MIRROR,        // raw p0,p1 -> min, max, (max-min)
  D_ROTATE,        // raw angle,gain -> cos*gain, sin*gain
  D_PAN,           // raw pan -> l, r
  D_RAMP_SR,       // P[0] = sr
  D_INV_SR,        // P[0] = (float)(1/sr)
  D_TAP,           // raw min,max -> min, max, sr
};
struct Deriver {
  uint32_t kind;
  uint32_t raw_base, n_raw;
  uint32_t p_base, n_p;
  int32_t mode, aux;
  float sr;
};

struct Tape {
  TapeHeader h;
  std::vector<Instr> code;
  std::vector<float> params;          // template parameter values (P)
  std::vector<uint32_t> state_init;   // default state words (bit patterns)
  std::vector<uint8_t> state_keep;    // 1: the word survives reset() — Seq::reset only resets its nets, the event list stays
                                      // (/root/reference/src/nodes.rs:116-120)
  std::vector<Ring> rings;
  std::vector<ResetRange> resets;
  std::vector<HashInit> hash_init;
  std::vector<float> tables;
  std::vector<uint16_t> out_x;        // X index of each net output
  std::vector<float> raw;             // template raw parameters, in lowering order
  std::vector<uint8_t> raw_structural;// 1: the raw value shapes the tape (delay lengths...) and must not vary per voice
  std::vector<Deriver> derivers;
  uint64_t signature = 0;             // structure hash: equal signatures <=> voices can share the tape

  // derive device parameters for one voice from its raw parameter vector
  void derive(const float* raw_in, float* p_out) const;
};

// Time-vector execution plan (k_interp_tv): possible when every op is stateless or one of the block-capable stateful
// ops (noise, wave, impulse, tick, delay, tap, rfft, ifft).  H = hop in samples.
struct TvPlan {
  bool ok = false;
  bool has_fft = false;
  int H = 0;
  int fft_n = 0;
  int align_s = -1;   // X index of one rfft/ifft counter
  int n_lti = 0;      // fixed-coefficient LTI filters evaluated by block-level scans
  bool sequential = false;   // the tape has phase accumulators that one thread steps through: only worth it for small banks
};

TvPlan plan_tv(const Tape& t, size_t smem_limit);

// Lower a graph.  Returns false (and fills `err`) when the graph contains something that has no GPU lowering —
// the product never falls back to a CPU path.
bool lower(const Graph& g, Tape* out, std::string* err);
void collect_raw(const Graph& g, std::vector<float>* raw);
uint64_t structure_signature(const Graph& g);

}  // namespace qg
