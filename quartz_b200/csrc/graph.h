// Host-side graph IR: the data-only counterpart of a FunDSP `Net`.
//
// The reference keeps an audio graph as boxed `AudioUnit` objects that are ticked through virtual calls
// (/root/reference/src/process.rs:1347-1351).  Here a graph is plain data — primitive nodes in evaluation
// order plus edges — so it can be (a) composed with the same algebra and arity rules as the reference's
// connective ops (process.rs:1669-1876) and (b) lowered to an op tape for the GPU (lower.cpp).
// "Units" record which nodes came from one str_to_net() call: a FunDSP static graph such as
// `constant(f) >> sine()` is ONE Net vertex, which is what `Net::size()` counts (process.rs:1752, 1833).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace qg {

static const double DEFAULT_SR = 44100.0;

enum NodeKind : uint16_t {
  NK_CONST, NK_PASS, NK_SINK, NK_SPLIT, NK_REVERSE, NK_JOIN, NK_PAN, NK_ZERO_SRC,
  NK_UNARY, NK_BIN, NK_TERN, NK_SPLINE, NK_NARY_CONST, NK_MAP22, NK_CLIP, NK_WRAP2, NK_WRAP1, NK_MIRROR, NK_ROTATE,
  NK_SINE, NK_NOISE, NK_IMPULSE, NK_RAMP, NK_WAVE, NK_WAVETABLE,
  NK_SVF, NK_BIQUAD, NK_ONEPOLE, NK_PINKPASS, NK_FIR,
  NK_TICK, NK_DELAY, NK_TAP, NK_SAMP_DELAY, NK_ENVELOPE, NK_DECLICK,
  NK_SHIFT_REG, NK_SNH, NK_QUANTIZE, NK_ARR_GET,
  NK_KR, NK_FEEDBACK, NK_SELECT, NK_SEQ, NK_RESET, NK_RFFT, NK_IFFT,
};

// Unit ids mixed into the structural hash (AttoHash).  In-tree ids are the reference's `const ID`
// (nodes.rs); FunDSP ids are placeholders because its source is not vendored — see DESIGN.md "hashes".
enum : uint64_t {
  ID_PASS = 48, ID_SINK = 47, ID_CONSTANT = 8, ID_MAP = 27, ID_SINE = 21, ID_NOISE = 20, ID_SVF = 36,
  ID_BIQUAD = 15, ID_LOWPOLE = 12, ID_HIGHPOLE = 14, ID_DCBLOCK = 22, ID_ALLPOLE = 46, ID_PINKPASS = 42,
  ID_FIR = 5, ID_TICK = 9, ID_DELAY = 13, ID_TAP = 50, ID_TAPLIN = 51, ID_ENVELOPE = 14001, ID_ENVELOPE_IN = 53,
  ID_JOIN = 41, ID_SPLIT = 40, ID_REVERSE = 45, ID_PAN = 49, ID_DECLICK = 23, ID_IMPULSE = 81,
  ID_PIPE = 2, ID_STACK = 3, ID_BRANCH = 4, ID_BUS = 10, ID_BINOP = 11, ID_NET = 63, ID_FEEDBACK = 79, ID_WAVE = 65, ID_WAVESYNTH = 34,
  ID_SELECT = 1213, ID_SEQ = 1729, ID_ARRGET = 1312, ID_SHIFTREG = 1110, ID_QUANTIZER = 1111, ID_KR = 1112,
  ID_RESET = 1113, ID_TRIGRESET = 1114, ID_RESETV = 1115, ID_RAMP = 1116, ID_INPUT = 1117,
  ID_VAR = 70, ID_MONITOR = 56,
  ID_RFFT = 1120, ID_IFFT = 1121, ID_SAMPDELAY = 1122, ID_BUFFIN = 1123, ID_BUFFOUT = 1124, ID_SNH = 1125,
};

uint64_t atto(uint64_t state, uint64_t data);

struct Graph;

struct Src {
  int node;   // >= 0: output `port` of nodes[node]; -1: graph input `port`; -2: constant zero
  int port;
};

struct Node {
  uint16_t kind = NK_PASS;
  uint16_t devop = 0;          // device opcode for the generic stateless kinds
  int n_in = 0, n_out = 0;
  int mode = 0, aux = 0;       // kind-specific small integers (filter mode, fixed-parameter count, ...)
  std::vector<Src> in;
  std::vector<float> raw;      // the op-string parameters exactly as the reference's constructor received them
  std::vector<float> table;    // array-fed nodes (quantize/get/wave)
  std::vector<Graph> kids;     // nested nets (kr, feedback, select, seq, reset*)
  double sr = DEFAULT_SR;
  uint64_t hash = 0;           // structural hash assigned by Graph::rehash() (FunDSP `ping`)
};

struct PingStep {
  enum Type : int { MIX, LEAF, KID } type;
  uint64_t id;
  int node;    // LEAF: node index; KID: owner node index (kid 0 is pinged)
};

struct Unit {
  std::vector<PingStep> ping;
};

struct Graph {
  std::vector<Node> nodes;
  std::vector<Unit> units;
  std::vector<Src> outs;
  int n_in = 0;
  double sr = DEFAULT_SR;
  std::string unsupported;     // non-empty: the op exists in the reference but has no GPU lowering yet

  Graph() {}
  Graph(int ni, int no);                    // Net::new(i, o)
  int inputs() const { return n_in; }
  int outputs() const { return (int)outs.size(); }
  int size() const { return (int)units.size(); }
  void set_sample_rate(double sr);          // AudioUnit::set_sample_rate over every vertex (process.rs:1573)
  uint64_t ping(bool probe, uint64_t h);
  void rehash();

  // one-unit graph from a list of primitive nodes; `wire` fills inputs/outputs (used by parse.cpp)
  int add_node(const Node& n) { nodes.push_back(n); return (int)nodes.size() - 1; }

  // Net algebra.  kind: '>' pipe, '|' stack, '&' bus, '^' branch, '+' '-' '*' binops.  No arity checks here:
  // the guards belong to the callers, exactly as in process.rs.
  static Graph combine(char kind, Graph a, Graph b);
  static Graph thru(Graph a);
};

// functions.rs:111 — same tokenisation, same arity-by-parameter-count rules, same silent fallbacks.
Graph str_to_net(const std::string& op);
bool parse_with_constants(const std::string& s, float* out);   // functions.rs:47-109

// graph-level constructors (process.rs:1450-1667)
Graph make_get(const std::vector<float>& arr);
Graph make_quantize(const std::vector<float>& arr);
Graph make_wave(const std::vector<float>& arr);
Graph make_feedback(const Graph& net, bool has_delay, double delay);
Graph make_kr(const Graph& net, double n, bool preserve_time);
Graph make_reset(const Graph& net, double s);
Graph make_trig_reset(const Graph& net, bool variable);
Graph make_seq_select(bool is_seq, const std::vector<const Graph*>& nets);
Graph make_live_io(const std::string& name);
Graph make_var(float value);
// connective circles (process.rs:1719-1876) and array ops (process.rs:1669-1717)
Graph connect(const std::string& op, const std::vector<const Graph*>& nets, double number, int node_limit);
Graph array_op(const std::string& kind, const std::string& op_str, const std::vector<float>& arr);

}  // namespace qg
