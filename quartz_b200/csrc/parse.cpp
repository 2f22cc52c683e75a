// Op string -> one-unit Graph.  Mirror of /root/reference/src/functions.rs:111-1226 `str_to_net`:
// identical tokenisation (:112-127), constant names (:47-109), arity-by-parameter-count rules and silent
// fallbacks to Net::new(0,0) (:124-127, :1225).  Ops whose FunDSP implementation has no GPU lowering yet
// produce a graph flagged `unsupported` so that rendering them fails loudly instead of going silent.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>

#include "graph.h"
#include "tape.h"

namespace qg {

static bool parse_f32(const std::string& s, float* out) {   // Rust `str::parse::<f32>` grammar
  if (s.empty()) return false;
  const char* p = s.c_str();
  size_t i = 0;
  if (p[i] == '+' || p[i] == '-') i++;
  std::string low;
  for (const char* q = p + i; *q; q++) low.push_back((char)std::tolower((unsigned char)*q));
  if (low == "inf" || low == "infinity" || low == "nan") {
    float v = low == "nan" ? std::numeric_limits<float>::quiet_NaN() : std::numeric_limits<float>::infinity();
    *out = p[0] == '-' ? -v : v;
    return true;
  }
  bool digits = false, dot = false;
  size_t j = i;
  for (; p[j]; j++) {
    if (p[j] >= '0' && p[j] <= '9') digits = true;
    else if (p[j] == '.' && !dot) dot = true;
    else break;
  }
  if (!digits) return false;
  if (p[j] == 'e' || p[j] == 'E') {
    size_t k = j + 1;
    if (p[k] == '+' || p[k] == '-') k++;
    if (!(p[k] >= '0' && p[k] <= '9')) return false;
    while (p[k] >= '0' && p[k] <= '9') k++;
    j = k;
  }
  if (p[j] != 0) return false;
  *out = std::strtof(p, nullptr);
  return true;
}

bool parse_with_constants(const std::string& s, float* out) {
  if (parse_f32(s, out)) return true;
  struct K { const char* name; float v; };
  static const K table[] = {
      {"E", 2.71828182845904523536f}, {"FRAC_1_PI", 0.318309886183790671538f},
      {"FRAC_1_SQRT_2", 0.707106781186547524401f}, {"FRAC_2_PI", 0.636619772367581343076f},
      {"FRAC_2_SQRT_PI", 1.12837916709551257390f}, {"FRAC_PI_2", 1.57079632679489661923f},
      {"FRAC_PI_3", 1.04719755119659774615f}, {"FRAC_PI_4", 0.785398163397448309616f},
      {"FRAC_PI_6", 0.52359877559829887308f}, {"FRAC_PI_8", 0.39269908169872415481f},
      {"LN_2", 0.693147180559945309417f}, {"LN_10", 2.30258509299404568402f},
      {"LOG2_10", 3.32192809488736234787f}, {"LOG2_E", 1.44269504088896340736f},
      {"LOG10_2", 0.301029995663981195214f}, {"LOG10_E", 0.434294481903251827651f},
      {"PI", 3.14159265358979323846f}, {"SQRT_2", 1.41421356237309504880f}, {"TAU", 6.28318530717958647692f},
      {"EGAMMA", 0.5772157f}, {"FRAC_1_SQRT_3", 0.57735026f}, {"FRAC_1_SQRT_PI", 0.5641896f},
      {"PHI", 1.618034f}, {"SQRT_3", 1.7320508f}};
  bool neg = !s.empty() && s[0] == '-';
  const char* key = s.c_str() + (neg ? 1 : 0);
  for (const K& k : table)
    if (!strcmp(k.name, key)) { *out = neg ? -k.v : k.v; return true; }
  if (s == "MAX") { *out = std::numeric_limits<float>::max(); return true; }
  if (s == "MIN") { *out = std::numeric_limits<float>::lowest(); return true; }
  if (s == "EPSILON") { *out = std::numeric_limits<float>::epsilon(); return true; }
  if (s == "MIN_POSITIVE") { *out = std::numeric_limits<float>::min(); return true; }
  return false;
}

namespace {

static uint64_t as_usize(float x) {
  if (!(x > 0.0f)) return 0;
  if (x >= 18446744073709551616.0f) return UINT64_MAX;
  return (uint64_t)x;
}

// Builds the single Net vertex ("unit") a str_to_net() call returns.  The order of leaf()/mix() calls is the
// order FunDSP's ping() visits the static graph: Pipe/Stack/Branch/Binop mix their id, then x, then y.
struct UnitBuilder {
  Graph g;
  Unit u;
  void mix(uint64_t id) { u.ping.push_back(PingStep{PingStep::MIX, id, 0}); }
  int node(uint16_t kind, int n_in, int n_out, std::vector<Src> in, std::vector<float> raw = {}) {
    Node n;
    n.kind = kind; n.n_in = n_in; n.n_out = n_out; n.in = std::move(in); n.raw = std::move(raw);
    return g.add_node(n);
  }
  int leaf(uint64_t id, uint16_t kind, int n_in, int n_out, std::vector<Src> in, std::vector<float> raw = {}) {
    int i = node(kind, n_in, n_out, std::move(in), std::move(raw));
    u.ping.push_back(PingStep{PingStep::LEAF, id, i});
    return i;
  }
  Node& N(int i) { return g.nodes[i]; }
  Graph finish(int n_in, std::vector<Src> outs) {
    g.n_in = n_in;
    g.outs = std::move(outs);
    g.units.push_back(u);
    g.rehash();
    return g;
  }
};
static std::vector<Src> gin(int n, int from = 0) {
  std::vector<Src> v;
  for (int i = 0; i < n; i++) v.push_back(Src{-1, from + i});
  return v;
}
static std::vector<Src> outs_of(int node, int n) {
  std::vector<Src> v;
  for (int i = 0; i < n; i++) v.push_back(Src{node, i});
  return v;
}
// a unit made of one node fed straight from the unit's inputs
static Graph single(uint64_t id, uint16_t kind, int n_in, int n_out, std::vector<float> raw = {}, uint16_t devop = 0,
                    int mode = 0, int aux = 0) {
  UnitBuilder b;
  int i = b.leaf(id, kind, n_in, n_out, gin(n_in), std::move(raw));
  b.N(i).devop = devop; b.N(i).mode = mode; b.N(i).aux = aux;
  return b.finish(n_in, outs_of(i, n_out));
}
// An op the reference has but whose FunDSP implementation is not restated here.  The placeholder carries the op's real
// arity (FunDSP's documented signature), so composition — arity guards, stacking, node counts — proceeds exactly as in the
// reference, and the `unsupported` mark travels with the graph: lowering refuses it by name instead of rendering
// something else.
static Graph unsupported(const std::string& name, int n_in, int n_out) {
  Graph g = single(ID_MAP, NK_ZERO_SRC, n_in, n_out);
  g.unsupported = name;
  return g;
}
static Graph svf(int mode, const std::vector<float>& p) {
  int npar = mode >= 6 ? 3 : 2;
  int nfixed;
  std::vector<float> raw;
  if ((int)p.size() >= npar) { nfixed = npar; raw.assign(p.begin(), p.begin() + npar); }
  else if ((int)p.size() >= npar - 1) { nfixed = npar - 1; raw.assign(p.begin(), p.begin() + (npar - 1)); }
  else nfixed = 0;
  return single(ID_SVF, NK_SVF, 1 + npar - nfixed, 1, raw, 0, mode, nfixed);
}
// `>(c)` style: one parameter => constant right-hand side, none => second input (functions.rs:824-935)
static Graph bin_or_const(uint16_t devop, const std::vector<float>& p) {
  if (!p.empty()) return single(ID_MAP, NK_BIN, 1, 1, {p[0]}, devop);
  return single(ID_MAP, NK_BIN, 2, 1, {}, devop);
}
static Graph tern_or_const(uint16_t devop, const std::vector<float>& p) {   // functions.rs:1002-1065
  if (p.size() >= 2) return single(ID_MAP, NK_TERN, 1, 1, {p[0], p[1]}, devop);
  return single(ID_MAP, NK_TERN, 3, 1, {}, devop);
}
static Graph nary_const(const std::vector<float>& p, uint16_t devop, bool recip) {   // functions.rs:709-804
  std::vector<float> c;
  for (size_t i = 0; i < p.size() && i < 8; i++) c.push_back(recip ? 1.0f / p[i] : p[i]);
  if (c.empty()) c.push_back(1.0f);
  int n = (int)c.size();
  return single(ID_MAP, NK_NARY_CONST, n, n, c, devop);
}

}  // namespace

Graph str_to_net(const std::string& op_in) {
  std::string op;
  for (char c : op_in) if (c != ' ') op.push_back(c);
  std::vector<std::string> args(1);
  for (char c : op) {
    if (c == '(' || c == ')') args.emplace_back();
    else args.back().push_back(c);
  }
  if (args.size() < 2) return Graph(0, 0);   // no parentheses (functions.rs:124-127)
  std::vector<float> p;
  {
    std::string cur;
    std::vector<std::string> toks;
    for (char c : args[1]) {
      if (c == ',') { toks.push_back(cur); cur.clear(); } else cur.push_back(c);
    }
    toks.push_back(cur);
    for (const std::string& t : toks) {
      float v;
      if (parse_with_constants(t, &v)) p.push_back(v);   // unparseable parameters are dropped (:119-123)
    }
  }
  const std::string& name = args[0];
  auto has = [&](size_t n) { return p.size() >= n; };
  const Graph EMPTY(0, 0);

  // -------------------- sources (functions.rs:130-230)
  if (name == "sine") {
    if (has(1)) {   // sine_hz(f) = constant(f) >> sine()
      UnitBuilder b;
      b.mix(ID_PIPE);
      int c = b.leaf(ID_CONSTANT, NK_CONST, 0, 1, {}, {p[0]});
      int s = b.leaf(ID_SINE, NK_SINE, 1, 1, {Src{c, 0}});
      return b.finish(0, {Src{s, 0}});
    }
    return single(ID_SINE, NK_SINE, 1, 1);
  }
  if (name == "white" || name == "noise") return single(ID_NOISE, NK_NOISE, 0, 1);
  if (name == "brown") {   // (white() >> lowpole_hz(10)) * dc(13.7)
    UnitBuilder b;
    b.mix(ID_BINOP);
    b.mix(ID_PIPE);
    int nz = b.leaf(ID_NOISE, NK_NOISE, 0, 1, {});
    int lp = b.leaf(ID_LOWPOLE, NK_ONEPOLE, 1, 1, {Src{nz, 0}}, {10.0f});
    int c = b.leaf(ID_CONSTANT, NK_CONST, 0, 1, {}, {13.7f});
    int m = b.node(NK_BIN, 2, 1, {Src{lp, 0}, Src{c, 0}});
    b.N(m).devop = OP_MUL;
    return b.finish(0, {Src{m, 0}});
  }
  if (name == "pink") {   // white() >> pinkpass()
    UnitBuilder b;
    b.mix(ID_PIPE);
    int nz = b.leaf(ID_NOISE, NK_NOISE, 0, 1, {});
    int pk = b.leaf(ID_PINKPASS, NK_PINKPASS, 1, 1, {Src{nz, 0}});
    return b.finish(0, {Src{pk, 0}});
  }
  if (name == "zero") return single(ID_CONSTANT, NK_CONST, 0, 1, {0.0f});
  if (name == "impulse") return single(ID_IMPULSE, NK_IMPULSE, 0, 1);
  if (name == "constant" || name == "dc") {   // functions.rs:180-196
    std::vector<float> v(p.begin(), p.begin() + std::min<size_t>(p.size(), 8));
    if (v.empty()) v.push_back(1.0f);
    int n = (int)v.size();
    return single(ID_CONSTANT, NK_CONST, 0, n, v);
  }
  if (name == "ramp") return single(ID_RAMP, NK_RAMP, 1, 1);
  if (name == "saw" || name == "square" || name == "triangle" || name == "soft_saw") {   // band-limited wavetable oscillators
    int shape = name == "saw" ? 0 : name == "square" ? 1 : name == "triangle" ? 2 : 3;
    if (has(1)) {   // saw_hz(f) = constant(f) >> saw()
      UnitBuilder b;
      b.mix(ID_PIPE);
      int c = b.leaf(ID_CONSTANT, NK_CONST, 0, 1, {}, {p[0]});
      int w = b.leaf(ID_WAVESYNTH, NK_WAVETABLE, 1, 1, {Src{c, 0}});
      b.N(w).mode = shape;
      return b.finish(0, {Src{w, 0}});
    }
    return single(ID_WAVESYNTH, NK_WAVETABLE, 1, 1, {}, 0, shape);
  }
  if (name == "organ" || name == "hammond") return unsupported(name, has(1) ? 0 : 1, 1);      // organ_hz(f) / organ()
  if (name == "pulse") return unsupported(name, 2, 1);                                          // frequency, duty cycle
  if (name == "lorenz" || name == "rossler") return unsupported(name, 1, 1);                    // frequency
  if (name == "dsf_saw" || name == "dsf_square") return unsupported(name, has(1) ? 1 : 2, 1);   // dsf_saw_r(r) / dsf_saw()
  if (name == "mls") return unsupported(name, 0, 1);
  if (name == "pluck") return has(3) ? unsupported(name, 1, 1) : EMPTY;

  // -------------------- filters (functions.rs:233-430)
  if (name == "lowpass") return svf(0, p);
  if (name == "highpass") return svf(1, p);
  if (name == "bandpass") return svf(2, p);
  if (name == "notch") return svf(3, p);
  if (name == "peak") return svf(4, p);
  if (name == "allpass") return svf(5, p);
  if (name == "bell") return svf(6, p);
  if (name == "lowshelf") return svf(7, p);
  if (name == "highshelf") return svf(8, p);
  if (name == "biquad") return has(5) ? single(ID_BIQUAD, NK_BIQUAD, 1, 1, {p[0], p[1], p[2], p[3], p[4]}, 0, 0, 0) : EMPTY;
  if (name == "butterpass")
    return has(1) ? single(ID_BIQUAD, NK_BIQUAD, 1, 1, {p[0]}, 0, 1, 0) : single(ID_BIQUAD, NK_BIQUAD, 2, 1, {}, 0, 1, 1);
  if (name == "resonator")
    return has(2) ? single(ID_BIQUAD, NK_BIQUAD, 1, 1, {p[0], p[1]}, 0, 2, 0) : single(ID_BIQUAD, NK_BIQUAD, 3, 1, {}, 0, 2, 2);
  if (name == "lowpole")
    return has(1) ? single(ID_LOWPOLE, NK_ONEPOLE, 1, 1, {p[0]}, 0, 0) : single(ID_LOWPOLE, NK_ONEPOLE, 2, 1, {}, 0, 0, 1);
  if (name == "highpole")
    return has(1) ? single(ID_HIGHPOLE, NK_ONEPOLE, 1, 1, {p[0]}, 0, 1) : single(ID_HIGHPOLE, NK_ONEPOLE, 2, 1, {}, 0, 1, 1);
  if (name == "dcblock") return single(ID_DCBLOCK, NK_ONEPOLE, 1, 1, {has(1) ? p[0] : 10.0f}, 0, 2);
  if (name == "allpole")
    return has(1) ? single(ID_ALLPOLE, NK_ONEPOLE, 1, 1, {p[0]}, 0, 3) : single(ID_ALLPOLE, NK_ONEPOLE, 2, 1, {}, 0, 3, 1);
  if (name == "pinkpass") return single(ID_PINKPASS, NK_PINKPASS, 1, 1);
  if (name == "fir") {
    if (!has(1)) return EMPTY;
    return single(ID_FIR, NK_FIR, 1, 1, std::vector<float>(p.begin(), p.begin() + std::min<size_t>(p.size(), 10)));
  }
  if (name == "fir3") {
    if (!has(1)) return EMPTY;
    float alpha = (p[0] + 1.0f) / 2.0f, beta = (1.0f - alpha) / 2.0f;
    return single(ID_FIR, NK_FIR, 1, 1, {beta, alpha, beta});
  }
  if (name == "follow") return has(1) ? unsupported(name, 1, 1) : EMPTY;                        // follow(t) / afollow(a, r)
  if (name == "moog" || name == "lowrez" || name == "bandrez")                                  // x_hz(f, q) / x_q(q) / x()
    return unsupported(name, has(2) ? 1 : (has(1) ? 2 : 3), 1);
  if (name == "morph") return unsupported(name, has(3) ? 1 : 4, 1);                             // morph_hz(f, q, m) / morph()

  // -------------------- channels (functions.rs:433-494)
  if (name == "sink") return single(ID_SINK, NK_SINK, 1, 0);
  if (name == "pass") return single(ID_PASS, NK_PASS, 1, 1);
  if (name == "chan") {
    Graph net(0, 0);
    for (float v : p)
      net = Graph::combine('|', std::move(net), v == 0.0f ? single(ID_SINK, NK_SINK, 1, 0) : single(ID_PASS, NK_PASS, 1, 1));
    return net;
  }
  if (name == "pan") return has(1) ? single(ID_PAN, NK_PAN, 1, 2, {p[0]}) : single(ID_PAN, NK_PAN, 2, 2);
  if (name == "join" || name == "split" || name == "reverse") {
    if (has(1)) {
      uint64_t n = as_usize(p[0]);
      if (n >= 2 && n <= 8) {
        if (name == "join") return single(ID_JOIN, NK_JOIN, (int)n, 1);
        if (name == "split") return single(ID_SPLIT, NK_SPLIT, 1, (int)n);
        return single(ID_REVERSE, NK_REVERSE, (int)n, (int)n);
      }
    }
    return EMPTY;
  }

  // -------------------- envelopes (functions.rs:497-578): mode = shape (0 xd, 1 xD, 2 ar, 3 t)
  if (name == "adsr") return has(4) ? unsupported(name, 1, 1) : EMPTY;                          // adsr_live: gate in
  if (name == "xd")
    return has(1) ? single(ID_ENVELOPE, NK_ENVELOPE, 0, 1, {p[0]}, 0, 0) : single(ID_ENVELOPE_IN, NK_ENVELOPE, 1, 1, {}, 0, 0);
  if (name == "xD") {
    if (has(2)) return single(ID_ENVELOPE, NK_ENVELOPE, 0, 1, {p[0], p[1]}, 0, 1);
    if (has(1)) return single(ID_ENVELOPE_IN, NK_ENVELOPE, 1, 1, {p[0]}, 0, 1);
    return single(ID_ENVELOPE_IN, NK_ENVELOPE, 2, 1, {}, 0, 1);
  }
  if (name == "ar") {
    if (has(4)) return single(ID_ENVELOPE, NK_ENVELOPE, 0, 1, {p[0], p[1], p[2], p[3]}, 0, 2);
    if (has(2)) return single(ID_ENVELOPE_IN, NK_ENVELOPE, 2, 1, {p[0], p[1]}, 0, 2);
    return single(ID_ENVELOPE_IN, NK_ENVELOPE, 4, 1, {}, 0, 2);
  }

  // -------------------- other (functions.rs:581-706)
  if (name == "tick") return single(ID_TICK, NK_TICK, 1, 1);
  if (name == "shift_reg") return single(ID_SHIFTREG, NK_SHIFT_REG, 2, 8);
  if (name == "snh") return single(ID_SNH, NK_SNH, 2, 1);
  if (name == "meter") {   // meter(peak|rms, t): the mode word is not a number, the time is (functions.rs:584-592)
    const bool mode = args[1].rfind("peak", 0) == 0 || args[1].rfind("rms", 0) == 0;
    return (mode && has(1)) ? unsupported(name, 1, 1) : EMPTY;
  }
  if (name == "chorus") return has(4) ? unsupported(name, 1, 1) : EMPTY;
  if (name == "hold") return has(2) ? unsupported(name, 1, 1) : (has(1) ? unsupported(name, 2, 1) : EMPTY);   // hold_hz / hold
  if (name == "limiter") return has(2) ? unsupported(name, 1, 1) : EMPTY;
  if (name == "limiter_stereo") return has(2) ? unsupported(name, 2, 2) : EMPTY;
  if (name == "reverb_stereo") return has(1) ? unsupported(name, 2, 2) : EMPTY;
  if (name == "reverb_mono") return has(1) ? unsupported(name, 1, 1) : EMPTY;
  if (name == "clip") {
    float lo = -1.0f, hi = 1.0f;
    if (has(2)) { lo = p[0] < p[1] ? p[0] : p[1]; hi = p[0] < p[1] ? p[1] : p[0]; }
    return single(ID_MAP, NK_CLIP, 1, 1, {lo, hi});
  }
  if (name == "declick") return single(ID_DECLICK, NK_DECLICK, 1, 1, {has(1) ? p[0] : 0.010f});
  if (name == "delay") return has(1) ? single(ID_DELAY, NK_DELAY, 1, 1, {p[0]}) : EMPTY;
  if (name == "tap" || name == "tap_linear") {
    if (!has(2)) return EMPTY;
    float p0 = std::fmax(p[0], 0.0f), p1 = std::fmax(p[1], 0.0f);
    bool cubic = name == "tap";
    return single(cubic ? ID_TAP : ID_TAPLIN, NK_TAP, 2, 1, {std::fmin(p0, p1), std::fmax(p0, p1)}, 0, cubic ? 1 : 0);
  }
  if (name == "samp_delay") {
    if (!has(1)) return EMPTY;
    uint64_t mx = as_usize(p[0]);
    if (mx > 0x7fffffffULL) mx = 0x7fffffffULL;
    return single(ID_SAMPDELAY, NK_SAMP_DELAY, 2, 1, {}, 0, 0, (int)mx);
  }
  if (name == "pdhalf_bi") return single(ID_MAP, NK_BIN, 2, 1, {}, OP_PDHALF_BI);
  if (name == "pdhalf_uni") return single(ID_MAP, NK_BIN, 2, 1, {}, OP_PDHALF_UNI);

  // -------------------- math (functions.rs:709-1222)
  if (name == "add") return nary_const(p, OP_ADD, false);
  if (name == "sub") return nary_const(p, OP_SUB, false);
  if (name == "mul") return nary_const(p, OP_MUL, false);
  if (name == "div") return nary_const(p, OP_MUL, true);
  if (name == "rotate") return has(2) ? single(ID_MAP, NK_ROTATE, 2, 2, {p[0], p[1]}) : EMPTY;
  if (name == "t") return single(ID_ENVELOPE, NK_ENVELOPE, 0, 1, {}, 0, 3);
  if (name == "rise" || name == "fall") {   // (pass() ^ tick()) >> map   (functions.rs:813-822)
    UnitBuilder b;
    b.mix(ID_PIPE);
    b.mix(ID_BRANCH);
    int ps = b.leaf(ID_PASS, NK_PASS, 1, 1, gin(1));
    int tk = b.leaf(ID_TICK, NK_TICK, 1, 1, gin(1));
    int m = b.leaf(ID_MAP, NK_BIN, 2, 1, {Src{ps, 0}, Src{tk, 0}});
    b.N(m).devop = name == "rise" ? OP_GT : OP_LT;
    return b.finish(1, {Src{m, 0}});
  }
  static const std::map<std::string, uint16_t> bin = {
      {">", OP_GT}, {"<", OP_LT}, {"==", OP_EQ}, {"!=", OP_NE}, {">=", OP_GE}, {"<=", OP_LE},
      {"min", OP_MIN}, {"max", OP_MAX}, {"pow", OP_POW}, {"mod", OP_REM}, {"rem", OP_REM}, {"log", OP_LOG},
      {"bitand", OP_BITAND}, {"bitor", OP_BITOR}, {"bitxor", OP_BITXOR}, {"shl", OP_SHL}, {"shr", OP_SHR}};
  {
    auto it = bin.find(name);
    if (it != bin.end()) return bin_or_const(it->second, p);
  }
  static const std::map<std::string, uint16_t> tern = {
      {"lerp", OP_LERP}, {"lerp11", OP_LERP11}, {"delerp", OP_DELERP}, {"delerp11", OP_DELERP11},
      {"xerp", OP_XERP}, {"xerp11", OP_XERP11}, {"dexerp", OP_DEXERP}, {"dexerp11", OP_DEXERP11}};
  {
    auto it = tern.find(name);
    if (it != tern.end()) return tern_or_const(it->second, p);
  }
  static const std::map<std::string, uint16_t> unary = {
      {"abs", OP_ABS}, {"signum", OP_SIGNUM}, {"floor", OP_FLOOR}, {"fract", OP_FRACT}, {"ceil", OP_CEIL},
      {"round", OP_ROUND}, {"sqrt", OP_SQRT}, {"exp", OP_EXP}, {"exp2", OP_EXP2}, {"exp10", OP_EXP10},
      // bug-for-bug with functions.rs:1077-1078: "exp_m1" computes ln_1p and "ln_1p" computes exp_m1
      {"exp_m1", OP_LN_1P_FN}, {"ln_1p", OP_EXP_M1_FN},
      {"ln", OP_LN}, {"log2", OP_LOG2}, {"log10", OP_LOG10}, {"sin", OP_SIN}, {"cos", OP_COS}, {"tan", OP_TAN},
      {"asin", OP_ASIN}, {"acos", OP_ACOS}, {"atan", OP_ATAN}, {"sinh", OP_SINH}, {"cosh", OP_COSH}, {"tanh", OP_TANH},
      {"asinh", OP_ASINH}, {"acosh", OP_ACOSH}, {"atanh", OP_ATANH}, {"squared", OP_SQUARED}, {"cubed", OP_CUBED},
      {"db_amp", OP_DB_AMP}, {"amp_db", OP_AMP_DB}, {"a_weight", OP_A_WEIGHT}, {"softsign", OP_SOFTSIGN},
      {"smooth3", OP_SMOOTH3}, {"smooth5", OP_SMOOTH5}, {"smooth7", OP_SMOOTH7}, {"smooth9", OP_SMOOTH9},
      {"uparc", OP_UPARC}, {"downarc", OP_DOWNARC}, {"sine_ease", OP_SINE_EASE}, {"semitone_ratio", OP_SEMITONE_RATIO},
      {"rnd1", OP_RND1}, {"rnd2", OP_RND2}, {"deg", OP_DEG}, {"rad", OP_RAD}, {"recip", OP_RECIP}, {"normal", OP_NORMAL}};
  {
    auto it = unary.find(name);
    if (it != unary.end()) return single(ID_MAP, NK_UNARY, 1, 1, {}, it->second);
  }
  static const std::map<std::string, uint16_t> bin2 = {
      {"hypot", OP_HYPOT}, {"atan2", OP_ATAN2}, {"dissonance", OP_DISSONANCE}, {"sin_hz", OP_SIN_HZ},
      {"cos_hz", OP_COS_HZ}, {"sqr_hz", OP_SQR_HZ}, {"tri_hz", OP_TRI_HZ}};
  {
    auto it = bin2.find(name);
    if (it != bin2.end()) return single(ID_MAP, NK_BIN, 2, 1, {}, it->second);
  }
  if (name == "spline") return single(ID_MAP, NK_SPLINE, 5, 1);
  if (name == "dissonance_max" || name == "m_weight" || name == "softexp") return unsupported(name, 1, 1);
  if (name == "spline_mono") return unsupported(name, 5, 1);
  if (name == "softmix") return unsupported(name, 3, 1);
  if (name == "spline_noise") return unsupported(name, 2, 1);
  if (name == "fractal_noise") return unsupported(name, 4, 1);
  if (name == "wrap") {
    if (has(2)) return single(ID_MAP, NK_WRAP2, 1, 1, {p[0], p[1]});
    if (has(1)) return single(ID_MAP, NK_WRAP1, 1, 1, {p[0]});
    return EMPTY;
  }
  if (name == "mirror") return has(2) ? single(ID_MAP, NK_MIRROR, 1, 1, {p[0], p[1]}) : EMPTY;
  if (name == "pol") return single(ID_MAP, NK_MAP22, 2, 2, {}, OP_POL);
  if (name == "car") return single(ID_MAP, NK_MAP22, 2, 2, {}, OP_CAR);
  if (name == "rfft" || name == "ifft") {   // functions.rs:1196-1217
    if (!has(2)) return EMPTY;
    uint64_t i = as_usize(p[0]);
    uint64_t x = std::min<uint64_t>(std::max<uint64_t>(i, 2), 32768), n = 1;
    while (n < x) n <<= 1;
    uint64_t start = std::min<uint64_t>(as_usize(p[1]), n - 1);
    bool r = name == "rfft";
    return single(r ? ID_RFFT : ID_IFFT, r ? NK_RFFT : NK_IFFT, r ? 1 : 2, 2, {}, 0, (int)start, (int)n);
  }
  return EMPTY;
}

}  // namespace qg
