// Graph IR: composition algebra, structural hashing and the graph-level constructors.
// Mirrors the audio-graph arms of the reference's patch interpreter, /root/reference/src/process.rs:1450-1876.
#include "graph.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "tape.h"

namespace qg {

uint64_t atto(uint64_t state, uint64_t data) {
  uint64_t r = (state << 5) | (state >> 59);
  return (r ^ data) * 0x517cc1b727220a95ULL;
}

// Rust `f32 as usize` / `as i32`: saturating, NaN -> 0.
static uint64_t as_usize(float x) {
  if (!(x > 0.0f)) return 0;
  if (x >= 18446744073709551616.0f) return UINT64_MAX;
  return (uint64_t)x;
}
static int32_t as_i32(float x) {
  if (x != x) return 0;
  if (x >= 2147483648.0f) return INT32_MAX;
  if (x <= -2147483648.0f) return INT32_MIN;
  return (int32_t)x;
}

Graph::Graph(int ni, int no) : n_in(ni) {
  for (int i = 0; i < no; i++) outs.push_back(i < ni ? Src{-1, i} : Src{-2, 0});
}

static void set_sr_node(Node& n, double sr) {
  n.sr = sr;
  if (n.kind == NK_KR && n.mode) {   // s(): inner net runs at sr / n (nodes.rs:263-269)
    for (Graph& k : n.kids) k.set_sample_rate(sr / (double)n.aux);
  } else {
    for (Graph& k : n.kids) k.set_sample_rate(sr);
  }
}
void Graph::set_sample_rate(double s) {
  sr = s;
  for (Node& n : nodes) set_sr_node(n, s);
}

uint64_t Graph::ping(bool probe, uint64_t h) {
  h = atto(h, ID_NET);
  for (Unit& u : units) {
    for (PingStep& st : u.ping) {
      switch (st.type) {
        case PingStep::MIX: h = atto(h, st.id); break;
        case PingStep::LEAF:
          if (!probe) nodes[st.node].hash = h;
          h = atto(h, st.id);
          break;
        case PingStep::KID: h = nodes[st.node].kids[0].ping(probe, h); break;
      }
    }
  }
  return h;
}
void Graph::rehash() {
  uint64_t h = ping(true, ID_NET);
  ping(false, h);
}

static Src remap(const Src& s, int off, const std::vector<Src>& gin) {
  if (s.node >= 0) return Src{s.node + off, s.port};
  if (s.node == -1) return s.port < (int)gin.size() ? gin[s.port] : Src{-2, 0};
  return s;
}
static std::vector<Src> ginputs(int from, int n) {
  std::vector<Src> g;
  for (int i = 0; i < n; i++) g.push_back(Src{-1, from + i});
  return g;
}
// append b's nodes/units to a; b's graph inputs resolve through `gin`; returns b's outputs in a's index space
static std::vector<Src> append(Graph& a, Graph& b, const std::vector<Src>& gin) {
  int off = (int)a.nodes.size();
  for (Node& n : b.nodes) {
    for (Src& s : n.in) s = remap(s, off, gin);
    a.nodes.push_back(std::move(n));
  }
  for (Unit& u : b.units) {
    for (PingStep& st : u.ping)
      if (st.type != PingStep::MIX) st.node += off;
    a.units.push_back(std::move(u));
  }
  std::vector<Src> bout;
  for (const Src& s : b.outs) bout.push_back(remap(s, off, gin));
  if (!b.unsupported.empty() && a.unsupported.empty()) a.unsupported = b.unsupported;
  return bout;
}

Graph Graph::combine(char kind, Graph a, Graph b) {
  switch (kind) {
    case '>': {
      std::vector<Src> bout = append(a, b, a.outs);
      a.outs = bout;
      break;
    }
    case '|': {
      std::vector<Src> bout = append(a, b, ginputs(a.n_in, b.n_in));
      a.n_in += b.n_in;
      a.outs.insert(a.outs.end(), bout.begin(), bout.end());
      break;
    }
    case '^': {
      std::vector<Src> bout = append(a, b, ginputs(0, b.n_in));
      a.outs.insert(a.outs.end(), bout.begin(), bout.end());
      break;
    }
    default: {
      // `&` mixes outputs over shared inputs; `+ - *` combine outputs over stacked inputs.  FunDSP's Net
      // pushes one 2-in/1-out vertex per output channel, which is why size() grows by outputs().
      bool bus = kind == '&';
      int bn = b.n_in;
      std::vector<Src> bout = append(a, b, bus ? ginputs(0, bn) : ginputs(a.n_in, bn));
      if (!bus) a.n_in += bn;
      uint16_t dev = (bus || kind == '+') ? OP_ADD : kind == '-' ? OP_SUB : OP_MUL;
      for (size_t i = 0; i < a.outs.size(); i++) {
        Node n;
        n.kind = NK_BIN;
        n.devop = dev;
        n.n_in = 2;
        n.n_out = 1;
        n.in = {a.outs[i], bout[i]};
        n.sr = a.sr;
        int idx = a.add_node(n);
        Unit u;
        u.ping.push_back(PingStep{PingStep::LEAF, ID_MAP, idx});
        a.units.push_back(u);
        a.outs[i] = Src{idx, 0};
      }
    }
  }
  a.rehash();
  return a;
}

Graph Graph::thru(Graph a) {
  int ni = a.n_in;
  if ((int)a.outs.size() > ni) a.outs.resize(ni);
  for (int i = (int)a.outs.size(); i < ni; i++) a.outs.push_back(Src{-1, i});
  a.rehash();
  return a;
}

// ------------------------------------------------------------------ graph-level constructors
static Graph wrap_node(Node n, uint64_t id, bool ping_kid = false) {
  Graph g;
  g.n_in = n.n_in;
  n.in = ginputs(0, n.n_in);
  int no = n.n_out;
  int idx = g.add_node(n);
  Unit u;
  if (ping_kid) {   // AudioUnit owners: self.x.ping(probe, hash.hash(ID))  (nodes.rs:316-318)
    u.ping.push_back(PingStep{PingStep::MIX, id, 0});
    u.ping.push_back(PingStep{PingStep::KID, 0, idx});
  } else {
    u.ping.push_back(PingStep{PingStep::LEAF, id, idx});
  }
  g.units.push_back(u);
  for (int i = 0; i < no; i++) g.outs.push_back(Src{idx, i});
  for (const Graph& k : g.nodes[idx].kids)
    if (!k.unsupported.empty()) g.unsupported = k.unsupported;
  g.rehash();
  return g;
}

Graph make_get(const std::vector<float>& arr) {   // process.rs:1455
  Node n;
  n.kind = NK_ARR_GET; n.n_in = 1; n.n_out = 1; n.table = arr;
  return wrap_node(n, ID_ARRGET);
}
Graph make_quantize(const std::vector<float>& arr) {   // process.rs:1468-1471
  if (arr.empty()) return Graph(0, 0);
  Node n;
  n.kind = NK_QUANTIZE; n.n_in = 1; n.n_out = 1; n.table = arr;
  n.raw = {arr.back() - arr.front()};
  return wrap_node(n, ID_QUANTIZER);
}
Graph make_wave(const std::vector<float>& arr) {   // process.rs:1658-1662
  Node n;
  n.kind = NK_WAVE; n.n_in = 0; n.n_out = 1; n.table = arr;
  return wrap_node(n, ID_WAVE);
}
Graph make_feedback(const Graph& net, bool has_delay, double delay) {   // process.rs:1497-1510
  if (net.outputs() != net.inputs()) return Graph(0, 0);
  Node n;
  n.kind = NK_FEEDBACK; n.n_in = net.inputs(); n.n_out = net.outputs();
  float d = has_delay ? (float)delay : 0.0f;   // `del.into()`: f32 -> f64
  n.raw = {d};
  n.kids.push_back(net);
  return wrap_node(n, ID_FEEDBACK, true);
}
Graph make_kr(const Graph& net, double num, bool preserve_time) {   // process.rs:1562-1566
  Node n;
  n.kind = NK_KR; n.n_in = net.inputs(); n.n_out = net.outputs();
  float nf = (float)num;
  uint64_t period = as_usize(std::fmax(nf, 1.0f));
  if (period > 0x7fffffffULL) period = 0x7fffffffULL;
  n.aux = (int)period;
  n.mode = preserve_time ? 1 : 0;
  n.kids.push_back(net);
  return wrap_node(n, ID_KR, true);
}
Graph make_reset(const Graph& net, double s) {   // process.rs:1568-1569
  if (!(net.inputs() == 0 && net.outputs() == 1)) return Graph(0, 0);
  Node n;
  n.kind = NK_RESET; n.mode = 0; n.n_in = 0; n.n_out = 1;
  n.raw = {(float)s};
  n.kids.push_back(net);
  return wrap_node(n, ID_RESET);
}
Graph make_trig_reset(const Graph& net, bool variable) {   // process.rs:1600-1606
  if (!(net.inputs() == 0 && net.outputs() == 1)) return Graph(0, 0);
  Node n;
  n.kind = NK_RESET; n.mode = variable ? 2 : 1; n.n_in = 1; n.n_out = 1;
  n.kids.push_back(net);
  return wrap_node(n, variable ? ID_RESETV : ID_TRIGRESET);
}
Graph make_seq_select(bool is_seq, const std::vector<const Graph*>& nets) {   // process.rs:1635-1647
  Node n;
  n.kind = is_seq ? NK_SEQ : NK_SELECT;
  n.n_in = is_seq ? 4 : 1;
  n.n_out = 1;
  std::string unsup;
  for (const Graph* g : nets) {
    if (g && !g->unsupported.empty() && unsup.empty()) unsup = g->unsupported;
    if (g && g->inputs() == 0 && g->outputs() == 1) n.kids.push_back(*g);
  }
  Graph r = wrap_node(n, is_seq ? ID_SEQ : ID_SELECT);
  if (r.unsupported.empty()) r.unsupported = unsup;
  return r;
}
Graph make_var(float value) {   // var(): a Shared-backed constant (process.rs:1373-1385), value = the circle's Number
  Node n;
  n.kind = NK_CONST; n.n_in = 0; n.n_out = 1; n.raw = {value};
  return wrap_node(n, ID_CONSTANT);
}
Graph make_live_io(const std::string& name) {
  // InputNode / BuffOut: every try_recv() misses offline -> 0.0 (nodes.rs:516-517, 786); BuffIn passes (nodes.rs:761-762)
  Node n;
  if (name == "in()" || name == "adc()") { n.kind = NK_ZERO_SRC; n.n_out = 2; return wrap_node(n, ID_INPUT); }
  if (name == "buffout()") { n.kind = NK_ZERO_SRC; n.n_out = 1; return wrap_node(n, ID_BUFFOUT); }
  if (name == "buffin()") { n.kind = NK_PASS; n.n_in = 1; n.n_out = 1; return wrap_node(n, ID_MAP); }
  // monitor(&s, Meter::Sample) passes its input through while publishing it (process.rs:1404); timer(&s) has no
  // inputs or outputs (process.rs:1406): offline they reduce to a pass-through / an empty net
  if (name == "monitor()") { n.kind = NK_PASS; n.n_in = 1; n.n_out = 1; return wrap_node(n, ID_PASS); }
  return Graph(0, 0);
}

// ------------------------------------------------------------------ connective circles
Graph connect(const std::string& op, const std::vector<const Graph*>& nets, double number, int node_limit) {
  if (op == "!") {   // process.rs:1868-1873
    Graph g = (!nets.empty() && nets[0]) ? *nets[0] : Graph(0, 0);
    std::string u = g.unsupported;
    Graph r = Graph::thru(g);
    if (r.unsupported.empty()) r.unsupported = u;
    return r;
  }
  if (op == "-") {   // process.rs:1787-1797
    std::string u;
    for (const Graph* np : nets) if (np && !np->unsupported.empty() && u.empty()) u = np->unsupported;
    if (nets.size() >= 2 && nets[0] && nets[1] && nets[0]->outputs() == nets[1]->outputs()) {
      Graph g = Graph::combine('-', *nets[0], *nets[1]);
      if (g.unsupported.empty()) g.unsupported = u;
      if (g.size() < node_limit) return g;
    }
    Graph e(0, 0);
    e.unsupported = u;
    return e;
  }
  Graph graph(0, 0);
  bool empty = true;
  // an input that mentions an op without a GPU lowering marks the result even when an arity guard skips it
  std::string unsup;
  for (const Graph* np : nets) if (np && !np->unsupported.empty() && unsup.empty()) unsup = np->unsupported;
  int reps = as_i32(std::fmax((float)number, 1.0f));   // `.max(1.) as i32` (process.rs:1744, 1825)
  // a tape addresses at most 2^16 words, so a graph beyond 2^16 vertices can never be lowered: a caller-supplied limit above
  // that only decides how much memory the composition burns before lowering refuses it
  node_limit = std::min(node_limit, 1 << 16);
  for (int r = 0; r < reps; r++) {
    // a pass that combines nothing (arity guards, node limit) leaves `graph` as it was, so every later pass would do the
    // same: stop instead of spinning through up to 2^31 repetitions (the result is the reference's, without its wait)
    bool changed = false;
    // ... and so does a pass over operands of size 0 without ports (a failed str_to_net, an empty connective): combining
    // them succeeds without growing anything, the node limit is never reached
    const int size0 = graph.size(), in0 = graph.inputs(), out0 = graph.outputs();
    const bool was_empty = empty;
    for (const Graph* np : nets) {
      if (!np) continue;
      if (empty) { graph = *np; empty = false; changed = true; continue; }
      int gi = graph.inputs(), go = graph.outputs(), ni = np->inputs(), no = np->outputs();
      if (op == "+" || op == "*") {   // process.rs:1751-1759
        if (go == no) {
          if (graph.size() >= node_limit) continue;
          graph = Graph::combine(op[0], std::move(graph), *np);
          changed = true;
        }
      } else {   // process.rs:1833-1844
        if (graph.size() >= node_limit) continue;
        if (op == ">>") { if (go == ni) { graph = Graph::combine('>', std::move(graph), *np); changed = true; } }
        else if (op == "|") { graph = Graph::combine('|', std::move(graph), *np); changed = true; }
        else if (op == "&") { if (gi == ni && go == no) { graph = Graph::combine('&', std::move(graph), *np); changed = true; } }
        else if (op == "^") { if (gi == ni) { graph = Graph::combine('^', std::move(graph), *np); changed = true; } }
      }
    }
    if (!changed) break;
    if (!was_empty && graph.size() == size0 && graph.inputs() == in0 && graph.outputs() == out0) break;
  }
  if (graph.unsupported.empty()) graph.unsupported = unsup;
  return graph;
}

// Rust `format!("{}", f32)`: shortest decimal that round-trips, never an exponent.
static std::string fmt_f32(float v) {
  char buf[512];
  if (v != v) return "NaN";                       // Rust's Display for f32
  if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
  for (int prec = 1; prec < 12; prec++) {
    snprintf(buf, sizeof buf, "%.*g", prec, (double)v);
    if (strtof(buf, nullptr) == v) break;
  }
  std::string s(buf);
  if (s.find('e') != std::string::npos) {
    snprintf(buf, sizeof buf, "%.60f", (double)v);
    s = buf;
    // trim to the shortest prefix that still round-trips
    size_t dot = s.find('.');
    for (size_t len = dot + 2; len <= s.size(); len++) {
      if (strtof(s.substr(0, len).c_str(), nullptr) == v) { s = s.substr(0, len); break; }
    }
  }
  return s;
}

Graph array_op(const std::string& kind, const std::string& op_str, const std::vector<float>& arr) {   // process.rs:1689-1714
  Graph graph(0, 0);
  bool empty = true;
  for (float v : arr) {
    std::string r, num = fmt_f32(v);
    for (char c : op_str) { if (c == '#') r += num; else r.push_back(c); }
    Graph net = str_to_net(r);
    if (empty) { graph = net; empty = false; continue; }
    int gi = graph.inputs(), go = graph.outputs(), ni = net.inputs(), no = net.outputs();
    if (kind == "branch()") { if (gi == ni) graph = Graph::combine('^', std::move(graph), net); }
    else if (kind == "bus()") { if (gi == ni && go == no) graph = Graph::combine('&', std::move(graph), net); }
    else if (kind == "pipe()") { if (go == ni) graph = Graph::combine('>', std::move(graph), net); }
    else if (kind == "stack()") graph = Graph::combine('|', std::move(graph), net);
    else if (kind == "sum()") { if (go == no) graph = Graph::combine('+', std::move(graph), net); }
    else if (kind == "product()") { if (go == no) graph = Graph::combine('*', std::move(graph), net); }
    if (!net.unsupported.empty() && graph.unsupported.empty()) graph.unsupported = net.unsupported;
  }
  return graph;
}

}  // namespace qg
