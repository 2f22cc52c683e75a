// TEST INFRASTRUCTURE — CPU oracle (see qo_math.h header).  Graph containers: static combinators used inside
// str_to_net (FunDSP `An<Pipe<..>>` style graphs count as ONE vertex of a Net), the dynamic `Net`
// (vertex list + edges, FunDSP net.rs semantics [U]) and the units that own nested nets
// (/root/reference/src/nodes.rs Select/Seq/Kr/Reset/TrigReset/ResetV [P], FunDSP FeedbackUnit [U]).
#pragma once
#include "qo_units.h"

namespace qo {

// static binary combinators: kind 0 pipe, 1 stack, 2 branch, 3 bus, 4 add, 5 sub, 6 mul
struct Comp : Unit {
  int kind;
  UnitP x, y;
  std::vector<float> tx, ty;
  Comp(int k, UnitP a, UnitP b) : kind(k), x(std::move(a)), y(std::move(b)) {
    tx.resize(std::max(1, x->outs()));
    ty.resize(std::max(1, y->outs()));
  }
  Comp(const Comp& o) : kind(o.kind), x(o.x->clone()), y(o.y->clone()), tx(o.tx), ty(o.ty) {}
  int ins() const override {
    switch (kind) {
      case 0: return x->ins();
      case 2: case 3: return x->ins();
      default: return x->ins() + y->ins();
    }
  }
  int outs() const override {
    switch (kind) {
      case 0: return y->outs();
      case 1: case 2: return x->outs() + y->outs();
      default: return x->outs();
    }
  }
  void reset() override { x->reset(); y->reset(); }
  void set_sr(double sr) override { x->set_sr(sr); y->set_sr(sr); }
  void tick(const float* in, float* out) override {
    switch (kind) {
      case 0: x->tick(in, tx.data()); y->tick(tx.data(), out); break;
      case 1: x->tick(in, out); y->tick(in + x->ins(), out + x->outs()); break;
      case 2: x->tick(in, out); y->tick(in, out + x->outs()); break;
      default: {
        x->tick(in, tx.data());
        y->tick(kind == 3 ? in : in + x->ins(), ty.data());
        int n = x->outs();
        for (int i = 0; i < n; i++)
          out[i] = kind == 5 ? tx[i] - ty[i] : kind == 6 ? tx[i] * ty[i] : tx[i] + ty[i];
      }
    }
  }
  uint64_t id() const override {
    static const uint64_t ids[7] = {ID_PIPE, ID_STACK, ID_BRANCH, ID_BUS, ID_BINOP, ID_BINOP, ID_BINOP};
    return ids[kind];
  }
  uint64_t ping(bool probe, uint64_t h) override { return y->ping(probe, x->ping(probe, atto(h, id()))); }
  void salt(uint64_t s) override { x->salt(s); y->salt(s); }
  QO_CLONE(Comp)
};

struct Net : Unit {
  struct Src { int v, port; };   // v >= 0 vertex output port; v == -1 global input `port`; v == -2 constant zero
  struct Vertex {
    UnitP u;
    std::vector<Src> in;
    std::vector<float> ibuf, obuf;
  };
  std::vector<Vertex> vs;
  std::vector<Src> osrc;
  int nin = 0;
  double sr = DEFAULT_SR;

  Net() {}
  Net(int ni, int no) : nin(ni) {
    // Net::new(i, o): outputs fed from inputs where available, otherwise zero
    for (int i = 0; i < no; i++) osrc.push_back(i < ni ? Src{-1, i} : Src{-2, 0});
  }
  Net(const Net& o) : osrc(o.osrc), nin(o.nin), sr(o.sr) {
    for (const Vertex& v : o.vs) {
      Vertex w;
      w.u = v.u->clone();
      w.in = v.in;
      w.ibuf = v.ibuf;
      w.obuf = v.obuf;
      vs.push_back(std::move(w));
    }
  }
  static std::unique_ptr<Net> wrap(UnitP u) {
    std::unique_ptr<Net> n(new Net());
    n->nin = u->ins();
    int no = u->outs();
    n->push(std::move(u));
    for (int i = 0; i < n->nin; i++) n->vs[0].in[i] = Src{-1, i};
    for (int i = 0; i < no; i++) n->osrc.push_back(Src{0, i});
    n->rehash();
    return n;
  }
  int push(UnitP u) {
    Vertex v;
    v.in.assign(u->ins(), Src{-2, 0});
    v.ibuf.assign(std::max(1, u->ins()), 0.0f);
    v.obuf.assign(std::max(1, u->outs()), 0.0f);
    v.u = std::move(u);
    vs.push_back(std::move(v));
    return (int)vs.size() - 1;
  }
  int size() const { return (int)vs.size(); }
  int ins() const override { return nin; }
  int outs() const override { return (int)osrc.size(); }
  void reset() override { for (Vertex& v : vs) v.u->reset(); }
  void set_sr(double s) override { sr = s; for (Vertex& v : vs) v.u->set_sr(s); }
  float fetch(const Src& s, const float* in) const {
    if (s.v >= 0) return vs[s.v].obuf[s.port];
    if (s.v == -1) return in[s.port];
    return 0.0f;
  }
  void tick(const float* in, float* out) override {
    for (Vertex& v : vs) {
      for (size_t i = 0; i < v.in.size(); i++) v.ibuf[i] = fetch(v.in[i], in);
      v.u->tick(v.ibuf.data(), v.obuf.data());
    }
    for (size_t i = 0; i < osrc.size(); i++) out[i] = fetch(osrc[i], in);
  }
  uint64_t id() const override { return ID_NET; }
  uint64_t ping(bool probe, uint64_t h) override {
    h = atto(h, ID_NET);
    for (Vertex& v : vs) h = v.u->ping(probe, h);
    return h;
  }
  // "the hash depends on vertices but not edges": probe pass, then the committing pass starts from its result
  void rehash() {
    uint64_t h = ping(true, ID_NET);
    ping(false, h);
  }
  void salt(uint64_t s) override { for (Vertex& v : vs) v.u->salt(s); }
  UnitP clone() const override { return UnitP(new Net(*this)); }

  // ---- Net algebra (operators on Net in FunDSP; arity guards live in the callers, process.rs)
  // remap a Src of `b` after b's vertices were appended at offset `off`; global inputs of b resolve via `gin`
  static Src remap(const Src& s, int off, const std::vector<Src>& gin) {
    if (s.v >= 0) return Src{s.v + off, s.port};
    if (s.v == -1) return s.port < (int)gin.size() ? gin[s.port] : Src{-2, 0};
    return s;
  }
  void append(Net& b, const std::vector<Src>& gin, std::vector<Src>& bout) {
    int off = (int)vs.size();
    for (Vertex& v : b.vs) {
      for (Src& s : v.in) s = remap(s, off, gin);
      vs.push_back(std::move(v));
    }
    bout.clear();
    for (const Src& s : b.osrc) bout.push_back(remap(s, off, gin));
  }
  static std::vector<Src> ginputs(int from, int n) {
    std::vector<Src> g;
    for (int i = 0; i < n; i++) g.push_back(Src{-1, from + i});
    return g;
  }
  // kind: '>' pipe, '|' stack, '&' bus, '^' branch, '+', '*', '-' binops
  static std::unique_ptr<Net> combine(char kind, std::unique_ptr<Net> a, std::unique_ptr<Net> b) {
    std::vector<Src> bout;
    switch (kind) {
      case '>': a->append(*b, a->osrc, bout); a->osrc = bout; break;
      case '|': a->append(*b, ginputs(a->nin, b->nin), bout); a->nin += b->nin;
                a->osrc.insert(a->osrc.end(), bout.begin(), bout.end()); break;
      case '^': a->append(*b, ginputs(0, b->nin), bout);
                a->osrc.insert(a->osrc.end(), bout.begin(), bout.end()); break;
      default: {
        bool bus = kind == '&';
        a->append(*b, bus ? ginputs(0, b->nin) : ginputs(a->nin, b->nin), bout);
        if (!bus) a->nin += b->nin;
        char k = bus ? '+' : kind;
        for (size_t i = 0; i < a->osrc.size(); i++) {
          int v = a->push(UnitP(new Map(2, 1, [k](const float* in, float* out) {
            out[0] = k == '+' ? in[0] + in[1] : k == '-' ? in[0] - in[1] : in[0] * in[1];
          })));
          a->vs[v].in[0] = a->osrc[i];
          a->vs[v].in[1] = bout[i];
          a->osrc[i] = Src{v, 0};
        }
      }
    }
    a->rehash();
    return a;
  }
  // !net: pass missing outputs through from the inputs, cut surplus outputs
  static std::unique_ptr<Net> thru(std::unique_ptr<Net> a) {
    int ni = a->nin;
    a->osrc.resize(std::min<size_t>(a->osrc.size(), ni));
    for (int i = (int)a->osrc.size(); i < ni; i++) a->osrc.push_back(Src{-1, i});
    a->rehash();
    return a;
  }
};
typedef std::unique_ptr<Net> NetP;

// ------------------------------------------------------------------ units that own nets
struct Kr : Unit {   // nodes.rs:235-327
  NetP x;
  size_t n, count = 0;
  std::vector<float> vals;
  bool preserve_time;
  Kr(NetP x_, size_t n_, bool pt) : x(std::move(x_)), n(n_), vals(std::max(1, x->outs()), 0.0f), preserve_time(pt) {}
  Kr(const Kr& o) : x(new Net(*o.x)), n(o.n), count(o.count), vals(o.vals), preserve_time(o.preserve_time) {}
  int ins() const override { return x->ins(); }
  int outs() const override { return x->outs(); }
  void reset() override { x->reset(); count = 0; std::fill(vals.begin(), vals.end(), 0.0f); }
  void set_sr(double sr) override { x->set_sr(preserve_time ? sr / (double)n : sr); }
  void tick(const float* in, float* out) override {
    if (count == 0) { count = n; x->tick(in, vals.data()); }
    count -= 1;
    for (int i = 0; i < x->outs(); i++) out[i] = vals[i];
  }
  uint64_t id() const override { return ID_KR; }
  uint64_t ping(bool probe, uint64_t h) override { return x->ping(probe, atto(h, ID_KR)); }
  void salt(uint64_t s) override { x->salt(s); }
  QO_CLONE(Kr)
};
struct Feedback : Unit {   // FunDSP FeedbackUnit [U]: out = x(in + out delayed by max(1, round(delay·sr)))
  NetP x;
  double delay;
  int ch;
  size_t samples, idx = 0;
  std::vector<std::vector<float>> buf;
  std::vector<float> tb;
  Feedback(double d, NetP x_) : x(std::move(x_)), delay(d), ch(x->outs()), tb(std::max(1, x->outs())) { set_len(DEFAULT_SR); }
  Feedback(const Feedback& o) : x(new Net(*o.x)), delay(o.delay), ch(o.ch), samples(o.samples), idx(o.idx), buf(o.buf), tb(o.tb) {}
  void set_len(double sr) {
    double s = std::round(delay * sr);
    samples = !(s >= 1.0) ? 1 : (s > 4.0e8 ? (size_t)400000000 : (size_t)s);   // NaN / negative / zero: one sample
    buf.assign(ch, std::vector<float>(samples, 0.0f));
    idx = 0;
  }
  int ins() const override { return ch; }
  int outs() const override { return ch; }
  void reset() override { for (auto& b : buf) std::fill(b.begin(), b.end(), 0.0f); idx = 0; x->reset(); }
  void set_sr(double sr) override { set_len(sr); x->set_sr(sr); }
  void tick(const float* in, float* out) override {
    for (int c = 0; c < ch; c++) tb[c] = in[c] + buf[c][idx];
    x->tick(tb.data(), out);
    for (int c = 0; c < ch; c++) buf[c][idx] = out[c];
    idx = idx + 1 == samples ? 0 : idx + 1;
  }
  uint64_t id() const override { return ID_FEEDBACK; }
  uint64_t ping(bool probe, uint64_t h) override { return x->ping(probe, atto(h, ID_FEEDBACK)); }
  void salt(uint64_t s) override { x->salt(s); }
  QO_CLONE(Feedback)
};
static inline std::vector<NetP> clone_nets(const std::vector<NetP>& v) {
  std::vector<NetP> r;
  for (const NetP& n : v) r.push_back(NetP(new Net(*n)));
  return r;
}
struct Select : Unit {   // nodes.rs:10-46
  std::vector<NetP> nets;
  explicit Select(std::vector<NetP> n) : nets(std::move(n)) {}
  Select(const Select& o) : nets(clone_nets(o.nets)) {}
  int ins() const override { return 1; }
  int outs() const override { return 1; }
  void reset() override { for (auto& n : nets) n->reset(); }
  void set_sr(double sr) override { for (auto& n : nets) n->set_sr(sr); }
  void tick(const float* in, float* out) override {
    float b = 0.0f;
    uint64_t i = as_usize(in[0]);
    if (i < nets.size()) nets[i]->tick(nullptr, &b);
    out[0] = b;
  }
  uint64_t id() const override { return ID_SELECT; }
  void salt(uint64_t s) override { for (auto& n : nets) n->salt(s); }
  QO_CLONE(Select)
};
struct Seq : Unit {   // nodes.rs:55-121
  struct Ev { uint64_t idx, delay, dur; };
  std::vector<NetP> nets;
  std::vector<Ev> events;
  float sr = 44100.f;
  explicit Seq(std::vector<NetP> n) : nets(std::move(n)) {}
  Seq(const Seq& o) : nets(clone_nets(o.nets)), events(o.events), sr(o.sr) {}
  int ins() const override { return 4; }
  int outs() const override { return 1; }
  void reset() override { for (auto& n : nets) n->reset(); }   // note: events survive a reset (nodes.rs:116-120)
  void set_sr(double s) override { sr = (float)s; for (auto& n : nets) n->set_sr(s); }
  void tick(const float* in, float* out) override {
    if (in[0] != 0.0f) {
      uint64_t k = as_usize(in[1]);
      events.erase(std::remove_if(events.begin(), events.end(), [k](const Ev& e) { return e.idx == k; }), events.end());
      if (k < nets.size()) nets[k]->reset();
      events.push_back(Ev{k, as_usize(std::round(in[2] * sr)), as_usize(std::round(in[3] * sr))});
    }
    events.erase(std::remove_if(events.begin(), events.end(), [](const Ev& e) { return e.dur == 0; }), events.end());
    float o = 0.0f;
    for (Ev& e : events) {
      if (e.delay == 0) {
        if (e.idx < nets.size()) {
          float b = 0.0f;
          nets[e.idx]->tick(nullptr, &b);
          o += b;
        }
        e.dur -= 1;
      } else e.delay -= 1;
    }
    out[0] = o;
  }
  uint64_t id() const override { return ID_SEQ; }
  void salt(uint64_t s) override { for (auto& n : nets) n->salt(s); }
  QO_CLONE(Seq)
};
// kind 0 Reset(net, s) nodes.rs:332-371; 1 TrigReset nodes.rs:377-409; 2 ResetV nodes.rs:415-453
struct Resetter : Unit {
  int kind;
  NetP net;
  float dur = 0, sr = 44100.f;
  uint64_t n = 0, count = 0;
  Resetter(int k, NetP x, float s) : kind(k), net(std::move(x)), dur(s) {
    n = as_usize(std::round(s * 44100.0f));
  }
  Resetter(const Resetter& o) : kind(o.kind), net(new Net(*o.net)), dur(o.dur), sr(o.sr), n(o.n), count(o.count) {}
  int ins() const override { return kind == 0 ? 0 : 1; }
  int outs() const override { return 1; }
  void reset() override { count = 0; net->reset(); }
  void set_sr(double s) override {
    if (kind == 0) n = as_usize(std::round(dur * (float)s));
    if (kind == 2) sr = (float)s;
    net->set_sr(s);
  }
  void tick(const float* in, float* out) override {
    if (kind == 0) {
      if (count >= n) { net->reset(); count = 0; }
    } else if (kind == 1) {
      if (in[0] != 0.0f) net->reset();
    } else {
      if (count >= as_usize(std::round(in[0] * sr))) { net->reset(); count = 0; }
    }
    float b = 0.0f;
    net->tick(nullptr, &b);
    count += 1;
    out[0] = b;
  }
  uint64_t id() const override { return kind == 0 ? ID_RESET : kind == 1 ? ID_TRIGRESET : ID_RESETV; }
  void salt(uint64_t s) override { net->salt(s); }
  QO_CLONE(Resetter)
};

NetP str_to_net(const std::string& op, int* status);   // status: 0 ok, 1 op exists in the reference but is not restated

}  // namespace qo
