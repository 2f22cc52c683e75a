// Graph -> op tape.  This is the "op-string to GPU op-tape lowering" that sits beside the reference's
// str_to_net (/root/reference/src/functions.rs:111): every primitive node becomes one or a few fixed-size
// instructions in evaluation order; nested nets (nodes.rs Kr/Select/Reset*/FeedbackUnit) are inlined as
// predicated / jumped-over instruction ranges.
#include "lower.h"

#include <mutex>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "coefs.h"

namespace qg {

namespace {

const uint16_t NONE = 0xFFFF;
const uint16_t R_STATE = 0x4000, R_TEMP = 0x8000;   // region tags, rewritten to absolute X indices at the end

uint64_t as_usize(float x) {
  if (!(x > 0.0f)) return 0;
  if (x >= 18446744073709551616.0f) return UINT64_MAX;
  return (uint64_t)x;
}
uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// ---- band-limited wavetables (FunDSP wavetable.rs, restated): one table per quarter octave, 20 Hz .. 20 kHz
void host_ifft(std::vector<float>& re, std::vector<float>& im) {   // in-place radix-2 inverse DFT, unscaled
  size_t n = re.size();
  for (size_t i = 1, j = 0; i < n; i++) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    for (size_t k = 0; k < len / 2; k++) {
      double ang = 2.0 * 3.14159265358979323846 * (double)k / (double)len;
      float wr = (float)std::cos(ang), wi = (float)std::sin(ang);
      for (size_t i = k; i < n; i += len) {
        float ur = re[i], ui = im[i], vr = re[i + len / 2], vi = im[i + len / 2];
        float tr = vr * wr - vi * wi, ti = vr * wi + vi * wr;
        re[i] = ur + tr; im[i] = ui + ti;
        re[i + len / 2] = ur - tr; im[i + len / 2] = ui - ti;
      }
    }
  }
}
std::vector<float> make_wave(double pitch, int shape) {
  size_t harmonics = (size_t)std::floor(22050.0 / pitch);
  size_t target = 4 * harmonics, length = 32;
  while (length < target && length < 8192) length <<= 1;
  std::vector<float> re(length, 0.0f), im(length, 0.0f);
  for (size_t i = 1; i <= harmonics && i < length / 2; i++) {
    double f = pitch * (double)i, w;
    bool odd = (i & 1) != 0;
    switch (shape) {
      case 0: w = 1.0 / (double)i; break;                               // saw
      case 1: w = odd ? 1.0 / (double)i : 0.0; break;                   // square
      case 2: w = odd ? 1.0 / ((double)i * (double)i) : 0.0; break;     // triangle
      default: w = 1.0 / ((double)i * (double)i); break;                // soft saw
    }
    double fade = (22050.0 - f) / (22050.0 - 20000.0);
    fade = fade < 0 ? 0 : (fade > 1 ? 1 : fade);
    w *= ((6.0 * fade - 15.0) * fade + 10.0) * fade * fade * fade;
    double ph = (shape == 2 && (i & 3) == 3) ? 0.5 : 0.0;
    double r = 0.5 * w * std::sin(2.0 * 3.14159265358979323846 * ph), m = -0.5 * w * std::cos(2.0 * 3.14159265358979323846 * ph);
    re[i] = (float)r; im[i] = (float)m;
    re[length - i] = (float)r; im[length - i] = (float)-m;
  }
  host_ifft(re, im);
  float mx = 0.0f;
  for (float x : re) mx = std::fmax(mx, std::fabs(x));
  if (mx > 0.0f) for (float& x : re) x /= mx;
  return re;
}
// table-set blob: [n][n x (limit, offset(bits), length(bits))][samples...]; offsets are relative to the blob start
const std::vector<float>& wavetable_blob(int shape) {
  // built once per shape; banks may be created from several host threads (one per GPU), so the lazy build is guarded
  static std::vector<float> blobs[4];
  static std::once_flag once[4];
  std::vector<float>& b = blobs[shape & 3];
  std::call_once(once[shape & 3], [&b, shape] {
    std::vector<float> limits;
    std::vector<std::vector<float>> tabs;
    for (int i = 0;; i++) {
      double pn = 20.0 * std::pow(2.0, (double)(i + 1) / 4.0);
      limits.push_back((float)pn);
      tabs.push_back(make_wave(pn, shape));
      if (pn >= 20000.0) break;
    }
    size_t n = tabs.size(), off = 1 + 3 * n;
    b.push_back((float)n);
    for (size_t i = 0; i < n; i++) {
      b.push_back(limits[i]);
      uint32_t o = (uint32_t)off, l = (uint32_t)tabs[i].size();
      float fo, fl;
      memcpy(&fo, &o, 4); memcpy(&fl, &l, 4);
      b.push_back(fo); b.push_back(fl);
      off += tabs[i].size();
    }
    for (auto& tb : tabs) b.insert(b.end(), tb.begin(), tb.end());
  });
  return b;
}

struct Lower {
  Tape& t;
  std::string err;
  int n_temps = 0;
  int zero_p = -1;
  int wt_offset[4] = {-1, -1, -1, -1};   // wavetable sets already placed in the tables region
  long tw_offset[32];                    // twiddle tables already placed, by log2 of the transform size
  explicit Lower(Tape& t_) : t(t_) { for (long& o : tw_offset) o = -1; }

  // 16-bit operand indices: 15 bits of temporaries, 14 of parameters and of state.  Checked at allocation — an index that
  // wrapped would alias an earlier word and the tape would render wrong audio instead of being refused (the temporary-reuse
  // pass compacts indices later, from whatever they already are).
  void too_large() { if (err.empty()) err = "graph too large for one tape (parameter/state/temporary index overflow)"; }
  uint16_t temp(int n = 1) {
    if (n_temps + n > 0x7fff) { too_large(); return (uint16_t)R_TEMP; }
    int i = n_temps; n_temps += n; return (uint16_t)(R_TEMP | i);
  }
  uint16_t param(float v) { return params(1, v); }
  uint16_t params(int n, float v = 0.0f) {
    if (t.params.size() + (size_t)n > 0x3fff) { too_large(); return 0; }
    uint16_t b = (uint16_t)t.params.size(); t.params.resize(t.params.size() + n, v); return b;
  }
  uint16_t state(int n = 1, uint32_t init = 0) {
    if (t.state_init.size() + (size_t)n > 0x3fff) { too_large(); return (uint16_t)R_STATE; }
    uint16_t b = (uint16_t)t.state_init.size();
    t.state_init.resize(t.state_init.size() + n, init);
    t.state_keep.resize(t.state_init.size(), 0);
    return (uint16_t)(R_STATE | b);
  }
  uint16_t zero() { if (zero_p < 0) zero_p = param(0.0f); return (uint16_t)zero_p; }
  uint32_t ring(uint32_t len) {
    if (t.rings.size() >= 0xffff) { too_large(); return 0; }     // reset ranges address rings with 16 bits
    Ring r;
    r.offset = t.h.ring_floats;
    r.length = len;
    t.h.ring_floats += len;
    t.rings.push_back(r);
    return (uint32_t)t.rings.size() - 1;
  }
  uint32_t table(const std::vector<float>& v) {
    uint32_t off = (uint32_t)t.tables.size();
    t.tables.insert(t.tables.end(), v.begin(), v.end());
    return off;
  }
  // raw parameters of a node enter the tape-wide raw vector; returns their base index
  uint32_t raw(const Node& n, bool structural = false) {
    uint32_t b = (uint32_t)t.raw.size();
    for (float v : n.raw) { t.raw.push_back(v); t.raw_structural.push_back(structural ? 1 : 0); }
    return b;
  }
  void deriver(uint32_t kind, uint32_t raw_base, uint32_t n_raw, uint32_t p_base, uint32_t n_p, int mode, int aux, double sr) {
    Deriver d{kind, raw_base, n_raw, p_base, n_p, mode, aux, (float)sr};
    t.derivers.push_back(d);
  }
  Instr& emit(uint16_t op) {
    Instr i;
    memset(&i, 0, sizeof i);
    i.op = op;
    i.out = NONE; i.p = NONE; i.s = NONE;
    for (uint16_t& x : i.in) x = NONE;
    t.code.push_back(i);
    return t.code.back();
  }
  void hash_init(uint16_t s, uint32_t kind, uint64_t hash) {
    HashInit h;
    h.state = (uint32_t)(s & 0x3fff);
    h.kind = kind;
    h.hash = hash;
    t.hash_init.push_back(h);
  }

  std::vector<uint16_t> graph(const Graph& g, const std::vector<uint16_t>& inputs);
  std::vector<uint16_t> node(const Node& n, const std::vector<uint16_t>& in);
};

std::vector<uint16_t> Lower::graph(const Graph& g, const std::vector<uint16_t>& inputs) {
  if (!g.unsupported.empty() && err.empty())
    err = "op '" + g.unsupported + "' exists in the reference but has no GPU lowering yet (FunDSP source unavailable)";
  std::vector<std::vector<uint16_t>> vals(g.nodes.size());
  auto resolve = [&](const Src& s) -> uint16_t {
    if (s.node >= 0) return vals[s.node][s.port];
    if (s.node == -1 && s.port < (int)inputs.size()) return inputs[s.port];
    return zero();
  };
  for (size_t i = 0; i < g.nodes.size(); i++) {
    const Node& n = g.nodes[i];
    std::vector<uint16_t> in;
    for (const Src& s : n.in) in.push_back(resolve(s));
    vals[i] = node(n, in);
    if (!err.empty()) return {};
  }
  std::vector<uint16_t> outs;
  for (const Src& s : g.outs) outs.push_back(resolve(s));
  return outs;
}

std::vector<uint16_t> Lower::node(const Node& n, const std::vector<uint16_t>& in) {
  std::vector<uint16_t> out;
  switch (n.kind) {
    case NK_CONST: {
      uint32_t rb = raw(n);
      uint16_t p = params((int)n.raw.size());
      deriver(D_COPY, rb, (uint32_t)n.raw.size(), p, (uint32_t)n.raw.size(), 0, 0, n.sr);
      for (size_t i = 0; i < n.raw.size(); i++) out.push_back((uint16_t)(p + i));
      break;
    }
    case NK_PASS: out.push_back(in[0]); break;
    case NK_SINK: break;
    case NK_SPLIT: for (int i = 0; i < n.n_out; i++) out.push_back(in[0]); break;
    case NK_REVERSE: for (int i = 0; i < n.n_out; i++) out.push_back(in[n.n_in - 1 - i]); break;
    case NK_ZERO_SRC: for (int i = 0; i < n.n_out; i++) out.push_back(zero()); break;
    case NK_JOIN: {
      // left-to-right sum, five operands per instruction; the last instruction divides by n
      uint16_t acc = NONE;
      int i = 0;
      do {
        uint16_t ops[5];
        int k = 0;
        if (acc != NONE) ops[k++] = acc;
        while (k < 5 && i < n.n_in) ops[k++] = in[i++];
        uint16_t o = temp();
        Instr& a = emit(OP_JOIN);
        for (int q = 0; q < k; q++) a.in[q] = ops[q];
        a.n = (uint16_t)k; a.out = o;
        a.aux = i >= n.n_in ? (uint32_t)n.n_in : 0u;
        acc = o;
      } while (i < n.n_in);
      out.push_back(acc);
      break;
    }
    case NK_PAN: {
      if (n.n_in == 1) {
        uint32_t rb = raw(n);
        uint16_t p = params(2);
        deriver(D_PAN, rb, 1, p, 2, 0, 0, n.sr);
        Instr& i = emit(OP_PAN);
        i.in[0] = in[0]; i.p = p; i.out = temp(2);
        out = {i.out, (uint16_t)(i.out + 1)};
      } else {
        uint16_t s = state(3);
        t.state_init[s & 0x3fff] = fbits(std::numeric_limits<float>::quiet_NaN());   // cached pan: force first update
        Instr& i = emit(OP_PAN_VAR);
        i.in[0] = in[0]; i.in[1] = in[1]; i.s = s; i.out = temp(2);
        out = {i.out, (uint16_t)(i.out + 1)};
      }
      break;
    }
    case NK_UNARY: {
      Instr& i = emit(n.devop);
      i.in[0] = in[0]; i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_BIN: {
      uint16_t b;
      if (n.n_in == 1) {
        uint32_t rb = raw(n);
        b = params(1);
        deriver(D_COPY, rb, 1, b, 1, 0, 0, n.sr);
      } else b = in[1];
      Instr& i = emit(n.devop);
      i.in[0] = in[0]; i.in[1] = b; i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_TERN: {
      Instr* ip;
      if (n.n_in == 1) {
        uint32_t rb = raw(n);
        uint16_t p = params(2);
        deriver(D_COPY, rb, 2, p, 2, 0, 0, n.sr);
        ip = &emit(n.devop);
        ip->in[0] = p; ip->in[1] = (uint16_t)(p + 1); ip->in[2] = in[0];
      } else {
        ip = &emit(n.devop);
        ip->in[0] = in[0]; ip->in[1] = in[1]; ip->in[2] = in[2];
      }
      ip->out = temp();
      out.push_back(ip->out);
      break;
    }
    case NK_SPLINE: {
      Instr& i = emit(OP_SPLINE);
      for (int k = 0; k < 5; k++) i.in[k] = in[k];
      i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_NARY_CONST: {
      uint32_t rb = raw(n);
      int c = (int)n.raw.size();
      uint16_t p = params(c);
      deriver(D_COPY, rb, c, p, c, 0, 0, n.sr);
      for (int k = 0; k < c; k++) {
        Instr& i = emit(n.devop);
        i.in[0] = in[k]; i.in[1] = (uint16_t)(p + k); i.out = temp();
        out.push_back(i.out);
      }
      break;
    }
    case NK_MAP22: {
      Instr& i = emit(n.devop);
      i.in[0] = in[0]; i.in[1] = in[1]; i.out = temp(2);
      out = {i.out, (uint16_t)(i.out + 1)};
      break;
    }
    case NK_CLIP: case NK_WRAP1: {
      uint32_t rb = raw(n);
      int c = (int)n.raw.size();
      uint16_t p = params(c);
      deriver(D_COPY, rb, c, p, c, 0, 0, n.sr);
      Instr& i = emit(n.kind == NK_CLIP ? OP_CLIP : OP_WRAP1);
      i.in[0] = in[0]; i.p = p; i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_WRAP2: case NK_MIRROR: {
      uint32_t rb = raw(n);
      int np = n.kind == NK_WRAP2 ? 2 : 3;
      uint16_t p = params(np);
      deriver(n.kind == NK_WRAP2 ? D_WRAP2 : D_MIRROR, rb, 2, p, np, 0, 0, n.sr);
      Instr& i = emit(n.kind == NK_WRAP2 ? OP_WRAP2 : OP_MIRROR);
      i.in[0] = in[0]; i.p = p; i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_ROTATE: {
      uint32_t rb = raw(n);
      uint16_t p = params(2);
      deriver(D_ROTATE, rb, 2, p, 2, 0, 0, n.sr);
      Instr& i = emit(OP_ROTATE);
      i.in[0] = in[0]; i.in[1] = in[1]; i.p = p; i.out = temp(2);
      out = {i.out, (uint16_t)(i.out + 1)};
      break;
    }
    case NK_SINE: {
      uint16_t p = params(1), s = state(1);
      deriver(D_INV_SR, 0, 0, p, 1, 0, 0, n.sr);
      hash_init(s, INIT_SINE_PHASE, n.hash);
      Instr& i = emit(OP_SINE);
      i.in[0] = in[0]; i.p = p; i.s = s; i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_NOISE: {
      uint16_t s = state(1);
      hash_init(s, INIT_NOISE_SEED, n.hash);
      Instr& i = emit(OP_NOISE);
      i.s = s; i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_IMPULSE: {
      Instr& i = emit(OP_IMPULSE);
      i.s = state(1); i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_RAMP: {
      uint16_t p = params(1);
      deriver(D_RAMP_SR, 0, 0, p, 1, 0, 0, n.sr);
      Instr& i = emit(OP_RAMP);
      i.in[0] = in[0]; i.p = p; i.s = state(1); i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_WAVETABLE: {
      uint16_t p = params(1), s = state(2);
      deriver(D_INV_SR, 0, 0, p, 1, 0, 0, n.sr);
      hash_init(s, INIT_SINE_PHASE, n.hash);
      if (wt_offset[n.mode & 3] < 0) wt_offset[n.mode & 3] = (int)table(wavetable_blob(n.mode));
      Instr& i = emit(OP_WAVETABLE);
      i.in[0] = in[0]; i.p = p; i.s = s; i.aux = (uint32_t)wt_offset[n.mode & 3]; i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_WAVE: {
      if (n.table.empty()) { out.push_back(zero()); break; }
      Instr& i = emit(OP_WAVE);
      i.s = state(1); i.aux = table(n.table); i.aux2 = (uint32_t)n.table.size(); i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_SVF: {
      int npar = n.mode >= 6 ? 3 : 2, nfixed = n.aux;
      uint32_t rb = raw(n);
      if (nfixed == npar) {
        uint16_t p = params(6);
        deriver(D_SVF, rb, npar, p, 6, n.mode, 0, n.sr);
        Instr& i = emit(OP_SVF);
        i.in[0] = in[0]; i.p = p; i.s = state(2); i.out = temp();
        out.push_back(i.out);
      } else {
        uint16_t p = params(4);
        deriver(D_SVF_DEFAULTS, rb, nfixed, p, 4, n.mode, npar, n.sr);
        uint16_t s = state(11);
        t.state_init[(s & 0x3fff) + 8] = fbits(std::numeric_limits<float>::quiet_NaN());   // cached hz: force first update
        Instr& i = emit(OP_SVF_VAR);
        int nvar = npar - nfixed;
        for (int k = 0; k <= nvar; k++) i.in[k] = in[k];
        i.n = (uint16_t)(n.mode | (nvar << 8));
        i.p = p; i.s = s; i.out = temp();
        out.push_back(i.out);
      }
      break;
    }
    case NK_BIQUAD: {
      int kind = n.mode, nvar = n.aux;
      uint32_t rb = raw(n);
      if (nvar == 0) {
        uint16_t p = params(5);
        if (kind == 0) deriver(D_COPY, rb, 5, p, 5, 0, 0, n.sr);
        else deriver(D_BIQUAD, rb, kind == 1 ? 1 : 2, p, 5, kind, 0, n.sr);
        Instr& i = emit(OP_BIQUAD);
        i.in[0] = in[0]; i.p = p; i.s = state(4); i.out = temp();
        out.push_back(i.out);
      } else {
        uint16_t p = params(1);
        deriver(D_RAMP_SR, 0, 0, p, 1, 0, 0, n.sr);
        uint16_t s = state(11);
        t.state_init[(s & 0x3fff) + 9] = fbits(std::numeric_limits<float>::quiet_NaN());
        t.state_init[(s & 0x3fff) + 10] = fbits(110.0f);   // resonator() default bandwidth until the input arrives
        Instr& i = emit(OP_BIQUAD_VAR);
        for (int k = 0; k <= nvar; k++) i.in[k] = in[k];
        i.n = (uint16_t)(kind | (nvar << 8));
        i.p = p; i.s = s; i.out = temp();
        out.push_back(i.out);
      }
      break;
    }
    case NK_ONEPOLE: {
      int kind = n.mode;
      if (n.n_in == 1) {
        uint32_t rb = raw(n);
        uint16_t p = params(1);
        deriver(D_ONEPOLE, rb, 1, p, 1, kind, 0, n.sr);
        Instr& i = emit(OP_ONEPOLE);
        i.in[0] = in[0]; i.n = (uint16_t)kind; i.p = p; i.s = state(2); i.out = temp();
        out.push_back(i.out);
      } else {
        uint16_t p = params(1);
        deriver(D_RAMP_SR, 0, 0, p, 1, 0, 0, n.sr);
        uint16_t s = state(4);
        t.state_init[(s & 0x3fff) + 3] = fbits(std::numeric_limits<float>::quiet_NaN());
        Instr& i = emit(OP_ONEPOLE_VAR);
        i.in[0] = in[0]; i.in[1] = in[1]; i.n = (uint16_t)kind; i.p = p; i.s = s; i.out = temp();
        out.push_back(i.out);
      }
      break;
    }
    case NK_PINKPASS: {
      Instr& i = emit(OP_PINKPASS);
      i.in[0] = in[0]; i.s = state(7); i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_FIR: {
      uint32_t rb = raw(n);
      int c = (int)n.raw.size();
      uint16_t p = params(c);
      deriver(D_COPY, rb, c, p, c, 0, 0, n.sr);
      Instr& i = emit(OP_FIR);
      i.in[0] = in[0]; i.n = (uint16_t)c; i.p = p; i.s = state(c); i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_TICK: {
      Instr& i = emit(OP_TICK);
      i.in[0] = in[0]; i.s = state(1); i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_DELAY: {
      raw(n, true);
      double len = std::round((double)n.raw[0] * n.sr);
      uint32_t L = !(len >= 1.0) ? 1u : (len > 4.0e8 ? 400000000u : (uint32_t)len);   // NaN / negative / zero: one sample
      Instr& i = emit(OP_DELAY);
      i.in[0] = in[0]; i.aux = ring(L); i.s = state(1); i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_TAP: {
      uint32_t rb = raw(n, true);
      uint16_t p = params(3);
      deriver(D_TAP, rb, 2, p, 3, 0, 0, n.sr);
      double need = std::ceil((double)n.raw[1] * n.sr) + 4.0;
      uint32_t L = 4;
      while ((double)L < need && L < (1u << 30)) L <<= 1;
      Instr& i = emit(OP_TAP);
      i.in[0] = in[0]; i.in[1] = in[1]; i.n = (uint16_t)n.mode; i.aux = ring(L); i.p = p; i.s = state(1); i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_SAMP_DELAY: {
      if (n.aux <= 0) { out.push_back(zero()); break; }
      Instr& i = emit(OP_SAMP_DELAY);
      i.in[0] = in[0]; i.in[1] = in[1]; i.aux = ring((uint32_t)n.aux); i.s = state(1); i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_ENVELOPE: {
      uint32_t rb = raw(n);
      int c = (int)n.raw.size();
      uint16_t p = params(5);
      deriver(D_COPY, rb, c, p, c, 0, 0, n.sr);
      deriver(D_INV_SR, 0, 0, p + 4, 1, 0, 0, n.sr);
      uint16_t s = state(8);
      t.state_init[(s & 0x3fff) + 7] = 1u;   // first
      hash_init((uint16_t)(s + 5), INIT_HASH_LO, n.hash);
      hash_init((uint16_t)(s + 6), INIT_HASH_HI, n.hash);
      Instr& i = emit(OP_ENVELOPE);
      for (int k = 0; k < n.n_in && k < 4; k++) i.in[k] = in[k];
      i.n = (uint16_t)(n.mode | (n.n_in << 8));
      i.p = p; i.s = s; i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_DECLICK: {
      uint32_t rb = raw(n);
      uint16_t p = params(2);
      deriver(D_COPY, rb, 1, p, 1, 0, 0, n.sr);
      deriver(D_INV_SR, 0, 0, p + 1, 1, 0, 0, n.sr);
      Instr& i = emit(OP_DECLICK);
      i.in[0] = in[0]; i.p = p; i.s = state(1); i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_SHIFT_REG: {
      Instr& i = emit(OP_SHIFT_REG);
      i.in[0] = in[0]; i.in[1] = in[1]; i.s = state(8); i.out = temp(8);
      for (int k = 0; k < 8; k++) out.push_back((uint16_t)(i.out + k));
      break;
    }
    case NK_SNH: {
      Instr& i = emit(OP_SNH);
      i.in[0] = in[0]; i.in[1] = in[1]; i.s = state(1); i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_QUANTIZE: {
      uint32_t rb = raw(n);
      uint16_t p = params(1);
      deriver(D_COPY, rb, 1, p, 1, 0, 0, n.sr);
      Instr& i = emit(OP_QUANTIZE);
      i.in[0] = in[0]; i.p = p; i.aux = table(n.table); i.aux2 = (uint32_t)n.table.size(); i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_ARR_GET: {
      Instr& i = emit(OP_ARR_GET);
      i.in[0] = in[0]; i.aux = table(n.table); i.aux2 = (uint32_t)n.table.size(); i.out = temp();
      out.push_back(i.out);
      break;
    }
    case NK_KR: {   // nodes.rs:271-278
      t.h.flags |= TAPE_DIVERGENT;
      uint16_t cnt = state(1);
      uint16_t vals = n.n_out ? state(n.n_out) : NONE;
      size_t begin = t.code.size();
      {
        Instr& b = emit(OP_KR_BEGIN);
        b.s = cnt; b.aux2 = (uint32_t)n.aux;
      }
      std::vector<uint16_t> o = graph(n.kids[0], in);
      if (!err.empty()) return {};
      for (int k = 0; k < n.n_out; k++) {
        Instr& st = emit(OP_ST_STATE);
        st.in[0] = o[k]; st.s = (uint16_t)(vals + k);
      }
      t.code[begin].aux = (uint32_t)t.code.size();
      Instr& e = emit(OP_KR_END);
      e.s = cnt;
      for (int k = 0; k < n.n_out; k++) out.push_back((uint16_t)(vals + k));   // the held values ARE the outputs
      break;
    }
    case NK_FEEDBACK: {   // FeedbackUnit: out = x(in + out delayed by max(1, round(delay*sr)) samples)
      raw(n, true);
      double len = std::round((double)n.raw[0] * n.sr);
      uint32_t L = !(len >= 1.0) ? 1u : (len > 4.0e8 ? 400000000u : (uint32_t)len);   // NaN / negative / zero: one sample
      int ch = n.n_in;
      if (L == 1) {   // one value per channel: state, not a ring (a dependent HBM round trip per sample otherwise)
        std::vector<uint16_t> held, mixed1;
        for (int c = 0; c < ch; c++) {
          held.push_back(state(1));
          Instr& r = emit(OP_FB1_READ);
          r.in[0] = in[c]; r.s = held[c]; r.out = temp();
          mixed1.push_back(r.out);
        }
        std::vector<uint16_t> o1 = graph(n.kids[0], mixed1);
        if (!err.empty()) return {};
        for (int c = 0; c < ch; c++) {
          Instr& w = emit(OP_FB1_WRITE);
          w.in[0] = o1[c]; w.s = held[c];
        }
        out = o1;
        break;
      }
      uint16_t idx = state(1);
      std::vector<uint32_t> rg;
      std::vector<uint16_t> mixed;
      for (int c = 0; c < ch; c++) {
        rg.push_back(ring(L));
        Instr& r = emit(OP_FB_READ);
        r.in[0] = in[c]; r.aux = rg[c]; r.s = idx; r.out = temp();
        mixed.push_back(r.out);
      }
      std::vector<uint16_t> o = graph(n.kids[0], mixed);
      if (!err.empty()) return {};
      for (int c = 0; c < ch; c++) {
        Instr& w = emit(OP_FB_WRITE);
        w.in[0] = o[c]; w.aux = rg[c]; w.s = idx; w.n = (uint16_t)(c == ch - 1 ? 1 : 0);
      }
      out = o;
      break;
    }
    case NK_SELECT: {   // nodes.rs:27-33
      t.h.flags |= TAPE_DIVERGENT;
      uint16_t o = temp();
      { Instr& z = emit(OP_ZERO); z.out = o; }
      for (size_t k = 0; k < n.kids.size(); k++) {
        size_t j = t.code.size();
        { Instr& b = emit(OP_JNE_IDX); b.in[0] = in[0]; b.aux2 = (uint32_t)k; }
        std::vector<uint16_t> ko = graph(n.kids[k], {});
        if (!err.empty()) return {};
        { Instr& m = emit(OP_MOV); m.in[0] = ko[0]; m.out = o; }
        t.code[j].aux = (uint32_t)t.code.size();
      }
      out.push_back(o);
      break;
    }
    case NK_RESET: {   // Reset / TrigReset / ResetV, nodes.rs:332-453
      uint16_t cnt = n.mode == 1 ? NONE : state(1);
      size_t j = t.code.size();
      if (n.mode == 0) {
        Instr& r = emit(OP_RESET_EVERY);
        r.s = cnt;
        r.aux = (uint32_t)std::min<uint64_t>(as_usize(std::round(n.raw[0] * (float)n.sr)), 0xffffffffULL);
      } else if (n.mode == 1) {
        Instr& r = emit(OP_RESET_IF);
        r.in[0] = in[0];
      } else {
        uint16_t p = params(1);
        deriver(D_RAMP_SR, 0, 0, p, 1, 0, 0, n.sr);
        Instr& r = emit(OP_RESET_V);
        r.in[0] = in[0]; r.s = cnt; r.p = p;
      }
      if (n.mode == 0) raw(n, true);
      ResetRange rr;
      rr.s_lo = (uint16_t)t.state_init.size();
      rr.ring_lo = (uint16_t)t.rings.size();
      std::vector<uint16_t> o = graph(n.kids[0], {});
      if (!err.empty()) return {};
      rr.s_hi = (uint16_t)t.state_init.size();
      rr.ring_hi = (uint16_t)t.rings.size();
      t.code[j].aux2 = (uint32_t)t.resets.size();
      t.resets.push_back(rr);
      out.push_back(o[0]);
      break;
    }
    case NK_RFFT: case NK_IFFT: {   // nodes.rs:601-700, per-lane path
      uint32_t N = (uint32_t)n.aux;
      int lg = 0;
      while ((1u << lg) < N) lg++;
      bool r = n.kind == NK_RFFT;
      uint32_t first = ring(N);
      ring(N); ring(N);
      if (!r) ring(N);
      // twiddle table: w_k = exp(-2*pi*i*k/N), k < N/2, rounded from f64 like a precomputed f32 table
      // (one table per transform size and tape, at an even offset: the frame kernels read a twiddle as one float2)
      if (lg < 32 && tw_offset[lg] < 0) {
        std::vector<float> tw(N);
        for (uint32_t k = 0; k < N / 2; k++) {
          double a = -2.0 * 3.14159265358979323846 * (double)k / (double)N;
          tw[2 * k] = (float)std::cos(a);
          tw[2 * k + 1] = (float)std::sin(a);
        }
        if (t.tables.size() & 1) t.tables.push_back(0.0f);
        tw_offset[lg] = (long)table(tw);
      }
      uint16_t s = state(1, (uint32_t)n.mode);   // count starts at `start`
      Instr& i = emit(r ? OP_RFFT : OP_IFFT);
      i.in[0] = in[0];
      if (!r) i.in[1] = in[1];
      i.n = (uint16_t)lg; i.aux = first; i.aux2 = (uint32_t)tw_offset[lg]; i.s = s; i.out = temp(2);
      out = {i.out, (uint16_t)(i.out + 1)};
      break;
    }
    case NK_SEQ: {   // nodes.rs:74-106.  State: event counter, then per net: active, delay, dur, order stamp, ticking flag
      t.h.flags |= TAPE_DIVERGENT;
      int nk = (int)n.kids.size();
      if (nk == 0) { out.push_back(zero()); break; }
      if (nk > 64) { err = "seq(): more than 64 nets"; break; }
      uint16_t p = params(1);
      deriver(D_RAMP_SR, 0, 0, p, 1, 0, 0, n.sr);
      uint16_t st = state(1 + 5 * nk);
      for (int k = 0; k < 1 + 5 * nk; k++) t.state_keep[(st & 0x3fff) + k] = 1;   // the event list is not touched by reset()
      uint16_t vals = temp(nk);
      uint32_t first_range = (uint32_t)t.resets.size();
      t.resets.resize(t.resets.size() + nk);           // one reset range per net, filled in below
      {
        Instr& tr = emit(OP_SEQ_TRIG);
        for (int k = 0; k < 4; k++) tr.in[k] = in[k];
        tr.s = st; tr.p = p; tr.n = (uint16_t)nk; tr.aux2 = first_range;
      }
      for (int k = 0; k < nk; k++) {
        { Instr& z = emit(OP_ZERO); z.out = (uint16_t)(vals + k); }
        size_t j = t.code.size();
        { Instr& gt = emit(OP_SEQ_GATE); gt.s = st; gt.n = (uint16_t)k; }
        ResetRange rr;
        rr.s_lo = (uint16_t)t.state_init.size();
        rr.ring_lo = (uint16_t)t.rings.size();
        std::vector<uint16_t> ko = graph(n.kids[k], {});
        if (!err.empty()) return {};
        rr.s_hi = (uint16_t)t.state_init.size();
        rr.ring_hi = (uint16_t)t.rings.size();
        t.resets[first_range + k] = rr;
        { Instr& m = emit(OP_MOV); m.in[0] = ko[0]; m.out = (uint16_t)(vals + k); }
        t.code[j].aux = (uint32_t)t.code.size();
      }
      Instr& e = emit(OP_SEQ_END);
      e.in[0] = vals; e.s = st; e.n = (uint16_t)nk; e.out = temp();
      out.push_back(e.out);
      break;
    }
    default:
      err = "internal: unknown node kind";
  }
  return out;
}

uint64_t mix(uint64_t h, uint64_t v) { return atto(h, v ^ 0x9e3779b97f4a7c15ULL); }

uint64_t sig_graph(const Graph& g, uint64_t h) {
  h = mix(h, 0xabc0 + g.nodes.size());
  h = mix(h, (uint64_t)g.n_in);
  for (const Node& n : g.nodes) {
    h = mix(h, n.kind); h = mix(h, n.devop); h = mix(h, (uint64_t)n.n_in); h = mix(h, (uint64_t)n.n_out);
    h = mix(h, (uint64_t)(uint32_t)n.mode); h = mix(h, (uint64_t)(uint32_t)n.aux); h = mix(h, n.raw.size());
    uint64_t srb; double sr = n.sr; memcpy(&srb, &sr, 8); h = mix(h, srb);
    bool structural = n.kind == NK_DELAY || n.kind == NK_TAP || n.kind == NK_FEEDBACK || (n.kind == NK_RESET && n.mode == 0);
    if (structural) for (float v : n.raw) h = mix(h, fbits(v));
    for (float v : n.table) h = mix(h, fbits(v));
    for (const Src& s : n.in) { h = mix(h, (uint64_t)(uint32_t)s.node); h = mix(h, (uint64_t)(uint32_t)s.port); }
    for (const Graph& k : n.kids) h = sig_graph(k, h);
  }
  for (const Src& s : g.outs) { h = mix(h, (uint64_t)(uint32_t)s.node); h = mix(h, (uint64_t)(uint32_t)s.port); }
  return h;
}

void raw_graph(const Graph& g, std::vector<float>* raw) {
  for (const Node& n : g.nodes) {
    // must visit in exactly the order Lower::node() calls raw(): own parameters first, then nested nets —
    // except reset(), whose duration is registered before its kid like everything else.
    switch (n.kind) {
      case NK_SAMP_DELAY: case NK_KR: case NK_SELECT: case NK_SEQ: case NK_ARR_GET: case NK_WAVE: break;
      case NK_RESET: if (n.mode == 0) raw->insert(raw->end(), n.raw.begin(), n.raw.end()); break;
      default: raw->insert(raw->end(), n.raw.begin(), n.raw.end());
    }
    for (const Graph& k : n.kids) raw_graph(k, raw);
  }
}

}  // namespace

void collect_raw(const Graph& g, std::vector<float>* raw) { raw_graph(g, raw); }
uint64_t structure_signature(const Graph& g) { return sig_graph(g, 0x51475450ULL); }

void Tape::derive(const float* r, float* P) const {
  for (const Deriver& d : derivers) {
    const float* in = r + d.raw_base;
    float* o = P + d.p_base;
    switch (d.kind) {
      case D_COPY: for (uint32_t i = 0; i < d.n_raw; i++) o[i] = in[i]; break;
      case D_SVF: svf_coefs(d.mode, in[0], in[1], d.n_raw >= 3 ? in[2] : 1.0f, d.sr, o); break;
      case D_SVF_DEFAULTS: {
        // P = hz q gain sr; the fixed parameters are the trailing ones of (hz, q[, gain])
        int npar = d.aux, nfixed = (int)d.n_raw;
        o[0] = 440.0f; o[1] = 1.0f; o[2] = 1.0f; o[3] = d.sr;
        for (int k = 0; k < nfixed; k++) o[npar - nfixed + k] = in[k];
        break;
      }
      case D_BIQUAD: biquad_coefs(d.mode, in[0], d.n_raw >= 2 ? in[1] : 0.0f, d.sr, o); break;
      case D_ONEPOLE: o[0] = onepole_coef(d.mode, in[0], d.sr); break;
      case D_WRAP2: { float lo = fminf(in[0], in[1]), hi = fmaxf(in[0], in[1]); o[0] = lo; o[1] = hi - lo; break; }
      case D_MIRROR: { float lo = fminf(in[0], in[1]), hi = fmaxf(in[0], in[1]); o[0] = lo; o[1] = hi; o[2] = hi - lo; break; }
      case D_ROTATE: o[0] = cosf(in[0]) * in[1]; o[1] = sinf(in[0]) * in[1]; break;
      case D_PAN: pan_weights(in[0], &o[0], &o[1]); break;
      case D_RAMP_SR: o[0] = d.sr; break;
      case D_INV_SR: o[0] = (float)(1.0 / (double)d.sr); break;
      case D_TAP: o[0] = in[0]; o[1] = in[1]; o[2] = d.sr; break;
    }
  }
}

TvPlan plan_tv(const Tape& t, size_t smem_limit) {
  TvPlan pl;
  if (t.h.flags & TAPE_DIVERGENT) return pl;
  int H = 512, fft_n = 0;
  uint32_t min_tap_ring = 0xffffffffu;
  auto gcd = [](int a, int b) { while (b) { int r = a % b; a = b; b = r; } return a; };
  for (const Instr& i : t.code) {
    if (op_is_stateless(i.op)) continue;
    switch (i.op) {
      case OP_NOISE: case OP_WAVE: case OP_IMPULSE: case OP_TICK: case OP_DELAY: break;
      case OP_SVF: case OP_BIQUAD: case OP_ONEPOLE: pl.n_lti++; break;       // block-level scan over the hop
      case OP_SINE: case OP_RAMP: pl.sequential = true; break;              // exact phase recurrence on one thread
      // scalar-state ops without a block form: one thread steps them through the hop with the generic per-sample code
      case OP_SVF_VAR: case OP_BIQUAD_VAR: case OP_ONEPOLE_VAR: case OP_PINKPASS: case OP_FIR: case OP_ENVELOPE:
      case OP_DECLICK: case OP_SHIFT_REG: case OP_SNH: case OP_PAN_VAR: case OP_WAVETABLE:
        pl.sequential = true; break;
      case OP_TAP: min_tap_ring = std::min(min_tap_ring, t.rings[i.aux].length); break;
      case OP_RFFT: case OP_IFFT: {
        int N = 1 << i.n;
        if (N > 8192) return pl;                       // transform buffer must fit in shared memory
        int start = (int)t.state_init[i.s - t.h.n_params];
        H = gcd(H, N);
        if (start) H = gcd(H, start);
        fft_n = std::max(fft_n, N);
        pl.has_fft = true;
        pl.align_s = i.s;
        break;
      }
      default: return pl;
    }
  }
  auto bytes = [&](int h) {
    return (size_t)t.h.n_instr * sizeof(Instr) + (size_t)(t.h.n_params + t.h.n_state + 4) * 4 + (size_t)t.h.n_temps * h * 4 +
           (size_t)h * 4 + (size_t)(fft_n + fft_n / 32) * 8 + (size_t)pl.n_lti * TV_LTI_FLOATS * 4 + (size_t)h * 4 + 2560;
  };
  while (H >= 8 && (bytes(H) > smem_limit || (uint32_t)H > min_tap_ring)) H >>= 1;
  if (H < 8) return pl;
  pl.ok = true; pl.H = H; pl.fft_n = fft_n;
  return pl;
}

// Structural self-check of a finished tape: every index an interpreter will dereference without looking is in range.  The
// kernels trust the tape (no bounds checks in the sample loop), so a lowering bug must stop here, by name, on the host.
static bool validate_tape(const Tape& t, std::string* err) {
  const uint32_t P = t.h.n_params, NS = t.h.n_state, NT = t.h.n_temps, NX = P + NS + NT, N = t.h.n_instr;
  auto bad = [&](size_t k, const char* what) {
    if (err) *err = "internal error: invalid tape (instruction " + std::to_string(k) + ": " + what + ")";
    return false;
  };
  if (t.code.size() != N || t.params.size() != P || t.state_init.size() != NS || t.state_keep.size() != NS ||
      t.rings.size() != t.h.n_rings || t.out_x.size() != t.h.n_outputs)
    return bad(0, "header does not match the tables");
  for (size_t k = 0; k < t.code.size(); k++) {
    const Instr& i = t.code[k];
    if (i.op >= OP_COUNT_) return bad(k, "opcode");
    if (i.out >= std::max(NX, 1u) || i.p >= std::max(NX, 1u) || i.s >= std::max(NX, 1u)) return bad(k, "operand index");
    for (uint16_t x : i.in) if (x >= std::max(NX, 1u)) return bad(k, "input index");
    switch (i.op) {
      case OP_DELAY: case OP_TAP: case OP_SAMP_DELAY: case OP_FB_READ: case OP_FB_WRITE:
        if (i.aux >= t.rings.size() || t.rings[i.aux].length == 0) return bad(k, "ring id");
        break;
      case OP_RFFT: if (i.aux + 2 >= t.rings.size()) return bad(k, "ring id"); break;
      case OP_IFFT: if (i.aux + 3 >= t.rings.size()) return bad(k, "ring id"); break;
      case OP_KR_BEGIN: case OP_JNE_IDX: case OP_SEQ_GATE:
        if (i.aux > N) return bad(k, "jump target");
        break;
      case OP_QUANTIZE: case OP_ARR_GET: case OP_WAVE:
        if ((uint64_t)i.aux + i.aux2 > t.tables.size()) return bad(k, "table range");
        break;
      case OP_SVF: case OP_BIQUAD: case OP_ONEPOLE:
        if (i.aux >= t.h.n_lti) return bad(k, "filter index");
        break;
      default: break;
    }
  }
  for (uint16_t o : t.out_x) if (o >= std::max(NX, 1u)) return bad(0, "output index");
  uint64_t ring_floats = 0;
  for (const Ring& r : t.rings) {
    if (r.offset != ring_floats) return bad(0, "ring offsets are not contiguous");
    ring_floats += r.length;
  }
  if (ring_floats != t.h.ring_floats) return bad(0, "ring_floats");
  for (const ResetRange& r : t.resets)
    if (r.s_lo > r.s_hi || r.s_hi > P + NS || r.s_lo < P || r.ring_lo > r.ring_hi || r.ring_hi > t.rings.size()) return bad(0, "reset range");
  for (const HashInit& h : t.hash_init) if (h.state >= NS) return bad(0, "hash-initialised state index");
  return true;
}

bool lower(const Graph& g, Tape* out, std::string* err) {
  Tape& t = *out;
  t = Tape();
  memset(&t.h, 0, sizeof t.h);
  Lower L(t);
  std::vector<uint16_t> ins;
  for (int i = 0; i < g.n_in; i++) ins.push_back(L.temp());
  std::vector<uint16_t> outs = L.graph(g, ins);
  if (!L.err.empty()) { if (err) *err = L.err; return false; }
  // ---- temporary reuse (uniform tapes only): a temporary dies after its last reader; outputs of multi-output ops
  // stay contiguous.  Shrinks the per-voice working set (shared memory) of long feed-forward graphs.
  if (!(t.h.flags & TAPE_DIVERGENT) && L.n_temps > 0) {
    const int NTv = L.n_temps;
    auto n_out_of = [](const Instr& i) -> int {
      switch (i.op) {
        case OP_POL: case OP_CAR: case OP_PAN: case OP_PAN_VAR: case OP_ROTATE: case OP_RFFT: case OP_IFFT: return 2;
        case OP_SHIFT_REG: return 8;
        default: return i.out == NONE ? 0 : 1;
      }
    };
    std::vector<int> last(NTv, -1);
    const int END = (int)t.code.size();
    for (int k = 0; k < (int)t.code.size(); k++)
      for (uint16_t x : t.code[k].in)
        if (x != NONE && (x & R_TEMP)) last[x & 0x7fff] = k;
    for (uint16_t o : outs)
      if (o & R_TEMP) last[o & 0x7fff] = END;
    std::vector<int> phys(NTv, -1), free_at;      // free_at[slot] = instruction index after which the slot is free
    for (int i = 0; i < g.n_in; i++) { phys[i] = i; free_at.push_back(END); }   // net inputs are refilled every sample
    for (int k = 0; k < (int)t.code.size(); k++) {
      Instr& ins = t.code[k];
      int no = n_out_of(ins);
      if (no == 0 || !(ins.out & R_TEMP)) continue;
      int v0 = ins.out & 0x7fff;
      int die = k;
      for (int q = 0; q < no; q++) die = std::max(die, last[v0 + q]);
      // first fit: `no` contiguous slots all free strictly before k (a slot read by instruction k may not be reused by k)
      int base = -1;
      for (int s0 = 0; s0 + no <= (int)free_at.size() && base < 0; s0++) {
        bool ok = true;
        for (int q = 0; q < no; q++) ok = ok && free_at[s0 + q] < k;
        if (ok) base = s0;
      }
      if (base < 0) { base = (int)free_at.size(); free_at.resize(base + no, -1); }
      for (int q = 0; q < no; q++) { phys[v0 + q] = base + q; free_at[base + q] = die; }
    }
    auto rn = [&](uint16_t x) -> uint16_t {
      if (x == NONE || !(x & R_TEMP)) return x;
      int p = phys[x & 0x7fff];
      return (uint16_t)(R_TEMP | (p < 0 ? 0 : p));
    };
    for (Instr& ins : t.code) { ins.out = rn(ins.out); for (uint16_t& x : ins.in) x = rn(x); }
    for (uint16_t& o : outs) o = rn(o);
    L.n_temps = std::max<int>((int)free_at.size(), 1);
  }
  size_t P = t.params.size(), NS = t.state_init.size();
  if (P >= 0x4000 || NS >= 0x4000 || (size_t)L.n_temps >= 0x7fff || P + NS + L.n_temps >= 0xfff0) {
    if (err) *err = "graph too large for one tape (parameter/state/temporary index overflow)";
    return false;
  }
  auto fix = [&](uint16_t x) -> uint16_t {
    if (x == NONE) return 0;
    if (x & R_TEMP) return (uint16_t)(P + NS + (x & 0x7fff));
    if (x & R_STATE) return (uint16_t)(P + (x & 0x3fff));
    return x;
  };
  for (Instr& i : t.code) {
    i.out = fix(i.out); i.p = fix(i.p); i.s = fix(i.s);
    for (uint16_t& x : i.in) x = fix(x);
  }
  for (ResetRange& r : t.resets) { r.s_lo = (uint16_t)(P + r.s_lo); r.s_hi = (uint16_t)(P + r.s_hi); }
  for (uint16_t o : outs) t.out_x.push_back(fix(o));
  // stateless runs (time-vector kernel): `pad` of a stateless instruction = index of the first stateful one after it
  for (int k = (int)t.code.size() - 1, end = (int)t.code.size(); k >= 0; k--) {
    if (op_is_stateless(t.code[k].op)) t.code[k].pad = (uint32_t)end;
    else { t.code[k].pad = 0; end = k; }
  }
  // fixed-coefficient LTI filters get an index into the time-vector kernel's table of scan matrices
  t.h.n_lti = 0;
  for (Instr& i : t.code) if (op_is_lti(i.op)) i.aux = t.h.n_lti++;
  t.h.magic = TAPE_MAGIC; t.h.version = TAPE_VERSION;
  t.h.n_instr = (uint32_t)t.code.size();
  t.h.n_params = (uint32_t)P; t.h.n_state = (uint32_t)NS; t.h.n_temps = (uint32_t)L.n_temps;
  t.h.n_inputs = (uint32_t)g.n_in; t.h.n_outputs = (uint32_t)outs.size();
  t.h.n_rings = (uint32_t)t.rings.size();
  t.h.n_resets = (uint32_t)t.resets.size();
  t.h.n_hash_init = (uint32_t)t.hash_init.size();
  t.h.table_floats = (uint32_t)t.tables.size();
  t.h.n_raw = (uint32_t)t.raw.size();
  t.h.sample_rate = (float)g.sr;
  t.signature = structure_signature(g);
  // the derivers define the template parameter values
  std::vector<float> p(P, 0.0f);
  for (size_t i = 0; i < P; i++) p[i] = t.params[i];
  t.derive(t.raw.data(), p.data());
  t.params = p;
  return validate_tape(t, err);
}

}  // namespace qg
