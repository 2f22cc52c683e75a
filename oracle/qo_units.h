// TEST INFRASTRUCTURE — CPU oracle (see qo_math.h header).  Object-graph restatement of the reference's
// architecture: every node is a boxed unit with a virtual per-sample tick(), a Net is an ordered vertex
// list evaluated once per sample (the shape of `Net::tick` as driven by /root/reference/src/process.rs:
// 1347-1351).  In-tree nodes follow /root/reference/src/nodes.rs line by line [P]; FunDSP 0.18.2 units are
// restated from the published algorithms [U] (source not vendored; Cargo.lock:1844-1847).
#pragma once
#include <algorithm>
#include <cstdio>
#include <deque>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "qo_math.h"

namespace qo {

static const double DEFAULT_SR = 44100.0;

// Unit ids fed to AttoHash::hash.  In-tree ids are the reference's (nodes.rs `const ID`); FunDSP ids are
// placeholders [U] — the derived phases/seeds are injectable through salts for exactly that reason.
enum : uint64_t {
  ID_PASS = 48, ID_SINK = 47, ID_CONSTANT = 8, ID_MAP = 27, ID_SINE = 21, ID_NOISE = 20, ID_SVF = 36,
  ID_BIQUAD = 15, ID_LOWPOLE = 12, ID_HIGHPOLE = 14, ID_DCBLOCK = 22, ID_ALLPOLE = 46, ID_PINKPASS = 42,
  ID_FIR = 5, ID_TICK = 9, ID_DELAY = 13, ID_TAP = 50, ID_TAPLIN = 51, ID_ENVELOPE = 14001, ID_ENVELOPE_IN = 53,
  ID_JOIN = 41, ID_SPLIT = 40, ID_REVERSE = 45, ID_PAN = 49, ID_CLIP = 88, ID_DECLICK = 23, ID_IMPULSE = 81,
  ID_MIXER = 17, ID_PIPE = 2, ID_STACK = 3, ID_BRANCH = 4, ID_BUS = 10, ID_BINOP = 11, ID_THRU = 31, ID_NET = 63,
  ID_FEEDBACK = 79, ID_WAVE = 65, ID_WAVESYNTH = 34,
  // in-tree (nodes.rs)
  ID_SELECT = 1213, ID_SEQ = 1729, ID_ARRGET = 1312, ID_SHIFTREG = 1110, ID_QUANTIZER = 1111, ID_KR = 1112,
  ID_RESET = 1113, ID_TRIGRESET = 1114, ID_RESETV = 1115, ID_RAMP = 1116, ID_INPUT = 1117, ID_SWAP = 1118,
  ID_RFFT = 1120, ID_IFFT = 1121, ID_SAMPDELAY = 1122, ID_BUFFIN = 1123, ID_BUFFOUT = 1124, ID_SNH = 1125,
};

struct Unit;
typedef std::unique_ptr<Unit> UnitP;

struct Unit {
  virtual ~Unit() {}
  virtual int ins() const = 0;
  virtual int outs() const = 0;
  virtual void tick(const float* in, float* out) = 0;
  virtual void reset() {}
  virtual void set_sr(double) {}
  virtual uint64_t id() const = 0;
  virtual void set_hash(uint64_t) {}
  // AudioNode::ping default: a leaf stores the incoming hash, then mixes in its id.
  virtual uint64_t ping(bool probe, uint64_t h) {
    if (!probe) set_hash(h);
    return atto(h, id());
  }
  // per-voice salt (extension, not in the reference): hash-seeded leaves re-derive hash = atto(base, salt)
  virtual void salt(uint64_t) {}
  virtual UnitP clone() const = 0;
};

#define QO_CLONE(T) \
  UnitP clone() const override { return UnitP(new T(*this)); }

// ------------------------------------------------------------------ stateless closures (`map`)
struct Map : Unit {
  int ni, no;
  std::function<void(const float*, float*)> f;
  Map(int ni_, int no_, std::function<void(const float*, float*)> f_) : ni(ni_), no(no_), f(std::move(f_)) {}
  int ins() const override { return ni; }
  int outs() const override { return no; }
  void tick(const float* in, float* out) override { f(in, out); }
  uint64_t id() const override { return ID_MAP; }
  QO_CLONE(Map)
};

struct Constant : Unit {
  std::vector<float> v;
  explicit Constant(std::vector<float> v_) : v(std::move(v_)) {}
  int ins() const override { return 0; }
  int outs() const override { return (int)v.size(); }
  void tick(const float*, float* out) override { for (size_t i = 0; i < v.size(); i++) out[i] = v[i]; }
  uint64_t id() const override { return ID_CONSTANT; }
  QO_CLONE(Constant)
};

struct Pass : Unit {
  int ins() const override { return 1; }
  int outs() const override { return 1; }
  void tick(const float* in, float* out) override { out[0] = in[0]; }
  uint64_t id() const override { return ID_PASS; }
  QO_CLONE(Pass)
};
struct Sink : Unit {
  int ins() const override { return 1; }
  int outs() const override { return 0; }
  void tick(const float*, float*) override {}
  uint64_t id() const override { return ID_SINK; }
  QO_CLONE(Sink)
};
struct Join : Unit {   // mean of n inputs [U]
  int n;
  explicit Join(int n_) : n(n_) {}
  int ins() const override { return n; }
  int outs() const override { return 1; }
  void tick(const float* in, float* out) override {
    float s = in[0];
    for (int i = 1; i < n; i++) s += in[i];
    out[0] = s / (float)n;
  }
  uint64_t id() const override { return ID_JOIN; }
  QO_CLONE(Join)
};
struct Split : Unit {
  int n;
  explicit Split(int n_) : n(n_) {}
  int ins() const override { return 1; }
  int outs() const override { return n; }
  void tick(const float* in, float* out) override { for (int i = 0; i < n; i++) out[i] = in[0]; }
  uint64_t id() const override { return ID_SPLIT; }
  QO_CLONE(Split)
};
struct Reverse : Unit {
  int n;
  explicit Reverse(int n_) : n(n_) {}
  int ins() const override { return n; }
  int outs() const override { return n; }
  void tick(const float* in, float* out) override { for (int i = 0; i < n; i++) out[i] = in[n - 1 - i]; }
  uint64_t id() const override { return ID_REVERSE; }
  QO_CLONE(Reverse)
};
// equal-power panner [U]: angle = (clamp(pan,-1,1)+1)·π/4, (l, r) = (cos, sin)·x
struct Pan : Unit {
  bool fixed;
  float pan, l, r;
  Pan(bool fixed_, float p) : fixed(fixed_), pan(p) { set(p); }
  void set(float p) {
    pan = p;
    float a = (rclamp(p, -1.0f, 1.0f) + 1.0f) * (PI_F * 0.25f);
    l = std::cos(a);
    r = std::sin(a);
  }
  int ins() const override { return fixed ? 1 : 2; }
  int outs() const override { return 2; }
  void tick(const float* in, float* out) override {
    if (!fixed && in[1] != pan) set(in[1]);
    out[0] = l * in[0];
    out[1] = r * in[0];
  }
  uint64_t id() const override { return ID_PAN; }
  QO_CLONE(Pan)
};

// ------------------------------------------------------------------ sources
// sine(): phase accumulator in f32, output taken before the increment [U]
struct Sine : Unit {
  float phase = 0, sd = (float)(1.0 / DEFAULT_SR);
  uint64_t hash = 0, base = 0;
  int ins() const override { return 1; }
  int outs() const override { return 1; }
  void reset() override { phase = (float)rnd1(hash); }
  void set_sr(double sr) override { sd = (float)(1.0 / sr); }
  void set_hash(uint64_t h) override { base = hash = h; reset(); }
  void salt(uint64_t s) override { hash = s ? atto(base, s) : base; reset(); }
  void tick(const float* in, float* out) override {
    float p = phase;
    phase += in[0] * sd;
    phase -= std::floor(phase);
    out[0] = std::sin(p * TAU_F);
  }
  uint64_t id() const override { return ID_SINE; }
  QO_CLONE(Sine)
};
// white(): counter-based noise, u32 state seeded from the node hash [U]
struct Noise : Unit {
  uint32_t state = 0;
  uint64_t hash = 0, base = 0;
  int ins() const override { return 0; }
  int outs() const override { return 1; }
  void reset() override { state = (uint32_t)hash; }
  void set_hash(uint64_t h) override { base = hash = h; reset(); }
  void salt(uint64_t s) override { hash = s ? atto(base, s) : base; reset(); }
  void tick(const float*, float* out) override {
    state += 1u;
    out[0] = (float)(int32_t)hash32x(state) * (1.0f / 2147483648.0f);
  }
  uint64_t id() const override { return ID_NOISE; }
  QO_CLONE(Noise)
};
struct Impulse : Unit {
  bool fired = false;
  int ins() const override { return 0; }
  int outs() const override { return 1; }
  void reset() override { fired = false; }
  void tick(const float*, float* out) override { out[0] = fired ? 0.0f : 1.0f; fired = true; }
  uint64_t id() const override { return ID_IMPULSE; }
  QO_CLONE(Impulse)
};
// Ramp [P] nodes.rs:459-492
struct Ramp : Unit {
  float val = 0, sr = 44100.f;
  int ins() const override { return 1; }
  int outs() const override { return 1; }
  void reset() override { val = 0; }
  void set_sr(double s) override { sr = (float)s; }
  void tick(const float* in, float* out) override {
    out[0] = val;
    val += in[0] / sr;
    if (val >= 1.0f) val -= 1.0f;
  }
  uint64_t id() const override { return ID_RAMP; }
  QO_CLONE(Ramp)
};

// Band-limited wavetable oscillators saw / square / triangle / soft_saw (FunDSP wavetable.rs + WaveSynth) [U]:
// one table per quarter octave from 20 Hz to 20 kHz, each the inverse FFT of the partials below Nyquist (upper
// partials faded out between 20 kHz and 22.05 kHz), 4x oversampled power-of-two length, peak-normalised, read with
// Niemitalo's optimal 4x 4-point 4th-order interpolator.  shape: 0 saw (1/i), 1 square (odd 1/i), 2 triangle (odd 1/i^2,
// alternating sign), 3 soft_saw (1/i^2).
struct WaveTableSet {
  std::vector<float> limit;                 // table i serves frequencies below limit[i]
  std::vector<std::vector<float>> table;
};
const WaveTableSet& wavetable_set(int shape);
static inline float optimal4x44(float a0, float a1, float a2, float a3, float x) {
  float z = x - 0.5f;
  float even1 = a2 + a1, odd1 = a2 - a1, even2 = a3 + a0, odd2 = a3 - a0;
  float c0 = even1 * 0.46567255120778489f + even2 * 0.03432729708429672f;
  float c1 = odd1 * 0.53743830753560162f + odd2 * 0.15429462557307461f;
  float c2 = even1 * -0.25194210134021744f + even2 * 0.25194744935939062f;
  float c3 = odd1 * -0.46896069955075126f + odd2 * 0.15578800670302476f;
  float c4 = even1 * 0.00986988334359864f + even2 * -0.00989340017126506f;
  return (((c4 * z + c3) * z + c2) * z + c1) * z + c0;
}
struct WaveSynth : Unit {
  int shape;
  float phase = 0, sd = (float)(1.0 / DEFAULT_SR);
  size_t hint = 0;
  uint64_t hash = 0, base = 0;
  explicit WaveSynth(int shape_) : shape(shape_) {}
  int ins() const override { return 1; }
  int outs() const override { return 1; }
  void reset() override { phase = (float)rnd1(hash); hint = 0; }
  void set_sr(double sr) override { sd = (float)(1.0 / sr); }
  void set_hash(uint64_t h) override { base = hash = h; reset(); }
  void salt(uint64_t s) override { hash = s ? atto(base, s) : base; reset(); }
  void tick(const float* in, float* out) override {
    const WaveTableSet& ts = wavetable_set(shape);
    float f = in[0];
    phase += f * sd;
    phase -= std::floor(phase);
    float af = std::fabs(f);
    while (hint + 1 < ts.table.size() && af >= ts.limit[hint]) hint++;
    while (hint > 0 && af < ts.limit[hint - 1]) hint--;
    const std::vector<float>& t = ts.table[hint];
    float p = (float)t.size() * phase;
    size_t i1 = (size_t)p, mask = t.size() - 1;
    float w = p - (float)i1;
    size_t i0 = (i1 + t.size() - 1) & mask;
    i1 &= mask;
    out[0] = optimal4x44(t[i0], t[i1], t[(i1 + 1) & mask], t[(i1 + 2) & mask], w);
  }
  uint64_t id() const override { return ID_WAVESYNTH; }
  QO_CLONE(WaveSynth)
};

// ------------------------------------------------------------------ filters
// Simper SVF [U].  mode: 0 lowpass 1 highpass 2 bandpass 3 notch 4 peak 5 allpass 6 bell 7 lowshelf 8 highshelf.
// nfixed = number of trailing parameters frozen at construction (the rest arrive as inputs after the signal).
struct Svf : Unit {
  int mode, npar, nfixed;
  float cutoff = 440.f, q = 1.f, gain = 1.f, sr = (float)DEFAULT_SR;
  float a1 = 0, a2 = 0, a3 = 0, m0 = 0, m1 = 0, m2 = 0, ic1 = 0, ic2 = 0;
  Svf(int mode_, int nfixed_, float c, float q_, float g) : mode(mode_), nfixed(nfixed_), cutoff(c), q(q_), gain(g) {
    npar = mode >= 6 ? 3 : 2;
    update();
  }
  void update() {
    float g, k, A;
    switch (mode) {
      case 6: A = std::sqrt(gain); g = std::tan(PI_F * cutoff / sr); k = 1.0f / (q * A); break;
      case 7: A = std::sqrt(gain); g = std::tan(PI_F * cutoff / sr) / std::sqrt(A); k = 1.0f / q; break;
      case 8: A = std::sqrt(gain); g = std::tan(PI_F * cutoff / sr) * std::sqrt(A); k = 1.0f / q; break;
      default: A = 1.0f; g = std::tan(PI_F * cutoff / sr); k = 1.0f / q; break;
    }
    a1 = 1.0f / (1.0f + g * (g + k));
    a2 = g * a1;
    a3 = g * a2;
    switch (mode) {
      case 0: m0 = 0; m1 = 0; m2 = 1; break;
      case 1: m0 = 1; m1 = -k; m2 = -1; break;
      case 2: m0 = 0; m1 = 1; m2 = 0; break;
      case 3: m0 = 1; m1 = -k; m2 = 0; break;
      case 4: m0 = 1; m1 = -k; m2 = -2; break;
      case 5: m0 = 1; m1 = -2.0f * k; m2 = 0; break;
      case 6: m0 = 1; m1 = k * (A * A - 1.0f); m2 = 0; break;
      case 7: m0 = 1; m1 = k * (A - 1.0f); m2 = A * A - 1.0f; break;
      case 8: m0 = A * A; m1 = k * (1.0f - A) * A; m2 = 1.0f - A * A; break;
    }
  }
  int ins() const override { return 1 + npar - nfixed; }
  int outs() const override { return 1; }
  void reset() override { ic1 = ic2 = 0; }
  void set_sr(double s) override { sr = (float)s; update(); }
  void tick(const float* in, float* out) override {
    int nvar = npar - nfixed;
    if (nvar > 0) {
      float c = cutoff, qq = q, gg = gain;
      // parameter order (hz, q, gain); the fixed ones are the trailing ones
      if (nvar >= 1) c = in[1];
      if (nvar >= 2) qq = in[2];
      if (nvar >= 3) gg = in[3];
      if (c != cutoff || qq != q || gg != gain) { cutoff = c; q = qq; gain = gg; update(); }
    }
    float v0 = in[0];
    float v3 = v0 - ic2;
    float v1 = a1 * ic1 + a2 * v3;
    float v2 = ic2 + a2 * ic1 + a3 * v3;
    ic1 = 2.0f * v1 - ic1;
    ic2 = 2.0f * v2 - ic2;
    out[0] = m0 * v0 + m1 * v1 + m2 * v2;
  }
  uint64_t id() const override { return ID_SVF; }
  QO_CLONE(Svf)
};

// direct-form-I biquad [U]; kind 0 fixed coefficients, 1 butterpass(hz), 2 resonator(hz, bw); var = inputs
struct Biquad : Unit {
  int kind, nvar;
  float p0, p1, sr = (float)DEFAULT_SR;
  float a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0, x1 = 0, x2 = 0, y1 = 0, y2 = 0;
  Biquad(int kind_, int nvar_, float p0_, float p1_) : kind(kind_), nvar(nvar_), p0(p0_), p1(p1_) { update(); }
  static Biquad* fixed(float a1, float a2, float b0, float b1, float b2) {
    Biquad* b = new Biquad(0, 0, 0, 0);
    b->a1 = a1; b->a2 = a2; b->b0 = b0; b->b1 = b1; b->b2 = b2;
    return b;
  }
  void update() {
    if (kind == 1) {
      float f = std::tan(p0 * PI_F / sr);
      float a0r = 1.0f / (1.0f + 1.41421356237309504880f * f + f * f);
      a1 = (2.0f * f * f - 2.0f) * a0r;
      a2 = (1.0f - 1.41421356237309504880f * f + f * f) * a0r;
      b0 = f * f * a0r;
      b1 = 2.0f * b0;
      b2 = b0;
    } else if (kind == 2) {
      float r = std::exp(-PI_F * p1 / sr);
      a1 = -2.0f * r * std::cos(TAU_F * p0 / sr);
      a2 = r * r;
      b0 = std::sqrt(1.0f - r * r) * 0.5f;
      b1 = 0.0f;
      b2 = -b0;
    }
  }
  int ins() const override { return 1 + nvar; }
  int outs() const override { return 1; }
  void reset() override { x1 = x2 = y1 = y2 = 0; }
  void set_sr(double s) override { sr = (float)s; update(); }
  void tick(const float* in, float* out) override {
    if (nvar >= 1) {
      float c = in[1], d = nvar >= 2 ? in[2] : p1;
      if (c != p0 || d != p1) { p0 = c; p1 = d; update(); }
    }
    float x0 = in[0];
    float y0 = b0 * x0 + b1 * x1 + b2 * x2 - a1 * y1 - a2 * y2;
    x2 = x1; x1 = x0; y2 = y1; y1 = y0;
    out[0] = y0;
  }
  uint64_t id() const override { return ID_BIQUAD; }
  QO_CLONE(Biquad)
};

// one-pole family [U]: kind 0 lowpole, 1 highpole, 2 dcblock, 3 allpole(delay)
struct OnePole : Unit {
  int kind;
  bool var;
  float p, sr = (float)DEFAULT_SR, coeff = 0, x1 = 0, y1 = 0;
  OnePole(int kind_, bool var_, float p_) : kind(kind_), var(var_), p(p_) { update(); }
  void update() {
    if (kind == 0 || kind == 1) coeff = std::exp(-TAU_F * p / sr);
    else if (kind == 2) coeff = 1.0f - TAU_F * p / sr;
    else coeff = (1.0f - p) / (1.0f + p);
  }
  int ins() const override { return var ? 2 : 1; }
  int outs() const override { return 1; }
  void reset() override { x1 = y1 = 0; }
  void set_sr(double s) override { sr = (float)s; update(); }
  void tick(const float* in, float* out) override {
    if (var && in[1] != p) { p = in[1]; update(); }
    float x = in[0], y;
    switch (kind) {
      case 0: y = (1.0f - coeff) * x + coeff * y1; break;
      case 1: y = coeff * (y1 + x - x1); break;
      case 2: y = x - x1 + coeff * y1; break;
      default: y = coeff * (x - y1) + x1; break;
    }
    x1 = x; y1 = y;
    out[0] = y;
  }
  uint64_t id() const override { return kind == 0 ? ID_LOWPOLE : kind == 1 ? ID_HIGHPOLE : kind == 2 ? ID_DCBLOCK : ID_ALLPOLE; }
  QO_CLONE(OnePole)
};

// Paul Kellet's pinking filter [U]
struct Pinkpass : Unit {
  float b[7] = {0, 0, 0, 0, 0, 0, 0};
  int ins() const override { return 1; }
  int outs() const override { return 1; }
  void reset() override { for (float& x : b) x = 0; }
  void tick(const float* in, float* out) override {
    float w = in[0];
    b[0] = 0.99886f * b[0] + w * 0.0555179f;
    b[1] = 0.99332f * b[1] + w * 0.0750759f;
    b[2] = 0.96900f * b[2] + w * 0.1538520f;
    b[3] = 0.86650f * b[3] + w * 0.3104856f;
    b[4] = 0.55000f * b[4] + w * 0.5329522f;
    b[5] = -0.7616f * b[5] - w * 0.0168980f;
    float pink = b[0] + b[1] + b[2] + b[3] + b[4] + b[5] + b[6] + w * 0.5362f;
    b[6] = w * 0.115926f;
    out[0] = pink * 0.11f;
  }
  uint64_t id() const override { return ID_PINKPASS; }
  QO_CLONE(Pinkpass)
};

struct Fir : Unit {   // y[n] = sum_i w[i] x[n-i], accumulated from i = 0 [U]
  std::vector<float> w, x;
  explicit Fir(std::vector<float> w_) : w(std::move(w_)), x(w.size(), 0.0f) {}
  int ins() const override { return 1; }
  int outs() const override { return 1; }
  void reset() override { std::fill(x.begin(), x.end(), 0.0f); }
  void tick(const float* in, float* out) override {
    for (size_t i = x.size() - 1; i > 0; i--) x[i] = x[i - 1];
    x[0] = in[0];
    float acc = 0.0f;
    for (size_t i = 0; i < w.size(); i++) acc += w[i] * x[i];
    out[0] = acc;
  }
  uint64_t id() const override { return ID_FIR; }
  QO_CLONE(Fir)
};

// ------------------------------------------------------------------ delays
struct Tick : Unit {
  float v = 0;
  int ins() const override { return 1; }
  int outs() const override { return 1; }
  void reset() override { v = 0; }
  void tick(const float* in, float* out) override { out[0] = v; v = in[0]; }
  uint64_t id() const override { return ID_TICK; }
  QO_CLONE(Tick)
};
struct Delay : Unit {   // fixed delay of max(1, round(t·sr)) samples [U]
  float t;
  std::vector<float> buf;
  size_t i = 0;
  explicit Delay(float t_) : t(t_) { set_sr(DEFAULT_SR); }
  static size_t length(float t, double sr) {
    double n = std::round((double)t * sr);
    return !(n >= 1.0) ? 1 : (n > 4.0e8 ? (size_t)400000000 : (size_t)n);   // NaN / negative / zero: one sample
  }
  int ins() const override { return 1; }
  int outs() const override { return 1; }
  void reset() override { std::fill(buf.begin(), buf.end(), 0.0f); i = 0; }
  void set_sr(double sr) override { buf.assign(length(t, sr), 0.0f); i = 0; }
  void tick(const float* in, float* out) override {
    out[0] = buf[i];
    buf[i] = in[0];
    i = i + 1 == buf.size() ? 0 : i + 1;
  }
  uint64_t id() const override { return ID_DELAY; }
  QO_CLONE(Delay)
};
// tap(min,max) cubic / tap_linear(min,max) [U]: write the current sample, then read `clamp(in[1])·sr` back
struct Tap : Unit {
  bool cubic;
  float mn, mx, sr = (float)DEFAULT_SR;
  std::vector<float> buf;
  size_t idx = 0;
  Tap(bool cubic_, float mn_, float mx_) : cubic(cubic_), mn(mn_), mx(mx_) { set_sr(DEFAULT_SR); }
  static size_t length(float mx, double sr) {
    size_t need = (size_t)std::ceil((double)mx * sr) + 4, n = 4;
    while (n < need) n <<= 1;
    return n;
  }
  int ins() const override { return 2; }
  int outs() const override { return 1; }
  void reset() override { std::fill(buf.begin(), buf.end(), 0.0f); idx = 0; }
  void set_sr(double s) override { sr = (float)s; buf.assign(length(mx, s), 0.0f); idx = 0; }
  void tick(const float* in, float* out) override {
    size_t mask = buf.size() - 1;
    buf[idx] = in[0];
    float tap = rclamp(in[1], mn, mx) * sr;
    if (tap != tap) tap = 0.0f;
    size_t fl = (size_t)tap;
    float d = tap - (float)fl;
    size_t i1 = (idx + buf.size() - fl) & mask;
    if (cubic) {
      size_t i0 = (i1 + 1) & mask, i2 = (i1 + buf.size() - 1) & mask, i3 = (i1 + buf.size() - 2) & mask;
      out[0] = spline(buf[i0], buf[i1], buf[i2], buf[i3], d);
    } else {
      size_t i2 = (i1 + buf.size() - 1) & mask;
      out[0] = lerp(buf[i1], buf[i2], d);
    }
    idx = (idx + 1) & mask;
  }
  uint64_t id() const override { return cubic ? ID_TAP : ID_TAPLIN; }
  QO_CLONE(Tap)
};
// SampDelay [P] nodes.rs:707-738
struct SampDelay : Unit {
  std::deque<float> buf;
  size_t mx;
  explicit SampDelay(size_t mx_) : buf(mx_, 0.0f), mx(mx_) {}
  int ins() const override { return 2; }
  int outs() const override { return 1; }
  void reset() override { buf.assign(mx, 0.0f); }
  void tick(const float* in, float* out) override {
    buf.push_front(in[0]);
    buf.pop_back();
    uint64_t i = as_usize(in[1]);
    out[0] = i < buf.size() ? buf[i] : 0.0f;
  }
  uint64_t id() const override { return ID_SAMPDELAY; }
  QO_CLONE(SampDelay)
};

// ------------------------------------------------------------------ envelopes: lfo / lfo_in [U sampling, P shapes]
// shape 0 xd, 1 xD, 2 ar, 3 t (closures at functions.rs:505-507, 517-540, 547-576, 811); the control
// function is sampled every ~2 ms (jittered 0.75..1.25) and linearly interpolated in between.
struct Envelope : Unit {
  int shape, nconst, nin;
  float c[4];
  float t = 0, t0 = 0, t1 = 0, v0 = 0, v1 = 0, sd = (float)(1.0 / DEFAULT_SR);
  uint64_t hash = 0, base = 0, thash = 0;
  bool first = true;
  Envelope(int shape_, int nin_, const std::vector<float>& cs) : shape(shape_), nconst((int)cs.size()), nin(nin_) {
    for (int i = 0; i < 4; i++) c[i] = i < nconst ? cs[i] : 0.0f;
    reset();
  }
  float eval(float tt, const float* in) const {
    switch (shape) {
      case 0: { float p = nin ? in[0] : c[0]; return std::exp(-tt * p); }
      case 1: {
        float d = nin >= 1 ? in[0] : c[0];
        float k = nin == 2 ? in[1] : (nin == 1 ? c[0] : c[1]);
        return tt < d ? std::pow((d - tt) / d, k) : 0.0f;
      }
      case 2: {
        float a, ak, r, rk;
        if (nin == 0) { a = c[0]; ak = c[1]; r = c[2]; rk = c[3]; }
        else if (nin == 2) { a = in[0]; r = in[1]; ak = c[0]; rk = c[1]; }
        else { a = in[0]; ak = in[1]; r = in[2]; rk = in[3]; }
        if (tt < a) return std::pow(tt / a, ak);
        if (tt < a + r) return std::pow((r - (tt - a)) / r, rk);
        return 0.0f;
      }
      default: return tt;
    }
  }
  int ins() const override { return nin; }
  int outs() const override { return 1; }
  void reset() override {
    t = t0 = t1 = 0; thash = hash; first = true;
    if (nin == 0) { v0 = v1 = eval(0.0f, nullptr); first = false; } else { v0 = v1 = 0; }
  }
  void set_sr(double sr) override { sd = (float)(1.0 / sr); }
  void set_hash(uint64_t h) override { base = hash = h; thash = h; }
  void salt(uint64_t s) override { hash = s ? atto(base, s) : base; thash = hash; }
  void tick(const float* in, float* out) override {
    if (t >= t1) {
      if (first) { v1 = eval(0.0f, in); first = false; }
      t0 = t1;
      v0 = v1;
      float next = lerp(0.75f, 1.25f, (float)rnd1(thash)) * 0.002f;
      t1 = t0 + next;
      v1 = eval(t1, in);
      thash += 1;
    }
    float u = delerp(t0, t1, t);
    t += sd;
    out[0] = lerp(v0, v1, u);
  }
  uint64_t id() const override { return nin ? ID_ENVELOPE_IN : ID_ENVELOPE; }
  QO_CLONE(Envelope)
};

// declick_s(t) [U]: fade in with smooth5 over t seconds
struct Declick : Unit {
  float dur, t = 0, sd = (float)(1.0 / DEFAULT_SR);
  explicit Declick(float d) : dur(d) {}
  int ins() const override { return 1; }
  int outs() const override { return 1; }
  void reset() override { t = 0; }
  void set_sr(double sr) override { sd = (float)(1.0 / sr); }
  void tick(const float* in, float* out) override {
    if (t < dur) {
      out[0] = in[0] * smooth5(t / dur);
      t += sd;
    } else out[0] = in[0];
  }
  uint64_t id() const override { return ID_DECLICK; }
  QO_CLONE(Declick)
};

// ------------------------------------------------------------------ in-tree nodes [P] (nodes.rs)
struct ShiftReg : Unit {   // nodes.rs:157-190
  float reg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int ins() const override { return 2; }
  int outs() const override { return 8; }
  void reset() override { for (float& r : reg) r = 0; }
  void tick(const float* in, float* out) override {
    if (in[1] != 0.0f) {
      for (int i = 7; i > 0; i--) reg[i] = reg[i - 1];
      reg[0] = in[0];
    }
    for (int i = 0; i < 8; i++) out[i] = reg[i];
  }
  uint64_t id() const override { return ID_SHIFTREG; }
  QO_CLONE(ShiftReg)
};
struct SnH : Unit {   // nodes.rs:795-821
  float val = 0;
  int ins() const override { return 2; }
  int outs() const override { return 1; }
  void reset() override { val = 0; }
  void tick(const float* in, float* out) override {
    if (in[1] != 0.0f) val = in[0];
    out[0] = val;
  }
  uint64_t id() const override { return ID_SNH; }
  QO_CLONE(SnH)
};
struct Quantizer : Unit {   // nodes.rs:196-229
  std::vector<float> arr;
  float range;
  Quantizer(std::vector<float> a, float r) : arr(std::move(a)), range(r) {}
  int ins() const override { return 1; }
  int outs() const override { return 1; }
  void tick(const float* in, float* out) override {
    float n = in[0];
    float wrapped = n - range * std::floor(n / range);
    float nearest = 0.0f, dist = std::numeric_limits<float>::max();
    for (float i : arr) {
      float d = std::fabs(wrapped - i);
      if (d < dist) { nearest = i; dist = d; }
    }
    out[0] = n + nearest - wrapped;
  }
  uint64_t id() const override { return ID_QUANTIZER; }
  QO_CLONE(Quantizer)
};
struct ArrGet : Unit {   // nodes.rs:127-150
  std::vector<float> arr;
  explicit ArrGet(std::vector<float> a) : arr(std::move(a)) {}
  int ins() const override { return 1; }
  int outs() const override { return 1; }
  void tick(const float* in, float* out) override {
    uint64_t i = as_usize(in[0]);
    out[0] = i < arr.size() ? arr[i] : 0.0f;
  }
  uint64_t id() const override { return ID_ARRGET; }
  QO_CLONE(ArrGet)
};
// live-I/O nodes have no offline meaning: every try_recv() misses -> 0.0 (nodes.rs:516-517, 786); BuffIn passes
struct ZeroSource : Unit {
  int no;
  uint64_t uid;
  ZeroSource(int no_, uint64_t uid_) : no(no_), uid(uid_) {}
  int ins() const override { return 0; }
  int outs() const override { return no; }
  void tick(const float*, float* out) override { for (int i = 0; i < no; i++) out[i] = 0.0f; }
  uint64_t id() const override { return uid; }
  QO_CLONE(ZeroSource)
};

// looping sample player: wavech(wave, 0, Some(0)) (process.rs:1658-1662) [U]
struct WavePlayer : Unit {
  std::vector<float> data;
  size_t idx = 0;
  explicit WavePlayer(std::vector<float> d) : data(std::move(d)) {}
  int ins() const override { return 0; }
  int outs() const override { return 1; }
  void reset() override { idx = 0; }
  void tick(const float*, float* out) override {
    if (data.empty()) { out[0] = 0.0f; return; }
    out[0] = data[idx];
    idx += 1;
    if (idx >= data.size()) idx = 0;
  }
  uint64_t id() const override { return ID_WAVE; }
  QO_CLONE(WavePlayer)
};

// ------------------------------------------------------------------ FFT (microfft-style radix-2, f32) [U scaling]
struct Cpx { float re, im; };
void fft_inplace(std::vector<Cpx>& a, bool inverse);   // forward unscaled, inverse scaled by 1/n
void real_fft(const std::vector<float>& in, std::vector<Cpx>& out);   // n -> n/2+1
void inverse_fft(const std::vector<Cpx>& in, std::vector<Cpx>& out);  // n -> n

struct Rfft : Unit {   // nodes.rs:601-649
  size_t n, count, start;
  std::vector<float> input;
  std::vector<Cpx> output;
  Rfft(size_t n_, size_t s) : n(n_), count(s), start(s), input(n_, 0.0f), output(n_ / 2 + 1, Cpx{0, 0}) {}
  int ins() const override { return 1; }
  int outs() const override { return 2; }
  void reset() override {
    count = start;
    std::fill(input.begin(), input.end(), 0.0f);
    std::fill(output.begin(), output.end(), Cpx{0, 0});
  }
  void tick(const float* in, float* out) override {
    size_t i = count;
    count += 1;
    if (count == n) count = 0;
    if (i == 0) real_fft(input, output);
    input[i] = in[0];
    if (i <= n / 2) { out[0] = output[i].re; out[1] = output[i].im; }
    else { out[0] = output[n - i].re; out[1] = -output[n - i].im; }
  }
  uint64_t id() const override { return ID_RFFT; }
  QO_CLONE(Rfft)
};
struct Ifft : Unit {   // nodes.rs:657-700
  size_t n, count, start;
  std::vector<Cpx> input, output;
  Ifft(size_t n_, size_t s) : n(n_), count(s), start(s), input(n_, Cpx{0, 0}), output(n_, Cpx{0, 0}) {}
  int ins() const override { return 2; }
  int outs() const override { return 2; }
  void reset() override {
    count = start;
    std::fill(input.begin(), input.end(), Cpx{0, 0});
    std::fill(output.begin(), output.end(), Cpx{0, 0});
  }
  void tick(const float* in, float* out) override {
    size_t i = count;
    count += 1;
    if (count == n) count = 0;
    if (i == 0) inverse_fft(input, output);
    input[i] = Cpx{in[0], in[1]};
    out[0] = output[i].re;
    out[1] = output[i].im;
  }
  uint64_t id() const override { return ID_IFFT; }
  QO_CLONE(Ifft)
};

}  // namespace qo
