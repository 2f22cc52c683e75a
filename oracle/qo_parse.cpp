// TEST INFRASTRUCTURE — CPU oracle (see qo_math.h header).
// str_to_net restated from /root/reference/src/functions.rs:111-1226 (grammar :112-127, constants :47-109).
// Ops whose FunDSP implementation is not restated return status 1 ("exists in the reference, unsupported
// here") instead of the reference's silent Net::new(0,0), so that a gap is never mistaken for parity.
#include <mutex>
#include <cstdlib>
#include <map>

#include "qo_net.h"

namespace qo {

// ------------------------------------------------------------------ FFT
void fft_inplace(std::vector<Cpx>& a, bool inverse) {
  size_t n = a.size();
  for (size_t i = 1, j = 0; i < n; i++) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    for (size_t k = 0; k < len / 2; k++) {
      double ang = -2.0 * 3.14159265358979323846 * (double)k / (double)len;
      float wr = (float)std::cos(ang), wi = (float)std::sin(ang);
      if (inverse) wi = -wi;
      for (size_t i = k; i < n; i += len) {
        Cpx u = a[i], v = a[i + len / 2];
        float tr = v.re * wr - v.im * wi, ti = v.re * wi + v.im * wr;
        a[i] = Cpx{u.re + tr, u.im + ti};
        a[i + len / 2] = Cpx{u.re - tr, u.im - ti};
      }
    }
  }
  if (inverse) {
    float s = 1.0f / (float)n;
    for (Cpx& c : a) { c.re *= s; c.im *= s; }
  }
}
void real_fft(const std::vector<float>& in, std::vector<Cpx>& out) {
  size_t n = in.size();
  std::vector<Cpx> a(n);
  for (size_t i = 0; i < n; i++) a[i] = Cpx{in[i], 0.0f};
  fft_inplace(a, false);
  for (size_t i = 0; i <= n / 2; i++) out[i] = a[i];
}
void inverse_fft(const std::vector<Cpx>& in, std::vector<Cpx>& out) {
  std::vector<Cpx> a(in);
  fft_inplace(a, true);
  out = a;
}

// ------------------------------------------------------------------ wavetables (FunDSP wavetable.rs) [U]
static std::vector<float> make_wave(double pitch, int shape) {
  size_t harmonics = (size_t)std::floor(22050.0 / pitch);
  size_t target = 4 * harmonics, length = 32;
  while (length < target && length < 8192) length <<= 1;
  std::vector<Cpx> a(length, Cpx{0, 0});
  for (size_t i = 1; i <= harmonics && i < length / 2; i++) {
    double f = pitch * (double)i, w;
    bool odd = (i & 1) != 0;
    switch (shape) {
      case 0: w = 1.0 / (double)i; break;
      case 1: w = odd ? 1.0 / (double)i : 0.0; break;
      case 2: w = odd ? 1.0 / ((double)i * (double)i) : 0.0; break;
      default: w = 1.0 / ((double)i * (double)i); break;
    }
    double fade = (22050.0 - f) / (22050.0 - 20000.0);
    fade = fade < 0 ? 0 : (fade > 1 ? 1 : fade);
    w *= ((6.0 * fade - 15.0) * fade + 10.0) * fade * fade * fade;   // smooth5
    double ph = (shape == 2 && (i & 3) == 3) ? 0.5 : 0.0;
    // a sine partial of amplitude w: X[i] = -i w/2 e^{i 2 pi ph}, X[N-i] = conj
    double re = 0.5 * w * std::sin(2.0 * 3.14159265358979323846 * ph), im = -0.5 * w * std::cos(2.0 * 3.14159265358979323846 * ph);
    a[i] = Cpx{(float)re, (float)im};
    a[length - i] = Cpx{(float)re, (float)-im};
  }
  fft_inplace(a, true);
  std::vector<float> wave(length);
  float mx = 0.0f;
  for (size_t k = 0; k < length; k++) { wave[k] = a[k].re * (float)length; mx = std::fmax(mx, std::fabs(wave[k])); }
  if (mx > 0.0f) for (float& x : wave) x /= mx;
  return wave;
}
const WaveTableSet& wavetable_set(int shape) {
  // built once per shape; render_bank() ticks voices from several threads, so the lazy build must be thread-safe
  static WaveTableSet sets[4];
  static std::once_flag once[4];
  WaveTableSet& ts = sets[shape & 3];
  std::call_once(once[shape & 3], [&ts, shape] {
    for (int i = 0;; i++) {
      double pn = 20.0 * std::pow(2.0, (double)(i + 1) / 4.0);
      ts.limit.push_back((float)pn);
      ts.table.push_back(make_wave(pn, shape));
      if (pn >= 20000.0) break;
    }
  });
  return ts;
}

// ------------------------------------------------------------------ parse_with_constants (functions.rs:47-109)
static bool parse_f32(const std::string& s, float* out) {
  if (s.empty()) return false;
  // Rust f32::from_str: optional sign, digits with optional '.', optional exponent, or inf/infinity/nan
  std::string t = s;
  const char* p = t.c_str();
  size_t i = 0;
  if (p[i] == '+' || p[i] == '-') i++;
  std::string rest = t.substr(i);
  std::string low;
  for (char c : rest) low.push_back((char)std::tolower((unsigned char)c));
  if (low == "inf" || low == "infinity" || low == "nan") {
    *out = low == "nan" ? std::numeric_limits<float>::quiet_NaN() : std::numeric_limits<float>::infinity();
    if (p[0] == '-') *out = -*out;
    return true;
  }
  bool digits = false, dot = false;
  size_t j = i;
  while (p[j]) {
    if (p[j] >= '0' && p[j] <= '9') digits = true;
    else if (p[j] == '.' && !dot) dot = true;
    else break;
    j++;
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
  static const std::map<std::string, float> K = {
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
  auto it = K.find(neg ? s.substr(1) : s);
  if (it != K.end()) { *out = neg ? -it->second : it->second; return true; }
  if (s == "MAX") { *out = std::numeric_limits<float>::max(); return true; }
  if (s == "MIN") { *out = std::numeric_limits<float>::lowest(); return true; }
  if (s == "EPSILON") { *out = std::numeric_limits<float>::epsilon(); return true; }
  if (s == "MIN_POSITIVE") { *out = std::numeric_limits<float>::min(); return true; }
  return false;
}

// ------------------------------------------------------------------ helpers
typedef std::function<void(const float*, float*)> Fn;
static UnitP U(Unit* u) { return UnitP(u); }
static UnitP map1(std::function<float(float)> f) {
  return U(new Map(1, 1, [f](const float* i, float* o) { o[0] = f(i[0]); }));
}
static UnitP map2(std::function<float(float, float)> f) {
  return U(new Map(2, 1, [f](const float* i, float* o) { o[0] = f(i[0], i[1]); }));
}
static UnitP map3(std::function<float(float, float, float)> f) {
  return U(new Map(3, 1, [f](const float* i, float* o) { o[0] = f(i[0], i[1], i[2]); }));
}
static UnitP pipe(UnitP a, UnitP b) { return U(new Comp(0, std::move(a), std::move(b))); }
[[maybe_unused]] static UnitP stack(UnitP a, UnitP b) { return U(new Comp(1, std::move(a), std::move(b))); }
static UnitP branch(UnitP a, UnitP b) { return U(new Comp(2, std::move(a), std::move(b))); }
static UnitP constant(std::vector<float> v) { return U(new Constant(std::move(v))); }
static NetP W(UnitP u) { return Net::wrap(std::move(u)); }
static NetP W(Unit* u) { return Net::wrap(UnitP(u)); }

// N-channel `x op constants` (FunDSP add/sub/mul((c0..cN-1))): N inputs, N outputs
static NetP nary_const(const std::vector<float>& p, char op, bool recip) {
  std::vector<float> c;
  for (size_t i = 0; i < p.size() && i < 8; i++) c.push_back(recip ? 1.0f / p[i] : p[i]);
  if (c.empty()) c.push_back(1.0f);
  int n = (int)c.size();
  return W(new Map(n, n, [c, op, n](const float* in, float* out) {
    for (int i = 0; i < n; i++) out[i] = op == '+' ? in[i] + c[i] : op == '-' ? in[i] - c[i] : in[i] * c[i];
  }));
}
// comparison / binary closure with "one param => constant rhs, else second input" (functions.rs:824-935)
static NetP bin_or_const(const std::vector<float>& p, std::function<float(float, float)> f) {
  if (!p.empty()) {
    float c = p[0];
    return W(map1([f, c](float a) { return f(a, c); }));
  }
  return W(map2(f));
}
static NetP tern_or_const(const std::vector<float>& p, std::function<float(float, float, float)> f) {
  if (p.size() >= 2) {
    float a = p[0], b = p[1];
    return W(map1([f, a, b](float t) { return f(a, b, t); }));
  }
  return W(map3(f));
}
static NetP svf(int mode, const std::vector<float>& p) {
  int npar = mode >= 6 ? 3 : 2;
  // lowpass(): all inputs; lowpass(q): hz input; lowpass(hz,q): fixed.  bell(q,gain) / bell(hz,q,gain).
  if ((int)p.size() >= npar) return W(new Svf(mode, npar, p[0], p[1], npar == 3 ? p[2] : 1.0f));
  if ((int)p.size() >= npar - 1) {
    if (npar == 2) return W(new Svf(mode, 1, 440.0f, p[0], 1.0f));
    return W(new Svf(mode, 2, 440.0f, p[0], p[1]));
  }
  return W(new Svf(mode, 0, 440.0f, 1.0f, 1.0f));
}
static const int UNSUPPORTED = 1;

NetP str_to_net(const std::string& op_in, int* status) {
  *status = 0;
  std::string op;
  for (char c : op_in) if (c != ' ') op.push_back(c);
  // split on '(' and ')'
  std::vector<std::string> args(1);
  for (char c : op) {
    if (c == '(' || c == ')') args.emplace_back();
    else args.back().push_back(c);
  }
  if (args.size() < 2) return NetP(new Net(0, 0));   // no parentheses
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
      if (parse_with_constants(t, &v)) p.push_back(v);
    }
  }
  const std::string& name = args[0];
  auto has = [&](size_t n) { return p.size() >= n; };
  auto usz = [&](float x) { return as_usize(x); };

  // -------------------- sources
  if (name == "sine") {
    if (has(1)) return W(pipe(constant({p[0]}), U(new Sine())));
    return W(new Sine());
  }
  if (name == "white" || name == "noise") return W(new Noise());
  if (name == "brown") {   // (white() >> lowpole_hz(10)) * dc(13.7)
    UnitP a = pipe(U(new Noise()), U(new OnePole(0, false, 10.0f)));
    return W(U(new Comp(6, std::move(a), constant({13.7f}))));
  }
  if (name == "pink") return W(pipe(U(new Noise()), U(new Pinkpass())));
  if (name == "zero") return W(constant({0.0f}));
  if (name == "impulse") return W(new Impulse());
  if (name == "constant" || name == "dc") {
    std::vector<float> v(p.begin(), p.begin() + std::min<size_t>(p.size(), 8));
    if (v.empty()) v.push_back(1.0f);
    return W(constant(v));
  }
  if (name == "ramp") return W(new Ramp());
  if (name == "saw" || name == "square" || name == "triangle" || name == "soft_saw") {
    int shape = name == "saw" ? 0 : name == "square" ? 1 : name == "triangle" ? 2 : 3;
    if (has(1)) return W(pipe(constant({p[0]}), U(new WaveSynth(shape))));
    return W(new WaveSynth(shape));
  }
  if (name == "organ" || name == "hammond" || name == "pulse" || name == "lorenz" || name == "rossler" || name == "dsf_saw" ||
      name == "dsf_square" || name == "mls") { *status = UNSUPPORTED; return nullptr; }
  if (name == "pluck") { if (has(3)) { *status = UNSUPPORTED; return nullptr; } return NetP(new Net(0, 0)); }

  // -------------------- filters
  if (name == "lowpass") return svf(0, p);
  if (name == "highpass") return svf(1, p);
  if (name == "bandpass") return svf(2, p);
  if (name == "notch") return svf(3, p);
  if (name == "peak") return svf(4, p);
  if (name == "allpass") return svf(5, p);
  if (name == "bell") return svf(6, p);
  if (name == "lowshelf") return svf(7, p);
  if (name == "highshelf") return svf(8, p);
  if (name == "biquad") {
    if (has(5)) return W(Biquad::fixed(p[0], p[1], p[2], p[3], p[4]));
    return NetP(new Net(0, 0));
  }
  if (name == "butterpass") return has(1) ? W(new Biquad(1, 0, p[0], 0)) : W(new Biquad(1, 1, 440.0f, 0));
  if (name == "resonator") return has(2) ? W(new Biquad(2, 0, p[0], p[1])) : W(new Biquad(2, 2, 440.0f, 110.0f));
  if (name == "lowpole") return has(1) ? W(new OnePole(0, false, p[0])) : W(new OnePole(0, true, 440.0f));
  if (name == "highpole") return has(1) ? W(new OnePole(1, false, p[0])) : W(new OnePole(1, true, 440.0f));
  if (name == "dcblock") return W(new OnePole(2, false, has(1) ? p[0] : 10.0f));
  if (name == "allpole") return has(1) ? W(new OnePole(3, false, p[0])) : W(new OnePole(3, true, 1.0f));
  if (name == "pinkpass") return W(new Pinkpass());
  if (name == "fir") {
    if (!has(1)) return NetP(new Net(0, 0));
    return W(new Fir(std::vector<float>(p.begin(), p.begin() + std::min<size_t>(p.size(), 10))));
  }
  if (name == "fir3") {
    if (!has(1)) return NetP(new Net(0, 0));
    float alpha = (p[0] + 1.0f) / 2.0f, beta = (1.0f - alpha) / 2.0f;
    return W(new Fir({beta, alpha, beta}));
  }
  if (name == "follow" || name == "moog" || name == "morph" || name == "lowrez" || name == "bandrez") {
    if (name == "follow" && !has(1)) return NetP(new Net(0, 0));
    *status = UNSUPPORTED; return nullptr;
  }

  // -------------------- channels
  if (name == "sink") return W(new Sink());
  if (name == "pass") return W(new Pass());
  if (name == "chan") {
    NetP net(new Net(0, 0));
    for (float v : p) net = Net::combine('|', std::move(net), v == 0.0f ? W(new Sink()) : W(new Pass()));
    return net;
  }
  if (name == "pan") return has(1) ? W(new Pan(true, p[0])) : W(new Pan(false, 0.0f));
  if (name == "join" || name == "split" || name == "reverse") {
    if (has(1)) {
      uint64_t n = usz(p[0]);
      if (n >= 2 && n <= 8) {
        if (name == "join") return W(new Join((int)n));
        if (name == "split") return W(new Split((int)n));
        return W(new Reverse((int)n));
      }
    }
    return NetP(new Net(0, 0));
  }

  // -------------------- envelopes
  if (name == "adsr") { if (has(4)) { *status = UNSUPPORTED; return nullptr; } return NetP(new Net(0, 0)); }
  if (name == "xd") return has(1) ? W(new Envelope(0, 0, {p[0]})) : W(new Envelope(0, 1, {}));
  if (name == "xD") {
    if (has(2)) return W(new Envelope(1, 0, {p[0], p[1]}));
    if (has(1)) return W(new Envelope(1, 1, {p[0]}));
    return W(new Envelope(1, 2, {}));
  }
  if (name == "ar") {
    if (has(4)) return W(new Envelope(2, 0, {p[0], p[1], p[2], p[3]}));
    if (has(2)) return W(new Envelope(2, 2, {p[0], p[1]}));
    return W(new Envelope(2, 4, {}));
  }

  // -------------------- other
  if (name == "tick") return W(new Tick());
  if (name == "shift_reg") return W(new ShiftReg());
  if (name == "snh") return W(new SnH());
  if (name == "meter" || name == "chorus" || name == "hold" || name == "limiter" || name == "limiter_stereo" ||
      name == "reverb_stereo" || name == "reverb_mono") {
    if (!has(1)) return NetP(new Net(0, 0));
    *status = UNSUPPORTED; return nullptr;
  }
  if (name == "clip") {
    float lo = -1.0f, hi = 1.0f;
    if (has(2)) { lo = p[0] < p[1] ? p[0] : p[1]; hi = p[0] < p[1] ? p[1] : p[0]; }
    return W(map1([lo, hi](float x) { return rclamp(x, lo, hi); }));
  }
  if (name == "declick") return W(new Declick(has(1) ? p[0] : 0.010f));
  if (name == "delay") return has(1) ? W(new Delay(p[0])) : NetP(new Net(0, 0));
  if (name == "tap" || name == "tap_linear") {
    if (!has(2)) return NetP(new Net(0, 0));
    float p0 = rmax(p[0], 0.0f), p1 = rmax(p[1], 0.0f);
    return W(new Tap(name == "tap", rmin(p0, p1), rmax(p0, p1)));
  }
  if (name == "samp_delay") return has(1) ? W(new SampDelay((size_t)usz(p[0]))) : NetP(new Net(0, 0));
  if (name == "pdhalf_bi") {   // functions.rs:677-688
    return W(map2([](float x, float m) {
      float mid = rclamp(m, -1.0f, 1.0f);
      if (x < mid) {
        float ls = mid != -1.0f ? 1.0f / (mid + 1.0f) : 0.0f;
        return ls * x;
      }
      float rs = mid != 1.0f ? 1.0f / (1.0f - mid) : 0.0f;
      return rs * (x - mid) + 0.5f;
    }));
  }
  if (name == "pdhalf_uni") {   // functions.rs:689-706
    return W(map2([](float x, float m) {
      float mid = m >= 1.0f ? 1.0f : (m <= -1.0f ? 0.0f : (m + 1.0f) / 2.0f);
      if (x < mid) {
        float ls = mid != 0.0f ? 0.5f / mid : 0.0f;
        return ls * x;
      }
      float rs = mid != 1.0f ? 0.5f / (1.0f - mid) : 0.0f;
      return rs * (x - mid) + 0.5f;
    }));
  }

  // -------------------- math
  if (name == "add") return nary_const(p, '+', false);
  if (name == "sub") return nary_const(p, '-', false);
  if (name == "mul") return nary_const(p, '*', false);
  if (name == "div") return nary_const(p, '*', true);
  if (name == "rotate") {
    if (!has(2)) return NetP(new Net(0, 0));
    float c = std::cos(p[0]) * p[1], s = std::sin(p[0]) * p[1];
    return W(new Map(2, 2, [c, s](const float* i, float* o) { o[0] = c * i[0] - s * i[1]; o[1] = s * i[0] + c * i[1]; }));
  }
  if (name == "t") return W(new Envelope(3, 0, {}));
  if (name == "rise" || name == "fall") {   // (pass() ^ tick()) >> map
    bool rise = name == "rise";
    UnitP m = map2([rise](float a, float b) { return (rise ? a > b : a < b) ? 1.0f : 0.0f; });
    return W(pipe(branch(U(new Pass()), U(new Tick())), std::move(m)));
  }
  if (name == ">") return bin_or_const(p, [](float a, float b) { return a > b ? 1.0f : 0.0f; });
  if (name == "<") return bin_or_const(p, [](float a, float b) { return a < b ? 1.0f : 0.0f; });
  if (name == "==") return bin_or_const(p, [](float a, float b) { return a == b ? 1.0f : 0.0f; });
  if (name == "!=") return bin_or_const(p, [](float a, float b) { return a != b ? 1.0f : 0.0f; });
  if (name == ">=") return bin_or_const(p, [](float a, float b) { return a >= b ? 1.0f : 0.0f; });
  if (name == "<=") return bin_or_const(p, [](float a, float b) { return a <= b ? 1.0f : 0.0f; });
  if (name == "min") return bin_or_const(p, [](float a, float b) { return rmin(a, b); });
  if (name == "max") return bin_or_const(p, [](float a, float b) { return rmax(a, b); });
  if (name == "pow") return bin_or_const(p, [](float a, float b) { return std::pow(a, b); });
  if (name == "mod" || name == "rem") return bin_or_const(p, [](float a, float b) { return rem_euclid(a, b); });
  if (name == "log") return bin_or_const(p, [](float a, float b) { return std::log(a) / std::log(b); });
  if (name == "bitand") return bin_or_const(p, [](float a, float b) { return (float)(as_i32(a) & as_i32(b)); });
  if (name == "bitor") return bin_or_const(p, [](float a, float b) { return (float)(as_i32(a) | as_i32(b)); });
  if (name == "bitxor") return bin_or_const(p, [](float a, float b) { return (float)(as_i32(a) ^ as_i32(b)); });
  if (name == "shl") return bin_or_const(p, [](float a, float b) {
    return (float)(int32_t)((uint32_t)as_i32(a) << (as_usize(b) & 31));
  });
  if (name == "shr") return bin_or_const(p, [](float a, float b) { return (float)(as_i32(a) >> (as_usize(b) & 31)); });
  if (name == "lerp") return tern_or_const(p, lerp);
  if (name == "lerp11") return tern_or_const(p, lerp11);
  if (name == "delerp") return tern_or_const(p, delerp);
  if (name == "delerp11") return tern_or_const(p, delerp11);
  if (name == "xerp") return tern_or_const(p, xerp);
  if (name == "xerp11") return tern_or_const(p, xerp11);
  if (name == "dexerp") return tern_or_const(p, dexerp);
  if (name == "dexerp11") return tern_or_const(p, dexerp11);

  static const std::map<std::string, float (*)(float)> unary = {
      {"abs", [](float x) { return std::fabs(x); }},
      {"signum", [](float x) { return signum(x); }},
      {"floor", [](float x) { return std::floor(x); }},
      {"fract", [](float x) { return fract(x); }},
      {"ceil", [](float x) { return std::ceil(x); }},
      {"round", [](float x) { return std::round(x); }},
      {"sqrt", [](float x) { return std::sqrt(x); }},
      {"exp", [](float x) { return std::exp(x); }},
      {"exp2", [](float x) { return std::exp2(x); }},
      {"exp10", [](float x) { return exp10f_(x); }},
      // bug-for-bug: functions.rs:1077-1078 swaps the two
      {"exp_m1", [](float x) { return std::log1p(x); }},
      {"ln_1p", [](float x) { return std::expm1(x); }},
      {"ln", [](float x) { return std::log(x); }},
      {"log2", [](float x) { return std::log2(x); }},
      {"log10", [](float x) { return std::log10(x); }},
      {"sin", [](float x) { return std::sin(x); }},
      {"cos", [](float x) { return std::cos(x); }},
      {"tan", [](float x) { return std::tan(x); }},
      {"asin", [](float x) { return std::asin(x); }},
      {"acos", [](float x) { return std::acos(x); }},
      {"atan", [](float x) { return std::atan(x); }},
      {"sinh", [](float x) { return std::sinh(x); }},
      {"cosh", [](float x) { return std::cosh(x); }},
      {"tanh", [](float x) { return std::tanh(x); }},
      {"asinh", [](float x) { return std::asinh(x); }},
      {"acosh", [](float x) { return std::acosh(x); }},
      {"atanh", [](float x) { return std::atanh(x); }},
      {"squared", [](float x) { return x * x; }},
      {"cubed", [](float x) { return x * x * x; }},
      {"db_amp", [](float x) { return db_amp(x); }},
      {"amp_db", [](float x) { return amp_db(x); }},
      {"a_weight", [](float x) { return a_weight(x); }},
      {"softsign", [](float x) { return softsign(x); }},
      {"smooth3", [](float x) { return smooth3(x); }},
      {"smooth5", [](float x) { return smooth5(x); }},
      {"smooth7", [](float x) { return smooth7(x); }},
      {"smooth9", [](float x) { return smooth9(x); }},
      {"uparc", [](float x) { return uparc(x); }},
      {"downarc", [](float x) { return downarc(x); }},
      {"sine_ease", [](float x) { return sine_ease(x); }},
      {"semitone_ratio", [](float x) { return semitone_ratio(x); }},
      {"rnd1", [](float x) { return (float)rnd1(as_usize(x)); }},
      {"rnd2", [](float x) { return (float)rnd2(as_usize(x)); }},
      {"deg", [](float x) { return x * 57.2957795130823208767981548141051703f; }},   // f32::to_degrees constant
      {"rad", [](float x) { return x * (PI_F / 180.0f); }},
      {"recip", [](float x) { return 1.0f / x; }},
      {"normal", [](float x) { return is_normal(x) ? x : 0.0f; }},
  };
  auto it = unary.find(name);
  if (it != unary.end()) { float (*f)(float) = it->second; return W(map1(f)); }
  if (name == "hypot") return W(map2([](float a, float b) { return std::hypot(a, b); }));
  if (name == "atan2") return W(map2([](float a, float b) { return std::atan2(a, b); }));
  if (name == "dissonance") return W(map2(dissonance));
  if (name == "sin_hz") return W(map2(sin_hz));
  if (name == "cos_hz") return W(map2(cos_hz));
  if (name == "sqr_hz") return W(map2(sqr_hz));
  if (name == "tri_hz") return W(map2(tri_hz));
  if (name == "spline")
    return W(new Map(5, 1, [](const float* i, float* o) { o[0] = spline(i[0], i[1], i[2], i[3], i[4]); }));
  if (name == "dissonance_max" || name == "m_weight" || name == "spline_mono" || name == "softexp" ||
      name == "softmix" || name == "spline_noise" || name == "fractal_noise") { *status = UNSUPPORTED; return nullptr; }
  if (name == "wrap") {   // functions.rs:1149-1162
    if (has(2)) {
      float p0 = rmin(p[0], p[1]), p1 = rmax(p[0], p[1]), r = p1 - p0;
      return W(map1([p0, r](float x) { return std::fmod(std::fmod(x - p0, r) + r, r) + p0; }));
    }
    if (has(1)) { float x0 = p[0]; return W(map1([x0](float x) { return x - x0 * std::floor(x / x0); })); }
    return NetP(new Net(0, 0));
  }
  if (name == "mirror") {   // functions.rs:1163-1182
    if (!has(2)) return NetP(new Net(0, 0));
    float p0 = rmin(p[0], p[1]), p1 = rmax(p[0], p[1]), r = p1 - p0;
    return W(map1([p0, p1, r](float x) {
      float n = is_normal(x) ? x : 0.0f;
      if (n >= p0 && n <= p1) return n;
      float distance = rmin(n - p1, p0 - n);
      float folds = std::floor(distance / r);
      if ((n > p1 && std::fmod(folds, 2.0f) == 0.0f) || (n < p0 && std::fmod(folds, 2.0f) != 0.0f))
        return p0 + (distance - folds * r);
      return p1 - (distance - folds * r);
    }));
  }
  if (name == "pol")
    return W(new Map(2, 2, [](const float* i, float* o) { o[0] = std::hypot(i[0], i[1]); o[1] = std::atan2(i[1], i[0]); }));
  if (name == "car")
    return W(new Map(2, 2, [](const float* i, float* o) { o[0] = i[0] * std::cos(i[1]); o[1] = i[0] * std::sin(i[1]); }));
  if (name == "rfft" || name == "ifft") {   // functions.rs:1196-1217
    if (!has(2)) return NetP(new Net(0, 0));
    uint64_t i = usz(p[0]);
    uint64_t x = std::min<uint64_t>(std::max<uint64_t>(i, 2), 32768), n = 1;
    while (n < x) n <<= 1;
    uint64_t start = std::min<uint64_t>(usz(p[1]), n - 1);
    if (name == "rfft") return W(new Rfft((size_t)n, (size_t)start));
    return W(new Ifft((size_t)n, (size_t)start));
  }
  return NetP(new Net(0, 0));
}

}  // namespace qo
