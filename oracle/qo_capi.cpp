// TEST INFRASTRUCTURE — CPU oracle (see qo_math.h header).  C entry points used by tests/ (ctypes), smoke()
// and bench.py's CPU-baseline legs.  The graph-construction calls restate the arms of the reference's
// patch interpreter that build audio graphs (/root/reference/src/process.rs:1311-1357, 1450-1907).
#include <atomic>
#include <cstring>
#include <thread>

#include "qo_net.h"

using namespace qo;

static thread_local std::string g_err;
static Net* N(void* h) { return static_cast<Net*>(h); }
static NetP copy(void* h) { return NetP(new Net(*N(h))); }

extern "C" {

const char* qo_last_error() { return g_err.c_str(); }

// str_to_net (functions.rs:111).  status: 0 ok, 1 op known to the reference but not restated in the oracle.
void* qo_str_to_net(const char* op, int* status) {
  int st = 0;
  NetP n = str_to_net(op, &st);
  if (status) *status = st;
  if (st) { g_err = std::string("oracle: unsupported op: ") + op; return nullptr; }
  return n.release();
}
void* qo_net_new(int ni, int no) { return new Net(ni, no); }
void* qo_clone(void* h) { return copy(h).release(); }
void qo_free(void* h) { delete N(h); }
int qo_inputs(void* h) { return N(h)->ins(); }
int qo_outputs(void* h) { return N(h)->outs(); }
int qo_size(void* h) { return N(h)->size(); }
void qo_reset(void* h) { N(h)->reset(); }
void qo_set_sample_rate(void* h, double sr) { N(h)->set_sr(sr); }   // sr(): process.rs:1571-1573
void qo_set_salt(void* h, uint64_t salt) { N(h)->salt(salt); }

// Connective ops "+ * - >> | & ^ !" (process.rs:1719-1876): inputs already ordered by link index, `number`
// is the circle's Number (repeat count), node_limit is NodeLimit (main.rs:72).
void* qo_connect(const char* opname, void** nets, int n_nets, double number, int node_limit) {
  std::string op(opname);
  if (op == "!") {   // process.rs:1868-1873
    NetP g = n_nets > 0 && nets[0] ? copy(nets[0]) : NetP(new Net(0, 0));
    return Net::thru(std::move(g)).release();
  }
  if (op == "-") {   // process.rs:1787-1797 (no result => the circle keeps its previous, here empty, net)
    if (n_nets >= 2 && nets[0] && nets[1]) {
      NetP l = copy(nets[0]), r = copy(nets[1]);
      if (l->outs() == r->outs()) {
        NetP g = Net::combine('-', std::move(l), std::move(r));
        if (g->size() < node_limit) return g.release();
      }
    }
    return new Net(0, 0);
  }
  NetP graph(new Net(0, 0));
  bool empty = true;
  float nf = (float)number;
  int reps = as_i32(rmax(nf, 1.0f));
  for (int r = 0; r < reps; r++) {
    for (int i = 0; i < n_nets; i++) {
      if (!nets[i]) continue;
      NetP net = copy(nets[i]);
      if (empty) { graph = std::move(net); empty = false; continue; }
      int gi = graph->ins(), go = graph->outs(), ni = net->ins(), no = net->outs();
      if (op == "+" || op == "*") {   // process.rs:1751-1759
        if (go == no) {
          if (graph->size() >= node_limit) continue;
          graph = Net::combine(op[0], std::move(graph), std::move(net));
        }
      } else {   // process.rs:1833-1844
        if (graph->size() >= node_limit) continue;
        if (op == ">>") { if (go == ni) graph = Net::combine('>', std::move(graph), std::move(net)); }
        else if (op == "|") graph = Net::combine('|', std::move(graph), std::move(net));
        else if (op == "&") { if (gi == ni && go == no) graph = Net::combine('&', std::move(graph), std::move(net)); }
        else if (op == "^") { if (gi == ni) graph = Net::combine('^', std::move(graph), std::move(net)); }
      }
    }
  }
  return graph.release();
}

// branch() bus() pipe() stack() sum() product() (process.rs:1669-1717): op string with '#' substituted per array
// element ("{}" formatting of f32), no node limit, no repeat.
static std::string fmt_f32(float v) {
  // Rust `{}` for f32: shortest repr that round-trips, no exponent, integers print without ".0"
  char buf[64];
  for (int prec = 1; prec < 12; prec++) {
    snprintf(buf, sizeof buf, "%.*g", prec, (double)v);
    if (strtof(buf, nullptr) == v) break;
  }
  std::string s(buf);
  if (s.find('e') != std::string::npos) { snprintf(buf, sizeof buf, "%.10f", (double)v); s = buf;
    while (s.find('.') != std::string::npos && (s.back() == '0' || s.back() == '.')) { bool dot = s.back() == '.'; s.pop_back(); if (dot) break; } }
  return s;
}
void* qo_array_op(const char* kind, const char* op_str, const float* arr, int n, int* status) {
  std::string k(kind);
  NetP graph(new Net(0, 0));
  bool empty = true;
  if (status) *status = 0;
  for (int i = 0; i < n; i++) {
    std::string s(op_str), r;
    std::string num = fmt_f32(arr[i]);
    for (char c : s) { if (c == '#') r += num; else r.push_back(c); }
    int st = 0;
    NetP net = str_to_net(r, &st);
    if (st) { if (status) *status = st; g_err = "oracle: unsupported op: " + r; return nullptr; }
    if (empty) { graph = std::move(net); empty = false; continue; }
    int gi = graph->ins(), go = graph->outs(), ni = net->ins(), no = net->outs();
    if (k == "branch()") { if (gi == ni) graph = Net::combine('^', std::move(graph), std::move(net)); }
    else if (k == "bus()") { if (gi == ni && go == no) graph = Net::combine('&', std::move(graph), std::move(net)); }
    else if (k == "pipe()") { if (go == ni) graph = Net::combine('>', std::move(graph), std::move(net)); }
    else if (k == "stack()") graph = Net::combine('|', std::move(graph), std::move(net));
    else if (k == "sum()") { if (go == no) graph = Net::combine('+', std::move(graph), std::move(net)); }
    else if (k == "product()") { if (go == no) graph = Net::combine('*', std::move(graph), std::move(net)); }
  }
  return graph.release();
}

// get() quantize() wave() (process.rs:1450-1477, 1652-1667)
void* qo_get(const float* arr, int n) { return Net::wrap(UnitP(new ArrGet(std::vector<float>(arr, arr + n)))).release(); }
void* qo_quantize(const float* arr, int n) {
  if (n < 1) return new Net(0, 0);
  float range = arr[n - 1] - arr[0];
  return Net::wrap(UnitP(new Quantizer(std::vector<float>(arr, arr + n), range))).release();
}
void* qo_wave(const float* arr, int n) { return Net::wrap(UnitP(new WavePlayer(std::vector<float>(arr, arr + n)))).release(); }
// feedback() (process.rs:1479-1515)
void* qo_feedback(void* h, int has_delay, double delay) {
  if (!h) return new Net(0, 0);
  NetP net = copy(h);
  if (net->outs() != net->ins()) return new Net(0, 0);
  float d = has_delay ? (float)delay : 0.0f;
  return Net::wrap(UnitP(new Feedback((double)d, std::move(net)))).release();
}
// kr() s() reset() (process.rs:1556-1580)
void* qo_kr(void* h, double n, int preserve_time) {
  if (!h) return new Net(0, 0);
  return Net::wrap(UnitP(new Kr(copy(h), (size_t)as_usize(rmax((float)n, 1.0f)), preserve_time != 0))).release();
}
void* qo_reset_every(void* h, double s) {
  if (!h) return new Net(0, 0);
  NetP net = copy(h);
  if (!(net->ins() == 0 && net->outs() == 1)) return new Net(0, 0);
  return Net::wrap(UnitP(new Resetter(0, std::move(net), (float)s))).release();
}
// trig_reset() reset_v() (process.rs:1597-1612)
void* qo_trig_reset(void* h, int variable) {
  if (!h) return new Net(0, 0);
  NetP net = copy(h);
  if (!(net->ins() == 0 && net->outs() == 1)) return new Net(0, 0);
  return Net::wrap(UnitP(new Resetter(variable ? 2 : 1, std::move(net), 0.0f))).release();
}
// seq() select() (process.rs:1634-1649): only 0-in/1-out nets are kept
void* qo_seq_select(int is_seq, void** nets, int n) {
  std::vector<NetP> v;
  for (int i = 0; i < n; i++)
    if (nets[i] && N(nets[i])->ins() == 0 && N(nets[i])->outs() == 1) v.push_back(copy(nets[i]));
  if (is_seq) return Net::wrap(UnitP(new Seq(std::move(v)))).release();
  return Net::wrap(UnitP(new Select(std::move(v)))).release();
}
// live I/O nodes offline: in()/adc() -> 2 silent channels; buffout() -> 1 silent channel; buffin() -> pass
void* qo_live_io(const char* name) {
  std::string s(name);
  if (s == "in()" || s == "adc()") return Net::wrap(UnitP(new ZeroSource(2, ID_INPUT))).release();
  if (s == "buffout()") return Net::wrap(UnitP(new ZeroSource(1, ID_BUFFOUT))).release();
  if (s == "monitor()") return Net::wrap(UnitP(new Pass())).release();
  if (s == "buffin()") return Net::wrap(UnitP(new Map(1, 1, [](const float* i, float* o) { o[0] = i[0]; }))).release();
  return new Net(0, 0);
}

// var() (process.rs:1373-1385): Shared-backed constant holding the circle's Number
void* qo_var(float value) { return Net::wrap(UnitP(new Constant({value}))).release(); }

// apply (process.rs:1322-1325): one frame
int qo_tick(void* h, const float* in, int n_in, float* out, int n_out) {
  Net* n = N(h);
  if (n->ins() != n_in || n->outs() != n_out) { g_err = "oracle: arity mismatch in tick"; return 1; }
  n->tick(in, out);
  return 0;
}
// render (process.rs:1345-1351), generalised to `outs` channels: out is frame-major [n][outs]
int qo_render(void* h, long n_samples, float* out) {
  Net* n = N(h);
  if (n->ins() != 0) { g_err = "oracle: render needs a 0-input net"; return 1; }
  int no = n->outs();
  for (long i = 0; i < n_samples; i++) n->tick(nullptr, out + (size_t)i * no);
  return 0;
}
// block path with inputs: in is frame-major [n][ins]
int qo_process(void* h, long n_samples, const float* in, float* out) {
  Net* n = N(h);
  int ni = n->ins(), no = n->outs();
  for (long i = 0; i < n_samples; i++) n->tick(in + (size_t)i * ni, out + (size_t)i * no);
  return 0;
}

// Bank render used as the CPU baseline: `n_voices` independent single-output nets, each ticked sample by
// sample exactly like the render op; voices are distributed over `n_threads` std::threads (the reference
// itself never uses more than one thread per graph — BASELINE.md section 3).  out is voice-major [V][T];
// if group > 1, consecutive voices are summed left to right in groups and scaled by 1/group: out [V/group][T].
int qo_render_bank(void** nets, int n_voices, long n_samples, int group, int n_threads, float* out) {
  if (group < 1) group = 1;
  int n_groups = n_voices / group;
  std::atomic<int> next(0);
  auto work = [&]() {
    std::vector<float> tmp((size_t)n_samples);
    for (;;) {
      int g = next.fetch_add(1);
      if (g >= n_groups) break;
      float* o = out + (size_t)g * n_samples;
      for (int k = 0; k < group; k++) {
        Net* n = N(nets[g * group + k]);
        float s;
        if (group == 1) {
          for (long i = 0; i < n_samples; i++) { n->tick(nullptr, &s); o[i] = s; }
        } else {
          for (long i = 0; i < n_samples; i++) { n->tick(nullptr, &s); tmp[i] = s; }
          if (k == 0) std::memcpy(o, tmp.data(), sizeof(float) * n_samples);
          else for (long i = 0; i < n_samples; i++) o[i] += tmp[i];
        }
      }
      if (group > 1) { float sc = 1.0f / (float)group; for (long i = 0; i < n_samples; i++) o[i] *= sc; }
    }
  };
  if (n_threads <= 1) { work(); return 0; }
  std::vector<std::thread> th;
  for (int i = 0; i < n_threads; i++) th.emplace_back(work);
  for (auto& t : th) t.join();
  return 0;
}

// direct access to the restated transforms (fundsp::fft::{real_fft, inverse_fft} call sites nodes.rs:632, 688)
void qo_real_fft(const float* in, int n, float* out_re_im) {
  std::vector<float> a(in, in + n);
  std::vector<Cpx> o(n / 2 + 1);
  real_fft(a, o);
  for (int i = 0; i <= n / 2; i++) { out_re_im[2 * i] = o[i].re; out_re_im[2 * i + 1] = o[i].im; }
}
void qo_inverse_fft(const float* in_re_im, int n, float* out_re_im) {
  std::vector<Cpx> a(n), o(n);
  for (int i = 0; i < n; i++) a[i] = Cpx{in_re_im[2 * i], in_re_im[2 * i + 1]};
  inverse_fft(a, o);
  for (int i = 0; i < n; i++) { out_re_im[2 * i] = o[i].re; out_re_im[2 * i + 1] = o[i].im; }
}

}  // extern "C"
