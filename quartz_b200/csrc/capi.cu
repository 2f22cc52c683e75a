// C ABI (include/quartz_gpu.h).  Host glue only: owns device buffers, lowers graphs, launches kernels.
// No CPU evaluation path exists here by design — if CUDA is unusable every device entry point fails loudly.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/quartz_gpu.h"
#include "fused.h"
#include "graph.h"
#include "kernels.h"
#include "lower.h"
#include "spec.h"
#include "spectral.h"

using namespace qg;

struct qg_net { Graph g; };
struct qg_ctx {
  int device;
  cudaStream_t stream;
  bool own_stream;
  long launches;
};
struct qg_bank {
  qg_ctx* ctx;
  Tape tape;
  long V;
  int Vp;
  int path;
  Instr* d_code = nullptr;
  uint16_t* d_out_x = nullptr;
  float* d_params = nullptr;
  float* d_state = nullptr;
  float* d_state_init = nullptr;
  uint8_t* d_state_keep = nullptr;
  bool state_ready = false;   // d_state holds a valid state (the first reset is a plain copy of the init table)
  float* d_rings = nullptr;
  Ring* d_ring_tab = nullptr;
  ResetRange* d_resets = nullptr;
  float* d_tables = nullptr;
  float* d_scratch = nullptr;
  size_t scratch_bytes = 0;
  float* d_in = nullptr;
  size_t in_bytes = 0;
  float* d_fused_scratch = nullptr;
  size_t fused_scratch_bytes = 0;
  std::vector<float> raw;      // host copy of the per-voice raw parameters [V][R] (empty: every voice = template)
  FusedPlan fused;
  TvPlan tv;
  bool biquad_scan_ok = true;   // every direct-form biquad of every voice may be re-associated (scans) within the tolerance
  bool block_ok = false;   // the tape may run on the block-mode lane interpreter (k_interp_blk)
  int ring_mode = 0;   // 0: rings laid out [pos][voice] (lane kernels); 1: [voice][pos] (time-vector kernel)
  SpecKernel spec;     // K1s: lane kernel compiled for this tape (qg_bank_set_path(QG_PATH_SPECIALISED), or AUTO once it pays)
  bool spec_auto_ok = false, spec_auto_tried = false;
  double lane_work = 0.0;   // voice-samples this bank has rendered on a lane interpreter (AUTO specialises past a threshold)
  // K5, the frame-parallel spectral path: evaluates from the state at reset plus an absolute sample time (it neither reads
  // nor writes d_state / d_rings), so a bank renders EITHER on K5 or on the other kernels between two resets
  SpPlan sp;
  Instr* d_sp_code = nullptr;
  SpSegment* d_sp_segs = nullptr;
  SpItem* d_sp_items = nullptr;
  uint16_t* d_sp_out_x = nullptr;
  long sp_time = 0;             // samples K5 has rendered since the last reset
  bool sp_started = false;      // K5 has rendered since the last reset
  bool other_started = false;   // another kernel has
  bool last_k5 = false;         // the previous render ran on K5 (survives a reset: qg_bank_kernel names what a repeat would use)
  SpectralKernels sp_spec;      // K5s: the plan compiled into the kernels (NVRTC, cached process-wide by plan)
  bool sp_spec_tried = false, sp_cache_checked = false;
  double sp_work = 0.0;         // voice-samples this bank has rendered on K5 (AUTO compiles K5s past a threshold)
};

static thread_local std::string g_err;
static thread_local int g_code = QG_OK;   // status of the last failure on this thread (entry points that return a handle)
static int fail(int code, const std::string& msg) { g_err = msg; g_code = code; return code; }
#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      g_err = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call;                       \
      g_code = QG_ERR_CUDA;                                                                              \
      return QG_ERR_CUDA;                                                                                \
    }                                                                                                    \
  } while (0)
// nothing may unwind across the C boundary (the reference builds with panic = 'abort', Cargo.toml:57): entry points that
// run host code which can allocate wrap their body in guard_int
template <typename F>
static int guard_int(F f) {
  try {
    return f();
  } catch (const std::exception& e) {
    return fail(QG_ERR_ARG, std::string("internal error: ") + e.what());
  } catch (...) {
    return fail(QG_ERR_ARG, "internal error");
  }
}
// Device memory comes from the device's stream-ordered pool (cudaMallocAsync / cudaFreeAsync on the context's stream, release
// threshold = never): building and dropping a bank per render — what quartz does on every patch edit — then costs no
// cudaMalloc / cudaFree round trip to the driver.  Measured on configs[3]: cudaFree of a bank's buffers took 9 ms .. 1.5 s per
// bank (the e2e step was 144 ms .. 1.68 s for the same 131 ms of render + copy).
static cudaError_t dev_malloc(void** p, size_t bytes, cudaStream_t s) { return cudaMallocAsync(p, bytes ? bytes : 1, s); }
static void dev_free(void* p, cudaStream_t s) { if (p) cudaFreeAsync(p, s); }
// a device allocation that lives for one call
template <typename T>
struct DevTmp {
  T* p = nullptr;
  cudaStream_t s;
  explicit DevTmp(cudaStream_t stream) : s(stream) {}
  ~DevTmp() { dev_free(p, s); }
};

template <typename F>
static qg_net* guard_net(F f) {
  try {
    return new qg_net{f()};
  } catch (const std::exception& e) {
    fail(QG_ERR_ARG, std::string("internal error: ") + e.what());
    return nullptr;
  } catch (...) {
    fail(QG_ERR_ARG, "internal error");
    return nullptr;
  }
}
static std::vector<float> fv(const float* arr, int n) { return (arr && n > 0) ? std::vector<float>(arr, arr + n) : std::vector<float>(); }
static std::vector<const Graph*> gv(const qg_net* const* nets, int n) {
  std::vector<const Graph*> v;
  for (int i = 0; i < n; i++) v.push_back(nets && nets[i] ? &nets[i]->g : nullptr);
  return v;
}

template <typename T>
static int upload(T** dst, const std::vector<T>& src, cudaStream_t s) {
  *dst = nullptr;
  if (src.empty()) { CU(dev_malloc((void**)dst, sizeof(T), s)); return QG_OK; }
  CU(dev_malloc((void**)dst, src.size() * sizeof(T), s));
  CU(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return QG_OK;
}

// K2 and the time-vector kernel's block scans re-associate recurrences (zero-state runs + scan).  A direct-form biquad with poles close to z = 1 amplifies any
// rounding difference by its round-off noise gain; when that gain is large even a perfect evaluation differs from the
// reference's own f32 rounding by more than the parity tolerance, so such banks stay on the interpreter, which performs
// the reference's operations in the reference's order.  Power gain of 1 / (1 + a1 z^-1 + a2 z^-2):
//   (1 + a2) / ((1 - a2) ((1 + a2)^2 - a1^2))
// Per-voice parameter derivation (coefficients through the host libm, exactly like the reference at construction) for
// banks of up to a million voices: voices are independent, so the table is filled by all host threads.
static void derive_table(const Tape& t, const float* raw_matrix, int R, long V, int Vp, std::vector<float>* host) {
  const int P = (int)t.h.n_params;
  host->assign((size_t)P * Vp, 0.0f);
  auto work = [&](long lo, long hi) {
    std::vector<float> pv(P);
    for (long v = lo; v < hi; v++) {
      const long src = v < V ? v : V - 1;
      for (int p = 0; p < P; p++) pv[p] = t.params[p];
      t.derive(raw_matrix + (size_t)src * R, pv.data());
      for (int p = 0; p < P; p++) (*host)[(size_t)p * Vp + v] = pv[p];
    }
  };
  unsigned nthr = std::thread::hardware_concurrency();
  if (nthr == 0) nthr = 1;
  const long chunk = 8192;
  if (Vp <= chunk || nthr == 1) { work(0, Vp); return; }
  nthr = (unsigned)std::min<long>(nthr, (Vp + chunk - 1) / chunk);
  std::vector<std::thread> pool;
  const long per = (Vp + nthr - 1) / nthr;
  for (unsigned k = 0; k < nthr; k++) {
    const long lo = (long)k * per, hi = std::min<long>(Vp, lo + per);
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  for (std::thread& th : pool) th.join();
}

template <typename ParamAt>
static bool biquads_well_conditioned(const Tape& t, ParamAt param, long V) {
  for (const Instr& i : t.code) {
    if (i.op != OP_BIQUAD) continue;
    for (long v = 0; v < V; v++) {
      const double a1 = param(i.p, v), a2 = param(i.p + 1, v);
      const double den = (1.0 - a2) * ((1.0 + a2) * (1.0 + a2) - a1 * a1);
      if (!(den > 0.0) || (1.0 + a2) / den > 1.0e4) return false;
    }
  }
  return true;
}

extern "C" {

const char* qg_last_error(void) { return g_err.c_str(); }
const char* qg_version(void) { return "quartz_gpu 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------ graph construction
qg_net* qg_str_to_net(const char* op) { return guard_net([&] { return str_to_net(op ? op : ""); }); }
qg_net* qg_net_new(int ni, int no) { return guard_net([&] { return Graph(ni, no); }); }
qg_net* qg_net_clone(const qg_net* n) { return n ? guard_net([&] { return n->g; }) : nullptr; }
void qg_net_free(qg_net* n) { delete n; }
int qg_net_inputs(const qg_net* n) { return n ? n->g.inputs() : 0; }
int qg_net_outputs(const qg_net* n) { return n ? n->g.outputs() : 0; }
int qg_net_size(const qg_net* n) { return n ? n->g.size() : 0; }
int qg_net_set_sample_rate(qg_net* n, double sr) {
  if (!n) return fail(QG_ERR_ARG, "null net");
  return guard_int([&] { n->g.set_sample_rate(sr); return (int)QG_OK; });
}
const char* qg_net_unsupported(const qg_net* n) { return (n && !n->g.unsupported.empty()) ? n->g.unsupported.c_str() : nullptr; }
qg_net* qg_connect(const char* op, const qg_net* const* nets, int n, double number, int node_limit) {
  return guard_net([&] { return connect(op ? op : "", gv(nets, n), number, node_limit); });
}
qg_net* qg_array_op(const char* kind, const char* op_str, const float* arr, int n) {
  return guard_net([&] { return array_op(kind ? kind : "", op_str ? op_str : "", fv(arr, n)); });
}
qg_net* qg_get(const float* arr, int n) { return guard_net([&] { return make_get(fv(arr, n)); }); }
qg_net* qg_quantize(const float* arr, int n) { return guard_net([&] { return make_quantize(fv(arr, n)); }); }
qg_net* qg_wave(const float* arr, int n) { return guard_net([&] { return make_wave(fv(arr, n)); }); }
qg_net* qg_feedback(const qg_net* net, int has_delay, double delay) {
  return guard_net([&] { return net ? make_feedback(net->g, has_delay != 0, delay) : Graph(0, 0); });
}
qg_net* qg_kr(const qg_net* net, double n, int preserve_time) {
  return guard_net([&] { return net ? make_kr(net->g, n, preserve_time != 0) : Graph(0, 0); });
}
qg_net* qg_reset_every(const qg_net* net, double s) { return guard_net([&] { return net ? make_reset(net->g, s) : Graph(0, 0); }); }
qg_net* qg_trig_reset(const qg_net* net, int variable) {
  return guard_net([&] { return net ? make_trig_reset(net->g, variable != 0) : Graph(0, 0); });
}
qg_net* qg_seq_select(int is_seq, const qg_net* const* nets, int n) {
  return guard_net([&] { return make_seq_select(is_seq != 0, gv(nets, n)); });
}
qg_net* qg_live_io(const char* name) { return guard_net([&] { return make_live_io(name ? name : ""); }); }
qg_net* qg_var(float value) { return guard_net([&] { return make_var(value); }); }

int qg_net_raw_count(const qg_net* n) {
  if (!n) return 0;
  int count = 0;
  guard_int([&] { std::vector<float> r; collect_raw(n->g, &r); count = (int)r.size(); return (int)QG_OK; });
  return count;
}
int qg_net_raw_params(const qg_net* n, float* out, int cap) {
  if (!n) return 0;
  int count = 0;
  guard_int([&] {
    std::vector<float> r;
    collect_raw(n->g, &r);
    for (int i = 0; out && i < (int)r.size() && i < cap; i++) out[i] = r[i];
    count = (int)r.size();
    return (int)QG_OK;
  });
  return count;
}
uint64_t qg_net_signature(const qg_net* n) {
  if (!n) return 0;
  uint64_t sig = 0;
  guard_int([&] { sig = structure_signature(n->g); return (int)QG_OK; });
  return sig;
}
int qg_net_tape_info(const qg_net* n, int* n_instr, int* n_params, int* n_state, int* n_temps, int* divergent) {
  if (!n) return fail(QG_ERR_ARG, "null net");
  return guard_int([&] {
    Tape t;
    std::string err;
    if (!lower(n->g, &t, &err)) return fail(QG_ERR_UNSUPPORTED, err);
    if (n_instr) *n_instr = (int)t.h.n_instr;
    if (n_params) *n_params = (int)t.h.n_params;
    if (n_state) *n_state = (int)t.h.n_state;
    if (n_temps) *n_temps = (int)t.h.n_temps;
    if (divergent) *divergent = (t.h.flags & TAPE_DIVERGENT) ? 1 : 0;
    return (int)QG_OK;
  });
}

// Device parameters of the template voice, as the lowering derives them (filter coefficients, pan weights, 1/sr ...): what
// every kernel reads as X[0, P).  Returns P (copies at most cap values) or a negated status.
int qg_net_device_params(const qg_net* n, float* out, int cap) {
  if (!n) return -fail(QG_ERR_ARG, "null net");
  int count = 0;
  int rc = guard_int([&] {
    Tape t;
    std::string err;
    if (!lower(n->g, &t, &err)) return fail(QG_ERR_UNSUPPORTED, err);
    count = (int)t.params.size();
    for (int i = 0; i < count && i < cap && out; i++) out[i] = t.params[i];
    return (int)QG_OK;
  });
  return rc == QG_OK ? count : -rc;
}

// Frame-parallel spectral plan of a graph (spectral.h): 1 and the plan's shape when the tape qualifies, else 0
int qg_net_spectral_info(const qg_net* n, int* n_segments, int* n_streams, int* n_instr, int* round_len) {
  if (!n) return -fail(QG_ERR_ARG, "null net");
  int ok = 0;
  int rc = guard_int([&] {
    Tape t;
    std::string err;
    if (!lower(n->g, &t, &err)) return fail(QG_ERR_UNSUPPORTED, err);
    SpPlan p = plan_spectral(t);
    ok = p.ok ? 1 : 0;
    if (n_segments) *n_segments = (int)p.segs.size();
    if (n_streams) *n_streams = p.n_streams;
    if (n_instr) *n_instr = (int)p.code.size();
    if (round_len) *round_len = p.C;
    return (int)QG_OK;
  });
  return rc == QG_OK ? ok : -rc;
}

// The translation unit the tape specialiser hands to NVRTC for this graph (spec.cpp); returns the length needed
// (excluding the terminator) or a negative status.  Copies at most cap - 1 characters.
long qg_net_spec_source(const qg_net* n, char* buf, long cap) {
  if (!n) return -(long)fail(QG_ERR_ARG, "null net");
  long need = 0;
  int rc = guard_int([&] {
    Tape t;
    std::string err;
    if (!lower(n->g, &t, &err)) return fail(QG_ERR_UNSUPPORTED, err);
    if (!spec_supported(t, &err)) return fail(QG_ERR_UNSUPPORTED, "tape cannot be specialised: " + err);
    std::string src = spec_source(t);
    need = (long)src.size();
    if (buf && cap > 0) {
      long k = std::min<long>(need, cap - 1);
      memcpy(buf, src.data(), (size_t)k);
      buf[k] = 0;
    }
    return (int)QG_OK;
  });
  return rc == QG_OK ? need : -(long)rc;
}

// The translation unit compiled for the graph's frame-parallel spectral plan (K5s, spectral_kernel.cuh); same contract
long qg_net_spectral_spec_source(const qg_net* n, char* buf, long cap) {
  if (!n) return -(long)fail(QG_ERR_ARG, "null net");
  long need = 0;
  int rc = guard_int([&] {
    Tape t;
    std::string err;
    if (!lower(n->g, &t, &err)) return fail(QG_ERR_UNSUPPORTED, err);
    SpPlan p = plan_spectral(t);
    if (!p.ok) return fail(QG_ERR_UNSUPPORTED, "the graph has no frame-parallel spectral plan");
    std::string src = spectral_spec_source(p, t);
    need = (long)src.size();
    if (buf && cap > 0) {
      long k = std::min<long>(need, cap - 1);
      memcpy(buf, src.data(), (size_t)k);
      buf[k] = 0;
    }
    return (int)QG_OK;
  });
  return rc == QG_OK ? need : -(long)rc;
}

// ------------------------------------------------------------------------------------ device
qg_ctx* qg_ctx_create(int device, void* stream) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    g_err = std::string("no usable CUDA device (") + cudaGetErrorString(e) + "); quartz_gpu has no CPU fallback";
    return nullptr;
  }
  if (device < 0 || device >= n) { g_err = "device index out of range"; return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { g_err = "cudaSetDevice failed"; return nullptr; }
  qg_ctx* c = new (std::nothrow) qg_ctx();
  if (!c) return nullptr;
  c->device = device;
  c->launches = 0;
  if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
  else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; g_err = "cudaStreamCreate failed"; return nullptr; }
    c->own_stream = true;
  }
  // banks allocate from the device's stream-ordered pool (dev_malloc): keep what they return instead of handing it back to the
  // driver at the next synchronisation
  cudaMemPool_t pool = nullptr;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  return c;
}
void qg_ctx_destroy(qg_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaMemPool_t pool = nullptr;
  if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess && pool) cudaMemPoolTrimTo(pool, 0);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}
int qg_ctx_synchronize(qg_ctx* c) {
  if (!c) return fail(QG_ERR_ARG, "null ctx");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  return QG_OK;
}
long qg_ctx_launch_count(const qg_ctx* c) { return c ? c->launches : 0; }
void* qg_device_alloc(qg_ctx* c, size_t bytes) {
  if (!c) return nullptr;
  void* p = nullptr;
  cudaSetDevice(c->device);
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) { g_err = std::string("cudaMalloc failed: ") + cudaGetErrorString(e); return nullptr; }
  return p;
}
void qg_device_free(qg_ctx* c, void* p) { if (c) cudaSetDevice(c->device); if (p) cudaFree(p); }
void* qg_host_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) { g_err = "cudaMallocHost failed"; return nullptr; }
  return p;
}
void qg_host_free_pinned(void* p) { if (p) cudaFreeHost(p); }

static void bank_release(qg_bank* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStream_t s = b->ctx->stream;   // stream-ordered: everything this bank launched on s is ahead of these frees
  void* bufs[] = {b->d_code, b->d_out_x, b->d_params, b->d_state, b->d_state_init, b->d_state_keep, b->d_rings, b->d_ring_tab, b->d_resets,
                  b->d_tables, b->d_scratch, b->d_in, b->d_fused_scratch, b->d_sp_code, b->d_sp_segs, b->d_sp_items, b->d_sp_out_x};
  for (void* q : bufs) dev_free(q, s);
  spec_release(&b->spec);
  delete b;
}

static int bank_init_state(qg_bank* b, const uint64_t* salts) {
  qg_ctx* c = b->ctx;
  const Tape& t = b->tape;
  int NS = (int)t.h.n_state;
  DevTmp<uint32_t> d_def(c->stream);
  DevTmp<HashInit> d_hi(c->stream);
  DevTmp<uint64_t> d_salts(c->stream);
  int rc = upload(&d_def.p, t.state_init, c->stream);
  if (rc) return rc;
  rc = upload(&d_hi.p, t.hash_init, c->stream);
  if (rc) return rc;
  if (salts) {
    std::vector<uint64_t> s((size_t)b->Vp, 0);
    for (long v = 0; v < b->V; v++) s[v] = salts[v];
    rc = upload(&d_salts.p, s, c->stream);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));   // `s` is about to go out of scope
  }
  if (NS > 0) {
    CU(launch_init_state(b->d_state_init, d_def.p, NS, b->Vp, d_hi.p, (int)t.hash_init.size(), d_salts.p, c->stream));
    c->launches++;
  }
  CU(cudaStreamSynchronize(c->stream));
  return QG_OK;
}

static int bank_build(qg_bank* b, const float* raw_matrix /* [V][R] or null */, const uint64_t* salts) {
  qg_ctx* c = b->ctx;
  Tape& t = b->tape;
  CU(cudaSetDevice(c->device));
  const int P = (int)t.h.n_params, NS = (int)t.h.n_state, R = (int)t.h.n_raw;
  b->Vp = (int)((b->V + 127) / 128 * 128);
  size_t need = ((size_t)P + 2 * (size_t)NS + (size_t)t.h.ring_floats) * b->Vp * sizeof(float);
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  if (need > free_b) {   // the pool may be holding what earlier banks returned: hand it back to the driver and look again
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess && pool) {
      cudaStreamSynchronize(c->stream);
      cudaMemPoolTrimTo(pool, 0);
      cudaMemGetInfo(&free_b, &total_b);
    }
  }
  if (need > free_b) return fail(QG_ERR_CUDA, "bank needs " + std::to_string(need >> 20) + " MiB of HBM, only " +
                                                  std::to_string(free_b >> 20) + " MiB free");
  int rc;
  if ((rc = upload(&b->d_code, t.code, c->stream))) return rc;
  if ((rc = upload(&b->d_out_x, t.out_x, c->stream))) return rc;
  if ((rc = upload(&b->d_ring_tab, t.rings, c->stream))) return rc;
  if ((rc = upload(&b->d_resets, t.resets, c->stream))) return rc;
  if ((rc = upload(&b->d_tables, t.tables, c->stream))) return rc;
  if ((rc = upload(&b->d_state_keep, t.state_keep, c->stream))) return rc;
  CU(dev_malloc((void**)&b->d_params, std::max<size_t>(1, (size_t)P * b->Vp) * sizeof(float), c->stream));
  CU(dev_malloc((void**)&b->d_state, std::max<size_t>(1, (size_t)NS * b->Vp) * sizeof(float), c->stream));
  CU(dev_malloc((void**)&b->d_state_init, std::max<size_t>(1, (size_t)NS * b->Vp) * sizeof(float), c->stream));
  CU(dev_malloc((void**)&b->d_rings, std::max<size_t>(1, (size_t)t.h.ring_floats * b->Vp) * sizeof(float), c->stream));
  b->fused = plan_fused(t);
  // ---- parameters: derive per voice on the host (same libm as the reference would use), or broadcast the template
  if (raw_matrix && R > 0) b->raw.assign(raw_matrix, raw_matrix + (size_t)b->V * R);
  if (raw_matrix && R > 0) {   // also for tapes without device parameters (pass() >> delay(x)): the delay still shapes the tape
    for (int r = 0; r < R; r++) {
      if (!t.raw_structural[r]) continue;
      for (long v = 0; v < b->V; v++)
        if (memcmp(&raw_matrix[(size_t)v * R + r], &t.raw[r], 4) != 0)
          return fail(QG_ERR_MISMATCH, "raw parameter " + std::to_string(r) + " shapes the tape (delay length / reset period) and must be equal for every voice");
    }
  }
  if (raw_matrix && R > 0 && P > 0) {
    std::vector<float> host;
    derive_table(t, raw_matrix, R, b->V, b->Vp, &host);
    CU(cudaMemcpyAsync(b->d_params, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    b->biquad_scan_ok = biquads_well_conditioned(t, [&](int p, long v) { return host[(size_t)p * b->Vp + v]; }, b->V);
  } else if (P > 0) {
    DevTmp<float> d_tmpl(c->stream);
    if ((rc = upload(&d_tmpl.p, t.params, c->stream))) return rc;
    CU(launch_broadcast_params(b->d_params, d_tmpl.p, P, b->Vp, c->stream));
    c->launches++;
    CU(cudaStreamSynchronize(c->stream));
  }
  if ((rc = bank_init_state(b, salts))) return rc;
  if (!(raw_matrix && R > 0 && P > 0)) {
    b->biquad_scan_ok = biquads_well_conditioned(t, [&](int p, long) { return t.params[p]; }, 1);
  }
  if (b->fused.id == FUSED_NOISE_SVF && b->fused.p[1] == 2 && !b->biquad_scan_ok) b->fused.id = FUSED_NONE;
  b->tv = plan_tv(t, (size_t)180 * 1024);
  b->sp = plan_spectral(t);
  if (b->sp.ok) {
    if ((rc = upload(&b->d_sp_code, b->sp.code, c->stream))) return rc;
    if ((rc = upload(&b->d_sp_segs, b->sp.segs, c->stream))) return rc;
    if ((rc = upload(&b->d_sp_items, b->sp.items, c->stream))) return rc;
    if ((rc = upload(&b->d_sp_out_x, b->sp.out_x, c->stream))) return rc;
  }
  if (!b->biquad_scan_ok) b->tv.sequential = true;   // those biquads are stepped by one thread in the reference's operation order
  // block mode evaluates an instruction for a whole block of samples before the next one: valid for feed-forward tapes,
  // and for feedback loops whose delay line is at least one block long
  b->spec_auto_ok = spec_auto_ok(t);
  b->block_ok = !(t.h.flags & TAPE_DIVERGENT);
  for (const Instr& i : t.code) {
    if ((i.op == OP_FB_READ || i.op == OP_FB_WRITE) && t.rings[i.aux].length < (uint32_t)interp_block_len()) b->block_ok = false;
    if (i.op == OP_FB1_READ || i.op == OP_FB1_WRITE) b->block_ok = false;
    // reset()/trig_reset()/reset_v() rewind OTHER ops' state in the middle of a block: sample-by-sample only
    if (i.op >= OP_KR_BEGIN && i.op <= OP_SEQ_END) b->block_ok = false;
  }
  return qg_bank_reset(b);
}

qg_bank* qg_bank_create(qg_ctx* ctx, const qg_net* tmpl, long n_voices, const float* raw, const uint64_t* salts) {
  if (!ctx || !tmpl || n_voices <= 0) { fail(QG_ERR_ARG, "qg_bank_create: bad arguments"); return nullptr; }
  qg_bank* b = new (std::nothrow) qg_bank();
  if (!b) return nullptr;
  b->ctx = ctx; b->V = n_voices; b->path = QG_PATH_AUTO;
  std::string err;
  try {
    if (!lower(tmpl->g, &b->tape, &err)) { fail(QG_ERR_UNSUPPORTED, err); delete b; return nullptr; }
    if (bank_build(b, raw, salts) != QG_OK) { bank_release(b); return nullptr; }
  } catch (const std::exception& e) {
    fail(QG_ERR_ARG, std::string("internal error: ") + e.what());
    bank_release(b);
    return nullptr;
  } catch (...) {
    fail(QG_ERR_ARG, "internal error");
    bank_release(b);
    return nullptr;
  }
  return b;
}

qg_bank* qg_bank_from_nets(qg_ctx* ctx, const qg_net* const* nets, long n, const uint64_t* salts) {
  if (!ctx || !nets || n <= 0 || !nets[0]) { fail(QG_ERR_ARG, "qg_bank_from_nets: bad arguments"); return nullptr; }
  try {
    uint64_t sig = structure_signature(nets[0]->g);
    std::vector<float> r0;
    collect_raw(nets[0]->g, &r0);
    size_t R = r0.size();
    std::vector<float> raw((size_t)n * std::max<size_t>(R, 1));
    for (long v = 0; v < n; v++) {
      if (!nets[v] || structure_signature(nets[v]->g) != sig) {
        fail(QG_ERR_MISMATCH, "qg_bank_from_nets: net " + std::to_string(v) + " does not have the structure of net 0");
        return nullptr;
      }
      std::vector<float> r;
      collect_raw(nets[v]->g, &r);
      if (r.size() != R) { fail(QG_ERR_MISMATCH, "qg_bank_from_nets: parameter count mismatch"); return nullptr; }
      for (size_t k = 0; k < R; k++) raw[(size_t)v * R + k] = r[k];
    }
    return qg_bank_create(ctx, nets[0], n, R ? raw.data() : nullptr, salts);
  } catch (const std::exception& e) {
    fail(QG_ERR_ARG, std::string("internal error: ") + e.what());
    return nullptr;
  } catch (...) {
    fail(QG_ERR_ARG, "internal error");
    return nullptr;
  }
}

void qg_bank_free(qg_bank* b) { bank_release(b); }

int qg_bank_reset(qg_bank* b) {
  if (!b) return fail(QG_ERR_ARG, "null bank");
  qg_ctx* c = b->ctx;
  CU(cudaSetDevice(c->device));
  size_t ns = (size_t)b->tape.h.n_state * b->Vp * sizeof(float);
  bool any_keep = false;
  for (uint8_t k : b->tape.state_keep) any_keep = any_keep || k;
  if (ns && b->state_ready && any_keep) {   // Seq::reset leaves the event list alone (nodes.rs:116-120)
    CU(launch_reset_state(b->d_state, b->d_state_init, b->d_state_keep, (int)b->tape.h.n_state, b->Vp, c->stream));
    c->launches++;
  } else if (ns) {
    CU(cudaMemcpyAsync(b->d_state, b->d_state_init, ns, cudaMemcpyDeviceToDevice, c->stream));
  }
  b->state_ready = true;
  b->sp_time = 0; b->sp_started = false; b->other_started = false;
  size_t rb = (size_t)b->tape.h.ring_floats * b->Vp * sizeof(float);
  if (rb) CU(cudaMemsetAsync(b->d_rings, 0, rb, c->stream));
  return QG_OK;
}
// K5 serves a bank from its reset on (see qg_bank): forced with QG_PATH_SPECTRAL, or chosen by AUTO when the first render after
// a reset is a bulk one (a block-wise stream of short calls stays on the time-vector kernel, where var() updates apply)
static const double kSpectralAutoWork = 262144.0;   // voice-samples
static bool bank_spectral(const qg_bank* b, long T) {
  if (!b->sp.ok || b->other_started) return false;
  if (b->path == QG_PATH_SPECTRAL || b->sp_started) return true;
  return b->path == QG_PATH_AUTO && (T <= 0 || ((double)b->V * (double)T >= kSpectralAutoWork && T >= 4L * b->sp.C));
}
// which kernel family serves voice-major, group-1 renders: 0 lane interpreter, 1 fused, 2 time-vector interpreter
static int bank_family(const qg_bank* b) {
  if (b->path == QG_PATH_SPECTRAL) return 2;   // (render_impl asks bank_spectral() first)
  if (b->path == QG_PATH_SPECIALISED) return b->spec.fn ? 3 : 0;
  if (b->path == QG_PATH_AUTO && b->spec.fn && b->fused.id == FUSED_NONE && !(b->tv.ok && (b->tv.has_fft || b->V <= (b->tv.sequential ? 256 : 2048))))
    return 3;   // AUTO specialised this lane-interpreter bank (bank_auto_specialise)
  if (b->path == QG_PATH_INTERP || b->path == QG_PATH_INTERP_SAMPLE) return 0;
  if (b->path == QG_PATH_TV) return b->tv.ok ? 2 : 0;
  if (b->fused.id != FUSED_NONE) return 1;
  if (b->tv.ok && (b->tv.has_fft || b->V <= (b->tv.sequential ? 256 : 2048))) return 2;
  return 0;
}
int qg_bank_set_path(qg_bank* b, int path) {
  if (!b) return fail(QG_ERR_ARG, "null bank");
  int before = bank_family(b) == 2;
  if (path < QG_PATH_AUTO || path > QG_PATH_SPECTRAL) return fail(QG_ERR_ARG, "unknown path");
  if (path == QG_PATH_SPECTRAL && !b->sp.ok)
    return fail(QG_ERR_UNSUPPORTED, "the frame-parallel spectral path needs rfft -> stateless bin chain -> ifft segments fed by "
                                    "signals that are pure functions of time (noise, wave tables, delays of those)");
  const bool was_k5 = b->sp_started;
  if (path == QG_PATH_SPECIALISED && !b->spec.fn) {
    CU(cudaSetDevice(b->ctx->device));
    std::string err;
    bool ok = false;
    int rc = guard_int([&] { ok = spec_compile(b->tape, &b->spec, &err); return (int)QG_OK; });
    if (rc != QG_OK) return rc;
    if (!ok) return fail(QG_ERR_UNSUPPORTED, err);
  }
  b->path = path;
  if ((was_k5 && path != QG_PATH_SPECTRAL && path != QG_PATH_AUTO) || (path == QG_PATH_SPECTRAL && b->other_started))
    return qg_bank_reset(b);                   // K5 keeps its state as a sample time, the other kernels as state words
  if ((bank_family(b) == 2) != before) return qg_bank_reset(b);   // the two interpreters lay delay lines out differently
  return QG_OK;
}
const char* qg_bank_kernel(const qg_bank* b) {
  if (!b) return "";
  if (b->sp.ok && !b->other_started && (b->path == QG_PATH_SPECTRAL || b->sp_started || (b->path == QG_PATH_AUTO && b->last_k5)))
    return b->sp_spec.frames ? "k_sp_frames" : "k_spectral_frames";
  int f = bank_family(b);
  if (f == 1) return fused_name(b->fused.id);
  if (f == 2) return "k_interp_tv";
  if (f == 3) return "k_spec";
  if (b->tape.h.flags & TAPE_DIVERGENT) return "k_interp<divergent>";
  return (b->block_ok && b->path != QG_PATH_INTERP_SAMPLE) ? "k_interp_blk" : "k_interp<uniform>";
}
long qg_bank_out_rows(const qg_bank* b, int group) {
  if (!b) return 0;
  if (group < 1) group = 1;
  return (b->V / group) * (long)b->tape.h.n_outputs;
}

// var() semantics (process.rs:1382-1385): the control plane rewrites one op-string parameter while the graph runs.
// The new value applies to every voice from the next render call on; state is untouched.
static int bank_set_raw(qg_bank* b, int raw_index, float value);
int qg_bank_set_raw(qg_bank* b, int raw_index, float value) {
  if (!b) return fail(QG_ERR_ARG, "null bank");
  return guard_int([&] { return bank_set_raw(b, raw_index, value); });
}
// a parameter update can move a direct-form biquad next to z = 1: from then on the bank keeps the reference's operation
// order for its biquads (no scans), exactly like a bank built with those coefficients
static void bank_biquads_now_ill_conditioned(qg_bank* b) {
  b->biquad_scan_ok = false;
  if (b->fused.id == FUSED_NOISE_SVF && b->fused.p[1] == 2) b->fused.id = FUSED_NONE;
}
static int bank_set_raw(qg_bank* b, int raw_index, float value) {
  Tape& t = b->tape;
  const int R = (int)t.h.n_raw, P = (int)t.h.n_params;
  if (raw_index < 0 || raw_index >= R) return fail(QG_ERR_ARG, "raw parameter index out of range");
  if (t.raw_structural[raw_index]) return fail(QG_ERR_MISMATCH, "this parameter shapes the tape (delay length / reset period): rebuild the bank");
  if (b->sp_started)
    return fail(QG_ERR_MISMATCH, "this bank is rendering on the frame-parallel spectral path, which evaluates from the state at reset: "
                                 "reset it first, or select QG_PATH_TV before the first render to stream with var() updates");
  qg_ctx* c = b->ctx;
  CU(cudaSetDevice(c->device));
  t.raw[raw_index] = value;
  if (P == 0) return QG_OK;
  if (b->raw.empty()) {
    std::vector<float> pv(t.params);
    t.derive(t.raw.data(), pv.data());
    t.params = pv;
    DevTmp<float> d_tmpl(c->stream);
    int rc = upload(&d_tmpl.p, t.params, c->stream);
    if (rc) return rc;
    CU(launch_broadcast_params(b->d_params, d_tmpl.p, P, b->Vp, c->stream));
    c->launches++;
    CU(cudaStreamSynchronize(c->stream));
    if (!biquads_well_conditioned(t, [&](int p, long) { return t.params[p]; }, 1)) bank_biquads_now_ill_conditioned(b);
  } else {
    std::vector<float> host;
    for (long v = 0; v < b->V; v++) b->raw[(size_t)v * R + raw_index] = value;
    derive_table(t, b->raw.data(), R, b->V, b->Vp, &host);
    CU(cudaMemcpyAsync(b->d_params, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (!biquads_well_conditioned(t, [&](int p, long v) { return host[(size_t)p * b->Vp + v]; }, b->V)) bank_biquads_now_ill_conditioned(b);
  }
  return QG_OK;
}

static int check_group(const qg_bank* b, int layout, int group) {
  if (group < 1) return fail(QG_ERR_ARG, "group must be >= 1");
  if (group > 1) {
    if (layout != QG_LAYOUT_VOICE_MAJOR) return fail(QG_ERR_ARG, "group mixes need the voice-major layout");
    if (group > 32 || (32 % group) != 0) return fail(QG_ERR_ARG, "group must be one of 1,2,4,8,16,32");
    if (b->V % group) return fail(QG_ERR_ARG, "voice count must be a multiple of group");
  }
  return QG_OK;
}

// AUTO: a bank that runs on a lane interpreter gets its tape compiled into the kernel (K1s) as soon as a kernel for this tape
// is in the process-wide cache, or once the bank has enough work behind and ahead of it to pay for ~1 s of NVRTC
// (QG_SPEC_MIN_WORK voice-samples, default 1e10; QG_SPEC_AUTO=0 switches the policy off).  A failed compile leaves the bank
// where it was.
static void bank_auto_specialise(qg_bank* b, long T) {
  if (b->path != QG_PATH_AUTO || b->spec.fn || b->spec_auto_tried || !b->spec_auto_ok) return;
  const char* ea = getenv("QG_SPEC_AUTO");
  const bool enabled = !(ea && ea[0] == '0');
  const char* ew = getenv("QG_SPEC_MIN_WORK");
  const double wv = ew ? atof(ew) : 0.0, min_work = wv > 0.0 ? wv : 1.0e10;
  if (!enabled) return;
  if (spec_cached(b->tape, &b->spec)) return;
  b->lane_work += (double)b->V * (double)T;
  if (b->lane_work < min_work) return;
  b->spec_auto_tried = true;
  std::string err;
  try { spec_compile(b->tape, &b->spec, &err); } catch (...) {}
}

static int ensure(float** p, size_t* have, size_t need, cudaStream_t s);

// K5 / K5s: one render call on the frame-parallel spectral path.  *served = false (and QG_OK) when this call has to go to the
// general kernels instead (ring or shared memory too large) — only possible while the bank has not started on K5.
static int render_spectral(qg_bank* b, long T, int layout, float* d_out, bool* served) {
  qg_ctx* c = b->ctx;
  const Tape& t = b->tape;
  const bool committed = b->sp_started || b->path == QG_PATH_SPECTRAL;
  int ring = 0;
  const size_t yb = spectral_y_bytes(b->sp, b->V, &ring);
  if (yb > b->fused_scratch_bytes) {        // the ring of a very large bank (10^6 voices x 2048-point frames) may not fit
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && yb > free_b / 2) {
      if (committed) return fail(QG_ERR_CUDA, "the frame-parallel spectral path needs " + std::to_string(yb >> 20) + " MiB for its ring");
      return QG_OK;
    }
  }
  int rc = ensure(&b->d_fused_scratch, &b->fused_scratch_bytes, yb, c->stream);
  if (rc) return rc;
  SpArgs sa;
  memset(&sa, 0, sizeof sa);
  sa.code = b->d_sp_code; sa.n_code = (int)b->sp.code.size(); sa.segs = b->d_sp_segs; sa.n_segs = (int)b->sp.segs.size();
  sa.items = b->d_sp_items; sa.n_items = (int)b->sp.items.size(); sa.out_x = b->d_sp_out_x; sa.n_out = (int)t.h.n_outputs;
  sa.params = b->d_params; sa.state_init = b->d_state_init; sa.tables = b->d_tables;
  sa.P = (int)t.h.n_params; sa.NS = (int)t.h.n_state; sa.V = (int)b->V; sa.Vp = b->Vp;
  sa.y = b->d_fused_scratch; sa.ring = ring; sa.n_streams = b->sp.n_streams; sa.out = d_out; sa.T = T; sa.t0 = b->sp_time;
  sa.frame_major = layout == QG_LAYOUT_FRAME_MAJOR;
  sa.C = b->sp.C; sa.post_lo = b->sp.post_lo; sa.post_hi = b->sp.post_hi;
  sa.n_slots_frame = b->sp.n_slots_frame; sa.n_slots_post = b->sp.n_slots_post;
  // K5s: compiling the plan into the kernels costs ~2 s of NVRTC once per plan and process (cached after that) and makes a
  // render 2.2x faster: AUTO compiles once the bank has QG_SPECTRAL_MIN_WORK voice-samples (default 2e9) behind and ahead of
  // it, or at once when the plan's kernels are already in the cache.  QG_SPECTRAL_SPEC=1 specialises every K5 bank, =0 none;
  // a failed compile keeps the generic kernels.
  if (!b->sp_spec.frames) {
    const char* es = getenv("QG_SPECTRAL_SPEC");
    const bool force = es && es[0] == '1', never = es && es[0] == '0';
    if (!never && !b->sp_cache_checked) { b->sp_cache_checked = true; spectral_spec_cached(b->sp, t, &b->sp_spec); }
    b->sp_work += (double)b->V * (double)T;
    const char* ew = getenv("QG_SPECTRAL_MIN_WORK");
    const double wv = ew ? atof(ew) : 0.0, min_work = wv > 0.0 ? wv : 2.0e9;
    if (!never && !b->sp_spec.frames && !b->sp_spec_tried && (force || b->sp_work >= min_work)) {
      b->sp_spec_tried = true;
      std::string err;
      try { spectral_spec_compile(b->sp, t, &b->sp_spec, &err); } catch (...) {}
      if (!b->sp_spec.frames && force) return fail(QG_ERR_UNSUPPORTED, "QG_SPECTRAL_SPEC=1: " + err);
    }
  }
  int l = 0;
  cudaError_t se = launch_spectral(sa, b->sp, c->stream, &l, b->sp_spec.frames, b->sp_spec.post);
  c->launches += l;
  if (se == cudaErrorNotSupported) {
    if (committed) return fail(QG_ERR_UNSUPPORTED, "the spectral path needs more shared memory than one SM offers");
    return QG_OK;
  }
  CU(se);
  b->sp_started = true; b->sp_time += T; b->last_k5 = true;
  *served = true;
  return QG_OK;
}

static int render_impl(qg_bank* b, long T, int layout, int group, const float* d_in, float* d_out) {
  qg_ctx* c = b->ctx;
  const Tape& t = b->tape;
  CU(cudaSetDevice(c->device));
  if (T <= 0) return QG_OK;
  int rc = check_group(b, layout, group);
  if (rc) return rc;
  if (!d_in && group == 1 && bank_spectral(b, T)) {
    bool served = false;
    rc = render_spectral(b, T, layout, d_out, &served);
    if (rc || served) return rc;
  }
  if (b->sp_started) return fail(QG_ERR_MISMATCH, "this bank is rendering on the frame-parallel spectral path (no group mix, no inputs): reset it first");
  if (b->path == QG_PATH_SPECTRAL) return fail(QG_ERR_ARG, "QG_PATH_SPECTRAL renders group-1 banks without inputs");
  b->other_started = true; b->last_k5 = false;
  if (bank_family(b) == 0) bank_auto_specialise(b, T);
  const int family = bank_family(b);
  if (family == 2 && group == 1) {
    TvArgs ta;
    memset(&ta, 0, sizeof ta);
    ta.code = b->d_code; ta.n_instr = (int)t.h.n_instr;
    ta.P = (int)t.h.n_params; ta.NS = (int)t.h.n_state; ta.NT = (int)t.h.n_temps;
    ta.n_in = (int)t.h.n_inputs; ta.n_out = (int)t.h.n_outputs; ta.out_x = b->d_out_x;
    ta.params = b->d_params; ta.state = b->d_state; ta.rings = b->d_rings; ta.ring_floats = t.h.ring_floats;
    ta.ring_tab = b->d_ring_tab; ta.tables = b->d_tables; ta.in = d_in; ta.out = d_out;
    ta.V = (int)b->V; ta.Vp = b->Vp; ta.T = T; ta.H = b->tv.H; ta.fft_n = b->tv.fft_n; ta.align_s = b->tv.align_s; ta.n_lti = b->tv.n_lti; ta.biquad_scan = b->biquad_scan_ok ? 1 : 0; ta.frame_major = layout == QG_LAYOUT_FRAME_MAJOR;
    int l = 0;
    CU(launch_interp_tv(ta, c->stream, &l));
    c->launches += l;
    return QG_OK;
  }
  if (family == 2) return fail(QG_ERR_ARG, "this bank runs on the time-vector interpreter, which has no group mix "
                                           "(qg_bank_set_path(QG_PATH_INTERP) selects the lane interpreter)");
  if (family == 1 && !d_in && layout == QG_LAYOUT_VOICE_MAJOR) {
    FusedArgs fa;
    fa.params = b->d_params; fa.state = b->d_state; fa.V = (int)b->V; fa.Vp = b->Vp; fa.T = T; fa.group = group; fa.out = d_out;
    fa.scratch = &b->d_fused_scratch; fa.scratch_bytes = &b->fused_scratch_bytes; fa.sample_rate = t.h.sample_rate;
    fa.tables = b->d_tables;
    int l = 0;
    cudaError_t fe = launch_fused(b->fused, fa, c->stream, &l);
    c->launches += l;
    if (fe == cudaSuccess) return QG_OK;
    if (fe != cudaErrorNotSupported) CU(fe);
    // the fused kernel does not serve this call (group mix on K2, sample rate too low for the envelope look-ahead ...):
    // the interpreters below do
  }
  InterpArgs a;
  memset(&a, 0, sizeof a);
  a.code = b->d_code; a.n_instr = (int)t.h.n_instr;
  a.P = (int)t.h.n_params; a.NS = (int)t.h.n_state; a.NT = (int)t.h.n_temps;
  a.n_in = (int)t.h.n_inputs; a.n_out = (int)t.h.n_outputs; a.out_x = b->d_out_x;
  a.params = b->d_params; a.state = b->d_state; a.state_init = b->d_state_init; a.state_keep = b->d_state_keep; a.rings = b->d_rings;
  a.ring_tab = b->d_ring_tab; a.resets = b->d_resets; a.tables = b->d_tables;
  a.in = d_in; a.out = d_out; a.V = (int)b->V; a.Vp = b->Vp; a.T = T;
  a.in_frame_major = layout == QG_LAYOUT_FRAME_MAJOR; a.out_frame_major = layout == QG_LAYOUT_FRAME_MAJOR;
  a.group = group;
  int l = 0;
  if (family == 3) {
    cudaError_t se = spec_launch(b->spec, a, c->stream, &l);
    c->launches += l;
    CU(se);
    return QG_OK;
  }
  cudaError_t e = launch_interp(a, (t.h.flags & TAPE_DIVERGENT) != 0, b->block_ok && b->path != QG_PATH_INTERP_SAMPLE, c->stream, &l);
  c->launches += l;
  if (e == cudaErrorInvalidConfiguration) return fail(QG_ERR_UNSUPPORTED, "tape needs more shared memory per voice than one SM offers");
  CU(e);
  return QG_OK;
}

int qg_bank_render_device(qg_bank* b, long T, int layout, int group, float* d_out) {
  if (!b || !d_out) return fail(QG_ERR_ARG, "qg_bank_render_device: bad arguments");
  if (b->tape.h.n_inputs != 0) return fail(QG_ERR_ARITY, "render needs a net with 0 inputs (process.rs:1345)");
  return render_impl(b, T, layout, group, nullptr, d_out);
}

static int ensure(float** p, size_t* have, size_t need, cudaStream_t s) {
  if (*have >= need) return QG_OK;
  dev_free(*p, s);
  *p = nullptr; *have = 0;
  CU(dev_malloc((void**)p, need, s));
  *have = need;
  return QG_OK;
}

int qg_bank_render(qg_bank* b, long T, int layout, int group, float* h_out) {
  if (!b || !h_out) return fail(QG_ERR_ARG, "qg_bank_render: bad arguments");
  if (b->tape.h.n_inputs != 0) return fail(QG_ERR_ARITY, "render needs a net with 0 inputs (process.rs:1345)");
  if (T <= 0) return QG_OK;
  if (group < 1) group = 1;
  const size_t rows = (size_t)(b->V / group) * b->tape.h.n_outputs;
  const size_t bytes = rows * (size_t)T * sizeof(float);
  if (bytes == 0) return QG_OK;
  qg_ctx* c = b->ctx;
  CU(cudaSetDevice(c->device));
  // Large voice-major renders are produced in time chunks through two device staging buffers so that the
  // device->host copy of chunk k overlaps the render of chunk k+1 (state persists across chunks by construction).
  const size_t chunk_target = (size_t)512 << 20;
  if (layout == QG_LAYOUT_VOICE_MAJOR && bytes > 2 * chunk_target && rows > 0) {
    long Tc = (long)(chunk_target / (rows * sizeof(float)));
    Tc = Tc / 512 * 512;
    if (Tc >= 512 && Tc < T) {
      const size_t cb = rows * (size_t)Tc * sizeof(float);
      int rc = ensure(&b->d_scratch, &b->scratch_bytes, 2 * cb, b->ctx->stream);
      if (rc) return rc;
      cudaStream_t copy_stream = nullptr;
      cudaEvent_t rendered[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
      cudaError_t ce = cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking);
      for (int i = 0; i < 2 && ce == cudaSuccess; i++) {
        ce = cudaEventCreateWithFlags(&rendered[i], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming);
      }
      int k = 0;
      rc = QG_OK;
      // no early return inside the loop: the copy stream and the events are released below on every path
      for (long t0 = 0; t0 < T && rc == QG_OK && ce == cudaSuccess; t0 += Tc, k ^= 1) {
        long n = T - t0 < Tc ? T - t0 : Tc;
        float* d = b->d_scratch + (size_t)k * rows * Tc;
        if (t0 >= 2 * Tc) ce = cudaStreamWaitEvent(c->stream, copied[k], 0);      // staging buffer k is free again
        if (ce != cudaSuccess) break;
        rc = render_impl(b, n, layout, group, nullptr, d);                        // rows of n samples, pitch n
        if (rc) break;
        ce = cudaEventRecord(rendered[k], c->stream);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(copy_stream, rendered[k], 0);
        if (ce == cudaSuccess)
          ce = cudaMemcpy2DAsync(h_out + t0, (size_t)T * sizeof(float), d, (size_t)n * sizeof(float), (size_t)n * sizeof(float), rows,
                                 cudaMemcpyDeviceToHost, copy_stream);
        if (ce == cudaSuccess) ce = cudaEventRecord(copied[k], copy_stream);
      }
      cudaError_t e1 = copy_stream ? cudaStreamSynchronize(copy_stream) : cudaSuccess, e2 = cudaStreamSynchronize(c->stream);
      for (int i = 0; i < 2; i++) { if (rendered[i]) cudaEventDestroy(rendered[i]); if (copied[i]) cudaEventDestroy(copied[i]); }
      if (copy_stream) cudaStreamDestroy(copy_stream);
      if (rc) return rc;
      CU(ce); CU(e1); CU(e2);
      return QG_OK;
    }
  }
  int rc = ensure(&b->d_scratch, &b->scratch_bytes, bytes, b->ctx->stream);
  if (rc) return rc;
  rc = render_impl(b, T, layout, group, nullptr, b->d_scratch);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h_out, b->d_scratch, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return QG_OK;
}

int qg_bank_process(qg_bank* b, long T, int layout, const float* h_in, float* h_out) {
  if (!b) return fail(QG_ERR_ARG, "null bank");
  const Tape& t = b->tape;
  if (T <= 0) return QG_OK;
  size_t ib = (size_t)b->V * t.h.n_inputs * (size_t)T * sizeof(float);
  size_t ob = (size_t)b->V * t.h.n_outputs * (size_t)T * sizeof(float);
  if ((ib && !h_in) || (ob && !h_out)) return fail(QG_ERR_ARG, "qg_bank_process: null buffer");
  CU(cudaSetDevice(b->ctx->device));
  int rc;
  if ((rc = ensure(&b->d_in, &b->in_bytes, std::max<size_t>(ib, 4), b->ctx->stream))) return rc;
  if ((rc = ensure(&b->d_scratch, &b->scratch_bytes, std::max<size_t>(ob, 4), b->ctx->stream))) return rc;
  if (ib) CU(cudaMemcpyAsync(b->d_in, h_in, ib, cudaMemcpyHostToDevice, b->ctx->stream));
  rc = render_impl(b, T, layout, 1, b->d_in, b->d_scratch);
  if (rc) return rc;
  if (ob) CU(cudaMemcpyAsync(h_out, b->d_scratch, ob, cudaMemcpyDeviceToHost, b->ctx->stream));
  CU(cudaStreamSynchronize(b->ctx->stream));
  return QG_OK;
}

// Stream path: n sanitised, clamped, interleaved stereo frames of a one-voice bank (src/audio.rs:85-118) in the device
// sample type: f32, i16 or u16 (`T::from_sample`, audio.rs:115-116)
int qg_bank_render_stereo_as(qg_bank* b, long n, int sample_format, void* h_frames) {
  if (!b || !h_frames) return fail(QG_ERR_ARG, "qg_bank_render_stereo: bad arguments");
  if (sample_format != QG_SAMPLE_F32 && sample_format != QG_SAMPLE_I16 && sample_format != QG_SAMPLE_U16)
    return fail(QG_ERR_ARG, "unsupported sample format (cpal offers f32, i16, u16: audio.rs:47-59)");
  if (b->V != 1 || b->tape.h.n_inputs != 0) return fail(QG_ERR_ARITY, "stream path needs one 0-input graph (process.rs:1896)");
  if (n <= 0) return QG_OK;
  qg_ctx* c = b->ctx;
  CU(cudaSetDevice(c->device));
  const int no = (int)b->tape.h.n_outputs;
  const size_t sb = (size_t)std::max(no, 1) * n * sizeof(float), fb = (size_t)2 * n * sizeof(float);
  int rc = ensure(&b->d_scratch, &b->scratch_bytes, sb + fb, b->ctx->stream);
  if (rc) return rc;
  // frames first: k_stereo_frames stores float2, which needs 8-byte alignment whatever n and the channel count are
  float* d_frames = b->d_scratch;
  float* d_src = b->d_scratch + (size_t)2 * n;
  if (no == 1 || no == 2) {
    rc = render_impl(b, n, QG_LAYOUT_VOICE_MAJOR, 1, nullptr, d_src);
    if (rc) return rc;
  }
  const int n_ch = (no == 1 || no == 2) ? no : 0;
  size_t out_bytes = fb;
  if (sample_format == QG_SAMPLE_F32) CU(launch_stereo_frames(d_src, n_ch, n, d_frames, c->stream));
  else {
    CU(launch_stereo_frames_i16(d_src, n_ch, n, sample_format == QG_SAMPLE_U16 ? 1 : 0, d_frames, c->stream));
    out_bytes = (size_t)2 * n * sizeof(int16_t);
  }
  c->launches++;
  CU(cudaMemcpyAsync(h_frames, d_frames, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return QG_OK;
}
int qg_bank_render_stereo(qg_bank* b, long n, float* h_frames) { return qg_bank_render_stereo_as(b, n, QG_SAMPLE_F32, h_frames); }

// Value copy of a bank WITH its state (the reference deep-clones a Net, state included, on every hop: process.rs:1316, 1336,
// 1499, 1558, 1895; FunDSP's AudioUnit is DynClone): parameters, state, delay lines and the kernel choice are copied device
// to device; both banks continue independently and identically.
qg_bank* qg_bank_clone(const qg_bank* src) {
  if (!src) { fail(QG_ERR_ARG, "qg_bank_clone: null bank"); return nullptr; }
  qg_bank* b = nullptr;
  try {
    b = new qg_bank();
    b->ctx = src->ctx; b->tape = src->tape; b->V = src->V; b->Vp = src->Vp; b->path = src->path;
    b->raw = src->raw; b->fused = src->fused; b->tv = src->tv; b->biquad_scan_ok = src->biquad_scan_ok;
    b->block_ok = src->block_ok; b->ring_mode = src->ring_mode; b->state_ready = src->state_ready;
    b->spec_auto_ok = src->spec_auto_ok; b->spec_auto_tried = src->spec_auto_tried; b->lane_work = src->lane_work;
    if (src->spec.fn && src->spec.shared) b->spec = src->spec;     // kernels live in the process-wide cache
    b->sp = src->sp; b->sp_time = src->sp_time; b->sp_started = src->sp_started; b->other_started = src->other_started; b->last_k5 = src->last_k5; b->sp_spec = src->sp_spec; b->sp_spec_tried = src->sp_spec_tried; b->sp_cache_checked = src->sp_cache_checked; b->sp_work = src->sp_work;
  } catch (...) {
    delete b;
    fail(QG_ERR_ARG, "qg_bank_clone: out of memory");
    return nullptr;
  }
  qg_ctx* c = b->ctx;
  const Tape& t = b->tape;
  const size_t P = t.h.n_params, NS = t.h.n_state;
  auto dup = [&](auto** dst, const auto* from, size_t bytes) -> bool {
    if (dev_malloc((void**)dst, std::max<size_t>(bytes, 8), c->stream) != cudaSuccess) return false;   // never copy past a small source
    return !from || bytes == 0 || cudaMemcpyAsync(*dst, from, bytes, cudaMemcpyDeviceToDevice, c->stream) == cudaSuccess;
  };
  bool ok = cudaSetDevice(c->device) == cudaSuccess;
  ok = ok && dup(&b->d_code, src->d_code, t.code.size() * sizeof(Instr)) && dup(&b->d_out_x, src->d_out_x, t.out_x.size() * sizeof(uint16_t)) &&
       dup(&b->d_ring_tab, src->d_ring_tab, t.rings.size() * sizeof(Ring)) && dup(&b->d_resets, src->d_resets, t.resets.size() * sizeof(ResetRange)) &&
       dup(&b->d_tables, src->d_tables, t.tables.size() * sizeof(float)) && dup(&b->d_state_keep, src->d_state_keep, t.state_keep.size()) &&
       dup(&b->d_params, src->d_params, P * b->Vp * sizeof(float)) && dup(&b->d_state, src->d_state, NS * b->Vp * sizeof(float)) &&
       dup(&b->d_state_init, src->d_state_init, NS * b->Vp * sizeof(float)) &&
       dup(&b->d_rings, src->d_rings, (size_t)t.h.ring_floats * b->Vp * sizeof(float));
  if (ok && b->sp.ok)
    ok = dup(&b->d_sp_code, src->d_sp_code, b->sp.code.size() * sizeof(Instr)) && dup(&b->d_sp_segs, src->d_sp_segs, b->sp.segs.size() * sizeof(SpSegment)) &&
         dup(&b->d_sp_items, src->d_sp_items, b->sp.items.size() * sizeof(SpItem)) && dup(&b->d_sp_out_x, src->d_sp_out_x, b->sp.out_x.size() * sizeof(uint16_t));
  ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess;
  if (!ok) {
    fail(QG_ERR_CUDA, std::string("qg_bank_clone: ") + cudaGetErrorString(cudaGetLastError()));
    bank_release(b);
    return nullptr;
  }
  return b;
}

int qg_mix_rows_device(qg_ctx* c, const float* d_rows, long rows, long n, float scale, float* d_out) {
  if (!c || !d_rows || !d_out) return fail(QG_ERR_ARG, "qg_mix_rows_device: bad arguments");
  CU(cudaSetDevice(c->device));
  CU(launch_mix_rows(d_rows, (int)rows, n, scale, d_out, c->stream));
  c->launches++;
  return QG_OK;
}

// Measured FP32 (non-tensor) FFMA peak of this GPU in TFLOP/s: the roofline the compute-bound workloads are compared with.
double qg_ctx_measure_fp32_tflops(qg_ctx* c) {
  if (!c) { g_err = "null ctx"; return -1.0; }
  if (cudaSetDevice(c->device) != cudaSuccess) { g_err = "cudaSetDevice failed"; return -1.0; }
  float* d = nullptr;
  if (cudaMalloc((void**)&d, 4) != cudaSuccess) { g_err = "cudaMalloc failed"; return -1.0; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, c->device);
  const int blocks = prop.multiProcessorCount * 8, iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0, c->stream);
    launch_fp32_peak(d, blocks, iters, c->stream);
    cudaEventRecord(e1, c->stream);
    if (cudaEventSynchronize(e1) != cudaSuccess) { g_err = "fp32 probe failed"; best = -1.0; break; }
    c->launches++;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = (double)blocks * 256.0 * (double)iters * 64.0 * 2.0;
    if (rep > 0 && ms > 0.0f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}

// ------------------------------------------------------------------------------------ reference-shaped helpers
int qg_net_render(qg_ctx* ctx, const qg_net* net, long n, float* h_out) {
  if (!ctx || !net || !h_out) return fail(QG_ERR_ARG, "qg_net_render: bad arguments");
  if (net->g.inputs() != 0) return fail(QG_ERR_ARITY, "render needs a net with 0 inputs (process.rs:1345)");
  qg_bank* b = qg_bank_create(ctx, net, 1, nullptr, nullptr);
  if (!b) return g_code != QG_OK ? g_code : QG_ERR_UNSUPPORTED;
  int rc = qg_bank_render(b, n, QG_LAYOUT_FRAME_MAJOR, 1, h_out);
  qg_bank_free(b);
  return rc;
}
int qg_net_tick(qg_ctx* ctx, const qg_net* net, const float* in, int n_in, float* out, int n_out) {
  if (!ctx || !net) return fail(QG_ERR_ARG, "qg_net_tick: bad arguments");
  if (net->g.inputs() != n_in || net->g.outputs() != n_out) return fail(QG_ERR_ARITY, "tick: arity mismatch (process.rs:1322)");
  qg_bank* b = qg_bank_create(ctx, net, 1, nullptr, nullptr);
  if (!b) return g_code != QG_OK ? g_code : QG_ERR_UNSUPPORTED;
  int rc = qg_bank_process(b, 1, QG_LAYOUT_FRAME_MAJOR, in, out);
  qg_bank_free(b);
  return rc;
}

}  // extern "C"
