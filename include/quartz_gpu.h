/* quartz_gpu.h — C ABI of libquartz_gpu.so, the B200-native evaluator for quartz's audio-graph hot path.
 *
 * What each entry point replaces in the reference (syther-labs/quartz, paths relative to /root/reference):
 *   qg_str_to_net            src/functions.rs:111   str_to_net(op) -> Net
 *   qg_connect               src/process.rs:1719-1876  connective circles  + * - >> | & ^ !
 *   qg_array_op              src/process.rs:1669-1717  branch() bus() pipe() stack() sum() product()
 *   qg_get/quantize/wave     src/process.rs:1450-1477, 1652-1667
 *   qg_feedback              src/process.rs:1479-1515  (FunDSP FeedbackUnit)
 *   qg_kr / qg_reset_every   src/process.rs:1540-1580  kr() s() reset()          (src/nodes.rs:235-371)
 *   qg_trig_reset            src/process.rs:1582-1613  trig_reset() reset_v()    (src/nodes.rs:377-453)
 *   qg_seq_select            src/process.rs:1615-1650  seq() select()            (src/nodes.rs:10-121)
 *   qg_net_set_sample_rate   src/process.rs:1571-1573  sr()
 *   qg_net_inputs/outputs/size   AudioUnit::inputs/outputs, Net::size  (src/process.rs:1345, 1752; src/commands.rs:1054-1077)
 *   qg_net_render            src/process.rs:1332-1357  the `render` op loop  for _ in 0..len { net.tick(&[], &mut s) }
 *   qg_net_tick              src/process.rs:1311-1330  the `apply` op (one frame)
 *   qg_bank_*                many independent voices (one `render` circle each in the reference) evaluated at once;
 *                            qg_bank_process is the block path (AudioUnit::process, src/audio.rs:85-118)
 *
 * Conventions: plain pointers and sizes only; every function that can fail returns an int status (0 = ok) or a
 * NULL handle and records a message retrievable with qg_last_error(); nothing throws across the boundary
 * (the reference builds with panic='abort', Cargo.toml:57).  A context is bound to one GPU and one CUDA stream
 * and is thread-compatible (one call at a time per context).  There is NO CPU fallback: device entry points
 * fail with QG_ERR_CUDA when no GPU is usable, and graphs containing ops without a GPU lowering fail with
 * QG_ERR_UNSUPPORTED instead of rendering silence.
 */
#ifndef QUARTZ_GPU_H
#define QUARTZ_GPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qg_net qg_net;    /* host-side audio graph (value semantics: clone = deep copy, like Net::clone) */
typedef struct qg_ctx qg_ctx;    /* one GPU + one stream */
typedef struct qg_bank qg_bank;  /* V voices that share one op tape, resident on a context */

enum { QG_OK = 0, QG_ERR_ARG = 1, QG_ERR_UNSUPPORTED = 2, QG_ERR_CUDA = 3, QG_ERR_ARITY = 4, QG_ERR_MISMATCH = 5 };
enum { QG_LAYOUT_VOICE_MAJOR = 0, QG_LAYOUT_FRAME_MAJOR = 1 };
enum { QG_PATH_AUTO = 0, QG_PATH_INTERP = 1, QG_PATH_TV = 2, QG_PATH_INTERP_SAMPLE = 3, QG_PATH_SPECIALISED = 4, QG_PATH_SPECTRAL = 5 };   /* kernel
   selection: AUTO picks a fused kernel when the tape matches, the time-vector interpreter (one CTA per voice) for spectral /
   small feed-forward banks, else the lane interpreter (block mode for feed-forward tapes); INTERP forces the lane
   interpreter, INTERP_SAMPLE its sample-by-sample kernel; SPECIALISED compiles a lane kernel for this bank's tape with
   NVRTC (seconds, once per bank; uniform tapes only; fails with QG_ERR_UNSUPPORTED when NVRTC or the tape does not allow
   it); AUTO chooses it by itself for lane banks whose work pays for the compile; SPECTRAL forces the frame-parallel path
   for rfft -> bin chain -> ifft patches (src/nodes.rs:601-700), which AUTO takes for bulk renders of such patches */
enum { QG_SAMPLE_F32 = 0, QG_SAMPLE_I16 = 1, QG_SAMPLE_U16 = 2 };   /* cpal::SampleFormat as src/audio.rs:56-59 dispatches it */

const char* qg_last_error(void);
const char* qg_version(void);

/* ---- graph construction (host only) ---- */
qg_net* qg_str_to_net(const char* op);
qg_net* qg_net_new(int inputs, int outputs);
qg_net* qg_net_clone(const qg_net* net);
void qg_net_free(qg_net* net);
int qg_net_inputs(const qg_net* net);
int qg_net_outputs(const qg_net* net);
int qg_net_size(const qg_net* net);
int qg_net_set_sample_rate(qg_net* net, double sample_rate);
const char* qg_net_unsupported(const qg_net* net);   /* NULL, or the name of an op that cannot be lowered yet */
qg_net* qg_connect(const char* op, const qg_net* const* nets, int n_nets, double number, int node_limit);
qg_net* qg_array_op(const char* kind, const char* op_str, const float* arr, int n);
qg_net* qg_get(const float* arr, int n);
qg_net* qg_quantize(const float* arr, int n);
qg_net* qg_wave(const float* arr, int n);
qg_net* qg_feedback(const qg_net* net, int has_delay, double delay_seconds);
qg_net* qg_kr(const qg_net* net, double n, int preserve_time);
qg_net* qg_reset_every(const qg_net* net, double seconds);
qg_net* qg_trig_reset(const qg_net* net, int variable);
qg_net* qg_seq_select(int is_seq, const qg_net* const* nets, int n_nets);
qg_net* qg_live_io(const char* name);   /* in() adc() buffin() buffout() monitor(): offline equivalents */
qg_net* qg_var(float value);             /* var(): src/process.rs:1373-1385 */
/* lowering introspection */
int qg_net_raw_count(const qg_net* net);                       /* number of op-string parameters, lowering order */
int qg_net_raw_params(const qg_net* net, float* out, int cap);
uint64_t qg_net_signature(const qg_net* net);                  /* equal <=> same tape modulo parameter values */
int qg_net_tape_info(const qg_net* net, int* n_instr, int* n_params, int* n_state, int* n_temps, int* divergent);
/* the CUDA translation unit the tape specialiser compiles for this graph (NVRTC; see quartz_b200/csrc/spec_kernel.cuh):
   returns its length (copies at most cap - 1 characters into buf, which may be NULL) or a negated QG_ERR_* status */
long qg_net_spec_source(const qg_net* net, char* buf, long cap);
/* the template voice's DEVICE parameters (filter coefficients, pan weights, 1/sr ... as the lowering derives them from the
   op-string parameters): returns their count, copies at most cap values; < 0: -QG_ERR_* */
int qg_net_device_params(const qg_net* net, float* out, int cap);
/* 1 when the graph qualifies for the frame-parallel spectral path (QG_PATH_SPECTRAL): rfft -> stateless bin chain -> ifft
   segments (src/nodes.rs:601-700) fed by pure functions of time; fills the plan's shape.  0 when it does not, < 0: -QG_ERR_* */
int qg_net_spectral_info(const qg_net* net, int* n_segments, int* n_streams, int* n_instr, int* round_len);
/* the translation unit NVRTC compiles for that plan (quartz_b200/csrc/spectral_kernel.cuh); contract of qg_net_spec_source */
long qg_net_spectral_spec_source(const qg_net* net, char* buf, long cap);

/* ---- device ---- */
qg_ctx* qg_ctx_create(int device, void* cuda_stream /* cudaStream_t, or NULL for a private stream */);
void qg_ctx_destroy(qg_ctx* ctx);
int qg_ctx_synchronize(qg_ctx* ctx);
double qg_ctx_measure_fp32_tflops(qg_ctx* ctx);               /* FFMA micro-benchmark: the FP32-pipe roofline (TFLOP/s), < 0 on error */
long qg_ctx_launch_count(const qg_ctx* ctx);                   /* kernels launched by this library so far */
void* qg_device_alloc(qg_ctx* ctx, size_t bytes);
void qg_device_free(qg_ctx* ctx, void* p);
void* qg_host_alloc_pinned(size_t bytes);
void qg_host_free_pinned(void* p);

/* A bank evaluates n_voices copies of `tmpl`.  raw (optional): [n_voices][qg_net_raw_count] per-voice op-string
 * parameters; salts (optional): per-voice 64-bit salt mixed into every hash-seeded state (sine phase, noise seed). */
qg_bank* qg_bank_create(qg_ctx* ctx, const qg_net* tmpl, long n_voices, const float* raw, const uint64_t* salts);
/* Same, from n structurally identical nets (what n `render` circles would hold). */
qg_bank* qg_bank_from_nets(qg_ctx* ctx, const qg_net* const* nets, long n_nets, const uint64_t* salts);
void qg_bank_free(qg_bank* bank);
int qg_bank_reset(qg_bank* bank);                              /* AudioUnit::reset for every voice */
int qg_bank_set_path(qg_bank* bank, int path);                 /* QG_PATH_* */
const char* qg_bank_kernel(const qg_bank* bank);               /* name of the kernel family the next render uses */
long qg_bank_out_rows(const qg_bank* bank, int group);         /* rows of the voice-major output: (V/group)*outputs */

/* Render n_samples for every voice (nets with 0 inputs).  Output: voice-major [V/group][outputs][n_samples] or
 * frame-major [n_samples][V][outputs] (group must be 1).  group in {1,2,4,8,16,32}: consecutive voices are summed
 * left to right and scaled by 1/group.  State persists across calls (block-wise streaming). */
int qg_bank_render_device(qg_bank* bank, long n_samples, int layout, int group, float* d_out);
int qg_bank_render(qg_bank* bank, long n_samples, int layout, int group, float* h_out);
/* Block path with inputs: h_in voice-major [V][inputs][n] or frame-major [n][V][inputs] (same layout as the output). */
int qg_bank_process(qg_bank* bank, long n_samples, int layout, const float* h_in, float* h_out);
/* var() (src/process.rs:1382-1385): overwrite one op-string parameter for every voice; takes effect on the next render. */
int qg_bank_set_raw(qg_bank* bank, int raw_index, float value);
/* Stream path (src/audio.rs:85-118) for a one-voice bank: n frames, non-normal -> 0, clamp to [-1,1], interleaved L R;
 * mono graphs get a silent right channel, other arities play silence (src/process.rs:1896-1905). */
int qg_bank_render_stereo(qg_bank* bank, long n_frames, float* h_frames);
/* The same frames in the device's sample type (`T::from_sample`, src/audio.rs:47-59, 115-116): h_frames holds
 * 2 * n_frames samples of f32, i16 ((s * 32768) as i16, saturating) or u16 (offset binary). */
int qg_bank_render_stereo_as(qg_bank* bank, long n_frames, int sample_format /* QG_SAMPLE_* */, void* h_frames);
/* Value copy of a bank WITH its state (the reference deep-clones a Net, state included, on every hop: src/process.rs:1316,
 * 1336, 1499, 1558, 1895; AudioUnit is DynClone): both banks continue independently and identically. */
qg_bank* qg_bank_clone(const qg_bank* bank);
/* Sum the rows of a device buffer [rows][n] into d_out[n] (rows added in index order), scaled. */
int qg_mix_rows_device(qg_ctx* ctx, const float* d_rows, long rows, long n, float scale, float* d_out);

/* ---- single-graph conveniences with the reference's semantics ---- */
int qg_net_render(qg_ctx* ctx, const qg_net* net, long n_samples, float* h_out /* frame-major [n][outputs] */);
int qg_net_tick(qg_ctx* ctx, const qg_net* net, const float* in, int n_in, float* out, int n_out);

#ifdef __cplusplus
}
#endif
#endif
