//! quartz-gpu: Rust binding of libquartz_gpu.so (include/quartz_gpu.h).
//!
//! `GpuNet` mirrors the subset of `fundsp::net::Net` that quartz's patch interpreter uses when it builds audio
//! graphs (src/functions.rs:111 `str_to_net`, src/process.rs:1669-1876 connective ops) and `GpuRender` implements
//! `AudioUnit` (the trait quartz itself implements for `Kr`/`SwapUnit`, src/nodes.rs:256-327) by serving samples
//! from a GPU-rendered look-ahead buffer, so `render` (src/process.rs:1349) and the cpal callback
//! (src/audio.rs:85-118) can consume it unchanged.
use std::ffi::{c_char, c_double, c_float, c_int, c_long, c_void, CStr, CString};

#[repr(C)] pub struct qg_net { _p: [u8; 0] }
#[repr(C)] pub struct qg_ctx { _p: [u8; 0] }
#[repr(C)] pub struct qg_bank { _p: [u8; 0] }

#[allow(dead_code)]   // the whole C surface is declared; the safe wrappers below use what quartz needs
extern "C" {
    fn qg_last_error() -> *const c_char;
    fn qg_str_to_net(op: *const c_char) -> *mut qg_net;
    fn qg_net_clone(n: *const qg_net) -> *mut qg_net;
    fn qg_net_free(n: *mut qg_net);
    fn qg_net_inputs(n: *const qg_net) -> c_int;
    fn qg_net_outputs(n: *const qg_net) -> c_int;
    fn qg_net_size(n: *const qg_net) -> c_int;
    fn qg_net_set_sample_rate(n: *mut qg_net, sr: c_double) -> c_int;
    fn qg_connect(op: *const c_char, nets: *const *const qg_net, n: c_int, number: c_double, node_limit: c_int) -> *mut qg_net;
    fn qg_array_op(kind: *const c_char, op_str: *const c_char, arr: *const c_float, n: c_int) -> *mut qg_net;
    fn qg_get(arr: *const c_float, n: c_int) -> *mut qg_net;
    fn qg_quantize(arr: *const c_float, n: c_int) -> *mut qg_net;
    fn qg_wave(arr: *const c_float, n: c_int) -> *mut qg_net;
    fn qg_feedback(net: *const qg_net, has_delay: c_int, delay_seconds: c_double) -> *mut qg_net;
    fn qg_kr(net: *const qg_net, n: c_double, preserve_time: c_int) -> *mut qg_net;
    fn qg_reset_every(net: *const qg_net, seconds: c_double) -> *mut qg_net;
    fn qg_trig_reset(net: *const qg_net, variable: c_int) -> *mut qg_net;
    fn qg_seq_select(is_seq: c_int, nets: *const *const qg_net, n: c_int) -> *mut qg_net;
    fn qg_live_io(name: *const c_char) -> *mut qg_net;
    fn qg_var(value: c_float) -> *mut qg_net;
    fn qg_net_unsupported(n: *const qg_net) -> *const c_char;
    fn qg_net_tick(c: *mut qg_ctx, n: *const qg_net, input: *const c_float, n_in: c_int, out: *mut c_float, n_out: c_int) -> c_int;
    fn qg_bank_from_nets(c: *mut qg_ctx, nets: *const *const qg_net, n: c_long, salts: *const u64) -> *mut qg_bank;
    fn qg_bank_set_raw(b: *mut qg_bank, raw_index: c_int, value: c_float) -> c_int;
    fn qg_bank_render_stereo(b: *mut qg_bank, n_frames: c_long, frames: *mut c_float) -> c_int;
    fn qg_ctx_create(device: c_int, stream: *mut c_void) -> *mut qg_ctx;
    fn qg_ctx_destroy(c: *mut qg_ctx);
    fn qg_bank_create(c: *mut qg_ctx, t: *const qg_net, v: c_long, raw: *const c_float, salts: *const u64) -> *mut qg_bank;
    fn qg_bank_free(b: *mut qg_bank);
    fn qg_bank_reset(b: *mut qg_bank) -> c_int;
    fn qg_bank_render(b: *mut qg_bank, n: c_long, layout: c_int, group: c_int, out: *mut c_float) -> c_int;
    // the rest of include/quartz_gpu.h: introspection, device buffers, kernel selection, block path with inputs
    fn qg_version() -> *const c_char;
    fn qg_net_new(n_in: c_int, n_out: c_int) -> *mut qg_net;
    fn qg_net_render(c: *mut qg_ctx, n: *const qg_net, len: c_long, out: *mut c_float) -> c_int;
    fn qg_net_raw_count(n: *const qg_net) -> c_int;
    fn qg_net_raw_params(n: *const qg_net, out: *mut c_float, cap: c_int) -> c_int;
    fn qg_net_signature(n: *const qg_net) -> u64;
    fn qg_net_tape_info(n: *const qg_net, n_instr: *mut c_int, n_params: *mut c_int, n_state: *mut c_int, n_temps: *mut c_int,
                        divergent: *mut c_int) -> c_int;
    fn qg_net_spec_source(n: *const qg_net, buf: *mut c_char, cap: c_long) -> c_long;
    fn qg_net_device_params(n: *const qg_net, out: *mut c_float, cap: c_int) -> c_int;
    fn qg_net_spectral_spec_source(n: *const qg_net, buf: *mut c_char, cap: c_long) -> c_long;
    fn qg_net_spectral_info(n: *const qg_net, n_segments: *mut c_int, n_streams: *mut c_int, n_instr: *mut c_int, round_len: *mut c_int) -> c_int;
    fn qg_ctx_synchronize(c: *mut qg_ctx) -> c_int;
    fn qg_ctx_launch_count(c: *const qg_ctx) -> c_long;
    fn qg_ctx_measure_fp32_tflops(c: *mut qg_ctx) -> c_double;
    fn qg_device_alloc(c: *mut qg_ctx, bytes: usize) -> *mut c_void;
    fn qg_device_free(c: *mut qg_ctx, p: *mut c_void);
    fn qg_host_alloc_pinned(bytes: usize) -> *mut c_void;
    fn qg_host_free_pinned(p: *mut c_void);
    fn qg_bank_set_path(b: *mut qg_bank, path: c_int) -> c_int;
    fn qg_bank_kernel(b: *const qg_bank) -> *const c_char;
    fn qg_bank_out_rows(b: *const qg_bank, group: c_int) -> c_long;
    fn qg_bank_render_device(b: *mut qg_bank, n: c_long, layout: c_int, group: c_int, d_out: *mut c_float) -> c_int;
    fn qg_bank_process(b: *mut qg_bank, n: c_long, layout: c_int, input: *const c_float, out: *mut c_float) -> c_int;
    fn qg_bank_render_stereo_as(b: *mut qg_bank, n_frames: c_long, sample_format: c_int, frames: *mut c_void) -> c_int;
    fn qg_bank_clone(b: *const qg_bank) -> *mut qg_bank;
    fn qg_mix_rows_device(c: *mut qg_ctx, d_rows: *const c_float, rows: c_long, n: c_long, scale: c_float, d_out: *mut c_float) -> c_int;
}

/// kernel selection of a bank (QG_PATH_* in include/quartz_gpu.h)
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum Path { Auto = 0, Interp = 1, TimeVector = 2, InterpSample = 3, Specialised = 4, Spectral = 5 }

pub fn last_error() -> String {
    unsafe { CStr::from_ptr(qg_last_error()).to_string_lossy().into_owned() }
}

/// Host-side graph with value semantics (quartz deep-clones a Net on every hop).
pub struct GpuNet(*mut qg_net);
unsafe impl Send for GpuNet {}
unsafe impl Sync for GpuNet {}

impl GpuNet {
    /// src/functions.rs:111
    pub fn from_op(op: &str) -> GpuNet {
        let c = CString::new(op).unwrap_or_default();
        GpuNet(unsafe { qg_str_to_net(c.as_ptr()) })
    }
    pub fn inputs(&self) -> usize { unsafe { qg_net_inputs(self.0) as usize } }
    pub fn outputs(&self) -> usize { unsafe { qg_net_outputs(self.0) as usize } }
    pub fn size(&self) -> usize { unsafe { qg_net_size(self.0) as usize } }
    pub fn set_sample_rate(&mut self, sr: f64) { unsafe { qg_net_set_sample_rate(self.0, sr); } }
    /// connective circle (src/process.rs:1719-1876): op in {"+","*","-",">>","|","&","^","!"}
    pub fn connect(op: &str, nets: &[&GpuNet], number: f32, node_limit: usize) -> GpuNet {
        let c = CString::new(op).unwrap_or_default();
        let ptrs: Vec<*const qg_net> = nets.iter().map(|n| n.0 as *const qg_net).collect();
        GpuNet(unsafe { qg_connect(c.as_ptr(), ptrs.as_ptr(), ptrs.len() as c_int, number as c_double, node_limit as c_int) })
    }
}
/// graph-level constructors of the patch interpreter (src/process.rs:1450-1667); `arr` is the linked circle's `Arr`
impl GpuNet {
    fn wrap(p: *mut qg_net) -> GpuNet { GpuNet(p) }
    fn ptrs(nets: &[&GpuNet]) -> Vec<*const qg_net> { nets.iter().map(|n| n.0 as *const qg_net).collect() }
    pub fn get(arr: &[f32]) -> GpuNet { Self::wrap(unsafe { qg_get(arr.as_ptr(), arr.len() as c_int) }) }                 // :1455
    pub fn quantize(arr: &[f32]) -> GpuNet { Self::wrap(unsafe { qg_quantize(arr.as_ptr(), arr.len() as c_int) }) }       // :1468-1471
    pub fn wave(arr: &[f32]) -> GpuNet { Self::wrap(unsafe { qg_wave(arr.as_ptr(), arr.len() as c_int) }) }               // :1658-1662
    pub fn feedback(net: &GpuNet, delay: Option<f32>) -> GpuNet {                                                         // :1500-1507
        Self::wrap(unsafe { qg_feedback(net.0, delay.is_some() as c_int, delay.unwrap_or(0.) as c_double) })
    }
    pub fn kr(net: &GpuNet, n: f32, preserve_time: bool) -> GpuNet {                                                      // :1562-1566
        Self::wrap(unsafe { qg_kr(net.0, n as c_double, preserve_time as c_int) })
    }
    pub fn reset(net: &GpuNet, seconds: f32) -> GpuNet { Self::wrap(unsafe { qg_reset_every(net.0, seconds as c_double) }) }   // :1568-1569
    pub fn trig_reset(net: &GpuNet) -> GpuNet { Self::wrap(unsafe { qg_trig_reset(net.0, 0) }) }                          // :1602-1606
    pub fn reset_v(net: &GpuNet) -> GpuNet { Self::wrap(unsafe { qg_trig_reset(net.0, 1) }) }
    pub fn seq(nets: &[&GpuNet]) -> GpuNet {                                                                              // :1636-1647
        let p = Self::ptrs(nets);
        Self::wrap(unsafe { qg_seq_select(1, p.as_ptr(), p.len() as c_int) })
    }
    pub fn select(nets: &[&GpuNet]) -> GpuNet {
        let p = Self::ptrs(nets);
        Self::wrap(unsafe { qg_seq_select(0, p.as_ptr(), p.len() as c_int) })
    }
    /// branch() bus() pipe() stack() sum() product() with `#` substitution (src/process.rs:1669-1717)
    pub fn array_op(kind: &str, op_str: &str, arr: &[f32]) -> GpuNet {
        let (k, o) = (CString::new(kind).unwrap_or_default(), CString::new(op_str).unwrap_or_default());
        Self::wrap(unsafe { qg_array_op(k.as_ptr(), o.as_ptr(), arr.as_ptr(), arr.len() as c_int) })
    }
    /// var(): a constant the control plane rewrites (src/process.rs:1373-1385); see GpuRender::set_var
    pub fn var(value: f32) -> GpuNet { Self::wrap(unsafe { qg_var(value) }) }
    /// in() adc() buffin() buffout() monitor(): offline equivalents of the live-I/O units
    pub fn live_io(name: &str) -> GpuNet {
        let c = CString::new(name).unwrap_or_default();
        Self::wrap(unsafe { qg_live_io(c.as_ptr()) })
    }
    /// Some(op) when the graph mentions an op that has no GPU lowering: the caller keeps the CPU Net for this circle
    pub fn unsupported(&self) -> Option<String> {
        let p = unsafe { qg_net_unsupported(self.0) };
        if p.is_null() { None } else { Some(unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned()) }
    }
}
impl Clone for GpuNet {
    fn clone(&self) -> Self { GpuNet(unsafe { qg_net_clone(self.0) }) }
}
impl Drop for GpuNet {
    fn drop(&mut self) { unsafe { qg_net_free(self.0) } }
}

/// One GPU + stream.
pub struct GpuContext(*mut qg_ctx);
unsafe impl Send for GpuContext {}
impl GpuContext {
    pub fn new(device: i32) -> Result<GpuContext, String> {
        let p = unsafe { qg_ctx_create(device, std::ptr::null_mut()) };
        if p.is_null() { Err(last_error()) } else { Ok(GpuContext(p)) }
    }
}
impl Drop for GpuContext {
    fn drop(&mut self) { unsafe { qg_ctx_destroy(self.0) } }
}

/// `render` op replacement (src/process.rs:1345-1351): `len` ticks of a 0-input, 1-output net.
pub fn render(ctx: &GpuContext, net: &GpuNet, len: usize) -> Result<Vec<f32>, String> {
    if net.inputs() != 0 || net.outputs() != 1 { return Ok(Vec::new()); }   // the reference's arity guard (process.rs:1345)
    let len = len.min(10_000_000);                                          // and its length cap (process.rs:1341-1342)
    let bank = unsafe { qg_bank_create(ctx.0, net.0, 1, std::ptr::null(), std::ptr::null()) };
    if bank.is_null() { return Err(last_error()); }
    let mut out = vec![0f32; len];
    let rc = unsafe { qg_bank_render(bank, len as c_long, 0, 1, out.as_mut_ptr()) };
    unsafe { qg_bank_free(bank) };
    if rc != 0 { Err(last_error()) } else { Ok(out) }
}

/// `apply` op replacement (src/process.rs:1311-1330): one frame through the graph
pub fn apply(ctx: &GpuContext, net: &GpuNet, input: &[f32]) -> Result<Vec<f32>, String> {
    if net.inputs() != input.len() { return Ok(Vec::new()); }              // the reference's arity guard (process.rs:1322)
    let mut out = vec![0f32; net.outputs()];
    let rc = unsafe { qg_net_tick(ctx.0, net.0, input.as_ptr(), input.len() as c_int, out.as_mut_ptr(), out.len() as c_int) };
    if rc != 0 { Err(last_error()) } else { Ok(out) }
}

/// Many `render` circles whose graphs differ only in their constants: one bank, one launch.  `salts` re-seeds the
/// hash-derived state (sine phase, noise seed) per voice; returns voice-major rows [nets.len()][len].
pub fn render_many(ctx: &GpuContext, nets: &[&GpuNet], salts: Option<&[u64]>, len: usize) -> Result<Vec<f32>, String> {
    let len = len.min(10_000_000);
    let p: Vec<*const qg_net> = nets.iter().map(|n| n.0 as *const qg_net).collect();
    let bank = unsafe { qg_bank_from_nets(ctx.0, p.as_ptr(), p.len() as c_long, salts.map_or(std::ptr::null(), |s| s.as_ptr())) };
    if bank.is_null() { return Err(last_error()); }
    let mut out = vec![0f32; nets.len() * len];
    let rc = unsafe { qg_bank_render(bank, len as c_long, 0, 1, out.as_mut_ptr()) };
    unsafe { qg_bank_free(bank) };
    if rc != 0 { Err(last_error()) } else { Ok(out) }
}

/// V voices that share one graph structure, resident on one GPU: the unit of throughput work (BASELINE configs[1..4]).
/// `raw` is the per-voice op-string parameter table [V][raw_count] (None: every voice = the template), `salts` re-seed the
/// hash-derived state per voice.
pub struct GpuBank { bank: *mut qg_bank, voices: usize, inputs: usize, outputs: usize }
unsafe impl Send for GpuBank {}
impl GpuBank {
    pub fn new(ctx: &GpuContext, template: &GpuNet, voices: usize, raw: Option<&[f32]>, salts: Option<&[u64]>) -> Result<GpuBank, String> {
        let raw_count = unsafe { qg_net_raw_count(template.0) } as usize;
        if let Some(r) = raw { if r.len() != voices * raw_count { return Err("raw table must be [voices][raw_count]".into()); } }
        if let Some(s) = salts { if s.len() != voices { return Err("one salt per voice".into()); } }
        let bank = unsafe { qg_bank_create(ctx.0, template.0, voices as c_long, raw.map_or(std::ptr::null(), |r| r.as_ptr()),
                                           salts.map_or(std::ptr::null(), |s| s.as_ptr())) };
        if bank.is_null() { return Err(last_error()); }
        Ok(GpuBank { bank, voices, inputs: template.inputs(), outputs: template.outputs() })
    }
    /// name of the kernel family the next render uses ("k_noise_svf_scan", "k_interp_blk", "k_interp_tv", "k_spec", ...)
    pub fn kernel(&self) -> String { unsafe { CStr::from_ptr(qg_bank_kernel(self.bank)).to_string_lossy().into_owned() } }
    /// `Path::Specialised` compiles a kernel for this bank's tape with NVRTC (once; fails if NVRTC or the tape do not allow it)
    pub fn set_path(&mut self, path: Path) -> Result<(), String> {
        if unsafe { qg_bank_set_path(self.bank, path as c_int) } != 0 { Err(last_error()) } else { Ok(()) }
    }
    pub fn reset(&mut self) { unsafe { qg_bank_reset(self.bank); } }
    /// voice-major rows [voices / group][outputs][len]; `group` > 1 mixes consecutive voices (sum left to right, scaled by 1/group)
    pub fn render(&mut self, len: usize, group: usize) -> Result<Vec<f32>, String> {
        let rows = unsafe { qg_bank_out_rows(self.bank, group as c_int) } as usize;
        let mut out = vec![0f32; rows * len];
        let rc = unsafe { qg_bank_render(self.bank, len as c_long, 0, group as c_int, out.as_mut_ptr()) };
        if rc != 0 { Err(last_error()) } else { Ok(out) }
    }
    /// block path with external inputs (AudioUnit::process): `input` voice-major [voices][inputs][len]
    pub fn process(&mut self, len: usize, input: &[f32]) -> Result<Vec<f32>, String> {
        if input.len() != self.voices * self.inputs * len { return Err("input must be [voices][inputs][len]".into()); }
        let mut out = vec![0f32; self.voices * self.outputs * len];
        let rc = unsafe { qg_bank_process(self.bank, len as c_long, 0, input.as_ptr(), out.as_mut_ptr()) };
        if rc != 0 { Err(last_error()) } else { Ok(out) }
    }
}
/// value copy WITH state (parameters, filter state, delay lines), like `Net::clone`
impl GpuBank {
    pub fn try_clone(&self) -> Result<GpuBank, String> {
        let bank = unsafe { qg_bank_clone(self.bank) };
        if bank.is_null() { Err(last_error()) } else { Ok(GpuBank { bank, voices: self.voices, inputs: self.inputs, outputs: self.outputs }) }
    }
}
impl Drop for GpuBank {
    fn drop(&mut self) { unsafe { qg_bank_free(self.bank) } }
}

/// AudioUnit served from GPU-rendered blocks (0 inputs, C outputs); what `slot.set(.., Box::new(..))`
/// (src/process.rs:1897) receives instead of the CPU Net.  A failed device call latches `error` and the unit plays
/// silence from then on (nothing may unwind: the reference builds with panic='abort', Cargo.toml:57).
pub struct GpuRender {
    ctx: *mut qg_ctx,    // borrowed: the GpuContext must outlive every unit made from it
    net: GpuNet,         // the graph this unit renders (set_sample_rate rebuilds the bank from it)
    bank: *mut qg_bank,
    outputs: usize,
    block: Vec<f32>,     // voice-major [outputs][BLOCK]
    pos: usize,
    pub error: Option<String>,
}
unsafe impl Send for GpuRender {}
unsafe impl Sync for GpuRender {}
const BLOCK: usize = 4096;

/// sample types of the cpal stream (src/audio.rs:56-59); QG_SAMPLE_* in include/quartz_gpu.h
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum SampleFormat { F32 = 0, I16 = 1, U16 = 2 }

impl GpuRender {
    pub fn new(ctx: &GpuContext, net: &GpuNet) -> Result<GpuRender, String> {
        let bank = unsafe { qg_bank_create(ctx.0, net.0, 1, std::ptr::null(), std::ptr::null()) };
        if bank.is_null() { return Err(last_error()); }
        let outputs = net.outputs();
        Ok(GpuRender { ctx: ctx.0, net: net.clone(), bank, outputs, block: vec![0.0; outputs * BLOCK], pos: BLOCK, error: None })
    }
    /// var() update from the control plane (src/process.rs:1382-1385): takes effect at the next block
    pub fn set_var(&mut self, raw_index: usize, value: f32) -> Result<(), String> {
        if unsafe { qg_bank_set_raw(self.bank, raw_index as c_int, value) } != 0 { Err(last_error()) } else { Ok(()) }
    }
    /// `n` sanitised, clamped, interleaved stereo frames straight from the device (src/audio.rs:85-118)
    pub fn stereo_frames(&mut self, n: usize) -> Result<Vec<f32>, String> {
        let mut frames = vec![0f32; 2 * n];
        if unsafe { qg_bank_render_stereo(self.bank, n as c_long, frames.as_mut_ptr()) } != 0 { Err(last_error()) } else { Ok(frames) }
    }
    /// the same frames as the stream's own sample type (`T::from_sample`, src/audio.rs:115-116)
    pub fn stereo_frames_i16(&mut self, n: usize) -> Result<Vec<i16>, String> {
        let mut frames = vec![0i16; 2 * n];
        let rc = unsafe { qg_bank_render_stereo_as(self.bank, n as c_long, SampleFormat::I16 as c_int, frames.as_mut_ptr() as *mut c_void) };
        if rc != 0 { Err(last_error()) } else { Ok(frames) }
    }
    pub fn stereo_frames_u16(&mut self, n: usize) -> Result<Vec<u16>, String> {
        let mut frames = vec![0u16; 2 * n];
        let rc = unsafe { qg_bank_render_stereo_as(self.bank, n as c_long, SampleFormat::U16 as c_int, frames.as_mut_ptr() as *mut c_void) };
        if rc != 0 { Err(last_error()) } else { Ok(frames) }
    }
    fn refill(&mut self) {
        if self.error.is_none() && unsafe { qg_bank_render(self.bank, BLOCK as c_long, 0, 1, self.block.as_mut_ptr()) } != 0 {
            self.error = Some(last_error());
        }
        if self.error.is_some() { self.block.iter_mut().for_each(|s| *s = 0.0); }   // never replay a stale block
        self.pos = 0;
    }
}
impl Drop for GpuRender {
    fn drop(&mut self) { unsafe { qg_bank_free(self.bank) } }
}
/// FunDSP's AudioUnit is DynClone and quartz deep-clones a Net, state included, on every hop (src/process.rs:1316, 1336,
/// 1499, 1558, 1895): the clone owns a device-to-device copy of parameters, state and delay lines and resumes at the same
/// position of the same look-ahead block.  If the device copy fails the clone plays silence with `error` set.
impl Clone for GpuRender {
    fn clone(&self) -> Self {
        let mut bank = unsafe { qg_bank_clone(self.bank) };
        let mut error = self.error.clone();
        if bank.is_null() {
            error = Some(last_error());
            bank = std::ptr::null_mut();   // qg_bank_free(NULL) and every render on NULL are harmless (QG_ERR_ARG)
        }
        GpuRender { ctx: self.ctx, net: self.net.clone(), bank, outputs: self.outputs, block: self.block.clone(), pos: self.pos, error }
    }
}

impl fundsp::audiounit::AudioUnit for GpuRender {
    fn reset(&mut self) { unsafe { qg_bank_reset(self.bank); } self.pos = BLOCK; }
    /// `sr()` reaches every unit of the graph (src/process.rs:1571-1573).  Coefficients and delay lengths are derived from
    /// the sample rate when the tape is lowered, so the bank is rebuilt from the stored graph; like FunDSP's own
    /// `set_sample_rate` on delay lines, that restarts the state.
    fn set_sample_rate(&mut self, sample_rate: f64) {
        self.net.set_sample_rate(sample_rate);
        let bank = unsafe { qg_bank_create(self.ctx, self.net.0, 1, std::ptr::null(), std::ptr::null()) };
        if bank.is_null() { self.error = Some(last_error()); return; }
        unsafe { qg_bank_free(self.bank) };
        self.bank = bank;
        self.pos = BLOCK;
    }
    fn tick(&mut self, _input: &[f32], output: &mut [f32]) {
        if self.pos == BLOCK { self.refill(); }
        for c in 0..self.outputs { output[c] = self.block[c * BLOCK + self.pos]; }
        self.pos += 1;
    }
    fn process(&mut self, size: usize, _input: &fundsp::buffer::BufferRef, output: &mut fundsp::buffer::BufferMut) {
        for i in 0..size {
            if self.pos == BLOCK { self.refill(); }
            for c in 0..self.outputs { output.set_f32(c, i, self.block[c * BLOCK + self.pos]); }
            self.pos += 1;
        }
    }
    fn inputs(&self) -> usize { 0 }
    fn outputs(&self) -> usize { self.outputs }
    fn route(&mut self, input: &fundsp::signal::SignalFrame, _frequency: f64) -> fundsp::signal::SignalFrame {
        fundsp::signal::Routing::Arbitrary(0.0).route(input, self.outputs)
    }
    fn get_id(&self) -> u64 { 1130 }
    fn ping(&mut self, _probe: bool, hash: fundsp::math::AttoHash) -> fundsp::math::AttoHash { hash.hash(self.get_id()) }
    fn footprint(&self) -> usize { core::mem::size_of::<Self>() }
    fn allocate(&mut self) {}
}
