"""Host-side mirror of the reference's operator interface for the audio-graph path.

`Net` plays the role of FunDSP's `Net`/`AudioUnit` as quartz uses it: `str_to_net` (src/functions.rs:111), the
connective circles (src/process.rs:1719-1876), graph-level constructors (src/process.rs:1450-1667), the arity
queries everything else relies on, and `tick` / `render` / `process` which here run on the GPU through
libquartz_gpu.so.  `Bank` evaluates many structurally identical voices in one launch."""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import QuartzGpuError, check, lib

LAYOUT_VOICE_MAJOR, LAYOUT_FRAME_MAJOR = 0, 1
PATH_AUTO, PATH_INTERP, PATH_TV, PATH_INTERP_SAMPLE, PATH_SPECIALISED, PATH_SPECTRAL = 0, 1, 2, 3, 4, 5
SAMPLE_F32, SAMPLE_I16, SAMPLE_U16 = 0, 1, 2      # cpal::SampleFormat as audio.rs:56-59 dispatches it
NODE_LIMIT_DEFAULT = 500   # src/main.rs:72

_CTX = {}


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Context:
    """One GPU + one CUDA stream (qg_ctx)."""

    def __init__(self, device=0, stream=None):
        self.h = lib().qg_ctx_create(int(device), C.c_void_p(stream) if stream else None)
        if not self.h:
            raise QuartzGpuError(_ffi.last_error())
        self.device = device

    def synchronize(self):
        check(lib().qg_ctx_synchronize(self.h))

    def launch_count(self):
        return lib().qg_ctx_launch_count(self.h)

    def measure_fp32_tflops(self):
        """FFMA micro-benchmark on this GPU: the FP32-pipe roofline for compute-bound workloads"""
        v = lib().qg_ctx_measure_fp32_tflops(self.h)
        if v < 0:
            raise QuartzGpuError(_ffi.last_error())
        return v

    def close(self):
        if self.h:
            lib().qg_ctx_destroy(self.h)
            self.h = None


def default_context(device=0):
    if device not in _CTX:
        _CTX[device] = Context(device)
    return _CTX[device]


class Net:
    def __init__(self, h):
        if not h:
            raise QuartzGpuError(_ffi.last_error())
        self.h = h
        self._bank = None

    def __del__(self):
        try:
            self._bank = None
            lib().qg_net_free(self.h)
        except Exception:
            pass

    # ------------------------------------------------------------------ construction (host only)
    @staticmethod
    def str_to_net(op):
        return Net(lib().qg_str_to_net(op.encode()))

    @staticmethod
    def empty(ni=0, no=0):
        return Net(lib().qg_net_new(ni, no))

    @staticmethod
    def connect(op, nets, number=0.0, node_limit=NODE_LIMIT_DEFAULT):
        arr = (C.c_void_p * max(len(nets), 1))(*[n.h for n in nets])
        return Net(lib().qg_connect(op.encode(), arr, len(nets), float(number), int(node_limit)))

    @staticmethod
    def array_op(kind, op_str, arr):
        a = _f32(arr)
        return Net(lib().qg_array_op(kind.encode(), op_str.encode(), _fptr(a), len(a)))

    @staticmethod
    def get(arr):
        a = _f32(arr)
        return Net(lib().qg_get(_fptr(a), len(a)))

    @staticmethod
    def quantize(arr):
        a = _f32(arr)
        return Net(lib().qg_quantize(_fptr(a), len(a)))

    @staticmethod
    def wave(arr):
        a = _f32(arr)
        return Net(lib().qg_wave(_fptr(a), len(a)))

    @staticmethod
    def feedback(net, delay=None):
        return Net(lib().qg_feedback(net.h, 0 if delay is None else 1, 0.0 if delay is None else float(delay)))

    @staticmethod
    def kr(net, n, preserve_time=False):
        return Net(lib().qg_kr(net.h, float(n), 1 if preserve_time else 0))

    @staticmethod
    def reset_every(net, s):
        return Net(lib().qg_reset_every(net.h, float(s)))

    @staticmethod
    def trig_reset(net):
        return Net(lib().qg_trig_reset(net.h, 0))

    @staticmethod
    def reset_v(net):
        return Net(lib().qg_trig_reset(net.h, 1))

    @staticmethod
    def seq(nets):
        arr = (C.c_void_p * max(len(nets), 1))(*[n.h for n in nets])
        return Net(lib().qg_seq_select(1, arr, len(nets)))

    @staticmethod
    def select(nets):
        arr = (C.c_void_p * max(len(nets), 1))(*[n.h for n in nets])
        return Net(lib().qg_seq_select(0, arr, len(nets)))

    @staticmethod
    def live_io(name):
        return Net(lib().qg_live_io(name.encode()))

    @staticmethod
    def var(value):
        return Net(lib().qg_var(float(value)))

    # operators, as on a FunDSP Net (no arity guard here: the guards belong to connect(), like in process.rs)
    def __rshift__(self, o):
        return Net.connect(">>", [self, o])

    def __or__(self, o):
        return Net.connect("|", [self, o])

    def __and__(self, o):
        return Net.connect("&", [self, o])

    def __xor__(self, o):
        return Net.connect("^", [self, o])

    def __add__(self, o):
        return Net.connect("+", [self, o])

    def __mul__(self, o):
        return Net.connect("*", [self, o])

    def __sub__(self, o):
        return Net.connect("-", [self, o])

    def __invert__(self):
        return Net.connect("!", [self])

    # ------------------------------------------------------------------ AudioUnit surface
    def clone(self):
        return Net(lib().qg_net_clone(self.h))

    def inputs(self):
        return lib().qg_net_inputs(self.h)

    def outputs(self):
        return lib().qg_net_outputs(self.h)

    def size(self):
        return lib().qg_net_size(self.h)

    def unsupported(self):
        s = lib().qg_net_unsupported(self.h)
        return s.decode() if s else None

    def set_sample_rate(self, sr):
        check(lib().qg_net_set_sample_rate(self.h, float(sr)))
        self._bank = None
        return self

    def raw_params(self):
        n = lib().qg_net_raw_count(self.h)
        out = np.zeros(max(n, 1), dtype=np.float32)
        lib().qg_net_raw_params(self.h, _fptr(out), n)
        return out[:n]

    def signature(self):
        return lib().qg_net_signature(self.h)

    def spec_source(self):
        """the CUDA translation unit the tape specialiser would compile for this graph (uniform tapes only)"""
        n = lib().qg_net_spec_source(self.h, None, 0)
        if n < 0:
            check(int(-n))
        buf = C.create_string_buffer(int(n) + 1)
        lib().qg_net_spec_source(self.h, buf, int(n) + 1)
        return buf.value.decode()

    def spectral_spec_source(self):
        """the CUDA translation unit compiled for this graph's frame-parallel spectral plan (K5s)"""
        n = lib().qg_net_spectral_spec_source(self.h, None, 0)
        if n < 0:
            check(int(-n))
        buf = C.create_string_buffer(int(n) + 1)
        lib().qg_net_spectral_spec_source(self.h, buf, int(n) + 1)
        return buf.value.decode()

    def tape_info(self):
        v = [C.c_int(0) for _ in range(5)]
        check(lib().qg_net_tape_info(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("n_instr", "n_params", "n_state", "n_temps", "divergent"), [x.value for x in v]))

    def device_params(self):
        """the template voice's device parameters as the lowering derives them (filter coefficients, pan weights ...)"""
        n = lib().qg_net_device_params(self.h, None, 0)
        if n < 0:
            raise QuartzGpuError(_ffi.last_error())
        out = np.zeros(max(n, 1), np.float32)
        lib().qg_net_device_params(self.h, _fptr(out), n)
        return out[:n]

    def spectral_info(self):
        """shape of the frame-parallel spectral plan (QG_PATH_SPECTRAL), or None when the graph does not qualify"""
        v = [C.c_int(0) for _ in range(4)]
        rc = lib().qg_net_spectral_info(self.h, *[C.byref(x) for x in v])
        if rc < 0:
            raise QuartzGpuError(_ffi.last_error())
        return dict(zip(("n_segments", "n_streams", "n_instr", "round_len"), [x.value for x in v])) if rc else None

    def _voice(self, ctx=None):
        if self._bank is None:
            self._bank = Bank(self, 1, ctx=ctx)
        return self._bank

    def reset(self):
        if self._bank is not None:
            self._bank.reset()

    def tick(self, inp, ctx=None):
        """One frame (the `apply` op, process.rs:1322-1325).  State advances, like AudioUnit::tick."""
        a = _f32(inp).reshape(-1)
        if len(a) != self.inputs():
            raise QuartzGpuError("tick: arity mismatch (process.rs:1322)")
        out = self._voice(ctx).process(a.reshape(1, 1, -1) if len(a) else None, 1, layout=LAYOUT_FRAME_MAJOR)
        return out.reshape(-1)

    def render(self, n, ctx=None):
        """The `render` op (process.rs:1345-1351): n ticks of a 0-input net.  Returns frame-major [n, outputs]."""
        if self.inputs() != 0:
            raise QuartzGpuError("render needs a net with 0 inputs (process.rs:1345)")
        return self._voice(ctx).render(n, layout=LAYOUT_FRAME_MAJOR).reshape(n, self.outputs())

    def process(self, inp, ctx=None):
        """Block path (AudioUnit::process): inp frame-major [n, inputs] -> [n, outputs]."""
        a = _f32(inp).reshape(-1, max(self.inputs(), 1))
        n = a.shape[0]
        out = self._voice(ctx).process(a.reshape(n, 1, -1), n, layout=LAYOUT_FRAME_MAJOR)
        return out.reshape(n, self.outputs())


RENDER_LEN_CAP = 10_000_000   # process.rs:1342


def render_op(net, number, arr=None, ctx=None):
    """The `render` circle (process.rs:1339-1353) with its guards: `len = min(number as usize, 10^7)`; a net that is not
    0-in / 1-out leaves the circle's array untouched (returns `arr`); otherwise the array is cleared and filled with `len`
    ticks (state advances, as `net.tick` does on the circle's own Net)."""
    x = float(number)
    n = 0 if (x != x or x <= 0.0) else int(min(x, float(RENDER_LEN_CAP)))      # saturating `as usize`, NaN -> 0
    n = min(n, RENDER_LEN_CAP)
    if net.inputs() != 0 or net.outputs() != 1:
        return arr
    if n == 0:
        return np.zeros(0, dtype=np.float32)
    return net.render(n, ctx=ctx)[:, 0]


def apply_op(net, input_arr, arr=None, ctx=None):
    """The `apply` circle (process.rs:1318-1326): one frame through the graph when the input array has exactly
    `net.inputs()` entries (output resized to `net.outputs()`), else the circle's array stays as it was."""
    a = _f32(input_arr).reshape(-1)
    if net.inputs() != len(a):
        return arr
    if net.outputs() == 0:
        return np.zeros(0, dtype=np.float32)
    return net.tick(a, ctx=ctx)


def str_to_net(op):
    return Net.str_to_net(op)


class Bank:
    """V structurally identical voices evaluated together (qg_bank)."""

    def __init__(self, template, n_voices=None, raw=None, salts=None, nets=None, ctx=None, device=0):
        self.ctx = ctx or default_context(device)
        L = lib()
        sp = None
        if salts is not None:
            salts = np.ascontiguousarray(np.asarray(salts, dtype=np.uint64))
            sp = salts.ctypes.data_as(C.POINTER(C.c_uint64))
        if nets is not None:
            arr = (C.c_void_p * len(nets))(*[n.h for n in nets])
            self.h = L.qg_bank_from_nets(self.ctx.h, arr, len(nets), sp)
            template = nets[0]
            n_voices = len(nets)
        else:
            rp = None
            if raw is not None:
                raw = _f32(raw)
                if raw.shape != (n_voices, L.qg_net_raw_count(template.h)):
                    raise QuartzGpuError(f"raw must be [n_voices, {L.qg_net_raw_count(template.h)}], got {raw.shape}")
                rp = _fptr(raw)
            self.h = L.qg_bank_create(self.ctx.h, template.h, int(n_voices), rp, sp)
        if not self.h:
            raise QuartzGpuError(_ffi.last_error())
        self.n_voices = int(n_voices)
        self.n_in = template.inputs()
        self.n_out = template.outputs()

    def __del__(self):
        try:
            if self.h:
                lib().qg_bank_free(self.h)
        except Exception:
            pass

    def reset(self):
        check(lib().qg_bank_reset(self.h))

    def set_path(self, path):
        check(lib().qg_bank_set_path(self.h, path))
        return self

    def kernel(self):
        return lib().qg_bank_kernel(self.h).decode()

    def render(self, n, layout=LAYOUT_VOICE_MAJOR, group=1, out=None):
        """voice-major -> [V/group, outputs, n]; frame-major -> [n, V, outputs]"""
        if layout == LAYOUT_VOICE_MAJOR:
            shape = (self.n_voices // group, self.n_out, n)
        else:
            shape = (n, self.n_voices, self.n_out)
        if out is None:
            out = np.zeros(shape, dtype=np.float32)
        check(lib().qg_bank_render(self.h, int(n), layout, group, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_raw(self, raw_index, value):
        """var(): rewrite one op-string parameter for every voice (process.rs:1382-1385)"""
        check(lib().qg_bank_set_raw(self.h, int(raw_index), float(value)))

    def render_stereo(self, n, sample_format=SAMPLE_F32):
        """stream path (audio.rs:85-118): [n, 2] sanitised, clamped, interleaved frames of a one-voice bank, in the device's
        sample type (`T::from_sample`, audio.rs:56-59, 115-116): f32, i16 or u16"""
        out = np.zeros((n, 2), dtype={SAMPLE_I16: np.int16, SAMPLE_U16: np.uint16}.get(sample_format, np.float32))
        check(lib().qg_bank_render_stereo_as(self.h, int(n), int(sample_format), out.ctypes.data_as(C.c_void_p)))
        return out

    def clone(self):
        """value copy WITH state, like `Net::clone` (process.rs:1316, 1336, 1499, 1558, 1895): both continue identically"""
        h = lib().qg_bank_clone(self.h)
        if not h:
            raise QuartzGpuError(_ffi.last_error())
        b = Bank.__new__(Bank)
        b.ctx, b.h, b.n_voices, b.n_in, b.n_out = self.ctx, h, self.n_voices, self.n_in, self.n_out
        return b

    def render_device(self, n, d_out, layout=LAYOUT_VOICE_MAJOR, group=1):
        check(lib().qg_bank_render_device(self.h, int(n), layout, group, C.c_void_p(d_out)))

    def process(self, inp, n, layout=LAYOUT_VOICE_MAJOR):
        """inp voice-major [V, inputs, n] or frame-major [n, V, inputs]"""
        if layout == LAYOUT_VOICE_MAJOR:
            shape = (self.n_voices, self.n_out, n)
        else:
            shape = (n, self.n_voices, self.n_out)
        out = np.zeros(shape, dtype=np.float32)
        a = _f32(inp) if (inp is not None and self.n_in) else None
        check(lib().qg_bank_process(self.h, int(n), layout, a.ctypes.data_as(C.c_void_p) if a is not None else None,
                                    out.ctypes.data_as(C.c_void_p)))
        return out
