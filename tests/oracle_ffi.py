"""ctypes binding of the CPU oracle (oracle/libqoracle.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs — never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", _ORACLE_DIR])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_ORACLE_DIR, "libqoracle.so")
        if not os.path.exists(so):
            build_oracle()
        L = C.CDLL(so)
        vp, ci, cd, cl = C.c_void_p, C.c_int, C.c_double, C.c_long
        fp = C.POINTER(C.c_float)
        sig = {
            "qo_last_error": (C.c_char_p, []),
            "qo_str_to_net": (vp, [C.c_char_p, C.POINTER(ci)]),
            "qo_net_new": (vp, [ci, ci]),
            "qo_clone": (vp, [vp]), "qo_free": (None, [vp]),
            "qo_inputs": (ci, [vp]), "qo_outputs": (ci, [vp]), "qo_size": (ci, [vp]),
            "qo_reset": (None, [vp]), "qo_set_sample_rate": (None, [vp, cd]),
            "qo_set_salt": (None, [vp, C.c_uint64]),
            "qo_connect": (vp, [C.c_char_p, C.POINTER(vp), ci, cd, ci]),
            "qo_array_op": (vp, [C.c_char_p, C.c_char_p, fp, ci, C.POINTER(ci)]),
            "qo_get": (vp, [fp, ci]), "qo_quantize": (vp, [fp, ci]), "qo_wave": (vp, [fp, ci]),
            "qo_feedback": (vp, [vp, ci, cd]), "qo_kr": (vp, [vp, cd, ci]),
            "qo_reset_every": (vp, [vp, cd]), "qo_trig_reset": (vp, [vp, ci]),
            "qo_seq_select": (vp, [ci, C.POINTER(vp), ci]), "qo_live_io": (vp, [C.c_char_p]),
            "qo_var": (vp, [C.c_float]),
            "qo_tick": (ci, [vp, fp, ci, fp, ci]), "qo_render": (ci, [vp, cl, fp]),
            "qo_process": (ci, [vp, cl, fp, fp]),
            "qo_render_bank": (ci, [C.POINTER(vp), ci, cl, ci, ci, fp]),
            "qo_real_fft": (None, [fp, ci, fp]), "qo_inverse_fft": (None, [fp, ci, fp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _LIB = L
    return _LIB


class OracleUnsupported(Exception):
    pass


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class ONet:
    """Handle on an oracle Net with the reference's construction vocabulary."""

    def __init__(self, h):
        if not h:
            raise OracleUnsupported(lib().qo_last_error().decode())
        self.h = h

    def __del__(self):
        try:
            lib().qo_free(self.h)
        except Exception:
            pass

    # ---- construction
    @staticmethod
    def str_to_net(op):
        st = C.c_int(0)
        return ONet(lib().qo_str_to_net(op.encode(), C.byref(st)))

    @staticmethod
    def empty(ni=0, no=0):
        return ONet(lib().qo_net_new(ni, no))

    @staticmethod
    def connect(op, nets, number=0.0, node_limit=500):
        arr = (C.c_void_p * len(nets))(*[n.h for n in nets])
        return ONet(lib().qo_connect(op.encode(), arr, len(nets), float(number), int(node_limit)))

    @staticmethod
    def array_op(kind, op_str, arr):
        a = _f32(arr)
        st = C.c_int(0)
        return ONet(lib().qo_array_op(kind.encode(), op_str.encode(), _fptr(a), len(a), C.byref(st)))

    @staticmethod
    def get(arr):
        a = _f32(arr)
        return ONet(lib().qo_get(_fptr(a), len(a)))

    @staticmethod
    def quantize(arr):
        a = _f32(arr)
        return ONet(lib().qo_quantize(_fptr(a), len(a)))

    @staticmethod
    def wave(arr):
        a = _f32(arr)
        return ONet(lib().qo_wave(_fptr(a), len(a)))

    @staticmethod
    def feedback(net, delay=None):
        return ONet(lib().qo_feedback(net.h, 0 if delay is None else 1, 0.0 if delay is None else float(delay)))

    @staticmethod
    def kr(net, n, preserve_time=False):
        return ONet(lib().qo_kr(net.h, float(n), 1 if preserve_time else 0))

    @staticmethod
    def reset_every(net, s):
        return ONet(lib().qo_reset_every(net.h, float(s)))

    @staticmethod
    def trig_reset(net):
        return ONet(lib().qo_trig_reset(net.h, 0))

    @staticmethod
    def reset_v(net):
        return ONet(lib().qo_trig_reset(net.h, 1))

    @staticmethod
    def seq(nets):
        arr = (C.c_void_p * len(nets))(*[n.h for n in nets])
        return ONet(lib().qo_seq_select(1, arr, len(nets)))

    @staticmethod
    def select(nets):
        arr = (C.c_void_p * len(nets))(*[n.h for n in nets])
        return ONet(lib().qo_seq_select(0, arr, len(nets)))

    @staticmethod
    def live_io(name):
        return ONet(lib().qo_live_io(name.encode()))

    @staticmethod
    def var(value):
        return ONet(lib().qo_var(float(value)))

    # ---- AudioUnit surface
    def clone(self):
        return ONet(lib().qo_clone(self.h))

    def inputs(self):
        return lib().qo_inputs(self.h)

    def outputs(self):
        return lib().qo_outputs(self.h)

    def size(self):
        return lib().qo_size(self.h)

    def reset(self):
        lib().qo_reset(self.h)

    def set_sample_rate(self, sr):
        lib().qo_set_sample_rate(self.h, float(sr))
        return self

    def set_salt(self, salt):
        lib().qo_set_salt(self.h, int(salt) & 0xFFFFFFFFFFFFFFFF)
        return self

    def tick(self, inp):
        a = _f32(inp)
        out = np.zeros(self.outputs(), dtype=np.float32)
        rc = lib().qo_tick(self.h, _fptr(a), len(a), _fptr(out), len(out))
        if rc:
            raise ValueError(lib().qo_last_error().decode())
        return out

    def render(self, n):
        """frame-major [n, outputs]"""
        out = np.zeros((n, max(self.outputs(), 1)), dtype=np.float32)
        rc = lib().qo_render(self.h, n, _fptr(out))
        if rc:
            raise ValueError(lib().qo_last_error().decode())
        return out[:, :self.outputs()]

    def process(self, inp):
        a = _f32(inp).reshape(-1, max(self.inputs(), 1))
        out = np.zeros((a.shape[0], max(self.outputs(), 1)), dtype=np.float32)
        lib().qo_process(self.h, a.shape[0], _fptr(a), _fptr(out))
        return out[:, :self.outputs()]


def render_bank(nets, n_samples, group=1, threads=1, out=None):
    """voice-major [V/group, T]; `out` lets a timing caller pass a buffer whose pages are already resident"""
    arr = (C.c_void_p * len(nets))(*[n.h for n in nets])
    if out is None:
        out = np.zeros((len(nets) // group, n_samples), dtype=np.float32)
    assert out.shape == (len(nets) // group, n_samples) and out.dtype == np.float32 and out.flags.c_contiguous
    lib().qo_render_bank(arr, len(nets), n_samples, group, threads, _fptr(out))
    return out


def real_fft(x):
    a = _f32(x)
    out = np.zeros((len(a) // 2 + 1, 2), dtype=np.float32)
    lib().qo_real_fft(_fptr(a), len(a), _fptr(out))
    return out[:, 0] + 1j * out[:, 1]


def inverse_fft(z):
    z = np.asarray(z, dtype=np.complex64)
    a = np.ascontiguousarray(np.stack([z.real, z.imag], axis=1).astype(np.float32))
    out = np.zeros_like(a)
    lib().qo_inverse_fft(_fptr(a), len(z), _fptr(out))
    return out[:, 0] + 1j * out[:, 1]
