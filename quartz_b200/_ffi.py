"""ctypes binding of libquartz_gpu.so (C ABI declared in include/quartz_gpu.h).  The library is built in-tree by
quartz_b200/build.sh; importing this module never compiles anything and never falls back to another backend."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libquartz_gpu.so")
_LIB = None

# name -> (restype, argtypes); kept in one table so tests can check it against include/quartz_gpu.h
vp, ci, cl, cd, cf, u64 = C.c_void_p, C.c_int, C.c_long, C.c_double, C.c_float, C.c_uint64
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)
SIGNATURES = {
    "qg_last_error": (C.c_char_p, []),
    "qg_version": (C.c_char_p, []),
    "qg_str_to_net": (vp, [C.c_char_p]),
    "qg_net_new": (vp, [ci, ci]),
    "qg_net_clone": (vp, [vp]),
    "qg_net_free": (None, [vp]),
    "qg_net_inputs": (ci, [vp]),
    "qg_net_outputs": (ci, [vp]),
    "qg_net_size": (ci, [vp]),
    "qg_net_set_sample_rate": (ci, [vp, cd]),
    "qg_net_unsupported": (C.c_char_p, [vp]),
    "qg_connect": (vp, [C.c_char_p, C.POINTER(vp), ci, cd, ci]),
    "qg_array_op": (vp, [C.c_char_p, C.c_char_p, fp, ci]),
    "qg_get": (vp, [fp, ci]),
    "qg_quantize": (vp, [fp, ci]),
    "qg_wave": (vp, [fp, ci]),
    "qg_feedback": (vp, [vp, ci, cd]),
    "qg_kr": (vp, [vp, cd, ci]),
    "qg_reset_every": (vp, [vp, cd]),
    "qg_trig_reset": (vp, [vp, ci]),
    "qg_seq_select": (vp, [ci, C.POINTER(vp), ci]),
    "qg_live_io": (vp, [C.c_char_p]),
    "qg_var": (vp, [cf]),
    "qg_net_raw_count": (ci, [vp]),
    "qg_net_raw_params": (ci, [vp, fp, ci]),
    "qg_net_signature": (u64, [vp]),
    "qg_net_tape_info": (ci, [vp, ip, ip, ip, ip, ip]),
    "qg_net_spec_source": (cl, [vp, C.c_char_p, cl]),
    "qg_net_spectral_info": (ci, [vp, ip, ip, ip, ip]),
    "qg_net_spectral_spec_source": (cl, [vp, C.c_char_p, cl]),
    "qg_net_device_params": (ci, [vp, fp, ci]),
    "qg_ctx_create": (vp, [ci, vp]),
    "qg_ctx_destroy": (None, [vp]),
    "qg_ctx_synchronize": (ci, [vp]),
    "qg_ctx_launch_count": (cl, [vp]),
    "qg_ctx_measure_fp32_tflops": (cd, [vp]),
    "qg_device_alloc": (vp, [vp, C.c_size_t]),
    "qg_device_free": (None, [vp, vp]),
    "qg_host_alloc_pinned": (vp, [C.c_size_t]),
    "qg_host_free_pinned": (None, [vp]),
    "qg_bank_create": (vp, [vp, vp, cl, fp, C.POINTER(u64)]),
    "qg_bank_from_nets": (vp, [vp, C.POINTER(vp), cl, C.POINTER(u64)]),
    "qg_bank_free": (None, [vp]),
    "qg_bank_reset": (ci, [vp]),
    "qg_bank_set_path": (ci, [vp, ci]),
    "qg_bank_kernel": (C.c_char_p, [vp]),
    "qg_bank_out_rows": (cl, [vp, ci]),
    "qg_bank_render_device": (ci, [vp, cl, ci, ci, vp]),
    "qg_bank_render": (ci, [vp, cl, ci, ci, vp]),
    "qg_bank_process": (ci, [vp, cl, ci, vp, vp]),
    "qg_bank_set_raw": (ci, [vp, ci, cf]),
    "qg_bank_render_stereo": (ci, [vp, cl, vp]),
    "qg_bank_render_stereo_as": (ci, [vp, cl, ci, vp]),
    "qg_bank_clone": (vp, [vp]),
    "qg_mix_rows_device": (ci, [vp, vp, cl, cl, cf, vp]),
    "qg_net_render": (ci, [vp, vp, cl, vp]),
    "qg_net_tick": (ci, [vp, vp, fp, ci, fp, ci]),
}


class QuartzGpuError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise QuartzGpuError(
                f"{LIB_PATH} is missing: build it with quartz_b200/build.sh (or __graft_entry__.build()). "
                "quartz_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _LIB = L
    return _LIB


def last_error():
    return lib().qg_last_error().decode()


def check(rc):
    if rc != 0:
        raise QuartzGpuError(f"quartz_gpu error {rc}: {last_error()}")
