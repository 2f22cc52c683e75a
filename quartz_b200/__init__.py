"""quartz_b200 — B200-native evaluator for quartz's audio-graph hot path (op strings -> op tape -> sm_100a kernels).
Only what the path needs lives here: csrc/ (CUDA kernels + C ABI, built to libquartz_gpu.so) and the host-side
mirror of the reference's Net/AudioUnit interface (net.py)."""
from ._ffi import LIB_PATH, QuartzGpuError, lib  # noqa: F401
from .net import (LAYOUT_FRAME_MAJOR, LAYOUT_VOICE_MAJOR, NODE_LIMIT_DEFAULT, PATH_AUTO, PATH_INTERP, PATH_INTERP_SAMPLE, PATH_SPECIALISED, PATH_SPECTRAL, PATH_TV, SAMPLE_F32, SAMPLE_I16, SAMPLE_U16, Bank, Context,  # noqa: F401
                  Net, apply_op, default_context, render_op, str_to_net)
from .multi import render_sharded  # noqa: F401,E402
