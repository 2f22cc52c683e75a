"""Single-process multi-GPU render (quartz is ONE process: the patch interpreter owns every graph).

Voices are independent (each `render` circle owns its cloned Net, /root/reference/src/process.rs:1336-1337), so a bank
shards across the GPUs of a box with no collective: one context + one bank per device over a contiguous, group-aligned
voice range, one host thread per device (the C ABI is thread-compatible, one call at a time per context; ctypes releases
the GIL), each writing its rows straight into the caller's voice-major host buffer."""
import threading

import numpy as np

from . import net as _net
from .shard import voice_range


def render_sharded(template, n_voices, n_samples, raw=None, salts=None, group=1, devices=(0,), out=None):
    """Render `n_voices` copies of `template` for `n_samples` on `devices` (a device may be listed more than once).
    Returns voice-major [n_voices // group, outputs, n_samples]; row order does not depend on the device list."""
    n_out = template.outputs()
    rows = n_voices // group
    if out is None:
        out = np.zeros((rows, n_out, n_samples), dtype=np.float32)
    assert out.shape == (rows, n_out, n_samples) and out.dtype == np.float32 and out.flags.c_contiguous
    errors = []

    def work(rank, device):
        try:
            lo, hi = voice_range(n_voices, len(devices), rank, group)
            if hi <= lo:
                return
            ctx = _net.Context(device)
            try:
                bank = _net.Bank(template, hi - lo, raw=None if raw is None else raw[lo:hi],
                                 salts=None if salts is None else salts[lo:hi], ctx=ctx)
                bank.render(n_samples, group=group, out=out[lo // group: hi // group])
                del bank
            finally:
                ctx.close()
        except Exception as e:   # surfaced to the caller below
            errors.append((rank, device, e))

    threads = [threading.Thread(target=work, args=(r, d)) for r, d in enumerate(devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise _net.QuartzGpuError(f"device {errors[0][1]} (rank {errors[0][0]}): {errors[0][2]}")
    return out
