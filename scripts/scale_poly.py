"""k_polysynth time vs bank size (how the kernel scales with warps per scheduler): python scripts/scale_poly.py [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quartz_b200 as qb
from quartz_b200 import workloads
from quartz_b200.graphs import build
T = int(sys.argv[1]) if len(sys.argv) > 1 else 120000
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = qb.Context(0, stream=stream.cuda_stream)
for V in (9472, 18944, 37888, 65536, 75776, 113664, 151552, 303104):
    wl = workloads.c3_polysynth(V=V, T=T)
    bank = qb.Bank(build(wl.expr, qb.Net), wl.V, raw=wl.raw, salts=wl.salts, ctx=ctx)
    d = torch.empty((wl.V // wl.group) * T, dtype=torch.float32, device="cuda")
    ms = []
    for _ in range(3):
        bank.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bank.render_device(T, d.data_ptr(), group=wl.group); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = min(ms[1:])
    warps = V / 64
    print(f"V={V} warps={warps:.0f} ({warps / 592:.2f} per SMSP): {t:.3f} ms, {V * T / t / 1e6:.1f} G voice-samples/s, "
          f"{t * 1e-3 * 1.965e9 / T:.1f} cycles per sample-step of the slowest SMSP", flush=True)
    del bank, d
