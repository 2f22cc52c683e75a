"""A/B of k_polysynth_x2 development variants on one box: python scripts/ab_poly.py [T]
QG_POLY_VAR: bit 0 wrap (0 FSET.BF, 1 VIMNMX), bits 1-2 unroll (0: 8 samples, 1: 16, 2: 4); QG_POLY_SEGMENTS=1 forces one time segment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quartz_b200 as qb
from quartz_b200 import workloads
from quartz_b200.graphs import build
T = int(sys.argv[1]) if len(sys.argv) > 1 else 480000
wl = workloads.c3_polysynth(T=T)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = qb.Context(0, stream=stream.cuda_stream)
bank = qb.Bank(build(wl.expr, qb.Net), wl.V, raw=wl.raw, salts=wl.salts, ctx=ctx)
d = torch.empty((wl.V // wl.group) * T, dtype=torch.float32, device="cuda")
ref = None
for seg in ("1", "0"):
    for var in sys.argv[2:] or ["0", "1", "2", "3", "4", "5"]:
        os.environ["QG_POLY_VAR"] = var
        os.environ["QG_POLY_SEGMENTS"] = seg
        ms = []
        for _ in range(4):
            bank.reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); bank.render_device(T, d.data_ptr(), group=wl.group); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        chk = float(d[:100000].double().abs().sum())
        print(f"segments={'1' if seg == '1' else 'auto'} var={var}: {min(ms[1:]):.3f} ms (runs {['%.2f' % m for m in ms]}) checksum {chk:.6f}", flush=True)
