"""Render one workload at a reduced length (profiling helper): python scripts/prof_wl.py c4 96000 [V]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quartz_b200 as qb
from quartz_b200 import workloads
from quartz_b200.graphs import build
name, T = sys.argv[1], int(sys.argv[2])
kw = {"T": T}
if len(sys.argv) > 3: kw["V"] = int(sys.argv[3])
wls = workloads.WORKLOADS[name](**kw)
wls = wls if isinstance(wls, list) else [wls]
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = qb.Context(0, stream=stream.cuda_stream)
for wl in wls:
    net = build(wl.expr, qb.Net)
    bank = qb.Bank(net, wl.V, raw=wl.raw, salts=wl.salts, ctx=ctx)
    rows = (wl.V // wl.group) * net.outputs()
    d = torch.empty(rows * T, dtype=torch.float32, device="cuda")
    for _ in range(2):
        bank.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bank.render_device(T, d.data_ptr(), group=wl.group); e1.record(); torch.cuda.synchronize()
    print(wl.name, bank.kernel(), f"{e0.elapsed_time(e1):.3f} ms", f"{wl.V * T / e0.elapsed_time(e1) / 1e6:.2f} G voice-samples/s")
