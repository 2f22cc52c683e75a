"""Single-graph offline render (the reference's `render` op: ONE voice) on every kernel family: ms for 10 s at 48 kHz."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quartz_b200 as qb
from quartz_b200.graphs import build, pipe, stack, L
T = 480000
def sr(g): return {"op": "sr()", "net": g, "n": 48000.0}
GRAPHS = {
    "sine(440)": sr(L("sine(440)")),
    "white >> lowpass >> highpole": sr(pipe("white()", "lowpass(900,3)", "highpole(30)")),
    "saw >> lowpass * ar": sr({"op": "*", "n": 0.0, "inputs": [pipe("saw(110)", "lowpass(1200,2)"), L("ar(0.01,1,2,4)")]}),
    "fm: sine -> sine >> butterpass >> tanh": sr(pipe(pipe("sine(3)", "mul(200)", "add(440)"), "sine()", "butterpass(3000)", "tanh()")),
    "8 detuned sines summed >> resonator": sr(pipe({"op": "+", "n": 0.0, "inputs": [L(f"sine({220 + 1.3 * k})") for k in range(8)]}, "resonator(800,40)")),
}
for name, expr in GRAPHS.items():
    row = []
    for pname, path in (("auto", qb.PATH_AUTO), ("lane_block", qb.PATH_INTERP), ("lane_sample", qb.PATH_INTERP_SAMPLE)):
        bank = qb.Bank(build(expr, qb.Net), 1).set_path(path)
        d = qb.lib().qg_device_alloc(bank.ctx.h, T * 4 * bank.n_out)
        for _ in range(2):
            bank.reset(); bank.ctx.synchronize()
            t0 = time.perf_counter(); bank.render_device(T, d); bank.ctx.synchronize(); dt = time.perf_counter() - t0
        row.append(f"{pname}={bank.kernel()} {dt * 1e3:.1f} ms")
        qb.lib().qg_device_free(bank.ctx.h, d)
    print(f"{name:45s} " + " | ".join(row))
