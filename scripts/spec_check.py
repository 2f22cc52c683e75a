"""K1s (tape-specialised lane kernel, qg_bank_set_path(QG_PATH_SPECIALISED)) against the default kernels on the four
configs[4] archetypes: bit-for-bit comparison of the group-mixed outputs and device render times.
usage: python scripts/spec_check.py [V per archetype] [T]"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quartz_b200 as qb
from quartz_b200 import workloads
from quartz_b200.graphs import build

V = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 9600
ctx = qb.default_context()
lib = qb.lib()
rows = V // 32
d_out = lib.qg_device_alloc(ctx.h, rows * T * 4)
res = []
for wl in workloads.c5_mixed(V=4 * V, T=T):
    net = build(wl.expr, qb.Net)
    out = {}
    for name, path in (("default", qb.PATH_AUTO), ("specialised", qb.PATH_SPECIALISED)):
        bank = qb.Bank(net, wl.V, raw=wl.raw, salts=wl.salts, ctx=ctx)
        t0 = time.perf_counter()
        try:
            bank.set_path(path)
        except qb.QuartzGpuError as e:
            out[name] = {"error": str(e)[:400]}
            continue
        t_set = time.perf_counter() - t0
        host = bank.render(T, group=32)
        best = 1e9
        for _ in range(2):
            bank.reset(); ctx.synchronize()
            t0 = time.perf_counter()
            bank.render_device(T, d_out, group=32); ctx.synchronize()
            best = min(best, time.perf_counter() - t0)
        out[name] = {"kernel": bank.kernel(), "ms": round(best * 1e3, 3), "set_path_s": round(t_set, 2), "host": host}
    line = {"workload": wl.name, "V": wl.V, "T": T}
    for k, v in out.items():
        line[k] = {kk: vv for kk, vv in v.items() if kk != "host"}
    if all("host" in v for v in out.values()):
        a, b = out["default"]["host"], out["specialised"]["host"]
        line["max_abs_diff"] = float(np.abs(a - b).max()); line["bit_identical"] = bool((a.view(np.uint32) == b.view(np.uint32)).all())
        line["speedup"] = round(out["default"]["ms"] / out["specialised"]["ms"], 2)
    print(json.dumps(line), flush=True)
    res.append(line)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/spec_check.json", "w"), indent=1)
