"""CPU-side fuzz of the C ABI (no GPU): graph-level constructors with pathological numbers (NaN, inf, 2^64 ...) + lowering.\nRun under `timeout`: a hang IS a finding.  usage: fuzz_constructors.py [seed] [count]"""
import random, sys, ctypes as C, math
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np
from quartz_b200 import _ffi
lib = _ffi.lib()
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
W = [float("nan"), float("inf"), float("-inf"), 0.0, -0.0, -1.0, 1.0, 0.5, 1e-9, 1e9, 1e30, -1e30, 3.7, 44100.0, 2.0**31, 2.0**63, 2.0**64]
leafs = [b"white()", b"sine(440)", b"dc(1)", b"lowpass()", b"dc(1,2)", b"pass()", b"sink()", b"ramp()", b"mul(0.5)", b"delay(0.001)"]
errs = {}
def leaf(): return lib.qg_str_to_net(random.choice(leafs))
def arr():
    n = random.randint(0, 12)
    a = (C.c_float * max(n, 1))(*[random.choice(W) for _ in range(max(n, 1))])
    return a, n
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 20000):
    k = random.randint(0, 9)
    a = leaf(); b = leaf(); made = None
    if k == 0: made = lib.qg_feedback(a, random.randint(0, 1), random.choice(W))
    elif k == 1: made = lib.qg_kr(a, random.choice(W), random.randint(0, 1))
    elif k == 2: made = lib.qg_reset_every(a, random.choice(W))
    elif k == 3: made = lib.qg_trig_reset(a, random.randint(0, 1))
    elif k == 4:
        nets = (C.c_void_p * 2)(a, b); made = lib.qg_seq_select(random.randint(0, 1), nets, 2)
    elif k == 5:
        x, n = arr(); made = random.choice([lib.qg_get, lib.qg_quantize, lib.qg_wave])(x, n)
    elif k == 6:
        nets = (C.c_void_p * 2)(a, b)
        made = lib.qg_connect(random.choice([b">>", b"|", b"&", b"^", b"+", b"*", b"-", b"!", b"?", b""]), nets, random.randint(0, 2), random.choice(W), random.choice([0, 1, 10, 500, -1]))
    elif k == 7:
        x, n = arr(); made = lib.qg_array_op(random.choice([b"sum()", b"stack()", b"pipe()", b"branch()", b"bus()", b"product()", b"thru()", b"x"]), random.choice([b"sine(#)", b"dc(#)", b"delay(#)", b"#", b"lowpass(#,1)", b"kr(#)"]), x, n)
    elif k == 8: made = lib.qg_var(random.choice(W))
    else:
        lib.qg_net_set_sample_rate(a, random.choice(W)); made = lib.qg_net_clone(a)
    assert made, k
    ni, no, sz = lib.qg_net_inputs(made), lib.qg_net_outputs(made), lib.qg_net_size(made)
    assert 0 <= ni <= 4096 and 0 <= no <= 4096, (k, ni, no)
    if lib.qg_net_unsupported(made) is None:
        rc = lib.qg_net_tape_info(made, None, None, None, None, None)
        if rc: errs[lib.qg_last_error().decode()[:70]] = k
    for h in (a, b, made): lib.qg_net_free(h)
print("ok"); print(errs)
