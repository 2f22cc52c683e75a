"""CPU-side fuzz of the C ABI (no GPU): random op strings through qg_str_to_net + lowering.  usage: fuzz_op_strings.py [seed] [count]"""
import random, sys, ctypes as C
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from quartz_b200 import _ffi
from tests import cases
lib = _ffi.lib()
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
names = [c[0].split("(")[0] for c in cases.ARITY] + ["moog", "reverb_stereo", "dsf_saw", "constant", "dc", "fir", "join", "split", "chan", "rfft", "ifft", "samp_delay", "delay", "tap", "mls", "pluck", "meter"]
atoms = ["", "1", "-1", "0", "1e9", "-1e30", "1e-45", "nan", "inf", "-inf", "NaN", "TAU", "PI", "E", "440", "0.5", "999999999999999999999", "1e400", "-0", "0x10", "1,2", ",", ",,", " ", "\t", "peak", "rms", "abc", "-", "+", ".", "1.2.3", "1e", "e1", "1 2", "(", ")", "#", "é", "\x00x"]
n = 0
errs = {}
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 100000):
    k = random.random()
    if k < 0.7:
        nm = random.choice(names)
        args = ",".join(random.choice(atoms) for _ in range(random.randint(0, 12)))
        s = f"{nm}({args})"
        if random.random() < 0.2: s = s[:-1]
        if random.random() < 0.1: s = s.replace("(", "((", 1)
        if random.random() < 0.1: s += random.choice([")", " ", "x", "(1)"])
    else:
        s = "".join(random.choice("abcdefxyz_()0123456789,.-+e \t#") for _ in range(random.randint(0, 40)))
    b = s.encode("utf-8", "ignore").split(b"\x00")[0]
    h = lib.qg_str_to_net(b)
    assert h, s
    ni, no, sz = lib.qg_net_inputs(h), lib.qg_net_outputs(h), lib.qg_net_size(h)
    assert 0 <= ni <= 64 and 0 <= no <= 64 and 0 <= sz <= 1000, (s, ni, no, sz)
    if lib.qg_net_unsupported(h) is None and random.random() < 0.3:
        rc = lib.qg_net_tape_info(h, None, None, None, None, None)
        if rc: errs[lib.qg_last_error().decode()[:60]] = s
        lib.qg_net_signature(h); lib.qg_net_raw_count(h)
    lib.qg_net_free(h); n += 1
print("ok", n); print(errs)
