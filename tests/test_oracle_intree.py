"""A SECOND, independent restatement of the reference's in-tree nodes (pure Python / numpy scalars, written from
/root/reference/src/nodes.rs line by line) against the C++ oracle, bit for bit, on seeded inputs that hit the edge cases the
Rust semantics define: saturating `as usize` casts (negative, NaN, huge), strict `<` tie-breaking, out-of-range indices,
zero / negative-zero triggers.  CPU only — this hardens the pinned tier of oracle/ beyond the eight scene vectors."""
import math
import zlib

import numpy as np
import pytest

from tests.graphs import build, pipe
from tests.oracle_ffi import ONet

f32 = np.float32


def as_usize(x):
    """Rust `f32 as usize`: saturating, NaN -> 0"""
    x = float(x)
    if math.isnan(x) or x <= 0.0:
        return 0
    return 2**64 - 1 if x >= 2.0**64 else int(x)


def inputs(rng, n, cols):
    x = rng.uniform(-3, 40, (n, cols)).astype(np.float32)
    x[rng.integers(0, n, n // 10), rng.integers(0, cols, n // 10)] = 0.0
    x[rng.integers(0, n, 6), rng.integers(0, cols, 6)] = [np.nan, np.inf, -np.inf, -0.0, 3e9, 1e30]
    return x


def test_ramp_nodes_rs_476_483():
    rng = np.random.default_rng(1)
    fr = rng.uniform(-50, 30000, 4000).astype(np.float32)
    sr = f32(44100.0)                                     # default sample rate of the node (nodes.rs:466)
    val, exp = f32(0.0), []
    for f in fr:
        exp.append(val)
        val = f32(val + f32(f / sr))
        if val >= f32(1.0):
            val = f32(val - f32(1.0))
    got = build({"op": "ramp()"}, ONet).process(fr[:, None])[:, 0]
    assert (got.view(np.uint32) == np.array(exp, np.float32).view(np.uint32)).all()


def test_shift_reg_nodes_rs_173_185():
    x = inputs(np.random.default_rng(2), 3000, 2)
    reg, exp = [f32(0.0)] * 8, []
    for a, trig in x:
        if trig != 0.0:                                   # NaN != 0 is true, -0.0 != 0 is false
            reg = [a] + reg[:7]
        exp.append(list(reg))
    got = build({"op": "shift_reg()"}, ONet).process(x)
    assert (got.view(np.uint32) == np.array(exp, np.float32).view(np.uint32)).all()


def test_snh_nodes_rs_811_816():
    x = inputs(np.random.default_rng(3), 3000, 2)
    held, exp = f32(0.0), []
    for a, trig in x:
        if trig != 0.0:
            held = a
        exp.append(held)
    got = build({"op": "snh()"}, ONet).process(x)[:, 0]
    assert (got.view(np.uint32) == np.array(exp, np.float32).view(np.uint32)).all()


@pytest.mark.parametrize("arr", [[0, 2, 3, 5, 7, 8, 10, 12], [0.0, 4.0, 7.0], [1.5, 1.5, -2.0, 9.25], [0.0, 1.0, 2.5, 3.0, 4.0, 5.5, 6.0, 7.0, 8.5, 9.0, 10.0, 11.5, 12.0]],
                         ids=["minor", "triad", "dups_unsorted", "thirteen"])
def test_quantizer_nodes_rs_213_228(arr):
    rng = np.random.default_rng(4)
    x = rng.uniform(-60, 60, 3000).astype(np.float32)
    x[:4] = [0.0, -0.0, 1e9, -1e9]
    steps = [f32(v) for v in arr]
    rng_ = f32(steps[-1] - steps[0])                      # Quantizer::new(arr, last - first), process.rs:1468-1471
    exp = []
    with np.errstate(all="ignore"):
        for n in x:
            wrapped = f32(n - f32(rng_ * f32(np.floor(f32(n / rng_)))))
            nearest, dist = f32(0.0), np.finfo(np.float32).max
            for s in steps:
                d = f32(abs(f32(wrapped - s)))
                if d < dist:
                    nearest, dist = s, d
            exp.append(f32(f32(n + nearest) - wrapped))
    got = build({"op": "quantize()", "arr": arr}, ONet).process(x[:, None])[:, 0]
    e = np.array(exp, np.float32)
    assert ((got.view(np.uint32) == e.view(np.uint32)) | (np.isnan(got) & np.isnan(e))).all()


def test_arr_get_nodes_rs_143_149():
    arr = [1.0, -2.0, 3.5, 4.0, 5.0, 6.0, 7.25]
    x = inputs(np.random.default_rng(5), 2000, 1)
    x[:8, 0] = [0, 0.99, 1.0, 6.0, 6.999, 7.0, -1.0, 2.5]
    exp = [f32(arr[as_usize(v)]) if as_usize(v) < len(arr) else f32(0.0) for v in x[:, 0]]
    got = build({"op": "get()", "arr": arr}, ONet).process(x)[:, 0]
    assert (got.view(np.uint32) == np.array(exp, np.float32).view(np.uint32)).all()


def test_samp_delay_nodes_rs_727_730():
    mx = 32
    x = inputs(np.random.default_rng(6), 3000, 2)
    x[:, 1] = np.where(np.isfinite(x[:, 1]), np.abs(x[:, 1]), x[:, 1])
    buf, exp = [f32(0.0)] * mx, []
    for a, idx in x:
        buf = [a] + buf[:-1]                              # push_front + pop_back: index 0 is the CURRENT sample
        k = as_usize(idx)
        exp.append(buf[k] if k < mx else f32(0.0))
    got = build({"op": f"samp_delay({mx})"}, ONet).process(x)[:, 0]
    assert (got.view(np.uint32) == np.array(exp, np.float32).view(np.uint32)).all()


def test_rise_fall_functions_rs_813_822():
    """rise = (pass() ^ tick()) >> map(|i| if i[0] > i[1] {1} else {0}); fall: `<`"""
    x = inputs(np.random.default_rng(7), 2000, 1)[:, 0]
    prev = np.concatenate([[f32(0.0)], x[:-1]])
    with np.errstate(invalid="ignore"):
        rise, fall = (x > prev).astype(np.float32), (x < prev).astype(np.float32)
    assert (build({"op": "rise()"}, ONet).process(x[:, None])[:, 0] == rise).all()
    assert (build({"op": "fall()"}, ONet).process(x[:, None])[:, 0] == fall).all()


# ---------------------------------------------------------------- nested-net nodes around a `dc(f) >> ramp()` inner net
class PyRamp:
    """inner net `dc(f) >> ramp()` at 44.1 kHz: Ramp::reset sets val = 0 (nodes.rs:485-487)"""
    def __init__(self, f):
        self.inc, self.val = f32(f32(f) / f32(44100.0)), f32(0.0)

    def reset(self):
        self.val = f32(0.0)

    def tick(self):
        out = self.val
        self.val = f32(self.val + self.inc)
        if self.val >= f32(1.0):
            self.val = f32(self.val - f32(1.0))
        return out


def _bits_equal(got, exp):
    return (np.asarray(got, np.float32).view(np.uint32) == np.asarray(exp, np.float32).view(np.uint32)).all()


INNER = pipe("dc(700)", "ramp()")


@pytest.mark.parametrize("n", [1, 3, 8])
def test_kr_nodes_rs_271_278(n):
    inner, count, held, exp = PyRamp(700), 0, f32(0.0), []
    for _ in range(500):
        if count == 0:
            count, held = n, inner.tick()
        count -= 1
        exp.append(held)
    assert _bits_equal(build({"op": "kr()", "net": INNER, "n": n}, ONet).render(500)[:, 0], exp)


def test_reset_nodes_rs_351_360():
    secs = 0.0037
    n = int(round(float(f32(f32(secs) * f32(44100.0)))))               # Reset::new: (s * 44100.).round() as usize
    inner, count, exp = PyRamp(700), 0, []
    for _ in range(2000):
        if count >= n:
            inner.reset(); count = 0
        exp.append(inner.tick()); count += 1
    assert _bits_equal(build({"op": "reset()", "net": INNER, "n": secs}, ONet).render(2000)[:, 0], exp)


def test_trig_reset_nodes_rs_393_400():
    x = inputs(np.random.default_rng(8), 2000, 1)[:, 0]
    x[np.random.default_rng(9).uniform(size=2000) < 0.9] = 0.0
    inner, exp = PyRamp(700), []
    for t in x:
        if t != 0.0:
            inner.reset()
        exp.append(inner.tick())
    assert _bits_equal(build({"op": "trig_reset()", "net": INNER}, ONet).process(x[:, None])[:, 0], exp)


def test_reset_v_nodes_rs_433_442():
    rng = np.random.default_rng(10)
    x = np.repeat(rng.uniform(0.0005, 0.004, 40), 50).astype(np.float32)
    x[300:350] = np.nan; x[700:720] = -1.0                              # `as usize` of NaN / negative: 0 -> reset every sample
    inner, count, exp = PyRamp(700), 0, []
    for d in x:
        with np.errstate(invalid="ignore"):
            lim = as_usize(np.round(f32(d * f32(44100.0))))             # f32::round is half-away-from-zero; never a tie here
        if count >= lim:
            inner.reset(); count = 0
        exp.append(inner.tick()); count += 1
    assert _bits_equal(build({"op": "reset_v()", "net": INNER}, ONet).process(x[:, None])[:, 0], exp)


def test_select_nodes_rs_27_33():
    x = inputs(np.random.default_rng(11), 3000, 1)[:, 0]
    x = np.where(np.isfinite(x), x / f32(10.0), x).astype(np.float32)
    nets = [PyRamp(300), PyRamp(1100), PyRamp(50)]
    exp = []
    for s in x:
        k = as_usize(s)
        exp.append(nets[k].tick() if k < len(nets) else f32(0.0))      # only the selected net advances
    sel = {"op": "select()", "inputs": [pipe("dc(300)", "ramp()"), pipe("dc(1100)", "ramp()"), pipe("dc(50)", "ramp()")]}
    assert _bits_equal(build(sel, ONet).process(x[:, None])[:, 0], exp)


def rround(x):
    """f32::round: half away from zero"""
    x = float(x)
    if math.isnan(x) or math.isinf(x):
        return x
    return math.copysign(math.floor(abs(x) + 0.5), x)      # the sign survives: round(-0.3) = -0.0


def test_seq_nodes_rs_74_106():
    """event list (index, delay, duration) in samples; one live event per index; playing nets are summed in event order; an
    event for an index without a net still counts down"""
    rng = np.random.default_rng(12)
    n = 6000
    trig = (rng.uniform(size=n) < 0.01).astype(np.float32)
    trig[[0, 1, 2]] = 1.0
    idx = rng.choice(np.array([0, 1, 2, 2.7, 5, -1, np.nan], dtype=np.float32), n)
    idx[:3] = [0, 1, 2]
    delay = rng.uniform(0.0, 0.003, n).astype(np.float32)
    dur = rng.uniform(0.0, 0.02, n).astype(np.float32)
    dur[rng.integers(0, n, 200)] = 0.0
    nets = [PyRamp(300), PyRamp(1100), PyRamp(50)]
    sr, events, exp = f32(44100.0), [], []
    for t in range(n):
        if trig[t] != 0.0:
            k = as_usize(idx[t])
            events = [e for e in events if e[0] != k]
            if k < len(nets):
                nets[k].reset()
            events.append([k, as_usize(rround(f32(delay[t] * sr))), as_usize(rround(f32(dur[t] * sr)))])
        events = [e for e in events if e[2] != 0]
        out = f32(0.0)
        for e in events:
            if e[1] == 0:
                if e[0] < len(nets):
                    out = f32(out + nets[e[0]].tick())
                e[2] -= 1
            else:
                e[1] -= 1
        exp.append(out)
    seq = {"op": "seq()", "inputs": [pipe("dc(300)", "ramp()"), pipe("dc(1100)", "ramp()"), pipe("dc(50)", "ramp()")]}
    x = np.stack([trig, idx, delay, dur], axis=1)
    assert _bits_equal(build(seq, ONet).process(x)[:, 0], exp)


# ---------------------------------------------------------------- spectral nodes: the in-tree buffering around the FFTs
@pytest.mark.parametrize("n,start", [(64, 0), (64, 16), (256, 192), (8, 5)])
def test_rfft_node_buffering_nodes_rs_625_642(n, start):
    """count starts at `start`; the transform of the buffer runs when the counter is 0, BEFORE the current sample is stored;
    bin i for i <= n/2, conj(bin n-i) above.  The FFT itself is external (microfft): numpy f64 within tolerance."""
    rng = np.random.default_rng(20 + n + start)
    T = 5 * n + 7
    x = rng.uniform(-1, 1, T).astype(np.float32)
    buf, spec, count, exp = np.zeros(n, np.float32), np.zeros(n // 2 + 1, np.complex128), start, []
    for t in range(T):
        i = count
        count = 0 if count + 1 == n else count + 1
        if i == 0:
            spec = np.fft.rfft(buf.astype(np.float64))
        buf[i] = x[t]
        z = spec[i] if i <= n // 2 else np.conj(spec[n - i])
        exp.append([z.real, z.imag])
    got = build({"op": f"rfft({n},{start})"}, ONet).process(x[:, None])
    assert np.abs(got - np.array(exp)).max() <= 2e-6 * np.sqrt(n) * max(1.0, np.abs(np.array(exp)).max())


@pytest.mark.parametrize("n,start", [(64, 0), (64, 48), (8, 3)])
def test_ifft_node_buffering_nodes_rs_681_692(n, start):
    """full complex n-point inverse of the buffered frame, unity-gain round-trip convention (1/n)"""
    rng = np.random.default_rng(40 + n + start)
    T = 4 * n + 5
    x = rng.uniform(-1, 1, (T, 2)).astype(np.float32)
    buf, out, count, exp = np.zeros(n, np.complex128), np.zeros(n, np.complex128), start, []
    for t in range(T):
        i = count
        count = 0 if count + 1 == n else count + 1
        if i == 0:
            out = np.fft.ifft(buf)
        buf[i] = complex(x[t, 0], x[t, 1])
        exp.append([out[i].real, out[i].imag])
    got = build({"op": f"ifft({n},{start})"}, ONet).process(x)
    assert np.abs(got - np.array(exp)).max() <= 1e-6


# ---------------------------------------------------------------- in-tree closures (functions.rs) with Rust scalar semantics
def as_i32(x):
    x = float(x)
    if math.isnan(x):
        return 0
    return max(-2**31, min(2**31 - 1, int(x))) if math.isfinite(x) else (2**31 - 1 if x > 0 else -2**31)


def wrap_i32(v):
    v &= 0xFFFFFFFF
    return v - 2**32 if v >= 2**31 else v


def rust_min(a, b):
    return b if math.isnan(a) else (a if math.isnan(b) else (a if a < b else b))


def rust_max(a, b):
    return b if math.isnan(a) else (a if math.isnan(b) else (a if a > b else b))


def fmod32(a, b):
    with np.errstate(all="ignore"):
        return f32(np.fmod(f32(a), f32(b)))


def rem_euclid(a, b):
    r = fmod32(a, b)
    with np.errstate(all="ignore"):
        return f32(r + f32(abs(f32(b)))) if r < 0.0 else r


def mirror(x, p0, p1):
    p0, p1 = f32(min(p0, p1)), f32(max(p0, p1))
    r = f32(p1 - p0)
    n = f32(x) if (np.isfinite(x) and abs(float(x)) >= float(np.finfo(np.float32).tiny)) else f32(0.0)    # is_normal()
    if p0 <= n <= p1:
        return n
    distance = f32(min(f32(n - p1), f32(p0 - n)))
    folds = f32(np.floor(f32(distance / r)))
    rest = f32(distance - f32(folds * r))
    if (n > p1 and fmod32(folds, 2.0) == 0.0) or (n < p0 and fmod32(folds, 2.0) != 0.0):
        return f32(p0 + rest)
    return f32(p1 - rest)


def pdhalf_bi(a, b):
    mid = f32(max(-1.0, min(1.0, float(b)))) if not math.isnan(float(b)) else f32(b)
    if a < mid:
        slope = f32(f32(1.0) / f32(mid + f32(1.0))) if mid != -1.0 else f32(0.0)
        return f32(slope * a)
    slope = f32(f32(1.0) / f32(f32(1.0) - mid)) if mid != 1.0 else f32(0.0)
    return f32(f32(slope * f32(a - mid)) + f32(0.5))


def pdhalf_uni(a, b):
    mid = f32(1.0) if b >= 1.0 else (f32(0.0) if b <= -1.0 else f32(f32(b + f32(1.0)) / f32(2.0)))
    if a < mid:
        slope = f32(f32(0.5) / mid) if mid != 0.0 else f32(0.0)
        return f32(slope * a)
    slope = f32(f32(0.5) / f32(f32(1.0) - mid)) if mid != 1.0 else f32(0.0)
    return f32(f32(slope * f32(a - mid)) + f32(0.5))


def shift_amount(x):
    return as_usize(x) & 31          # Wrapping<i32> << usize masks the amount (functions.rs:966-990)


BINARY = {   # op string -> scalar model of the 2-input closure (functions.rs:824-1000, 677-706)
    ">()": lambda a, b: f32(a > b), "<()": lambda a, b: f32(a < b), "==()": lambda a, b: f32(a == b),
    "!=()": lambda a, b: f32(a != b), ">=()": lambda a, b: f32(a >= b), "<=()": lambda a, b: f32(a <= b),
    "min()": lambda a, b: f32(rust_min(float(a), float(b))), "max()": lambda a, b: f32(rust_max(float(a), float(b))),
    "rem()": rem_euclid,
    "bitand()": lambda a, b: f32(as_i32(a) & as_i32(b)), "bitor()": lambda a, b: f32(as_i32(a) | as_i32(b)),
    "bitxor()": lambda a, b: f32(as_i32(a) ^ as_i32(b)),
    "shl()": lambda a, b: f32(wrap_i32(as_i32(a) << shift_amount(b))), "shr()": lambda a, b: f32(as_i32(a) >> shift_amount(b)),
    "pdhalf_bi()": pdhalf_bi, "pdhalf_uni()": pdhalf_uni,
}


@pytest.mark.parametrize("op", sorted(BINARY))
def test_binary_closures_functions_rs(op):
    rng = np.random.default_rng(zlib.crc32(op.encode()))
    x = rng.uniform(-40, 40, (1500, 2)).astype(np.float32)
    if "pdhalf" in op:
        x = (x / f32(30.0)).astype(np.float32)
    x[:10] = [[1, 1], [0, 1], [1, 0], [3e9, 2], [-3e9, 2], [5, 33], [5, -1], [np.nan, 1], [2, np.nan], [np.inf, 3]]
    if op in ("min()", "max()", "==()", "!=()", ">=()", "<=()"):
        x[1:3] = [[0.5, 1], [1, 0.5]]        # keep +0 / -0 ordering out of min / max
    with np.errstate(all="ignore"):
        exp = np.array([BINARY[op](f32(a), f32(b)) for a, b in x], np.float32)
    got = build({"op": op}, ONet).process(x)[:, 0]
    same = (got.view(np.uint32) == exp.view(np.uint32)) | (np.isnan(got) & np.isnan(exp))
    assert same.all(), (op, x[~same][:4], got[~same][:4], exp[~same][:4])


def test_wrap_and_mirror_functions_rs_1149_1181():
    rng = np.random.default_rng(77)
    x = rng.uniform(-300, 300, 3000).astype(np.float32)
    x[:8] = [0.0, -12.0, 36.0, 36.5, -12.5, np.inf, np.nan, 1e-40]
    with np.errstate(all="ignore"):
        p0, r = f32(-12.0), f32(48.0)
        w2 = np.array([f32(fmod32(f32(fmod32(f32(v - p0), r) + r), r) + p0) for v in x], np.float32)
        w1 = np.array([f32(v - f32(f32(7.0) * f32(np.floor(f32(v / f32(7.0)))))) for v in x], np.float32)
        mi = np.array([mirror(v, -1.0, 2.0) for v in (x / f32(10.0)).astype(np.float32)], np.float32)
    for op, inp, exp in (("wrap(-12,36)", x, w2), ("wrap(7)", x, w1), ("mirror(-1,2)", (x / f32(10.0)).astype(np.float32), mi)):
        got = build({"op": op}, ONet).process(inp[:, None])[:, 0]
        same = (got.view(np.uint32) == exp.view(np.uint32)) | (np.isnan(got) & np.isnan(exp))
        assert same.all(), (op, inp[~same][:4], got[~same][:4], exp[~same][:4])


UNARY = {   # Rust std f32 methods reached by functions.rs:1066-1075, 1183-1185
    "abs()": lambda v: f32(abs(v)),
    "signum()": lambda v: v if math.isnan(float(v)) else f32(math.copysign(1.0, float(v))),          # signum(±0) = ±1
    "floor()": lambda v: f32(np.floor(v)), "ceil()": lambda v: f32(np.ceil(v)),
    "fract()": lambda v: f32(v - f32(np.trunc(v))),
    "round()": lambda v: f32(rround(v)),                                                            # half away from zero
    "sqrt()": lambda v: f32(np.sqrt(v)), "recip()": lambda v: f32(f32(1.0) / v),
    "deg()": lambda v: f32(v * f32(57.2957795130823208767981548141051703)),                          # f32::to_degrees constant
    "rad()": lambda v: f32(v * f32(f32(np.pi) / f32(180.0))),
}


@pytest.mark.parametrize("op", sorted(UNARY))
def test_unary_std_closures_functions_rs(op):
    rng = np.random.default_rng(zlib.crc32(op.encode()))
    x = rng.uniform(-50, 50, 2000).astype(np.float32)
    x[:12] = [0.0, -0.0, 0.5, -0.5, 1.5, 2.5, -2.5, 1e-40, 3e9, np.inf, -np.inf, np.nan]
    with np.errstate(all="ignore"):
        exp = np.array([UNARY[op](f32(v)) for v in x], np.float32)
    got = build({"op": op}, ONet).process(x[:, None])[:, 0]
    same = (got.view(np.uint32) == exp.view(np.uint32)) | (np.isnan(got) & np.isnan(exp))
    assert same.all(), (op, x[~same][:4], got[~same][:4], exp[~same][:4])
