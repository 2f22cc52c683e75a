"""Parity of the hand-fused kernels (K2 k_noise_svf_scan, K1f k_polysynth) against the CPU oracle and against the
generic interpreter, including BASELINE-size renders on a subset of voices (the oracle is scalar CPU code)."""
import numpy as np
import pytest

import quartz_b200 as qb
from quartz_b200 import Bank, Net, workloads
from tests.graphs import build
from tests.oracle_ffi import ONet, render_bank
from tests.util import assert_parity

pytestmark = pytest.mark.gpu


def oracle_voices(wl, voices, T, group=1, threads=8):
    onets = [build(wl.voice_expr(v), ONet).set_salt(int(wl.salts[v])) for v in voices]
    return render_bank(onets, T, group=group, threads=threads)


def make_bank(wl, V=None, path=qb.PATH_AUTO):
    V = V or wl.V
    tmpl = build(wl.expr, Net)
    return Bank(tmpl, V, raw=wl.raw[:V], salts=wl.salts[:V]).set_path(path)


@pytest.mark.parametrize("V,T", [(64, 4096), (7, 1000), (300, 2047), (4, 70000), (1, 257)])
def test_noise_svf_fused_small(V, T):
    wl = workloads.c2_lowpass_bank(V=V, T=T)
    bank = make_bank(wl)
    assert bank.kernel() == "k_noise_svf_scan"
    got = bank.render(T)[:, 0, :]
    ref = oracle_voices(wl, range(V), T)
    assert_parity(got, ref, "float", f"c2 V={V} T={T}")
    # the interpreter on the same bank agrees too, and the persisted state continues identically
    interp = make_bank(wl, path=qb.PATH_INTERP)
    assert_parity(interp.render(T)[:, 0, :], ref, "float", "interp")
    more_f, more_i = bank.render(300)[:, 0, :], interp.render(300)[:, 0, :]
    assert_parity(more_f, more_i, "float", "continuation after fused vs after interp")


@pytest.mark.parametrize("mode", ["highpass", "bandpass", "notch", "peak", "allpass"])
def test_noise_svf_fused_other_modes(mode):
    V, T = 33, 3000
    rng = np.random.default_rng(1)
    hz = rng.uniform(100, 8000, V).astype(np.float32)
    q = rng.uniform(0.5, 6, V).astype(np.float32)
    tmpl = build({"op": ">>", "n": 0, "inputs": [{"op": "white()"}, {"op": f"{mode}(1000,1)"}]}, Net)
    bank = Bank(tmpl, V, raw=np.stack([hz, q], 1), salts=np.arange(1, V + 1, dtype=np.uint64))
    assert bank.kernel() == "k_noise_svf_scan"
    onets = [build({"op": ">>", "n": 0, "inputs": [{"op": "white()"}, {"op": f"{mode}({float(hz[v])!r},{float(q[v])!r})"}]}, ONet)
             .set_salt(v + 1) for v in range(V)]
    assert_parity(bank.render(T)[:, 0, :], render_bank(onets, T), "float", mode)


def test_noise_svf_full_length_subset():
    """BASELINE configs[1] length (60 s at 48 kHz) on the 16 lowest- and highest-Q voices of the first 512."""
    wl = workloads.c2_lowpass_bank(V=512)
    T = wl.T
    bank = make_bank(wl)
    d = bank.render(T)[:, 0, :]
    order = np.argsort(wl.raw[:, 1] / wl.raw[:, 0])   # q/hz: slowest-decaying filters last
    pick = list(order[:8]) + list(order[-8:])
    ref = oracle_voices(wl, pick, T)
    assert_parity(d[pick], ref, "float", "c2 full length")


def test_noise_svf_chunked_equals_one_shot():
    wl = workloads.c2_lowpass_bank(V=40, T=10000)
    a, b = make_bank(wl), make_bank(wl)
    one = a.render(10000)[:, 0, :]
    parts = np.concatenate([b.render(n)[:, 0, :] for n in (3000, 1, 255, 4096, 2648)], axis=1)
    assert_parity(parts, one, "float", "chunked")


@pytest.mark.parametrize("V,T,G", [(64, 4096, 32), (96, 1000, 1), (33, 2500, 1), (128, 3001, 8)])
def test_polysynth_fused_small(V, T, G):
    wl = workloads.c3_polysynth(V=V, T=T, G=G)
    bank = make_bank(wl)
    assert bank.kernel() == "k_polysynth"
    got = bank.render(T, group=G)[:, 0, :]
    ref = oracle_voices(wl, range(V), T, group=G)
    assert_parity(got, ref, "float", f"c3 V={V} T={T} G={G}")
    interp = make_bank(wl, path=qb.PATH_INTERP)
    assert_parity(interp.render(T, group=G)[:, 0, :], ref, "float", "interp")
    assert_parity(bank.render(500, group=G)[:, 0, :], interp.render(500, group=G)[:, 0, :], "float", "continuation")


def test_polysynth_full_length_subset():
    """BASELINE configs[2] length (10 s at 48 kHz), 4 groups of 32 voices."""
    wl = workloads.c3_polysynth(V=128)
    bank = make_bank(wl)
    got = bank.render(wl.T, group=32)[:, 0, :]
    ref = oracle_voices(wl, range(128), wl.T, group=32)
    assert_parity(got, ref, "float", "c3 full length")


def test_hello_440_full_length():
    """BASELINE configs[0]: sine(440), 10 s at 48 kHz through the render path."""
    wl = workloads.c1_hello()
    net = build(wl.expr, Net)
    got = net.render(wl.T)
    ref = build(wl.expr, ONet).render(wl.T)
    assert_parity(got, ref, "float", "hello_440")
    # spectral purity as a size-independent property: the render is a 440 Hz sinusoid of unit amplitude
    x = got[:, 0].astype(np.float64)
    n = np.arange(len(x))
    amp = 2 * abs(np.sum(x * np.exp(-2j * np.pi * 440.0 * n / 48000.0))) / len(x)
    assert abs(amp - 1.0) < 1e-3


@pytest.mark.parametrize("mode", ["lowpass", "bandpass", "highshelf"])
def test_noise_svf_fused_extreme_parameters(mode):
    """poles almost on the unit circle (5 Hz, Q = 200), cutoffs next to Nyquist, over-damped filters: the scan matrices
    A^(32 * 2^i) and the chained segment states must stay within the audio tolerance of the sequential oracle"""
    hz = np.array([5, 5, 20, 20, 80, 300, 1000, 6000, 20000, 23000, 23900, 50], dtype=np.float32)
    q = np.array([0.1, 200, 50, 0.5, 200, 100, 0.05, 30, 2, 0.7, 5, 1000], dtype=np.float32)
    V, T = len(hz), 300000
    extra = ",2" if mode == "highshelf" else ""
    tmpl = build({"op": "sr()", "n": 48000.0, "net": {"op": ">>", "n": 0, "inputs": [{"op": "white()"}, {"op": f"{mode}(1000,1{extra})"}]}}, Net)
    raw = np.stack([hz, q] + ([np.full(V, 2.0, np.float32)] if extra else []), 1)
    salts = np.arange(1, V + 1, dtype=np.uint64)
    bank = Bank(tmpl, V, raw=raw, salts=salts)
    assert bank.kernel() == "k_noise_svf_scan"
    got = bank.render(T)[:, 0, :]
    onets = [build({"op": "sr()", "n": 48000.0, "net": {"op": ">>", "n": 0, "inputs": [
        {"op": "white()"}, {"op": f"{mode}({float(hz[v])!r},{float(q[v])!r}{extra})"}]}}, ONet).set_salt(v + 1) for v in range(V)]
    ref = render_bank(onets, T, threads=8)
    for v in range(V):   # per voice: the resonant gains differ by orders of magnitude
        assert_parity(got[v:v + 1], ref[v:v + 1], "float", f"{mode} hz={hz[v]} q={q[v]}")


LTI_OPS = {
    # op name -> (template op string, per-voice raw parameter generator, op string for voice v)
    "butterpass": ("butterpass(1000)", lambda rng, V: np.exp(rng.uniform(np.log(500), np.log(15000), (V, 1))), lambda r: f"butterpass({r[0]!r})"),
    "resonator": ("resonator(1000,50)", lambda rng, V: np.stack([np.exp(rng.uniform(np.log(600), np.log(9000), V)), rng.uniform(60, 400, V)], 1),
                  lambda r: f"resonator({r[0]!r},{r[1]!r})"),
    "biquad": ("biquad(-1.2,0.5,0.2,0.3,0.2)", lambda rng, V: np.stack([rng.uniform(-1.7, 1.7, V), rng.uniform(0.1, 0.95, V) * 0 + 0.8, rng.uniform(0.1, 1, V),
                                                                         rng.uniform(-1, 1, V), rng.uniform(-1, 1, V)], 1),
               lambda r: f"biquad({r[0]!r},{r[1]!r},{r[2]!r},{r[3]!r},{r[4]!r})"),
    "lowpole": ("lowpole(500)", lambda rng, V: np.exp(rng.uniform(np.log(5), np.log(15000), (V, 1))), lambda r: f"lowpole({r[0]!r})"),
    "highpole": ("highpole(500)", lambda rng, V: np.exp(rng.uniform(np.log(5), np.log(15000), (V, 1))), lambda r: f"highpole({r[0]!r})"),
    "dcblock": ("dcblock(10)", lambda rng, V: rng.uniform(1, 200, (V, 1)), lambda r: f"dcblock({r[0]!r})"),
    "allpole": ("allpole(0.4)", lambda rng, V: rng.uniform(0.05, 3.0, (V, 1)), lambda r: f"allpole({r[0]!r})"),
}


@pytest.mark.parametrize("V,T", [(37, 5000), (3, 200000)])
@pytest.mark.parametrize("op", sorted(LTI_OPS))
def test_noise_lti_fused_families(op, V, T):
    """K2 on the other linear recurrences (one-pole family, direct-form biquads): one wave (V = 37) and chained time
    segments (V = 3, 200,000 samples), then a continuation call that has to pick up filter state AND input history"""
    tmpl_op, gen, fmt = LTI_OPS[op]
    rng = np.random.default_rng(sum(map(ord, op)) + V)
    raw = gen(rng, V).astype(np.float32)
    if op == "biquad":   # keep |a1| < 1 + a2 (stable)
        raw[:, 0] = np.clip(raw[:, 0], -1.7, 1.7)
    sr = lambda g: {"op": "sr()", "n": 48000.0, "net": g}
    tmpl = build(sr({"op": ">>", "n": 0, "inputs": [{"op": "white()"}, {"op": tmpl_op}]}), Net)
    salts = np.arange(1, V + 1, dtype=np.uint64)
    bank = Bank(tmpl, V, raw=raw, salts=salts)
    assert bank.kernel() == "k_noise_svf_scan"
    got = np.concatenate([bank.render(T)[:, 0, :], bank.render(777)[:, 0, :]], axis=1)
    onets = [build(sr({"op": ">>", "n": 0, "inputs": [{"op": "white()"}, {"op": fmt([float(x) for x in raw[v]])}]}), ONet).set_salt(v + 1)
             for v in range(V)]
    ref = render_bank(onets, T + 777, threads=8)
    for v in range(V):
        assert_parity(got[v:v + 1], ref[v:v + 1], "float", f"{op} raw={raw[v]}")


def test_ill_conditioned_biquads_stay_on_the_op_order_exact_path():
    """a direct-form biquad with poles next to z = 1 amplifies ANY rounding difference by its round-off noise gain (here
    ~4e2 in amplitude): re-associating it cannot stay inside the parity tolerance, so the bank is refused by K2 and runs on
    an interpreter that performs the reference's operations in the reference's order"""
    V, T = 9, 20000
    sr = lambda g: {"op": "sr()", "n": 48000.0, "net": g}
    hz = np.array([97.5, 800, 50, 2000, 30, 5000, 120, 60, 1000], dtype=np.float32)
    bw = np.array([169.4, 100, 20, 300, 10, 50, 40, 300, 80], dtype=np.float32)
    tmpl = build(sr({"op": ">>", "n": 0, "inputs": [{"op": "white()"}, {"op": "resonator(1000,50)"}]}), Net)
    salts = np.arange(1, V + 1, dtype=np.uint64)
    bank = Bank(tmpl, V, raw=np.stack([hz, bw], 1), salts=salts)
    assert bank.kernel() != "k_noise_svf_scan", bank.kernel()
    onets = [build(sr({"op": ">>", "n": 0, "inputs": [{"op": "white()"}, {"op": f"resonator({float(hz[v])!r},{float(bw[v])!r})"}]}), ONet)
             .set_salt(v + 1) for v in range(V)]
    ref = render_bank(onets, T, threads=4)
    got = bank.render(T)[:, 0, :]
    for v in range(V):
        assert_parity(got[v:v + 1], ref[v:v + 1], "float", f"resonator hz={hz[v]} bw={bw[v]} [{bank.kernel()}]")


@pytest.mark.parametrize("env", ["ar", "xd", "xD"])
@pytest.mark.parametrize("osc", ["sine", "saw", "square", "triangle", "soft_saw"])
def test_polysynth_fused_oscillator_and_envelope_variants(osc, env):
    """K1f serves <sine | band-limited wavetable oscillator>(f) >> <fixed SVF> * <xd | xD | ar> with per-voice constants:
    parity with the oracle, with the interpreter on the same bank, group mix, and a continuation call"""
    V, T, G = 64, 6000, 32
    rng = np.random.default_rng(sum(map(ord, osc + env)))
    f = np.exp(rng.uniform(np.log(40), np.log(5000), V)).astype(np.float32)
    hz = np.exp(rng.uniform(np.log(200), np.log(8000), V)).astype(np.float32)
    q = rng.uniform(0.5, 4, V).astype(np.float32)
    if env == "ar":
        ep = np.stack([rng.uniform(0.002, 0.05, V), np.ones(V), rng.uniform(0.02, 0.3, V), np.full(V, 4.0)], 1).astype(np.float32)
    elif env == "xd":
        ep = rng.uniform(2, 60, (V, 1)).astype(np.float32)
    else:
        ep = np.stack([rng.uniform(0.02, 0.4, V), rng.uniform(0.5, 4, V)], 1).astype(np.float32)
    sr = lambda g: {"op": "sr()", "n": 48000.0, "net": g}

    def voice(fv, hv, qv, e):
        osc_g = {"op": ">>", "n": 0.0, "inputs": [{"op": f"{osc}({fv!r})"}, {"op": f"highpass({hv!r},{qv!r})"}]}
        return sr({"op": "*", "n": 0.0, "inputs": [osc_g, {"op": f"{env}(" + ",".join(repr(float(x)) for x in e) + ")"}]})

    tmpl = build(voice(440.0, 1000.0, 1.0, ep[0]), Net)
    raw = np.concatenate([f[:, None], hz[:, None], q[:, None], ep], axis=1)
    salts = np.arange(1, V + 1, dtype=np.uint64)
    bank = Bank(tmpl, V, raw=raw, salts=salts)
    assert bank.kernel() == "k_polysynth"
    got = np.concatenate([bank.render(T, group=G)[:, 0, :], bank.render(999, group=G)[:, 0, :]], axis=1)
    onets = [build(voice(float(f[v]), float(hz[v]), float(q[v]), ep[v]), ONet).set_salt(v + 1) for v in range(V)]
    ref = render_bank(onets, T + 999, group=G, threads=8)
    assert_parity(got, ref, "float", f"{osc} * {env}")
    interp = Bank(tmpl, V, raw=raw, salts=salts).set_path(qb.PATH_INTERP)
    assert_parity(interp.render(T + 999, group=G)[:, 0, :], ref, "float", "interp")


def test_fused_kernels_hand_over_to_the_interpreters_when_they_do_not_serve_a_call():
    """K2 has no group mix; K1f needs an lfo segment of at least one 32-sample window (sample rate >= ~23 kHz): both banks
    still render, through the interpreters, and match the oracle"""
    V, T = 64, 3000
    wl = workloads.c2_lowpass_bank(V=V, T=T)
    bank = make_bank(wl)
    assert bank.kernel() == "k_noise_svf_scan"
    assert_parity(bank.render(T, group=2)[:, 0, :], oracle_voices(wl, range(V), T, group=2), "float", "K2 bank, group mix")
    # poly-synth voices at 8 kHz: the envelope's control points are 9-15 samples apart
    f = np.linspace(60, 900, V).astype(np.float32)

    def voice(fv):
        osc = {"op": ">>", "n": 0.0, "inputs": [{"op": f"sine({fv!r})"}, {"op": "lowpass(700.0,1.5)"}]}
        return {"op": "sr()", "n": 8000.0, "net": {"op": "*", "n": 0.0, "inputs": [osc, {"op": "ar(0.01,1.0,0.2,4.0)"}]}}

    tmpl = build(voice(220.0), Net)
    raw = np.stack([f, np.full(V, 700, np.float32), np.full(V, 1.5, np.float32), np.full(V, 0.01, np.float32), np.ones(V, np.float32),
                    np.full(V, 0.2, np.float32), np.full(V, 4, np.float32)], 1)
    salts = np.arange(1, V + 1, dtype=np.uint64)
    b2 = Bank(tmpl, V, raw=raw, salts=salts)
    assert b2.kernel() == "k_polysynth"
    ref = render_bank([build(voice(float(x)), ONet).set_salt(int(s)) for x, s in zip(f, salts)], T, group=32, threads=4)
    assert_parity(b2.render(T, group=32)[:, 0, :], ref, "float", "poly-synth bank at 8 kHz")


def test_parameter_update_that_ruins_the_conditioning_leaves_the_scan_kernel():
    """var() semantics (process.rs:1382-1385): a parameter rewritten while the bank runs.  Sweeping a resonator next to
    z = 1 makes its direct form ill-conditioned; from that update on the bank must evaluate it in the reference's operation
    order (same rule as at construction), with its filter state carried over"""
    V, T = 4096, 4096
    sr = lambda g: {"op": "sr()", "n": 48000.0, "net": g}
    tmpl = build(sr({"op": ">>", "n": 0, "inputs": [{"op": "white()"}, {"op": "resonator(1000,50)"}]}), Net)
    salts = np.arange(1, V + 1, dtype=np.uint64)
    raw = np.tile(np.array([[1000.0, 50.0]], dtype=np.float32), (V, 1))
    bank = Bank(tmpl, V, raw=raw, salts=salts)
    twin = Bank(tmpl, V, raw=raw, salts=salts).set_path(qb.PATH_INTERP)      # operation-order path from the start
    assert bank.kernel() == "k_noise_svf_scan"
    a0, b0 = bank.render(T)[:, 0, :], twin.render(T)[:, 0, :]
    assert_parity(a0[:64], b0[:64], "float", "before the update")
    for bk in (bank, twin):
        bk.set_raw(0, 30.0)
        bk.set_raw(1, 10.0)
    assert bank.kernel() != "k_noise_svf_scan", bank.kernel()
    a1, b1 = bank.render(T)[:, 0, :], twin.render(T)[:, 0, :]
    assert np.isfinite(a1).all() and np.abs(a1).max() > 1e-3
    # The two banks reach the update with filter states that differ by f32 rounding (scan vs operation order), and the new
    # filter amplifies exactly such differences (its round-off gain is what took it off the scan path), so the segments
    # agree to ~6e-4 of full scale, not to the 1e-4 parity bound; a lost or re-initialised state would differ by O(scale).
    scale = float(np.abs(b1[:64]).max())
    assert float(np.abs(a1[:64] - b1[:64]).max()) <= 1e-2 * scale
