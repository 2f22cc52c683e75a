"""Randomised graph parity: seeded random feed-forward graphs built from the supported op vocabulary are rendered by
every kernel family (auto selection, block-mode lane interpreter, sample-by-sample lane interpreter, time-vector kernel
when the tape is capable) for a small salted bank and compared with the CPU oracle.  Two families:
  * exact: only operations that round identically on both sides (IEEE add/mul/div/floor/compare, table lookups,
    in-tree nodes) -> bit-exact, including through delay lines, sample-and-hold and shift registers;
  * float: smooth DSP chains (oscillators, noise, LTI filters, waveshapers, delay taps) -> f32 audio tolerance.
The generator exercises what the fixed case list cannot enumerate: temporary reuse in lowering, in-place operands in
the block / time-vector kernels, diamonds (split -> parallel branches -> join), sums and products of sub-graphs."""
import os

import numpy as np
import pytest

import quartz_b200 as qb
from quartz_b200 import Bank, Net
from tests.graphs import L, add, branch, build, bus, mul, pipe, stack
from tests.oracle_ffi import ONet
from tests.util import assert_parity

pytestmark = pytest.mark.gpu

MINOR = [0.0, 2.0, 3.0, 5.0, 7.0, 8.0, 10.0, 12.0]
N_SEEDS = int(os.environ.get("QG_FUZZ_SEEDS", "40"))   # soak runs: QG_FUZZ_SEEDS=400 [QG_FUZZ_OFFSET=1000 for fresh graphs]
OFFSET = int(os.environ.get("QG_FUZZ_OFFSET", "0"))


def _c(rng, lo, hi, nd=3):
    return round(float(rng.uniform(lo, hi)), nd)


# ---------------------------------------------------------------- exact family
def exact_source(rng):
    k = rng.integers(0, 5)
    if k == 0:
        return pipe(f"dc({_c(rng, 20, 900)})", "ramp()")
    if k == 1:
        return L("white()")
    if k == 2:
        return {"op": "wave()", "arr": [round(float(x), 3) for x in rng.uniform(-2, 2, int(rng.integers(3, 40)))]}
    if k == 3:
        return pipe(f"dc({_c(rng, 1, 50)})", "ramp()", f"mul({_c(rng, 2, 24)})", {"op": "quantize()", "arr": MINOR})
    return pipe("impulse()", f"add({_c(rng, -1, 1)})")


def exact_stage(rng):
    k = rng.integers(0, 20)
    if k == 0:
        return L(f"add({_c(rng, -3, 3)})")
    if k == 1:
        return L(f"mul({_c(rng, -3, 3)})")
    if k == 2:
        return L("abs()")
    if k == 3:
        return L("floor()")
    if k == 4:
        return L(f">({_c(rng, -0.5, 0.8)})")
    if k == 5:
        return L(f"min({_c(rng, -1, 1)})")
    if k == 6:
        return L(f"clip({_c(rng, -1, 0)},{_c(rng, 0.1, 1)})")
    if k == 7:
        return L(f"wrap({_c(rng, -2, 0)},{_c(rng, 0.5, 3)})")
    if k == 8:
        return L("tick()")
    if k == 9:
        return L(f"delay({_c(rng, 0.0001, 0.01, 5)})")
    if k == 10:   # sample-and-hold clocked by an edge detector
        clock = pipe(f"dc({_c(rng, 30, 700)})", "ramp()", "<(0.5)", "rise()")
        return pipe(stack("pass()", clock), "snh()")
    if k == 11:   # shift register, a few taps summed
        clock = pipe(f"dc({_c(rng, 30, 700)})", "ramp()", "<(0.5)", "rise()")
        return pipe(stack("pass()", clock), "shift_reg()", "join(8)")
    if k == 12:   # diamond
        return pipe("split(2)", stack(exact_stage(rng), exact_stage(rng)), "join(2)")
    if k == 13:
        return L("squared()")
    if k == 14:
        return L("mirror(-1,1)")
    if k == 16:   # variable sample delay (nodes.rs:707-738), index from a slow ramp
        return pipe(stack("pass()", pipe(f"dc({_c(rng, 1, 40)})", "ramp()", f"mul({int(rng.integers(2, 60))})")), f"samp_delay({int(rng.integers(8, 64))})")
    if k == 17:   # table lookup (nodes.rs:127-150) of a scaled, folded index
        arr = [round(float(x), 3) for x in rng.uniform(-1, 1, int(rng.integers(4, 24)))]
        return pipe("abs()", f"mul({_c(rng, 1, 30)})", {"op": "get()", "arr": arr})
    if k == 18:
        return pipe(f"mul({_c(rng, 2, 200)})", f"bitand({int(rng.integers(1, 255))})")
    if k == 19:   # equal-power pan to stereo and back (FunDSP pan / join)
        return pipe(f"mul({_c(rng, 0.1, 2)})", "split(2)", stack("pass()", "tick()"), "reverse(2)", "join(2)")
    return L("round()")


def exact_graph(rng):
    g = exact_source(rng)
    for _ in range(int(rng.integers(1, 6))):
        g = pipe(g, exact_stage(rng))
    if rng.uniform() < 0.4:
        other = exact_source(rng)
        for _ in range(int(rng.integers(0, 3))):
            other = pipe(other, exact_stage(rng))
        g = (add if rng.uniform() < 0.5 else mul)(g, other)
    return g


# ---------------------------------------------------------------- float family
def float_source(rng):
    k = rng.integers(0, 9)
    if k == 6:
        return L(f"{rng.choice(['square', 'triangle', 'soft_saw'])}({_c(rng, 30, 2500)})")
    if k == 7:   # oscillator swept by an exponential map of a slow ramp
        return pipe(f"dc({_c(rng, 0.2, 3)})", "ramp()", f"xerp({_c(rng, 40, 200)},{_c(rng, 400, 6000)})", str(rng.choice(["sine()", "saw()"])))
    if k == 8:
        return L("brown()")
    if k == 0:
        return L("white()")
    if k == 1:
        return L(f"sine({_c(rng, 30, 3000)})")
    if k == 2:
        return L(f"saw({_c(rng, 30, 1500)})")
    if k == 3:
        return L("pink()")
    if k == 4:   # FM pair
        return pipe(pipe(f"sine({_c(rng, 0.5, 8)})", f"mul({_c(rng, 5, 80)})", f"add({_c(rng, 200, 900)})"), "sine()")
    return pipe(f"dc({_c(rng, 40, 800)})", "ramp()", "mul(TAU)", "sin()")


def float_stage(rng):
    k = rng.integers(0, 23)
    f, q = _c(rng, 80, 6000), _c(rng, 0.5, 5)
    if k == 0:
        return L(f"lowpass({f},{q})")
    if k == 1:
        return L(f"highpass({f},{q})")
    if k == 2:
        return L(f"bell({f},{q},{_c(rng, 0.5, 2)})")
    if k == 3:
        return L(f"butterpass({f})")
    if k == 4:
        return L(f"resonator({f},{_c(rng, 20, 300)})")
    if k == 5:
        return L(f"lowpole({f})")
    if k == 6:
        return L(f"highpole({_c(rng, 5, 400)})")
    if k == 7:
        return L("dcblock()")
    if k == 8:
        return L(f"allpole({_c(rng, 0.1, 0.9)})")
    if k == 9:
        return L("tanh()")
    if k == 10:
        return L(f"mul({_c(rng, 0.2, 1.5)})")
    if k == 11:
        return L(f"delay({_c(rng, 0.0002, 0.02, 5)})")
    if k == 12:   # modulated delay tap
        return pipe(stack("pass()", pipe(f"sine({_c(rng, 0.2, 5)})", "mul(0.002)", "add(0.004)")), "tap(0.001,0.01)")
    if k == 13:   # parallel filters summed
        return bus(float_stage(rng), float_stage(rng))
    if k == 14:   # diamond through branch
        return pipe(branch(float_stage(rng), "pass()"), "join(2)")
    if k == 15:
        return L("fir(0.2,0.5,0.2,0.1)")
    if k == 17:   # spectral round trip (nodes.rs:601-700): real part of ifft(rfft(x)), 2N samples late
        n, st = int(rng.choice([16, 64, 256])), 0
        st = int(rng.integers(0, n))
        return pipe(f"rfft({n},{st})", f"ifft({n},{st})", "chan(1,0)")
    if k == 18:   # stereo detour: pan, rotate, back to mono
        return pipe(f"pan({_c(rng, -1, 1)})", f"rotate({_c(rng, -3, 3)},{_c(rng, 0.5, 1.2)})", "join(2)")
    if k == 19:   # variable-cutoff lowpass driven by an lfo
        return pipe(stack("pass()", pipe(f"sine({_c(rng, 0.3, 9)})", f"mul({_c(rng, 50, 400)})", f"add({_c(rng, 600, 3000)})")), f"lowpass({q})")
    if k == 20:
        return L("pinkpass()")
    if k == 21:
        return pipe(f"mul({_c(rng, 0.5, 3)})", "atan()")
    return L("softsign()")


def float_graph(rng):
    g = float_source(rng)
    for _ in range(int(rng.integers(1, 6))):
        g = pipe(g, float_stage(rng))
    if rng.uniform() < 0.35:
        g = add(g, pipe(float_source(rng), float_stage(rng)))
    if rng.uniform() < 0.3:
        g = mul(g, L(f"xd({_c(rng, 1, 20)})"))
    return g


PATHS = [("auto", qb.PATH_AUTO), ("lane_block", qb.PATH_INTERP), ("lane_sample", qb.PATH_INTERP_SAMPLE), ("time_vector", qb.PATH_TV)]
if os.environ.get("QG_FUZZ_SPEC") == "1":   # also through K1s, the tape-specialised kernel (opt-in: ~0.5 s of NVRTC per graph)
    PATHS.append(("specialised", qb.PATH_SPECIALISED))


def _bank_on(net, V, salts, path):
    bank = Bank(net, V, salts=salts)
    try:
        return bank.set_path(path)
    except qb.QuartzGpuError as e:
        if path == qb.PATH_SPECIALISED and "cannot be specialised" in str(e):
            return None      # nested-net control flow / spectral nodes stay on the interpreters
        raise


def _check(expr, tol, n, seed, n_out=1):
    V = 3
    salts = np.arange(1, V + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15 + seed)
    net = build(expr, Net)
    assert net.inputs() == 0 and net.outputs() == n_out, (net.inputs(), net.outputs(), expr)
    ref = np.stack([build(expr, ONet).set_salt(int(s)).render(n).T for s in salts])     # [V, outputs, n]
    seen = set()
    for pname, path in PATHS:
        bank = _bank_on(net, V, salts, path)
        if bank is None:
            continue
        if (pname == "time_vector" and bank.kernel() != "k_interp_tv") or bank.kernel() in seen and pname != "auto":
            continue
        seen.add(bank.kernel())
        # two calls of uneven length: state, rings and counters carry over
        got = np.concatenate([bank.render(n // 3 + 1), bank.render(n - n // 3 - 1)], axis=2)
        assert_parity(got, ref, tol, f"seed {seed} [{pname}: {bank.kernel()}] {expr}", relative=True)
        if n_out > 1:   # the frame-major interleave of audio.rs:113-117
            bank.reset()
            fm = bank.render(n, layout=qb.LAYOUT_FRAME_MAJOR)                              # [n, V, outputs]
            assert_parity(fm.transpose(1, 2, 0), ref, tol, f"seed {seed} frame-major [{pname}: {bank.kernel()}]", relative=True)


@pytest.mark.parametrize("seed", range(OFFSET, OFFSET + N_SEEDS))
def test_random_exact_graphs(seed):
    rng = np.random.default_rng(1000 + seed)
    _check(exact_graph(rng), "exact", 2500, seed)


@pytest.mark.parametrize("seed", range(OFFSET, OFFSET + N_SEEDS))
def test_random_float_graphs(seed):
    rng = np.random.default_rng(5000 + seed)
    _check(float_graph(rng), "float", 3000, seed)


# ---------------------------------------------------------------- control-flow family (nested nets: nodes.rs:10-453)
def _clock(rng):
    return pipe(f"dc({_c(rng, 20, 400)})", "ramp()", "<(0.5)", "rise()")


def _gen0(rng, depth):
    """a 0-input, 1-output exact sub-graph, possibly wrapped in nested-net control nodes"""
    if depth <= 0 or rng.uniform() < 0.35:
        g = exact_source(rng)
        for _ in range(int(rng.integers(0, 3))):
            g = pipe(g, exact_stage(rng))
        return g
    k = rng.integers(0, 7)
    if k == 0:     # kr(): tick the inner net every n-th sample, hold in between
        return {"op": "kr()", "net": _gen0(rng, depth - 1), "n": float(rng.integers(2, 9))}
    if k == 1:     # s(): same, inner net runs at sr / n
        return {"op": "s()", "net": _gen0(rng, depth - 1), "n": float(rng.integers(2, 6))}
    if k == 2:     # reset(): periodic reset of the inner net
        return {"op": "reset()", "net": _gen0(rng, depth - 1), "n": _c(rng, 0.001, 0.02, 4)}
    if k == 3:     # trig_reset(): reset on a trigger input
        return pipe(_clock(rng), {"op": "trig_reset()", "net": _gen0(rng, depth - 1)})
    if k == 4:     # reset_v(): reset period read from an input
        return pipe(f"dc({_c(rng, 0.002, 0.02, 4)})", {"op": "reset_v()", "net": _gen0(rng, depth - 1)})
    if k == 5:     # select(): index input picks one of the nets
        kids = [_gen0(rng, depth - 1) for _ in range(int(rng.integers(2, 4)))]
        return pipe(f"dc({_c(rng, 5, 60)})", "ramp()", f"mul({len(kids)})", {"op": "select()", "inputs": kids})
    kids = [_gen0(rng, depth - 1) for _ in range(int(rng.integers(2, 4)))]   # seq(): triggered, possibly overlapping events
    ctl = stack(_clock(rng), pipe(f"dc({_c(rng, 3, 40)})", "ramp()", f"mul({len(kids)})"), f"dc({_c(rng, 0, 0.004, 4)})",
                f"dc({_c(rng, 0.002, 0.03, 4)})")
    return pipe(ctl, {"op": "seq()", "inputs": kids})


def control_graph(rng):
    g = _gen0(rng, 2)
    if rng.uniform() < 0.5:     # a 1-in/1-out feedback loop around exact stages (1-sample or longer delay line)
        inner = pipe(f"mul({_c(rng, -0.9, 0.9)})", exact_stage(rng))
        delay = None if rng.uniform() < 0.5 else _c(rng, 0.0001, 0.004, 5)
        g = pipe(g, {"op": "feedback()", "net": inner, "delay": delay})
    if rng.uniform() < 0.4:
        g = add(g, _gen0(rng, 1))
    return g


@pytest.mark.parametrize("seed", range(OFFSET, OFFSET + N_SEEDS))
def test_random_control_flow_graphs(seed):
    """kr / s / reset / trig_reset / reset_v / select / seq / feedback nested up to two levels around exact sub-graphs:
    the SIMT-stack interpreter must reproduce the oracle's nested-net semantics bit for bit"""
    rng = np.random.default_rng(9000 + seed)
    _check(control_graph(rng), "exact", 3000, seed)


# ---------------------------------------------------------------- block path with inputs (AudioUnit::process, audio.rs:85-118)
@pytest.mark.parametrize("seed", range(OFFSET, OFFSET + N_SEEDS))
def test_random_process_chains(seed):
    """1-input chains driven with external buffers through qg_bank_process, voice-major and frame-major, two calls"""
    rng = np.random.default_rng(13000 + seed)
    exact = bool(seed % 2)
    stage = exact_stage if exact else float_stage
    g = stage(rng)
    for _ in range(int(rng.integers(0, 4))):
        g = pipe(g, stage(rng))
    net = build(g, Net)
    assert (net.inputs(), net.outputs()) == (1, 1), g
    V, n = 3, 1800
    x = rng.uniform(-1, 1, (V, 1, n)).astype(np.float32)
    onets = [build(g, ONet).set_salt(v + 1) for v in range(V)]
    ref = np.stack([o.process(x[v, 0][:, None])[:, 0] for v, o in enumerate(onets)])
    tol = "exact" if exact else "float"
    salts = np.arange(1, V + 1, dtype=np.uint64)
    seen = set()
    for pname, path in PATHS:
        bank = _bank_on(net, V, salts, path)
        if bank is None:
            continue
        if (pname == "time_vector" and bank.kernel() != "k_interp_tv") or bank.kernel() in seen and pname != "auto":
            continue
        seen.add(bank.kernel())
        a = bank.process(x[:, :, :700], 700)[:, 0, :]
        b = bank.process(np.ascontiguousarray(x[:, :, 700:]), n - 700)[:, 0, :]
        assert_parity(np.concatenate([a, b], axis=1), ref, tol, f"seed {seed} voice-major [{pname}: {bank.kernel()}] {g}", relative=True)
        bank.reset()
        fm = bank.process(np.ascontiguousarray(x.transpose(2, 0, 1)), n, layout=qb.LAYOUT_FRAME_MAJOR)[:, :, 0].T
        assert_parity(fm, ref, tol, f"seed {seed} frame-major [{pname}: {bank.kernel()}] {g}", relative=True)


# ---------------------------------------------------------------- multi-output graphs (stereo patches end in out())
@pytest.mark.parametrize("seed", range(OFFSET, OFFSET + N_SEEDS))
def test_random_stereo_graphs(seed):
    """two float chains stacked, cross-fed through pan / rotate / reverse: 2 outputs, voice-major and frame-major"""
    rng = np.random.default_rng(17000 + seed)
    g = stack(float_graph(rng), float_graph(rng))
    k = rng.integers(0, 4)
    if k == 0:
        g = pipe(g, f"rotate({_c(rng, -3, 3)},{_c(rng, 0.5, 1.1)})")
    elif k == 1:
        g = pipe(g, "reverse(2)", stack(float_stage(rng), "pass()"))
    elif k == 2:
        g = pipe(g, "join(2)", f"pan({_c(rng, -1, 1)})")
    _check(g, "float", 2400, seed, n_out=2)


@pytest.mark.parametrize("seed", range(OFFSET, OFFSET + max(8, N_SEEDS // 5)))
def test_random_graphs_in_ragged_banks(seed):
    """banks whose voice count is not a multiple of a warp / CTA (130 voices), rows and pairwise group mixes, salted voices:
    padding lanes, partial tiles and the left-to-right mix on the lane kernels"""
    from tests.oracle_ffi import render_bank
    rng = np.random.default_rng(21000 + seed)
    exact = bool(seed % 2)
    g = exact_graph(rng) if exact else float_graph(rng)
    V, n = 130, 1100
    salts = (np.arange(V, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    net = build(g, Net)
    for G in (1, 2):
        ref = render_bank([build(g, ONet).set_salt(int(s)) for s in salts], n, group=G, threads=4)
        for pname, path in (("lane_block", qb.PATH_INTERP), ("lane_sample", qb.PATH_INTERP_SAMPLE)):
            bank = Bank(net, V, salts=salts).set_path(path)
            got = np.concatenate([bank.render(401, group=G)[:, 0, :], bank.render(n - 401, group=G)[:, 0, :]], axis=1)
            assert_parity(got, ref, "exact" if exact else "float", f"seed {seed} G={G} [{pname}: {bank.kernel()}] {g}", relative=True)


# ---------------------------------------------------------------- spectral family (the frame-parallel path, K5 / K5s)
def _ra_source(rng):
    """a pure function of time: counter noise, a wave table, an impulse, constants, delays / ticks and stateless maps of those"""
    k = rng.integers(0, 5)
    if k == 0:
        g = L("white()")
    elif k == 1:
        g = {"op": "wave()", "arr": [round(float(x), 3) for x in rng.uniform(-1, 1, int(rng.choice([7, 16, 33, 64])))]}
    elif k == 2:
        g = pipe("impulse()", f"mul({_c(rng, 0.5, 3)})")
    elif k == 3:
        g = mul("white()", {"op": "wave()", "arr": [round(float(x), 3) for x in rng.uniform(0, 1, int(rng.choice([16, 32])))]})
    else:
        g = add("white()", L(f"dc({_c(rng, -0.5, 0.5)})"))
    for _ in range(int(rng.integers(0, 3))):
        j = rng.integers(0, 5)
        if j == 0:
            g = pipe(g, f"delay({_c(rng, 0.0001, 0.004, 5)})")
        elif j == 1:
            g = pipe(g, "tick()")
        elif j == 2:
            g = pipe(g, f"mul({_c(rng, -2, 2)})")
        elif j == 3:
            g = pipe(g, "tanh()")
        else:
            g = add(g, pipe("white()", f"delay({_c(rng, 0.0002, 0.002, 5)})", "mul(0.25)"))
    return g


def _bin_chain(rng):
    """stateless (re, im) -> (re, im) maps: conjugation-equivariant ones (the mirrored half is derived) and others (every bin
    is evaluated), including chains with chain(0) != 0"""
    k = rng.integers(0, 8)
    thr = _c(rng, 0.1, 3)
    if k == 0:
        return []
    if k == 1:   # the gate of assets/spectral-gate
        return ["pol()", stack(pipe(branch(f">({thr})", "pass()"), mul("pass()", "pass()")), "pass()"), "car()"]
    if k == 2:
        return [stack(f"mul({_c(rng, -2, 2)})", f"mul({_c(rng, -2, 2)})")]
    if k == 3:   # not equivariant: a constant lands in the imaginary part
        return [stack(f"add({_c(rng, -0.3, 0.3)})", f"add({_c(rng, -0.3, 0.3)})")]
    if k == 4:
        return ["pol()", stack("sqrt()", f"mul({_c(rng, 0.5, 2)})"), "car()"]
    if k == 5:   # swaps the parts
        return ["reverse()"] if rng.uniform() < 0.5 else [f"rotate({_c(rng, -3, 3)},{_c(rng, 0.5, 1.5)})"]
    if k == 6:
        return [stack("abs()", "pass()")]
    return [stack(pipe("squared()", f"min({thr})"), "cubed()")]


def spectral_graph(rng):
    n = int(rng.choice([8, 16, 64, 128, 256]))
    J = int(rng.choice([1, 2, 4]))
    insts = []
    for j in range(J):
        st = int(rng.integers(0, n)) if rng.uniform() < 0.5 else (j * n // J) % n
        src = _ra_source(rng) if (j == 0 or rng.uniform() < 0.4) else insts[0][0]
        insts.append((src, pipe(src, f"rfft({n},{st})", *_bin_chain(rng), f"ifft({n},{st})")))
    outs = []
    for src, seg in insts:
        k = rng.integers(0, 3)
        if k == 0:
            outs.append(pipe(seg, "chan(1,0)"))
        elif k == 1:
            outs.append(pipe(seg, "join(2)"))                       # real and imaginary parts both read
        else:
            outs.append(mul(pipe(seg, "chan(1,0)"), pipe(_ra_source(rng), "softsign()")))
    g = outs[0]
    for o in outs[1:]:
        g = add(g, o)
    return pipe(g, f"mul({_c(rng, 0.2, 1)})")


@pytest.mark.parametrize("seed", range(OFFSET, OFFSET + N_SEEDS))
def test_random_spectral_graphs_on_the_frame_parallel_path(seed, monkeypatch):
    """K5 against the time-vector kernel BIT FOR BIT and against the oracle within the float tolerance, across uneven calls;
    every fifth seed also through K5s (the plan compiled into the kernels, ~2 s of NVRTC)"""
    rng = np.random.default_rng(7000 + seed)
    expr = spectral_graph(rng)
    V, n = 3, int(rng.integers(300, 1500))
    salts = np.arange(1, V + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15 + seed)
    net = build(expr, Net)
    assert (net.inputs(), net.outputs()) == (0, 1), expr
    assert net.spectral_info() is not None, expr
    ref = np.stack([build(expr, ONet).set_salt(int(s)).render(n).T for s in salts])
    # the time-vector kernel when the tape is hop-alignable (start offsets with a common divisor >= 8), else a lane kernel
    # with its per-lane transforms: the same butterfly network either way
    tv = Bank(net, V, salts=salts).set_path(qb.PATH_TV)
    assert tv.kernel() in ("k_interp_tv", "k_interp_blk", "k_interp<uniform>")
    want = tv.render(n)
    assert_parity(want, ref, "float", f"seed {seed} [{tv.kernel()}] {expr}", relative=True)
    monkeypatch.setenv("QG_SPECTRAL_SPEC", "0")
    k5 = Bank(net, V, salts=salts).set_path(qb.PATH_SPECTRAL)
    cut = int(rng.integers(1, n - 1))
    got = np.concatenate([k5.render(cut), k5.render(n - cut)], axis=2)
    assert k5.kernel() == "k_spectral_frames"
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"seed {seed} K5 vs time-vector {expr}"
    if seed % 5 == 0:
        monkeypatch.setenv("QG_SPECTRAL_SPEC", "1")
        k5s = Bank(net, V, salts=salts).set_path(qb.PATH_SPECTRAL)
        try:
            gs = np.concatenate([k5s.render(cut), k5s.render(n - cut)], axis=2)
        except qb.QuartzGpuError as e:
            if "NVRTC" in str(e):
                pytest.skip("NVRTC not available on this box")
            raise
        assert k5s.kernel() == "k_sp_frames"
        assert np.array_equal(gs.view(np.uint32), want.view(np.uint32)), f"seed {seed} K5s vs time-vector {expr}"
