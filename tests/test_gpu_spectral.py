"""Spectral chain (rfft -> bin chain -> ifft, nodes.rs:601-700) on the time-vector interpreter (K3/K4): parity with the
oracle at sizes it finishes in seconds, agreement with the per-lane path, and a round-trip property at larger sizes."""
import numpy as np
import pytest

import quartz_b200 as qb
from quartz_b200 import Bank, Net, workloads
from tests import cases
from tests.graphs import L, build, pipe, stack
from tests.oracle_ffi import ONet, render_bank
from tests.util import assert_parity

pytestmark = pytest.mark.gpu


def oracle(wl, V, T):
    return render_bank([build(wl.expr, ONet).set_salt(int(wl.salts[v])) for v in range(V)], T, threads=8)


@pytest.mark.parametrize("N,J,V,T", [(64, 4, 5, 1000), (256, 4, 3, 3000), (2048, 4, 2, 14000), (128, 2, 4, 900)])
def test_spectral_gate_matches_oracle(N, J, V, T):
    wl = workloads.c4_spectral(V=V, T=T, N=N, J=J, thr=0.25 * np.sqrt(N))
    bank = Bank(build(wl.expr, Net), V, salts=wl.salts)
    assert bank.kernel() == "k_interp_tv"
    got = bank.render(T)[:, 0, :]
    ref = oracle(wl, V, T)
    assert_parity(got, ref, "float", f"spectral N={N}")
    # the per-lane interpreter (rings [pos][voice], per-lane FFT) agrees as well
    lane = Bank(build(wl.expr, Net), V, salts=wl.salts).set_path(qb.PATH_INTERP)
    assert lane.kernel() in ("k_interp_blk", "k_interp<uniform>")
    assert_parity(lane.render(T)[:, 0, :], ref, "float", "lane path")
    # block-wise continuation
    parts = np.concatenate([Bank(build(wl.expr, Net), V, salts=wl.salts).render(T // 2)[:, 0, :]] * 1, axis=1)
    b2 = Bank(build(wl.expr, Net), V, salts=wl.salts)
    both = np.concatenate([b2.render(T // 2)[:, 0, :], b2.render(T - T // 2)[:, 0, :]], axis=1)
    assert_parity(both, ref, "float", "two-call render")


@pytest.mark.parametrize("name,expr,n,tol", cases.RENDER, ids=[c[0] for c in cases.RENDER])
def test_time_vector_interpreter_on_generic_cases(name, expr, n, tol):
    """every render case whose tape is time-vector capable (feed-forward; stateless ops, counter sources, delay lines,
    transforms, fixed LTI filters as block-level scans, phase accumulators stepped by one thread)"""
    net = build(expr, Net)
    bank = Bank(net, 3).set_path(qb.PATH_TV)
    if bank.kernel() != "k_interp_tv":
        pytest.skip("graph is not time-vector capable")
    got = bank.render(n)      # [3, outs, n]
    ref = build(expr, ONet).render(n).T
    for v in range(3):
        assert_parity(got[v], ref, tol if tol in ("exact", "float") else "float", name)


def test_time_vector_covers_filters_and_oscillators():
    capable = {name for name, expr, n, tol in cases.RENDER if Bank(build(expr, Net), 1).set_path(qb.PATH_TV).kernel() == "k_interp_tv"}
    for must in ("white_lowpass", "biquad", "butterpass", "resonator", "lowpole", "highpole", "dcblock", "allpole", "white_bell"):
        assert must in capable, must


@pytest.mark.parametrize("expr", [
    pipe("white()", "lowpass(300,6)", "highpole(40)", "butterpass(2500)"),
    pipe("sine(220.5)", "resonator(900,30)", "allpole(0.3)"),
    pipe(pipe("dc(3.3)", "ramp()", "mul(TAU)", "sin()", "mul(300)", "add(500)"), "sine()", "dcblock()", "lowpole(1200)"),
], ids=["noise_svf_onepole_biquad", "sine_resonator_allpole", "fm_sine_dcblock_lowpole"])
def test_time_vector_scans_persist_state_across_odd_chunks(expr):
    """block-level scans + one-thread phase recurrences: hop-partial renders (1, 7, 511, 513 ...) continue exactly where the
    previous call stopped, for every voice of a salted bank"""
    V = 5
    salts = np.arange(11, 11 + V, dtype=np.uint64)
    bank = Bank(build(expr, Net), V, salts=salts).set_path(qb.PATH_TV)
    assert bank.kernel() == "k_interp_tv"
    chunks = (1, 7, 511, 513, 2, 1024, 700, 3000)
    parts = np.concatenate([bank.render(k)[:, 0, :] for k in chunks], axis=1)
    ref = render_bank([build(expr, ONet).set_salt(int(s)) for s in salts], sum(chunks))
    assert_parity(parts, ref, "float", "time-vector chunked")
    lane = Bank(build(expr, Net), V, salts=salts).set_path(qb.PATH_INTERP_SAMPLE)
    assert_parity(lane.render(sum(chunks))[:, 0, :], ref, "float", "lane path")


def test_round_trip_property_full_frame_size():
    """thr < 0 passes every bin: the patch is then a pure delay of 2N samples scaled by the Hann^2 overlap mean 0.375
    (encode -> decode round trip through K3/K4 with an external input, block path with host buffers)."""
    N, J, V, T = 2048, 4, 48, 40000
    expr = workloads.spectral_graph(N, J, -1.0, workloads.hann(N), source="pass()")
    net = build(expr, Net)
    assert (net.inputs(), net.outputs()) == (1, 1)
    bank = Bank(net, V)
    assert bank.kernel() == "k_interp_tv"
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, (V, 1, T)).astype(np.float32)
    y = bank.process(x, T)[:, 0, :]
    lat = 2 * N
    err = np.abs(y[:, lat + N:] - 0.375 * x[:, 0, N:T - lat]).max()
    assert err < 2e-4, err


# ---------------------------------------------------------------------------------------------- K5: frame-parallel path
def _gate(N, J, V, T, thr=None):
    return workloads.c4_spectral(V=V, T=T, N=N, J=J, thr=0.25 * np.sqrt(N) if thr is None else thr)


@pytest.mark.parametrize("N,J,V,T", [(64, 4, 5, 1000), (256, 4, 3, 3000), (2048, 4, 2, 14000), (128, 2, 4, 900), (8, 1, 2, 100)])
def test_frame_parallel_spectral_path_matches_oracle_and_time_vector_kernel(N, J, V, T):
    """K5 evaluates whole frames from the state at reset + absolute sample time; same transform arithmetic and the same
    op code as the time-vector kernel, so the two agree BIT FOR BIT, and with the oracle within the float tolerance"""
    wl = _gate(N, J, V, T)
    k5 = Bank(build(wl.expr, Net), V, salts=wl.salts).set_path(qb.PATH_SPECTRAL)
    assert k5.kernel() == "k_spectral_frames"
    got = k5.render(T)[:, 0, :]
    tv = Bank(build(wl.expr, Net), V, salts=wl.salts).set_path(qb.PATH_TV).render(T)[:, 0, :]
    assert np.array_equal(got, tv)
    assert_parity(got, oracle(wl, V, T), "float", f"K5 N={N}")
    # frame-major output
    fm = Bank(build(wl.expr, Net), V, salts=wl.salts).set_path(qb.PATH_SPECTRAL).render(T, layout=qb.LAYOUT_FRAME_MAJOR)
    assert np.array_equal(fm[:, :, 0].T, got)


def test_frame_parallel_path_continues_across_uneven_calls_and_resets():
    N, J, V = 256, 4, 6
    chunks = (1, 7, 255, 257, 2, 1024, 700, 3000)
    T = sum(chunks)
    wl = _gate(N, J, V, T)
    ref = Bank(build(wl.expr, Net), V, salts=wl.salts).set_path(qb.PATH_TV).render(T)
    b = Bank(build(wl.expr, Net), V, salts=wl.salts).set_path(qb.PATH_SPECTRAL)
    assert np.array_equal(np.concatenate([b.render(k) for k in chunks], axis=2), ref)
    b.reset()
    assert np.array_equal(b.render(1000), ref[:, :, :1000])
    c = b.clone()                                   # the clone carries the sample time
    assert np.array_equal(c.render(500), ref[:, :, 1000:1500]) and np.array_equal(b.render(500), ref[:, :, 1000:1500])
    # var() updates need real state: refused by name on a bank that is mid-render on K5, fine after a reset + PATH_TV
    with pytest.raises(qb.QuartzGpuError, match="reset"):
        b.set_raw(0, 1.0)
    # leaving K5 resets the bank (its state is a sample time, not state words)
    b.set_path(qb.PATH_TV)
    assert b.kernel() == "k_interp_tv" and np.array_equal(b.render(T), ref)


def test_auto_takes_the_frame_parallel_path_for_bulk_renders_only():
    N, J = 64, 4
    wl = _gate(N, J, 64, 8192)
    b = Bank(build(wl.expr, Net), 64, salts=wl.salts)
    assert b.kernel() == "k_interp_tv"
    small = b.render(1000)                          # a short first call: block-wise streaming stays on the time-vector kernel
    assert b.kernel() == "k_interp_tv"
    b.reset()
    bulk = b.render(8192)                           # 64 x 8192 voice-samples: bulk
    assert b.kernel() == "k_spectral_frames"
    assert np.array_equal(bulk[:, :, :1000], small)
    ref = Bank(build(wl.expr, Net), 64, salts=wl.salts).set_path(qb.PATH_TV).render(8192)
    assert np.array_equal(bulk, ref)
    with pytest.raises(qb.QuartzGpuError, match="spectral"):
        Bank(build(pipe("white()", "lowpass(500,1)"), Net), 4).set_path(qb.PATH_SPECTRAL)


@pytest.mark.parametrize("chain", [
    [],                                                            # rfft straight into ifft: a 2N delay
    ["pol()", "car()"],
    [stack("mul(0.5)", "add(0.25)")],                              # chain(0) != 0: zero-initialised positions must stay zero
    [stack(pipe("abs()", "sqrt()"), "mul(-1)")],
], ids=["identity", "pol_car", "affine", "nonlinear"])
@pytest.mark.parametrize("start", [0, 5, 32])
def test_frame_parallel_path_on_other_bin_chains(chain, start):
    N = 64
    expr = pipe(stack("white()"), f"rfft({N},{start})", *chain, f"ifft({N},{start})", stack(pipe("mul(0.7)", "tanh()"), "pass()"))
    V, T = 3, 700
    salts = np.arange(3, 3 + V, dtype=np.uint64)
    b = Bank(build(expr, Net), V, salts=salts).set_path(qb.PATH_SPECTRAL)
    got = np.concatenate([b.render(123), b.render(T - 123)], axis=2)
    tv = Bank(build(expr, Net), V, salts=salts).set_path(qb.PATH_TV).render(T)
    assert np.array_equal(got, tv)
    ref = np.stack([build(expr, ONet).set_salt(int(s)).render(T).T for s in salts])
    assert_parity(got, ref, "float", "chain")


@pytest.mark.parametrize("N,J,V,T", [(64, 4, 5, 1000), (2048, 4, 2, 14000), (128, 2, 4, 900)])
def test_specialised_spectral_kernels_are_bit_identical_to_the_generic_ones(N, J, V, T, monkeypatch):
    """K5s (the plan compiled into the kernels by NVRTC): same arithmetic in the same order as the generic K5 kernels"""
    wl = _gate(N, J, V, T)
    monkeypatch.setenv("QG_SPECTRAL_SPEC", "0")
    gen = Bank(build(wl.expr, Net), V, salts=wl.salts).set_path(qb.PATH_SPECTRAL)
    want = gen.render(T)
    assert gen.kernel() == "k_spectral_frames"
    monkeypatch.setenv("QG_SPECTRAL_SPEC", "1")
    sp = Bank(build(wl.expr, Net), V, salts=wl.salts).set_path(qb.PATH_SPECTRAL)
    try:
        got = np.concatenate([sp.render(T // 3), sp.render(T - T // 3)], axis=2)
    except qb.QuartzGpuError as e:
        if "NVRTC" in str(e):
            pytest.skip("NVRTC not available on this box")
        raise
    assert sp.kernel() == "k_sp_frames"
    assert np.array_equal(got, want)
    fm = Bank(build(wl.expr, Net), V, salts=wl.salts).set_path(qb.PATH_SPECTRAL).render(T, layout=qb.LAYOUT_FRAME_MAJOR)
    assert np.array_equal(fm[:, :, 0].T, want[:, 0, :])


@pytest.mark.parametrize("chain", [[], ["pol()", "car()"], [stack("mul(0.5)", "add(0.25)")]], ids=["identity", "pol_car", "affine"])
def test_specialised_spectral_kernels_on_other_bin_chains(chain, monkeypatch):
    N, start, V, T = 64, 5, 3, 700
    expr = pipe(stack("white()"), f"rfft({N},{start})", *chain, f"ifft({N},{start})", stack(pipe("mul(0.7)", "tanh()"), "pass()"))
    salts = np.arange(3, 3 + V, dtype=np.uint64)
    tv = Bank(build(expr, Net), V, salts=salts).set_path(qb.PATH_TV).render(T)
    monkeypatch.setenv("QG_SPECTRAL_SPEC", "1")
    b = Bank(build(expr, Net), V, salts=salts).set_path(qb.PATH_SPECTRAL)
    try:
        got = b.render(T)
    except qb.QuartzGpuError as e:
        if "NVRTC" in str(e):
            pytest.skip("NVRTC not available on this box")
        raise
    assert b.kernel() == "k_sp_frames" and np.array_equal(got, tv)


def test_single_graph_bulk_render_takes_the_frame_parallel_path_with_many_rounds_per_launch():
    """the `render` op on ONE spectral graph (process.rs:1345-1351, up to 10^7 samples): a one-voice bank fills the machine by
    taking up to 63 rounds per launch (Y ring of 64 rounds); same samples as the time-vector kernel"""
    N, J, T = 256, 4, 300000
    wl = _gate(N, J, 1, T)
    net = build(wl.expr, Net)
    b = Bank(net, 1, salts=wl.salts)
    got = b.render(T)
    assert b.kernel() in ("k_spectral_frames", "k_sp_frames")
    want = Bank(net, 1, salts=wl.salts).set_path(qb.PATH_TV).render(T)
    assert np.array_equal(got, want)
    # and through the single-graph convenience call (qg_net_render)
    one = build(wl.expr, Net)
    assert np.array_equal(one.render(T)[:, 0], Bank(one, 1).set_path(qb.PATH_TV).render(T)[0, 0])
