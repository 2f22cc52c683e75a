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


TV_CASES = ["white", "wave", "impulse", "tick", "delay", "tap_noise", "tap_zero_delay", "tap_linear_noise", "quantize", "arr_get", "rfft_ifft_roundtrip",
            "rfft_start", "spectral_gate_small", "split_join", "chan_pan", "rotate", "dc3", "live_io"]


@pytest.mark.parametrize("name", TV_CASES)
def test_time_vector_interpreter_on_generic_cases(name):
    _, expr, n, tol = next(c for c in cases.RENDER if c[0] == name)
    net = build(expr, Net)
    bank = Bank(net, 3).set_path(qb.PATH_TV)
    if bank.kernel() != "k_interp_tv":
        pytest.skip("graph is not time-vector capable")
    got = bank.render(n)      # [3, outs, n]
    ref = build(expr, ONet).render(n).T
    for v in range(3):
        assert_parity(got[v], ref, tol if tol in ("exact", "float") else "float", name)


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
