"""BASELINE.json configs at their FULL sizes, checked through size-independent properties (the scalar CPU oracle cannot
render 10^10 voice-samples in a test):
  * bank-size independence: a voice rendered inside the full bank (one wave of warps, fused kernel, group mixes) equals
    the same voice rendered in a small bank (different launch geometry: time segments / other block shapes), and a
    subset of those rows is also compared with the oracle at full length;
  * encode -> decode round trip of the spectral chain at full size (thr < 0 passes every bin: pure 2N delay x 0.375);
  * group mixes: rows of the full 1M-voice bank equal the left-to-right mix of the same voices rendered on their own.
Outputs stay on the device (torch only owns the buffers); rows are copied back selectively."""
import numpy as np
import pytest
import torch

import quartz_b200 as qb
from quartz_b200 import Bank, Net, workloads
from tests.graphs import build
from tests.oracle_ffi import ONet, render_bank
from tests.util import assert_parity

pytestmark = pytest.mark.gpu


def _render_device(bank, rows, T, group=1):
    d = torch.empty(rows * T, dtype=torch.float32, device="cuda")
    bank.render_device(T, d.data_ptr(), group=group)
    bank.ctx.synchronize()
    return d.view(rows, T)


def _sub_bank(wl, voices):
    return Bank(build(wl.expr, Net), len(voices), raw=None if wl.raw is None else wl.raw[voices], salts=wl.salts[voices])


def test_c2_full_size_rows_do_not_depend_on_the_bank():
    """configs[1]: 4,096 noise->lowpass voices x 2,880,000 samples (47.2 GB resident)"""
    wl = workloads.c2_lowpass_bank()
    full = Bank(build(wl.expr, Net), wl.V, raw=wl.raw, salts=wl.salts)
    assert full.kernel() == "k_noise_svf_scan"
    d = _render_device(full, wl.V, wl.T)
    order = np.argsort(wl.raw[:, 1] / wl.raw[:, 0])            # q/hz: slowest-decaying filters last
    pick = np.array(sorted(set(order[:3]) | set(order[-3:]) | {0, 1777, 4095}))
    got = d[torch.as_tensor(pick, device="cuda")].cpu().numpy()
    del d
    torch.cuda.empty_cache()
    small = _sub_bank(wl, pick)                                  # 9 voices: rendered in time segments (chained states)
    assert_parity(got, small.render(wl.T)[:, 0, :], "float", "c2: full bank vs 9-voice bank")
    ref = render_bank([build(wl.voice_expr(int(v)), ONet).set_salt(int(wl.salts[v])) for v in pick[:4]], wl.T, threads=4)
    assert_parity(got[:4], ref, "float", "c2 full size vs oracle")
    assert np.isfinite(got).all() and np.abs(got).max() < 50.0


def test_c3_full_size_group_rows_do_not_depend_on_the_bank():
    """configs[2]: 65,536 osc->lowpass->envelope voices x 480,000 samples, mixed in groups of 32"""
    wl = workloads.c3_polysynth()
    G = wl.group
    full = Bank(build(wl.expr, Net), wl.V, raw=wl.raw, salts=wl.salts)
    assert full.kernel() == "k_polysynth"
    d = _render_device(full, wl.V // G, wl.T, group=G)
    groups = [0, 1023, 2047]
    got = d[torch.as_tensor(groups, device="cuda")].cpu().numpy()
    del d
    for row, g in zip(got, groups):
        voices = np.arange(g * G, (g + 1) * G)
        small = _sub_bank(wl, voices)
        assert_parity(row[None, :], small.render(wl.T, group=G)[:, 0, :], "float", f"c3 group {g}: full bank vs 32-voice bank")
    ref = render_bank([build(wl.voice_expr(v), ONet).set_salt(int(wl.salts[v])) for v in range(G)], wl.T, group=G, threads=8)
    assert_parity(got[:1], ref, "float", "c3 full size vs oracle")


def test_c4_full_size_round_trip():
    """configs[3] shape (1,024 channels, 2048-point transforms, hop 512) on external input, 6 s: with every bin passing the
    patch is a pure delay of 2N samples scaled by the mean of Hann^2 over the 4 overlapping instances (0.375)"""
    N, J, V, T = 2048, 4, 1024, 288000
    net = build(workloads.spectral_graph(N, J, -1.0, workloads.hann(N), source="pass()"), Net)
    bank = Bank(net, V)
    assert bank.kernel() == "k_interp_tv"
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, (V, 1, T)).astype(np.float32)
    y = bank.process(x, T)[:, 0, :]
    lat = 2 * N
    err = np.abs(y[:, lat + N:] - 0.375 * x[:, 0, N:T - lat]).max()
    assert err < 2e-4, err


def test_c4_full_size_gate_rows_do_not_depend_on_the_bank_or_the_kernel():
    """configs[3] as benchmarked: 1,024 channels x 1,440,000 samples of the spectral gate on the frame-parallel path (K5s when
    NVRTC is there).  A channel rendered inside the full bank equals — bit for bit, the transforms share their arithmetic —
    the same channel rendered alone on the time-vector kernel, and the oracle within the float tolerance (first 60,000
    samples: the scalar oracle needs ~1 s per 10,000 channel-samples)."""
    wl = workloads.c4_spectral()
    full = Bank(build(wl.expr, Net), wl.V, salts=wl.salts)
    d = _render_device(full, wl.V, wl.T)
    assert full.kernel() in ("k_spectral_frames", "k_sp_frames")
    chans = [0, 511, 1023]
    got = d[torch.as_tensor(chans, device="cuda")].cpu().numpy()
    tail = d[:, -4096:].abs().max().item()
    del d
    assert 0.0 < tail < 4.0                                   # the last frames were rendered, and are audio
    small = _sub_bank(wl, np.asarray(chans)).set_path(qb.PATH_TV)
    assert small.kernel() == "k_interp_tv"
    assert np.array_equal(got, small.render(wl.T)[:, 0, :])
    n = 60000
    ref = render_bank([build(wl.expr, ONet).set_salt(int(wl.salts[v])) for v in chans[:2]], n, threads=8)
    assert_parity(got[:2, :n], ref, "float", "c4 full size vs oracle")


def test_c4_round_trip_on_the_frame_parallel_path():
    """every bin passing (thr < 0), a wave-table source: the patch is a pure 2N delay x 0.375 of its input — through K5"""
    N, J, V, T = 2048, 4, 64, 40000
    rng = np.random.default_rng(11)
    src = rng.uniform(-1, 1, 4096).astype(np.float32)
    expr = workloads.spectral_graph(N, J, -1.0, workloads.hann(N), source={"op": "wave()", "arr": [float(v) for v in src]})
    bank = Bank(build(expr, Net), V).set_path(qb.PATH_SPECTRAL)
    y = bank.render(T)[:, 0, :]
    x = np.tile(src, T // 4096 + 1)[:T]
    lat = 2 * N
    err = np.abs(y[:, lat + N:] - 0.375 * x[None, N:T - lat]).max()
    assert err < 2e-4, err


def test_c5_full_size_group_rows_do_not_depend_on_the_bank():
    """configs[4]: 1,048,576 voices (4 archetypes x 262,144), group-mixed by 32; 0.25 s per voice here (the property does
    not depend on the length; the 2 s render is the bench's)"""
    T = 12000
    for wl in workloads.c5_mixed(T=T):
        G = wl.group
        full = Bank(build(wl.expr, Net), wl.V, raw=wl.raw, salts=wl.salts)
        d = _render_device(full, wl.V // G, T, group=G)
        groups = [0, 4099, wl.V // G - 1]
        got = d[torch.as_tensor(groups, device="cuda")].cpu().numpy()
        del d, full
        torch.cuda.empty_cache()
        tol = "exact" if wl.name in ("c5b_shift_reg",) else "float"
        for row, g in zip(got, groups):
            voices = np.arange(g * G, (g + 1) * G)
            small = _sub_bank(wl, voices).set_path(qb.PATH_INTERP_SAMPLE)      # sample-by-sample kernel, 32 voices
            assert_parity(row[None, :], small.render(T, group=G)[:, 0, :], tol, f"{wl.name} group {g}")
        ref = render_bank([build(wl.voice_expr(v), ONet).set_salt(int(wl.salts[v])) for v in range(G)], T, group=G, threads=8)
        assert_parity(got[:1], ref, "float", f"{wl.name} vs oracle")
