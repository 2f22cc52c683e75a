"""The oracle against every known-answer vector the reference ships (scene-embedded arrays).
CPU only.  This is what pins the in-tree half of the oracle; FunDSP-internal ops stay parity-unpinned."""
import json
import os

import numpy as np
import pytest

from tests.graphs import build
from tests.oracle_ffi import ONet, inverse_fft, real_fft

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "quartz_assets.json")))


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_oracle_matches_asset_vector(case):
    net = build(case["net"], ONet)
    exp = np.array(case["output_bits"], dtype=np.uint32)
    if case["kind"] == "apply":
        out = net.tick(case["input"])
    else:
        net.set_sample_rate(case["sample_rate"])
        out = net.render(case["len"])[:, 0]
    assert (out.view(np.uint32) == exp).all()


def test_oracle_fft_matches_numpy():
    rng = np.random.default_rng(7)
    for n in (2, 8, 64, 2048):
        x = rng.uniform(-1, 1, n).astype(np.float32)
        X = real_fft(x)
        ref = np.fft.rfft(x.astype(np.float64))
        assert np.abs(X - ref).max() <= 2e-6 * np.sqrt(n) * max(1.0, np.abs(ref).max())
        z = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
        zi = inverse_fft(z)
        assert np.abs(zi - np.fft.ifft(z.astype(np.complex128))).max() <= 1e-6
        # unity-gain round trip (the asset gain structure relies on it)
        full = np.fft.fft(x.astype(np.float64)).astype(np.complex64)
        assert np.abs(inverse_fft(full).real - x).max() <= 1e-5


def test_oracle_in_tree_nodes_known_answers():
    # ShiftReg (nodes.rs:173-185): shifts only on a non-zero trigger
    sr = ONet.str_to_net("shift_reg()")
    assert sr.tick([1.0, 0.0]).tolist() == [0.0] * 8
    assert sr.tick([2.0, 1.0]).tolist() == [2.0] + [0.0] * 7
    assert sr.tick([3.0, -1.0]).tolist() == [3.0, 2.0] + [0.0] * 6
    # SnH (nodes.rs:811-816)
    snh = ONet.str_to_net("snh()")
    assert snh.tick([5.0, 0.0])[0] == 0.0 and snh.tick([5.0, 1.0])[0] == 5.0 and snh.tick([7.0, 0.0])[0] == 5.0
    # SampDelay (nodes.rs:726-731): index 0 is the sample just pushed
    sd = ONet.str_to_net("samp_delay(4)")
    assert sd.tick([1.0, 0.0])[0] == 1.0 and sd.tick([2.0, 1.0])[0] == 1.0 and sd.tick([3.0, 4.0])[0] == 0.0
    # ArrGet (nodes.rs:143-149): OOB and negative/NaN indices -> 0 / element 0
    g = ONet.get([10.0, 20.0, 30.0])
    assert [g.tick([i])[0] for i in (0.0, 1.9, 2.0, 3.0, -1.0, float("nan"))] == [10.0, 20.0, 30.0, 0.0, 10.0, 10.0]
    # Ramp (nodes.rs:476-483) at the default 44.1 kHz
    r = ONet.str_to_net("ramp()")
    assert r.tick([22050.0])[0] == 0.0 and r.tick([22050.0])[0] == 0.5 and r.tick([0.0])[0] == 0.0
    # exp_m1 / ln_1p are swapped in the reference (functions.rs:1077-1078)
    assert abs(ONet.str_to_net("exp_m1()").tick([1.0])[0] - np.log1p(1.0)) < 1e-6
    assert abs(ONet.str_to_net("ln_1p()").tick([1.0])[0] - np.expm1(1.0)) < 1e-6
    # Rfft emits the previous frame's bins, mirrored above n/2 (nodes.rs:625-642)
    f = ONet.str_to_net("rfft(8,0)")
    x = np.arange(1, 9, dtype=np.float32)
    for v in x:
        f.tick([v])
    bins = np.array([f.tick([0.0]) for _ in range(8)])
    ref = np.fft.fft(x.astype(np.float64))
    assert np.abs(bins[:, 0] + 1j * bins[:, 1] - ref).max() < 1e-4


DIGEST = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_exact_digest.json")))


@pytest.mark.parametrize("name", sorted(DIGEST))
def test_oracle_behaviour_is_frozen_on_the_exact_cases(name):
    """the oracle IS the parity reference of the GPU path: its output on every bit-exact case is pinned by digest
    (tests/golden/make_oracle_digest.py), so an edit of oracle/ that changes behaviour shows up here, on the CPU"""
    from tests import cases
    from tests.golden.make_oracle_digest import digest
    expr, n = next((c[1], c[2]) for c in cases.RENDER if c[0] == name)
    h, shape = digest(expr, n)
    assert shape == DIGEST[name]["shape"] and h == DIGEST[name]["sha1"], name
