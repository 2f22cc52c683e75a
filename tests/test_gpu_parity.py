"""GPU parity: every case goes through the C ABI (libquartz_gpu.so) and is compared with the CPU oracle on the
same graph, sample for sample."""
import json
import os

import numpy as np
import pytest

import quartz_b200 as qb
from quartz_b200 import Bank, Net
from tests import cases
from tests.graphs import L, build, pipe, stack
from tests.oracle_ffi import ONet, render_bank
from tests.util import assert_parity

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "quartz_assets.json")))


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_asset_golden_vectors_on_gpu(case):
    net = build(case["net"], Net)
    exp = np.array(case["output_bits"], dtype=np.uint32).view(np.float32)
    if case["kind"] == "apply":
        # the six known-answer `apply` pairs of assets/wip, curve and distance (wrap, Quantizer, semitone_ratio through the
        # correctly-rounded exp2, IEEE division): reproduced BIT FOR BIT on the device (measured: error 0.0 on all six)
        assert_parity(net.tick(case["input"]), exp, "exact", case["name"])
    else:
        net.set_sample_rate(case["sample_rate"])
        assert_parity(net.render(case["len"])[:, 0], exp, "float", case["name"])


@pytest.mark.parametrize("name,expr,n,tol", cases.RENDER, ids=[c[0] for c in cases.RENDER])
def test_render_matches_oracle(name, expr, n, tol):
    got = build(expr, Net).render(n)
    ref = build(expr, ONet).render(n)
    assert_parity(got, ref, tol if tol in ("exact", "float") else "float", name)


@pytest.mark.parametrize("op,dom,tol", cases.PROCESS, ids=[c[0] for c in cases.PROCESS])
def test_process_matches_oracle(op, dom, tol):
    a, b = Net.str_to_net(op), ONet.str_to_net(op)
    rng = np.random.default_rng(sum(map(ord, op)) * 7919 % (2 ** 31))
    n = 512
    x = rng.uniform(dom[0], dom[1], (n, a.inputs())).astype(np.float32)
    if a.inputs() >= 2 and op in ("shift_reg()", "snh()"):
        x[:, 1] = (rng.uniform(0, 1, n) < 0.2).astype(np.float32)
    # sprinkle special values on the stateless maps
    if tol == "exact" and a.inputs() == 1 and op not in ("tick()",):
        x[::37, 0] = [0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 2.5, -2.5, 1e-40, np.inf, -np.inf, np.nan, 3e9, -3e9][: len(x[::37, 0])]
    assert_parity(a.process(x), b.process(x), tol, op)


def test_state_persists_across_calls_like_audio_unit():
    expr = pipe("white()", "lowpass(800,2)")
    a, b = build(expr, Net), build(expr, ONet)
    got = np.concatenate([a.render(100), a.render(37), a.render(1), a.render(500)])
    assert_parity(got, b.render(638), "float", "chunked render")
    a.reset()
    b.reset()
    assert_parity(a.render(64), b.render(64), "float", "after reset")


def test_bank_per_voice_parameters_and_salts():
    V, T = 100, 1024
    rng = np.random.default_rng(3)
    hz = np.exp(rng.uniform(np.log(50), np.log(12000), V)).astype(np.float32)
    q = rng.uniform(0.5, 8, V).astype(np.float32)
    salts = (np.arange(V, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    tmpl = build(pipe("white()", "lowpass(1000,1)"), Net).set_sample_rate(48000)
    bank = Bank(tmpl, V, raw=np.stack([hz, q], axis=1), salts=salts).set_path(qb.PATH_INTERP)
    got = bank.render(T)[:, 0, :]
    onets = [build(pipe("white()", f"lowpass({float(hz[v])!r},{float(q[v])!r})"), ONet).set_sample_rate(48000).set_salt(int(salts[v]))
             for v in range(V)]
    ref = render_bank(onets, T)
    assert_parity(got, ref, "float", "bank")
    # same thing built from V separate nets (one `render` circle per voice in the reference)
    nets = [build(pipe("white()", f"lowpass({float(hz[v])!r},{float(q[v])!r})"), Net).set_sample_rate(48000) for v in range(V)]
    bank2 = Bank(None, nets=nets, salts=salts).set_path(qb.PATH_INTERP)
    assert_parity(bank2.render(T)[:, 0, :], ref, "float", "bank_from_nets")
    # frame-major layout is the transpose
    bank.reset()
    fm = bank.render(T, layout=qb.LAYOUT_FRAME_MAJOR)[:, :, 0]
    assert_parity(fm.T, ref, "float", "frame-major")


def test_bank_group_mix_is_left_to_right_sum():
    V, T, G = 64, 777, 32
    salts = np.arange(1, V + 1, dtype=np.uint64)
    tmpl = build(pipe("white()", "mul(0.5)"), Net)
    bank = Bank(tmpl, V, salts=salts).set_path(qb.PATH_INTERP)
    got = bank.render(T, group=G)[:, 0, :]
    onets = [build(pipe("white()", "mul(0.5)"), ONet).set_salt(int(s)) for s in salts]
    ref = render_bank(onets, T, group=G)
    assert_parity(got, ref, "exact", "group mix")


def test_render_refuses_nets_with_inputs():
    with pytest.raises(qb.QuartzGpuError):
        Net.str_to_net("lowpass(1000,1)").render(8)


# ---- every kernel family on the same graphs: AUTO picks per bank, the others are forced
PATHS = [("auto", qb.PATH_AUTO), ("lane_block", qb.PATH_INTERP), ("lane_sample", qb.PATH_INTERP_SAMPLE)]


@pytest.mark.parametrize("pname,path", PATHS[1:], ids=[p[0] for p in PATHS[1:]])
@pytest.mark.parametrize("name,expr,n,tol", cases.RENDER, ids=[c[0] for c in cases.RENDER])
def test_render_matches_oracle_on_lane_kernels(name, expr, n, tol, pname, path):
    net = build(expr, Net)
    bank = Bank(net, 1).set_path(path)
    n = n + 3   # not a multiple of the block length: exercises the partial last block
    got = bank.render(n, layout=qb.LAYOUT_FRAME_MAJOR).reshape(n, net.outputs())
    ref = build(expr, ONet).render(n)
    assert_parity(got, ref, tol if tol in ("exact", "float") else "float", f"{name} [{pname}: {bank.kernel()}]")


@pytest.mark.parametrize("pname,path", PATHS[1:], ids=[p[0] for p in PATHS[1:]])
def test_block_kernel_chunked_render_and_voice_major_tiles(pname, path):
    """state carried across calls of odd lengths; voice-major output tiles with a partial last column block"""
    V = 70
    expr = pipe("white()", "delay(0.001)", "lowpole(900)", "tick()")
    salts = np.arange(1, V + 1, dtype=np.uint64)
    bank = Bank(build(expr, Net), V, salts=salts).set_path(path)
    parts = np.concatenate([bank.render(k)[:, 0, :] for k in (1, 7, 8, 9, 33, 250, 64, 5)], axis=1)
    onets = [build(expr, ONet).set_salt(int(s)) for s in salts]
    ref = render_bank(onets, parts.shape[1])
    assert_parity(parts, ref, "float", f"chunked [{bank.kernel()}]")


@pytest.mark.parametrize("delay,blk", [(None, False), (0.0001, False), (0.001, True)])
def test_feedback_ring_shorter_than_a_block_stays_sample_by_sample(delay, blk):
    """FunDSP FeedbackUnit with 1-, 4- and 44-sample loops: only the last may run block-wise (ring >= block length)"""
    inner = pipe("mul(0.7)", "lowpole(3000)")
    expr = pipe("white()", {"op": "feedback()", "net": inner, "delay": delay})
    V, T = 40, 1500
    salts = np.arange(1, V + 1, dtype=np.uint64)
    bank = Bank(build(expr, Net), V, salts=salts).set_path(qb.PATH_INTERP)
    assert (bank.kernel() == "k_interp_blk") == blk, bank.kernel()
    onets = [build(expr, ONet).set_salt(int(s)) for s in salts]
    assert_parity(bank.render(T)[:, 0, :], render_bank(onets, T), "float", f"feedback delay={delay}")
