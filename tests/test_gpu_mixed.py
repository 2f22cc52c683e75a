"""BASELINE configs[4] archetypes (1M-voice mixed graph) at sizes the oracle finishes in seconds, per archetype and
group-mixed.  Integer/trigger paths are bit-exact; only the sin() oscillator archetype is held to the f32 tolerance."""
import numpy as np
import pytest

import quartz_b200 as qb
from quartz_b200 import Bank, Net, workloads
from tests.graphs import build
from tests.oracle_ffi import ONet, render_bank
from tests.util import assert_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("G", [1, 32])
def test_mixed_archetypes_match_oracle(G):
    V, T = 4 * 96, 6000
    tols = {"c5a_quantised_osc": "float", "c5b_shift_reg": "exact", "c5c_feedback": "exact", "c5d_delay_lowpole": "exact"}
    for wl in workloads.c5_mixed(V=V, T=T, G=G):
        tmpl = build(wl.expr, Net)
        assert tmpl.raw_params().shape[0] == wl.raw.shape[1], (wl.name, tmpl.raw_params(), wl.raw[0])
        bank = Bank(tmpl, wl.V, raw=wl.raw, salts=wl.salts).set_path(qb.PATH_INTERP)
        got = bank.render(T, group=G)[:, 0, :]
        onets = [build(wl.voice_expr(v), ONet).set_salt(int(wl.salts[v])) for v in range(wl.V)]
        ref = render_bank(onets, T, group=G, threads=8)
        # group mixes of bit-exact voices stay bit-exact (left-to-right sum, then * 1/G)
        assert_parity(got, ref, tols[wl.name], f"{wl.name} G={G}")


def test_structural_parameters_must_match_across_voices():
    wl = workloads.c5_mixed(V=8, T=16)[3]
    raw = wl.raw.copy()
    raw[1, 0] *= 2      # a different delay time changes the ring length: not batchable
    with pytest.raises(qb.QuartzGpuError):
        Bank(build(wl.expr, Net), wl.V, raw=raw, salts=wl.salts)


def test_structural_parameters_are_checked_for_tapes_without_device_parameters():
    """`pass() >> delay(x)` lowers to a tape with a structural raw parameter but NO device parameters: per-voice delay times
    that differ used to be accepted by qg_bank_create and every voice silently rendered with the template's length"""
    from tests.graphs import pipe
    net = build(pipe("white()", "delay(0.001)"), Net)
    assert net.tape_info()["n_params"] == 0
    raw = np.full((4, 1), np.float32(0.001))
    Bank(net, 4, raw=raw)                              # equal everywhere: fine
    raw[2, 0] = np.float32(0.002)
    with pytest.raises(qb.QuartzGpuError, match="shapes the tape"):
        Bank(net, 4, raw=raw)


def test_single_process_sharded_render_matches_one_bank():
    """quartz is one process: render_sharded() drives one context + bank per device from host threads (no collective).
    With one GPU the device list repeats it — the sharding, row placement and thread safety are what is under test;
    with several GPUs every device is used."""
    import torch
    from quartz_b200 import workloads
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev >= 2 else [0, 0, 0]
    wl = workloads.c3_polysynth(V=96 * len(devices) + 32, T=3000, G=32)
    tmpl = build(wl.expr, Net)
    one = Bank(tmpl, wl.V, raw=wl.raw, salts=wl.salts).render(wl.T, group=wl.group)
    got = qb.render_sharded(tmpl, wl.V, wl.T, raw=wl.raw, salts=wl.salts, group=wl.group, devices=devices)
    assert got.shape == one.shape
    assert_parity(got[:, 0, :], one[:, 0, :], "exact", "sharded vs single bank")
    # per-voice rows of a tape that only the interpreters serve, uneven split
    wl2 = workloads.c5_mixed(V=4 * 1000, T=700)[3]
    t2 = build(wl2.expr, Net)
    one2 = Bank(t2, wl2.V, raw=wl2.raw, salts=wl2.salts).render(wl2.T)
    got2 = qb.render_sharded(t2, wl2.V, wl2.T, raw=wl2.raw, salts=wl2.salts, devices=devices)
    assert_parity(got2[:, 0, :], one2[:, 0, :], "exact", "sharded delay/lowpole voices")
