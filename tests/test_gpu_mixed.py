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
