"""The in-tree nodes driven through qg_bank_process with the edge-case inputs of tests/test_oracle_intree.py (NaN / inf /
huge / negative indices and triggers, -0.0 triggers, zero durations) on every lane kernel, bit for bit against the oracle."""
import numpy as np
import pytest

import quartz_b200 as qb
from quartz_b200 import Bank, Net
from tests.graphs import build, pipe
from tests.oracle_ffi import ONet
from tests.test_oracle_intree import inputs

pytestmark = pytest.mark.gpu

RAMPS = [pipe("dc(300)", "ramp()"), pipe("dc(1100)", "ramp()"), pipe("dc(50)", "ramp()")]


def seq_inputs(n=4000):
    rng = np.random.default_rng(12)
    trig = (rng.uniform(size=n) < 0.01).astype(np.float32)
    trig[:3] = 1.0
    idx = rng.choice(np.array([0, 1, 2, 2.7, 5, -1, np.nan], dtype=np.float32), n)
    idx[:3] = [0, 1, 2]
    dur = rng.uniform(0.0, 0.02, n).astype(np.float32)
    dur[rng.integers(0, n, 150)] = 0.0
    return np.stack([trig, idx, rng.uniform(0.0, 0.003, n).astype(np.float32), dur], axis=1)


def select_inputs(n=3000):
    x = inputs(np.random.default_rng(11), n, 1)
    return np.where(np.isfinite(x), x / np.float32(10.0), x).astype(np.float32)


CASES = [
    ("shift_reg", {"op": "shift_reg()"}, lambda: inputs(np.random.default_rng(2), 3000, 2)),
    ("snh", {"op": "snh()"}, lambda: inputs(np.random.default_rng(3), 3000, 2)),
    ("get", {"op": "get()", "arr": [1.0, -2.0, 3.5, 4.0, 5.0, 6.0, 7.25]}, lambda: inputs(np.random.default_rng(5), 2000, 1)),
    ("samp_delay", {"op": "samp_delay(32)"}, lambda: inputs(np.random.default_rng(6), 3000, 2)),
    ("rise", {"op": "rise()"}, lambda: inputs(np.random.default_rng(7), 2000, 1)),
    ("trig_reset", {"op": "trig_reset()", "net": RAMPS[0]}, lambda: inputs(np.random.default_rng(8), 2000, 1)),
    ("reset_v", {"op": "reset_v()", "net": RAMPS[0]}, lambda: np.abs(inputs(np.random.default_rng(10), 2000, 1)) / np.float32(4000.0)),
    ("select", {"op": "select()", "inputs": RAMPS}, select_inputs),
    ("seq", {"op": "seq()", "inputs": RAMPS}, seq_inputs),
]


@pytest.mark.parametrize("name,expr,make", CASES, ids=[c[0] for c in CASES])
def test_in_tree_nodes_on_edge_case_inputs(name, expr, make):
    x = make()                                             # [n, inputs]
    n = x.shape[0]
    ref = build(expr, ONet).process(x)                     # [n, outputs]
    net = build(expr, Net)
    for pname, path in (("auto", qb.PATH_AUTO), ("lane_block", qb.PATH_INTERP), ("lane_sample", qb.PATH_INTERP_SAMPLE)):
        bank = Bank(net, 1).set_path(path)
        got = bank.process(np.ascontiguousarray(x.T[None]), n)[0].T        # voice-major [1, inputs, n] -> [n, outputs]
        same = (got.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(got) & np.isnan(ref))
        assert same.all(), f"{name} [{pname}: {bank.kernel()}] first mismatch at sample {int(np.argmin(same.all(axis=1)))}"
