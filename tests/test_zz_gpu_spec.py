"""K1s — the tape-specialised lane kernel (qg_bank_set_path(QG_PATH_SPECIALISED), csrc/spec.cpp + spec_kernel.cuh): the tape
is compiled into the kernel with NVRTC, exec() is the same code as in every interpreter, so results must be BIT-IDENTICAL to
the sample-by-sample interpreter.  Opt-in in round 1 (AUTO never selects it); this file sorts after every other test file on purpose."""
import numpy as np
import pytest

import quartz_b200 as qb
from quartz_b200 import Bank, Net, workloads
from tests import cases
from tests.graphs import build

pytestmark = pytest.mark.gpu


def specialised(bank):
    try:
        return bank.set_path(qb.PATH_SPECIALISED)
    except qb.QuartzGpuError as e:
        if "NVRTC" in str(e) and "not found" in str(e):
            pytest.skip(str(e))
        raise


@pytest.mark.parametrize("idx", range(4), ids=["quantised_osc", "shift_reg", "feedback", "delay_lowpole"])
def test_specialised_kernel_is_bit_identical_on_the_mixed_bank_archetypes(idx):
    """the configuration measured in profiles/r01_spec_check.json (configs[4] archetypes, group mix 32)"""
    wl = workloads.c5_mixed(V=4 * 4096, T=2048)[idx]
    net = build(wl.expr, Net)
    ref = Bank(net, wl.V, raw=wl.raw, salts=wl.salts).set_path(qb.PATH_INTERP_SAMPLE)
    spec = specialised(Bank(net, wl.V, raw=wl.raw, salts=wl.salts))
    assert spec.kernel() == "k_spec"
    for group in (32, 1):
        ref.reset(); spec.reset()
        a, b = ref.render(wl.T, group=group), spec.render(wl.T, group=group)
        assert (a.view(np.uint32) == b.view(np.uint32)).all(), f"{wl.name} group={group}: max diff {np.abs(a - b).max():.3e}"
    # a second call continues from the persisted state exactly like the interpreter does
    a, b = ref.render(777), spec.render(777)
    assert (a.view(np.uint32) == b.view(np.uint32)).all()


def test_tapes_that_cannot_be_specialised_are_refused_by_name():
    net = build({"op": "kr()", "net": {"op": "white()"}, "n": 8}, Net)
    with pytest.raises(qb.QuartzGpuError, match="control flow"):
        Bank(net, 64).set_path(qb.PATH_SPECIALISED)


UNIFORM = [c for c in cases.RENDER if not any(k in repr(c[1]) for k in ("kr()", "s()", "reset", "select()", "seq()", "rfft", "ifft"))]


@pytest.mark.xfail(strict=False, reason="broad sweep of K1s written after the round's GPU minutes were spent: first run is the round-end one")
@pytest.mark.parametrize("name,expr,n,tol", UNIFORM, ids=[c[0] for c in UNIFORM])
def test_specialised_kernel_matches_the_interpreter_on_every_uniform_case(name, expr, n, tol):
    net = build(expr, Net)
    if net.inputs() != 0:
        pytest.skip("render cases only")
    V = 40
    salts = np.arange(1, V + 1, dtype=np.uint64)
    ref = Bank(net, V, salts=salts).set_path(qb.PATH_INTERP_SAMPLE)
    spec = specialised(Bank(net, V, salts=salts))
    a, b = ref.render(min(n, 2048)), spec.render(min(n, 2048))
    assert (a.view(np.uint32) == b.view(np.uint32)).all(), f"{name}: max diff {np.nanmax(np.abs(a - b)):.3e}"
