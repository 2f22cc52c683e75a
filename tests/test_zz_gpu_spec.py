"""K1s — the tape-specialised lane kernel (qg_bank_set_path(QG_PATH_SPECIALISED), csrc/spec.cpp + spec_kernel.cuh): the tape
is compiled into the kernel with NVRTC, exec() is the same code as in every interpreter, so results must be BIT-IDENTICAL to
the sample-by-sample interpreter.  AUTO selects it for lane-interpreter banks once the bank's work pays for the compile, or at
once when the tape's kernel is already in the process-wide cache."""
import os
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


def test_auto_specialises_a_lane_bank_once_its_work_pays_for_the_compile_and_reuses_cached_kernels():
    """AUTO policy (capi.cu: bank_auto_specialise): below QG_SPEC_MIN_WORK voice-samples a bank stays on its interpreter; past
    it the tape is compiled into the kernel, the state carries over bit for bit, and a later bank of the same tape gets
    the cached kernel on its first render.  Taps / feedback rings keep the block-mode interpreter (dependent ring loads)."""
    wl = workloads.c5_mixed(V=4 * 4096, T=4096)[3]          # delay + lowpole: the prefetched delay line
    # a delay length no other test uses: kernels are cached process-wide BY TAPE, and in a whole-suite run an earlier test
    # has already compiled configs[4]'s own archetype — this bank would then start out specialised
    net = build({"op": "sr()", "n": 48000.0, "net": {"op": ">>", "n": 0.0, "inputs": [{"op": "white()"}, {"op": "delay(0.0201875)"}, {"op": "lowpole(1000)"}]}}, Net)
    wl.raw = np.stack([np.full(wl.V, np.float32(0.0201875)), wl.raw[:, 1]], axis=1)
    ref = Bank(net, wl.V, raw=wl.raw, salts=wl.salts).set_path(qb.PATH_INTERP_SAMPLE)
    os.environ["QG_SPEC_MIN_WORK"] = str(float(wl.V * 3000))
    try:
        bank = Bank(net, wl.V, raw=wl.raw, salts=wl.salts)
        assert bank.kernel() == "k_interp_blk"
        a0, b0 = ref.render(2000, group=32), bank.render(2000, group=32)      # 2000 < 3000: still the interpreter
        assert bank.kernel() == "k_interp_blk"
        a1, b1 = ref.render(2096, group=32), bank.render(2096, group=32)      # cumulative work passes the threshold
        if bank.kernel() != "k_spec":
            pytest.skip("NVRTC not available on this box: AUTO keeps the interpreter")
        a2, b2 = ref.render(1000, group=32), bank.render(1000, group=32)
        for x, y in ((a0, b0), (a1, b1), (a2, b2)):
            assert (x.view(np.uint32) == y.view(np.uint32)).all()
    finally:
        del os.environ["QG_SPEC_MIN_WORK"]
    again = Bank(net, wl.V, raw=wl.raw, salts=wl.salts)                        # default threshold (1e10): cache hit
    ref.reset()
    x, y = ref.render(500), again.render(500)
    assert again.kernel() == "k_spec"
    assert (x.view(np.uint32) == y.view(np.uint32)).all()
    tap = build({"op": ">>", "n": 0.0, "inputs": [{"op": "|", "n": 0.0, "inputs": [{"op": "white()"}, {"op": "dc(0.01)"}]}, {"op": "tap(0.001,0.05)"}]}, Net)
    os.environ["QG_SPEC_MIN_WORK"] = "1"
    try:
        tb = Bank(tap, 4096)
        tb.render(256)
        assert tb.kernel() != "k_spec", tb.kernel()
    finally:
        del os.environ["QG_SPEC_MIN_WORK"]
