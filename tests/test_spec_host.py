"""Host side of the tape specialiser (csrc/spec.cpp): which tapes it accepts, and that the translation unit it generates
compiles for sm_100a with NVRTC (no GPU needed for that)."""
import os

import pytest

import quartz_b200 as qb
from quartz_b200 import Net, workloads
from tests.graphs import build, pipe

CSRC = os.path.join(os.path.dirname(qb.LIB_PATH), "csrc")


def nvrtc_compile(src):
    nvrtc = pytest.importorskip("cuda.bindings.nvrtc")
    err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"quartz_spec.cu", 0, [], [])
    opts = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"--fmad=false", ("-I" + CSRC).encode(), b"--ptxas-options=-v"]
    res, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    _, n = nvrtc.nvrtcGetProgramLogSize(prog)
    log = b" " * n
    nvrtc.nvrtcGetProgramLog(prog, log)
    assert res == nvrtc.nvrtcResult.NVRTC_SUCCESS, log.decode()[-3000:]
    _, n = nvrtc.nvrtcGetCUBINSize(prog)
    return n, log.decode()


def test_source_carries_the_tape_as_constants():
    net = build(pipe("white()", "lowpole(1000)"), Net)
    src = net.spec_source()
    info = net.tape_info()
    assert f"#define QG_SPEC_N {info['n_instr']}" in src and f"#define QG_SPEC_NT {info['n_temps']}" in src
    assert src.count("\n  {") == info["n_instr"] and '#include "interp.cu"' in src and '#include "spec_kernel.cuh"' in src


def test_only_uniform_tapes_are_specialised():
    for expr, word in (({"op": "kr()", "net": {"op": "white()"}, "n": 8}, "control flow"),
                       (pipe("white()", "rfft(64,0)"), "spectral")):
        with pytest.raises(qb.QuartzGpuError, match=word):
            build(expr, Net).spec_source()
    with pytest.raises(qb.QuartzGpuError, match="moog"):
        Net.str_to_net("moog(1000,0.5)").spec_source()


@pytest.mark.parametrize("idx", [1, 3], ids=["shift_reg", "delay_lowpole"])
def test_generated_unit_compiles_for_sm100a(idx):
    wl = workloads.c5_mixed(V=4 * 128, T=64)[idx]
    size, log = nvrtc_compile(build(wl.expr, Net).spec_source())
    assert size > 0
    assert "k_spec" in log and "0 bytes spill stores" in log, log[-800:]   # X really lives in registers


SAMPLED = ["product", "env_ar", "shift_reg_join8", "lowpass_q_var", "tap_noise", "feedback_delay", "quantize_13", "pan_var"]


@pytest.mark.parametrize("name", SAMPLED)
def test_specialised_units_of_diverse_render_cases_compile_without_spills(name):
    """all 83 uniform render cases were compiled this way once (<= 55 registers, no spills); the suite keeps a sample"""
    from tests import cases
    expr = next(c[1] for c in cases.RENDER if c[0] == name)
    size, log = nvrtc_compile(build(expr, Net).spec_source())
    assert size > 0 and "0 bytes spill stores" in log, log[-600:]


@pytest.mark.parametrize("N,J", [(2048, 4), (64, 2)])
def test_spectral_plan_units_compile_for_sm100a(N, J):
    """K5s: the frame-parallel spectral plan of configs[3] compiled into its kernels (spectral_kernel.cuh)"""
    wl = workloads.c4_spectral(V=4, T=1000, N=N, J=J)
    net = build(wl.expr, Net)
    info = net.spectral_info()
    assert info == {"n_segments": J, "n_streams": J, "n_instr": info["n_instr"], "round_len": N}
    src = net.spectral_spec_source()
    assert f"#define SP_NSEG {J}" in src and f"#define SP_C {N}" in src and '#include "spectral_kernel.cuh"' in src
    size, log = nvrtc_compile(src)
    assert size > 0 and "k_sp_frames" in log and "k_sp_post" in log
    assert "0 bytes spill stores" in log.split("Function properties for k_sp_frames")[1][:200], log[-1200:]   # X in registers
    with pytest.raises(qb.QuartzGpuError, match="spectral plan"):
        build(pipe("white()", "lowpole(100)"), Net).spectral_spec_source()
