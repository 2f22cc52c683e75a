"""Host-side logic of the product (no GPU): op-string parsing, arity rules, composition guards and structural
bookkeeping must agree with the reference's documented behaviour and with the oracle's independent restatement."""
import ctypes
import os
import re

import numpy as np
import pytest

import quartz_b200 as qb
from quartz_b200 import Net, _ffi
from tests import cases
from tests.graphs import L, build, pipe, stack
from tests.oracle_ffi import ONet, OracleUnsupported

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "quartz_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(qg_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(qb.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/quartz_gpu.h but not exported"
    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)


@pytest.mark.parametrize("op,ni,no", cases.ARITY, ids=[c[0] for c in cases.ARITY])
def test_str_to_net_arity(op, ni, no):
    n = Net.str_to_net(op)
    assert (n.inputs(), n.outputs()) == (ni, no)
    try:
        o = ONet.str_to_net(op)
    except OracleUnsupported:
        return
    assert (o.inputs(), o.outputs()) == (ni, no)


@pytest.mark.parametrize("name,expr,n,tol", cases.RENDER, ids=[c[0] for c in cases.RENDER])
def test_graph_shape_matches_oracle(name, expr, n, tol):
    a, b = build(expr, Net), build(expr, ONet)
    assert (a.inputs(), a.outputs(), a.size()) == (b.inputs(), b.outputs(), b.size())
    assert a.unsupported() is None
    info = a.tape_info()
    assert info["n_instr"] >= 0 and info["n_temps"] >= a.inputs()


def test_connective_guards_follow_process_rs():
    s, lp2 = Net.str_to_net("sine(440)"), Net.str_to_net("lowpass()")
    # >> requires outputs == inputs, otherwise the rhs is skipped (process.rs:1839)
    g = Net.connect(">>", [s, lp2])
    assert (g.inputs(), g.outputs(), g.size()) == (0, 1, 1)
    # + requires equal outputs (process.rs:1751)
    g = Net.connect("+", [s, Net.str_to_net("dc(1,2)")])
    assert g.outputs() == 1 and g.size() == 1
    # repeat count (process.rs:1744-1745)
    g = Net.connect("|", [s], number=4)
    assert (g.outputs(), g.size()) == (4, 4)
    # node limit (process.rs:1752-1754): growth stops once size() reaches the limit
    g = Net.connect("|", [s], number=50, node_limit=10)
    assert g.size() == 10
    # `-` needs both sides and equal outputs (process.rs:1788-1797)
    assert Net.connect("-", [s]).outputs() == 0
    d = Net.connect("-", [s, s])
    assert (d.outputs(), d.size()) == (1, 3)
    # ! passes missing outputs through (process.rs:1873)
    t = Net.connect("!", [Net.str_to_net("sink()")])
    assert (t.inputs(), t.outputs()) == (1, 1)
    # feedback needs outs == ins (process.rs:1500); reset needs 0-in/1-out (process.rs:1568)
    assert Net.feedback(s).outputs() == 0
    assert Net.reset_every(lp2, 1.0).outputs() == 0
    assert Net.reset_every(s, 1.0).outputs() == 1
    # seq/select keep only 0-in/1-out nets (process.rs:1638)
    sel = Net.select([s, lp2, s])
    assert (sel.inputs(), sel.outputs()) == (1, 1)


def test_unsupported_ops_fail_loudly_not_silently():
    for op in ("organ(220)", "reverb_stereo(10,2)", "moog(1000,0.5)", "pluck(220,0.5,0.5)"):
        n = Net.str_to_net(op)
        assert n.unsupported() is not None
        with pytest.raises(qb.QuartzGpuError):
            n.tape_info()
        # the flag survives composition
        g = Net.connect(">>", [n, Net.str_to_net("mul(0.5)")])
        assert g.unsupported() is not None


UNSUPPORTED_ARITY = [   # FunDSP's documented signatures, per parameter count as functions.rs dispatches them
    ("organ()", 1, 1), ("organ(220)", 0, 1), ("hammond()", 1, 1), ("pulse()", 2, 1), ("lorenz()", 1, 1), ("rossler()", 1, 1),
    ("dsf_saw()", 2, 1), ("dsf_saw(0.5)", 1, 1), ("dsf_square()", 2, 1), ("dsf_square(0.4)", 1, 1), ("mls()", 0, 1), ("mls(12)", 0, 1),
    ("pluck(220,0.5,0.5)", 1, 1), ("follow(0.1)", 1, 1), ("follow(0.01,0.2)", 1, 1),
    ("moog()", 3, 1), ("moog(0.5)", 2, 1), ("moog(1000,0.5)", 1, 1), ("lowrez(0.3)", 2, 1), ("bandrez(800,0.3)", 1, 1),
    ("morph()", 4, 1), ("morph(1000,1,0)", 1, 1), ("adsr(0.1,0.1,0.5,0.2)", 1, 1), ("meter(peak,0.1)", 1, 1), ("meter(rms,0.1)", 1, 1),
    ("chorus(0,0.015,0.005,0.2)", 1, 1), ("hold(0.5)", 2, 1), ("hold(200,0.5)", 1, 1), ("limiter(0.01,0.1)", 1, 1),
    ("limiter_stereo(0.01,0.1)", 2, 2), ("reverb_stereo(10)", 2, 2), ("reverb_stereo(10,2,0.5)", 2, 2), ("reverb_mono(10,2)", 1, 1),
    ("dissonance_max()", 1, 1), ("m_weight()", 1, 1), ("softexp()", 1, 1), ("spline_mono()", 5, 1), ("softmix()", 3, 1),
    ("spline_noise()", 2, 1), ("fractal_noise()", 4, 1),
]


@pytest.mark.parametrize("op,ni,no", UNSUPPORTED_ARITY, ids=[u[0] for u in UNSUPPORTED_ARITY])
def test_unsupported_ops_keep_the_reference_arity_and_mark_every_graph_they_touch(op, ni, no):
    """an op without a GPU lowering composes like the reference's unit would (same arity, one vertex) and the mark travels
    through every connective, array op and nested-net constructor: the patch fails by name instead of rendering without it"""
    n = Net.str_to_net(op)
    assert (n.inputs(), n.outputs(), n.size()) == (ni, no, 1)
    name = op.split("(")[0]
    assert n.unsupported() == name
    src = Net.str_to_net("dc(" + ",".join(["0.5"] * ni) + ")") if ni else None
    g = Net.connect(">>", [src, n]) if src is not None else n
    assert (g.inputs(), g.outputs()) == (0, no) and g.unsupported() == name
    # arity guards that SKIP the net still carry the mark (here: stacking is fine, piping into a 7-input net is not)
    for combo in (Net.connect("|", [Net.str_to_net("sine(220)"), g]), Net.connect(">>", [g, Net.str_to_net("join(7)")]),
                  Net.connect("+", [Net.str_to_net("dc(1,2,3)"), g]), Net.connect("!", [g]), Net.connect("-", [Net.str_to_net("dc(1,2,3)"), g])):
        assert combo.unsupported() == name
        with pytest.raises(qb.QuartzGpuError):
            combo.tape_info()
    if (g.inputs(), g.outputs()) == (0, 1):
        for wrapped in (Net.kr(g, 4), Net.reset_every(g, 0.1), Net.trig_reset(g), Net.select([Net.str_to_net("white()"), g]),
                        Net.seq([g, Net.str_to_net("white()")])):
            assert wrapped.unsupported() == name


def test_raw_parameters_and_signature():
    a = build(pipe("sine(220)", "lowpass(800,2)"), Net)
    b = build(pipe("sine(330)", "lowpass(1200,0.7)"), Net)
    c = build(pipe("sine(330)", "highpass(1200,0.7)"), Net)
    assert a.raw_params().tolist() == [220.0, 800.0, 2.0]
    assert a.signature() == b.signature() != c.signature()
    # delay time shapes the tape (ring length) -> part of the signature
    d1, d2 = Net.str_to_net("delay(0.01)"), Net.str_to_net("delay(0.02)")
    assert d1.signature() != d2.signature()


def test_clone_is_a_deep_copy():
    a = build(pipe("white()", "lowpass(800,2)"), Net)
    b = a.clone()
    b.set_sample_rate(48000)
    assert a.signature() != b.signature()


def test_device_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(qb.QuartzGpuError):
        Net.str_to_net("sine(440)").render(16)


@pytest.mark.parametrize("family,base", [("exact", 1000), ("float", 5000), ("control", 9000)])
def test_random_graphs_agree_with_oracle_on_shape(family, base):
    """the generator of tests/test_gpu_fuzz.py, host side only: parser, connective algebra and lowering accept every
    graph, and the product and the oracle agree on inputs / outputs / size"""
    import numpy as np
    from tests import test_gpu_fuzz as fz
    gen = {"exact": fz.exact_graph, "float": fz.float_graph, "control": fz.control_graph}[family]
    for seed in range(40):
        e = gen(np.random.default_rng(base + seed))
        a, b = build(e, Net), build(e, ONet)
        assert (a.inputs(), a.outputs(), a.size()) == (b.inputs(), b.outputs(), b.size()) and a.outputs() == 1, e
        assert a.unsupported() is None
        assert a.tape_info()["n_instr"] > 0


def test_rust_binding_declares_only_exported_entry_points():
    """rust/quartz-gpu cannot be compiled here (no cargo): at least every `extern "C"` name it binds must be declared in
    include/quartz_gpu.h and exported by the built library"""
    src = open(os.path.join(ROOT, "rust", "quartz-gpu", "src", "lib.rs")).read()
    block = src[src.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    names = re.findall(r"\bfn (qg_\w+)\(", block)
    assert len(names) >= 25
    header = open(os.path.join(ROOT, "include", "quartz_gpu.h")).read()
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    for n in names:
        assert re.search(r"\b%s\(" % n, header), n
        assert hasattr(lib, n), n
    declared = set(re.findall(r"\b(qg_[a-z_0-9]+)\s*\(", re.sub(r"/\*.*?\*/", "", header, flags=re.S)))
    assert declared == set(names), declared ^ set(names)      # the binding covers the whole header


def test_c_abi_tolerates_null_and_negative_arguments():
    """the reference never fails on this path (failures degrade to silence, functions.rs:1225): the C ABI answers bad
    pointers / sizes with a status or an empty result, never with a crash or an exception across the boundary"""
    lib = _ffi.lib()   # status codes: include/quartz_gpu.h (QG_OK 0, QG_ERR_ARG 1, QG_ERR_UNSUPPORTED 2)
    for fn in (lib.qg_quantize, lib.qg_get, lib.qg_wave):
        for arr, n in ((None, 5), (None, 0), (None, -3)):
            h = fn(arr, n)
            assert h, fn
            assert lib.qg_net_inputs(h) >= 0 and lib.qg_net_outputs(h) >= 0
            lib.qg_net_free(h)
    n = Net.str_to_net("lowpass(800,2)")
    assert lib.qg_net_raw_params(n.h, None, 8) == 2          # count only
    assert lib.qg_net_raw_count(None) == 0 and lib.qg_net_signature(None) == 0
    assert lib.qg_net_tape_info(None, None, None, None, None, None) == 1
    assert lib.qg_net_tape_info(n.h, None, None, None, None, None) == 0      # every out-pointer is optional
    assert lib.qg_net_tape_info(Net.str_to_net("moog(1000,0.5)").h, None, None, None, None, None) == 2
    assert b"moog" in lib.qg_last_error()
    assert lib.qg_bank_reset(None) == 1 and lib.qg_bank_set_raw(None, 0, 1.0) == 1
    assert not lib.qg_bank_create(None, n.h, 4, None, None)


def test_render_and_apply_circles_keep_the_reference_guards():
    """process.rs:1339-1353 / 1318-1326: guard paths never touch the device, so they are checked here"""
    keep = np.array([1.0, 2.0], dtype=np.float32)
    two_out, one_in = Net.str_to_net("dc(1,2)"), Net.str_to_net("lowpass(800,2)")
    assert qb.render_op(two_out, 100, arr=keep) is keep            # not 0-in / 1-out: the array is left alone
    assert qb.render_op(one_in, 100, arr=keep) is keep
    osc = Net.str_to_net("sine(440)")
    for number in (0, -5, float("nan"), 0.9):                      # `as usize` saturates: len 0 -> cleared array
        out = qb.render_op(osc, number, arr=keep)
        assert out is not keep and out.shape == (0,)
    assert qb.net.RENDER_LEN_CAP == 10_000_000
    assert qb.apply_op(one_in, [0.1, 0.2], arr=keep) is keep       # input length != net.inputs()
    assert qb.apply_op(Net.str_to_net("sink()"), [0.5], arr=keep).shape == (0,)


def test_pathological_numbers_neither_hang_nor_corrupt_the_tape():
    """found by fuzzing the C ABI on the CPU: a NaN delay time used to lower to a zero-length ring (now: one sample, like the
    oracle), and a repeat count of 1e30 used to spin through 2^31 no-op passes"""
    import time
    for op in ("delay(NaN)", "delay(-1)", "delay(0)"):
        a, b = Net.str_to_net(op), ONet.str_to_net(op)
        assert (a.inputs(), a.outputs()) == (b.inputs(), b.outputs()) == (1, 1)
        assert a.tape_info()["n_instr"] == 1
        assert b.process(np.arange(1, 6, dtype=np.float32)[:, None])[:, 0].tolist() == [0.0, 1.0, 2.0, 3.0, 4.0]   # one-sample delay
    assert Net.feedback(Net.str_to_net("mul(0.5)"), delay=float("nan")).tape_info()["n_instr"] > 0
    t0 = time.perf_counter()
    s, lp = Net.str_to_net("sine(440)"), Net.str_to_net("lowpass()")
    g = Net.connect(">>", [s, lp], number=1e30)              # arity guard skips `lowpass()`: every pass is a no-op
    assert (g.inputs(), g.outputs(), g.size()) == (0, 1, 1)
    g = Net.connect("|", [s], number=float("inf"), node_limit=10)
    assert g.size() == 10
    assert time.perf_counter() - t0 < 20.0      # was minutes; generous for a loaded CI host


def test_empty_operands_do_not_spin_the_repeat_count():
    """combining size-0, port-less operands (a failed str_to_net, an empty connective) succeeds without growing anything, so
    the node limit never stops the repeat loop: a circle Number of 1e30 used to run 2^31 passes on the caller's thread"""
    import time
    t0 = time.perf_counter()
    e = Net.str_to_net("no_such_op(1)")
    assert (e.inputs(), e.outputs(), e.size()) == (0, 0, 0)
    for op in ("|", ">>", "&", "^", "+", "*"):
        g = Net.connect(op, [e, e], number=1e30)
        assert (g.inputs(), g.outputs(), g.size()) == (0, 0, 0), op
    assert time.perf_counter() - t0 < 10.0


def test_a_tape_that_overflows_its_index_space_is_refused_not_aliased():
    """16-bit operand indices: a sum of 20,000 sines needs more temporaries than 15 bits address before the reuse pass runs;
    the lowering has to stop at the allocation, with QG_ERR_UNSUPPORTED, instead of wrapping onto earlier words"""
    s = Net.str_to_net("sine(440)")
    big = Net.connect("+", [s], number=20000, node_limit=1 << 16)
    assert big.size() == 39999          # 20,000 sines + 19,999 adders
    with pytest.raises(qb.QuartzGpuError, match="too large"):
        big.tape_info()
    ok = Net.connect("+", [s], number=300)
    assert ok.tape_info()["n_instr"] >= 300
