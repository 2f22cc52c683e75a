"""Host side of the frame-parallel spectral path (csrc/spectral.cu: plan_spectral): which tapes it takes.  A tape qualifies
when everything that feeds an rfft is a pure function of time, the ops between an rfft and its ifft are stateless maps of
that rfft's two outputs, and what follows the iffts is stateless or again a pure function of time."""
import json
import os

import pytest

import quartz_b200 as qb
from quartz_b200 import Net, workloads
from tests.graphs import add, build, mul, pipe, stack

HERE = os.path.dirname(os.path.abspath(__file__))
SCENES = json.load(open(os.path.join(HERE, "golden", "scenes", "out_graphs.json")))


def plan(expr):
    return build(expr, Net).spectral_info()


def seg(src, chain=(), n=64, start=0, tail=("chan(1,0)",)):
    return pipe(src, f"rfft({n},{start})", *chain, f"ifft({n},{start})", *tail)


@pytest.mark.parametrize("src", [
    "white()", {"op": "wave()", "arr": [0.0, 0.5, -0.25]}, "impulse()", "dc(0.3)",
    pipe("white()", "delay(0.002)"), pipe("white()", "tick()", "tanh()"),
    mul("white()", pipe({"op": "wave()", "arr": [0.1] * 16}, "delay(0.0003)")),
], ids=["noise", "wave", "impulse", "constant", "delayed_noise", "tick_tanh", "windowed"])
def test_pure_functions_of_time_may_feed_an_rfft(src):
    info = plan(seg(src))
    assert info is not None and info["n_segments"] == 1 and info["n_streams"] == 1 and info["round_len"] == 64


@pytest.mark.parametrize("src", ["sine(440)", pipe("white()", "lowpass(800,1)"), pipe("dc(100)", "ramp()"), "pink()",
                                 pipe("white()", {"op": "feedback()", "net": {"op": "mul(0.5)"}, "delay": None})],
                         ids=["sine", "filtered", "ramp", "pink", "feedback"])
def test_recurrences_in_front_of_an_rfft_keep_the_general_path(src):
    assert plan(seg(src)) is None


def test_bin_chains_must_be_stateless_maps_of_their_own_rfft():
    assert plan(seg("white()", ["pol()", "car()"])) is not None
    assert plan(seg("white()", [stack("lowpole(100)", "pass()")])) is None                       # a filter across bins
    assert plan(seg("white()", [stack("tick()", "pass()")])) is None                             # one-bin delay
    two = pipe(stack(pipe("white()", "rfft(64,0)"), pipe("white()", "rfft(64,0)")),
               stack(add("pass()", "pass()"), "pass()", "sink()"))                               # mixes two rffts' bins
    assert plan(pipe(two, "ifft(64,0)", "chan(1,0)")) is None
    assert SCENES["spectral-delay"] and plan(SCENES["spectral-delay"]) is None                   # tap() between rfft and ifft
    assert plan(SCENES["spectral-gate"])["n_segments"] == 4


def test_rfft_and_ifft_must_agree_and_bins_must_not_escape():
    assert plan(pipe("white()", "rfft(64,0)", "ifft(32,0)", "chan(1,0)")) is None                # sizes differ
    assert plan(pipe("white()", "rfft(64,0)", "ifft(64,8)", "chan(1,0)")) is None                # start offsets differ
    assert plan(pipe("white()", "rfft(64,0)", "chan(1,0)")) is None                              # raw bins reach the output
    assert plan(pipe("white()", "rfft(64,5)", "ifft(64,5)", "chan(1,0)"))["round_len"] == 64


def test_what_follows_the_ifft_must_be_stateless_or_a_function_of_time():
    assert plan(seg("white()", tail=("chan(1,0)", "tanh()", "mul(0.5)"))) is not None
    assert plan(mul(seg("white()"), pipe({"op": "wave()", "arr": [0.5] * 8}, "delay(0.001)"))) is not None
    assert plan(seg("white()", tail=("chan(1,0)", "lowpole(500)"))) is None                      # a filter after resynthesis
    assert plan(seg("white()", tail=("chan(1,0)", "delay(0.001)"))) is None                      # a delay of the resynthesis
    assert plan(pipe(seg("white()"), "rfft(64,0)", "ifft(64,0)", "chan(1,0)")) is None           # a second analysis stage


def test_streams_only_for_the_components_the_post_graph_reads():
    assert plan(seg("white()", tail=("chan(1,0)",)))["n_streams"] == 1                           # real part only
    assert plan(seg("white()", tail=("join(2)",)))["n_streams"] == 2                             # both parts
    wl = workloads.c4_spectral(V=2, T=100)
    info = plan(wl.expr)
    assert (info["n_segments"], info["n_streams"], info["round_len"]) == (4, 4, 2048)
    mixed = add(seg("white()", n=64), seg("white()", n=256, start=64))
    assert plan(mixed)["round_len"] == 256                                                       # a round = the largest size


def test_graphs_with_inputs_stay_on_the_block_path():
    net = build(workloads.spectral_graph(64, 2, 1.0, workloads.hann(64), source="pass()"), Net)
    assert net.inputs() == 1 and net.spectral_info() is None
    with pytest.raises(qb.QuartzGpuError, match="spectral plan"):
        net.spectral_spec_source()
