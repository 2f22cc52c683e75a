"""Next rows of SURVEY.md section 8(f): scene (RON) importer and the block-rate streaming adapter."""
import json
import os

import numpy as np
import pytest

import quartz_b200 as qb
from quartz_b200 import Net
from quartz_b200.scene import Scene
from tests.graphs import L, build, pipe, stack
from tests.oracle_ffi import ONet
from tests.util import assert_parity

HERE = os.path.dirname(os.path.abspath(__file__))
GRAPHS = json.load(open(os.path.join(HERE, "golden", "scenes", "out_graphs.json")))

MINI_SCENE = '''(
  resources: {},
  entities: {
    1: (
      components: {
        "quartz::components::Order": (0),
        "quartz::components::Number": (0.0),
        "quartz::components::Op": ("sine(220)"),
        "quartz::components::Holes": ([10,]),
        "quartz::components::Arr": ([]),
      },
    ),
    2: (
      components: {
        "quartz::components::Order": (0),
        "quartz::components::Number": (0.0),
        "quartz::components::Op": ("mul(0.5)"),
        "quartz::components::Holes": ([12,]),
        "quartz::components::Arr": ([]),
      },
    ),
    3: (
      components: {
        "quartz::components::Order": (1),
        "quartz::components::Number": (0.0),
        "quartz::components::Op": (">>"),
        "quartz::components::Holes": ([11,13,14,]),
        "quartz::components::Arr": ([]),
      },
    ),
    4: (
      components: {
        "quartz::components::Order": (2),
        "quartz::components::Number": (0.0),
        "quartz::components::Op": ("out()"),
        "quartz::components::Holes": ([15,]),
        "quartz::components::Arr": ([]),
      },
    ),
    10: (
      components: {
        "quartz::components::BlackHole": (wh: 11, wh_parent: 3,),
      },
    ),
    11: (
      components: {
        "quartz::components::WhiteHole": (bh: 10, bh_parent: 1, link_types: (0, 1), open: false,),
      },
    ),
    12: (
      components: {
        "quartz::components::BlackHole": (wh: 13, wh_parent: 3,),
      },
    ),
    13: (
      components: {
        "quartz::components::WhiteHole": (bh: 12, bh_parent: 2, link_types: (0, 2), open: false,),
      },
    ),
    14: (
      components: {
        "quartz::components::BlackHole": (wh: 15, wh_parent: 4,),
      },
    ),
    15: (
      components: {
        "quartz::components::WhiteHole": (bh: 14, bh_parent: 3, link_types: (0, 1), open: false,),
      },
    ),
  },
)'''


def test_importer_on_a_minimal_scene():
    sc = Scene(MINI_SCENE)
    (out,) = sc.find("out()")
    e = sc.expr(out)
    assert e == {"op": ">>", "n": 0.0, "inputs": [{"op": "sine(220)"}, {"op": "mul(0.5)"}]}
    n = build(e, Net)
    assert (n.inputs(), n.outputs(), n.size()) == (0, 1, 2)


@pytest.mark.parametrize("name", sorted(GRAPHS))
def test_imported_scene_graphs_agree_with_oracle_on_shape(name):
    a, b = build(GRAPHS[name], Net), build(GRAPHS[name], ONet)
    assert (a.inputs(), a.outputs(), a.size()) == (b.inputs(), b.outputs(), b.size())
    assert a.unsupported() is None and a.inputs() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GRAPHS))
def test_imported_scenes_render_like_the_oracle(name):
    n = 6000
    got = build(GRAPHS[name], Net).render(n)
    ref = build(GRAPHS[name], ONet).render(n)
    assert_parity(got, ref, "float", name)


@pytest.mark.gpu
def test_stream_adapter_matches_audio_rs_semantics():
    from quartz_b200.stream import StreamAdapter
    # a graph that exceeds [-1, 1] and emits a non-normal value: sanitise + clamp like audio.rs:88-92
    expr = pipe("sine(300)", "mul(3)")
    ad = StreamAdapter(build(expr, Net), block=1000)
    frames = np.concatenate([ad.read(700), ad.read(1), ad.read(2299)])
    ref = build(expr, ONet).render(3000)[:, 0]
    ref = np.where(np.isfinite(ref) & (np.abs(ref) >= np.finfo(np.float32).tiny), np.clip(ref, -1, 1), 0).astype(np.float32)
    assert_parity(frames[:, 0], ref, "float", "left")
    assert (frames[:, 1] == 0).all()                       # mono -> net | dc(0)   (process.rs:1897)
    # stereo graphs pass both channels, anything else plays silence (process.rs:1899-1904)
    st = StreamAdapter(build(stack("dc(0.25)", "dc(-2)"), Net), block=64).read(10)
    assert (st[:, 0] == 0.25).all() and (st[:, 1] == -1.0).all()
    assert (StreamAdapter(Net.str_to_net("lowpass(1000,1)"), block=64).read(10) == 0).all()
    # 1e-40 is subnormal -> 0 ; NaN -> 0
    z = StreamAdapter(build(stack("dc(1e-40)", pipe("dc(-1)", "sqrt()")), Net), block=32).read(5)
    assert (z == 0).all()


@pytest.mark.gpu
def test_var_updates_apply_at_block_boundaries():
    from quartz_b200.stream import StreamAdapter
    g = pipe({"op": "var()", "n": 0.25}, "mul(2)")
    ad = StreamAdapter(build(g, Net), block=100)
    a = ad.read(100)
    ad.set_var(0, -0.125)
    b = ad.read(100)
    assert (a[:, 0] == 0.5).all() and (b[:, 0] == -0.25).all()
