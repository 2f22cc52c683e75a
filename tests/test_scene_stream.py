"""Next rows of SURVEY.md section 8(f): scene (RON) importer and the block-rate streaming adapter."""
import json
import os

import numpy as np
import pytest

import quartz_b200 as qb
from quartz_b200 import Net
from quartz_b200.scene import Scene
from tests.graphs import L, add, build, pipe, stack
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


def _rust_i16(x):
    """dasp_sample 0.11 `f32 -> i16`: `(s * 32768.0) as i16` — Rust's float -> int cast truncates toward zero and saturates"""
    return np.clip(np.trunc(x.astype(np.float32) * np.float32(32768.0)), -32768, 32767).astype(np.int16)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [441, 1, 1000])
def test_stream_frames_in_the_device_sample_types(n):
    """`T::from_sample` (audio.rs:56-59, 115-116) for i16 / u16 streams; odd frame counts and mono graphs as well (the frame
    buffer of a mono graph with an odd n used to be 4-byte aligned under an 8-byte store)"""
    for expr in (pipe("sine(300)", "mul(1.7)"), stack(pipe("sine(201)", "mul(0.999)"), "dc(-1)"), stack("dc(1)", "dc(0.99998)")):
        banks = [qb.Bank(build(expr, Net), 1) for _ in range(3)]
        f = banks[0].render_stereo(n)
        i = banks[1].render_stereo(n, qb.SAMPLE_I16)
        u = banks[2].render_stereo(n, qb.SAMPLE_U16)
        assert f.shape == i.shape == u.shape == (n, 2) and i.dtype == np.int16 and u.dtype == np.uint16
        assert np.array_equal(i, _rust_i16(f))
        assert np.array_equal(u.astype(np.int32), i.astype(np.int32) + 32768)      # i16 -> u16 is offset binary
    assert int(_rust_i16(np.float32([1.0]))[0]) == 32767 and int(_rust_i16(np.float32([-1.0]))[0]) == -32768
    with pytest.raises(qb.QuartzGpuError, match="sample format"):
        banks[0].render_stereo(4, 7)


@pytest.mark.gpu
@pytest.mark.parametrize("path", [qb.PATH_AUTO, qb.PATH_INTERP, qb.PATH_INTERP_SAMPLE, qb.PATH_TV])
def test_a_cloned_bank_carries_its_state_and_both_continue_like_the_oracle(path):
    """`Net::clone` copies state (process.rs:1316, 1336, 1499, 1558, 1895): clone a running bank mid-render; the original
    and the copy both produce exactly what the uninterrupted render produces, independently of each other"""
    V = 48
    salts = np.arange(7, 7 + V, dtype=np.uint64)
    expr = pipe(add("white()", pipe("sine(330)", "delay(0.003)")), "lowpass(1200,2)", "dcblock()")
    a = qb.Bank(build(expr, Net), V, salts=salts).set_path(path)
    first = a.render(777)
    b = a.clone()
    assert b.kernel() == a.kernel()
    ra1, rb1 = a.render(500), b.render(500)
    _ = a.render(123)                       # advancing one does not move the other
    rb2 = b.render(300)
    whole = qb.Bank(build(expr, Net), V, salts=salts).set_path(path).render(777 + 500 + 300)
    assert np.array_equal(ra1, rb1)
    # (block-level scans re-associate per hop, and hops follow the call boundaries: chunked vs whole is a float comparison)
    assert_parity(np.concatenate([first, rb1, rb2], axis=2), whole, "float", "clone, chunked")
    ref = np.stack([build(expr, ONet).set_salt(int(s)).render(1577).T for s in salts])
    assert_parity(whole, ref, "float", "clone")
    b.reset()                               # reset of the clone restores ITS initial state (same init table)
    assert np.array_equal(b.render(777), first)


@pytest.mark.gpu
def test_a_cloned_fused_bank_continues_identically():
    V = 256
    rng = np.random.default_rng(5)
    raw = np.stack([np.exp(rng.uniform(np.log(60), np.log(9000), V)), rng.uniform(0.5, 6, V)], axis=1).astype(np.float32)
    tmpl = build(pipe("white()", "lowpass(1000,1)"), Net)
    a = qb.Bank(tmpl, V, raw=raw, salts=np.arange(V, dtype=np.uint64))
    assert a.kernel() == "k_noise_svf_scan"
    a.render(4096)
    b = a.clone()
    assert np.array_equal(a.render(4096), b.render(4096))
