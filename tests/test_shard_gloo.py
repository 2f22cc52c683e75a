"""N>1 host logic on CPU: world_size-2 gloo processes shard a workload, agree on the partition, reduce a time with MAX
and gather disjoint output rows.  (The GPU path is the same code with backend nccl; no data-path collective exists.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quartz_b200 import shard, workloads
    from tests.graphs import build
    from tests.oracle_ffi import ONet, render_bank
    # strong scaling: 64 voices (groups of 32) split over 2 ranks
    wl = shard.shard_workload(workloads.c3_polysynth, world, rank, strong=True, V=64, T=256, G=32)
    lo, hi = shard.voice_range(64, world, rank, 32)
    assert wl.V == hi - lo == 32
    # each rank renders ITS voices with the oracle standing in for the GPU bank (CPU box), rows are gathered
    rows = render_bank([build(wl.voice_expr(v), ONet).set_salt(int(wl.salts[v])) for v in range(wl.V)], wl.T, group=32)
    allrows = shard.gather_rows(rows, dist)
    slow = shard.max_over_ranks(1.0 + rank, dist)
    # weak scaling: distinct voice ids per rank
    w = shard.shard_workload(workloads.c2_lowpass_bank, world, rank, V=8, T=16)
    q.put((rank, lo, hi, allrows.shape, float(np.abs(allrows).sum()), slow, w.salts[:2].tolist(), w.raw[:1].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_on_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=60) for _ in range(world))
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    (r0, lo0, hi0, shape0, sum0, slow0, salts0, raw0), (r1, lo1, hi1, shape1, sum1, slow1, salts1, raw1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 32, 32, 64)               # disjoint cover of the voices
    assert shape0 == shape1 == (2, 256) and sum0 == sum1          # both ranks see the same gathered rows
    assert slow0 == slow1 == 2.0                                 # max over ranks
    assert salts0 != salts1 and raw0 != raw1                     # weak scaling: differently seeded banks per rank


def test_voice_range_partitions():
    from quartz_b200.shard import voice_range
    for V, G, W in [(65536, 32, 8), (4096, 1, 3), (96, 32, 2), (1048576, 32, 8)]:
        spans = [voice_range(V, W, r, G) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == V
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all((hi - lo) % G == 0 for lo, hi in spans)


def test_bench_arms_share_config_keys_and_strong_shards_cover_the_bank():
    """the reference arm and the B200 arm describe the workload with the same `config` object (the driver compares them), and
    a strong split of configs[4] hands every voice to exactly one rank"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for name in ("c1", "c2", "c3", "c4", "c5"):
        wls = bench.make_workloads(name, 0)
        cfg = bench.config_of(name, wls)
        assert set(cfg) == {"workload", "voices", "samples", "sample_rate", "group", "layout", "note"}, name
        assert cfg["voices"] == sum(w.V for w in wls) and cfg["sample_rate"] == 48000
    assert bench.config_of("c3", bench.make_workloads("c3", 0))["workload"] == "c3_polysynth_65536"      # the north-star target
    total = sum(w.V for w in bench.make_workloads("c5", 0))
    for world in (2, 4, 8):
        seen = 0
        for rank in range(world):
            share = bench.make_workloads("c5", rank, world, strong=True)
            assert len(share) == 4 and all(w.V % w.group == 0 for w in share)
            seen += sum(w.V for w in share)
        assert seen == total == 1048576
        weak = bench.make_workloads("c3", 1, world, strong=False)
        assert weak[0].V == 65536                                   # weak scaling: a full-size bank per rank
