#!/usr/bin/env python3
"""bench.py — voice-samples/s @ 48 kHz for quartz's audio-graph hot path on N B200s (one process per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c5|c1] [--only] [--impl reference]

A "step" is one offline render of the workload (all voices x all samples).  The headline workload is BASELINE.json's
north-star target, configs[2]: 65,536 osc -> lowpass -> envelope voices, 10 s at 48 kHz, mixed in groups of 32.
Prints ONE JSON line (contract in the task statement):
  value     kernel throughput with every input resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same job through the public API with host buffers: bank build from host parameter tables + render +
            device->host copy of every output sample
  roofline  algorithmic bytes or flops / kernel time against the measured peak that binds the workload
  cpu_baseline  the CPU oracle (a restatement of the reference — it is Rust and cannot be built here) on the host cores
  configs   the same four sub-records for the other BASELINE configs (c1, c2, c4, c5), each timed in this run
With --gpus N the headline is weak scaling of configs[2] (every rank a full-size, differently seeded bank) and
`configs.c5` is the 1,048,576-voice mixed graph of configs[4] SHARDED over the N ranks (strong scaling); no data-path
collective exists — NCCL carries the barrier and the max-over-ranks time only.
`--impl reference` times the CPU oracle on a bounded sample of the headline workload (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "voice-samples/s @48 kHz"
UNIT = "voice-samples/s"
HEADLINE = "c3"
SUB_CONFIGS = ["c1", "c2", "c4", "c5"]   # timed in the same run, reported under "configs"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def config_of(name, wls):
    """the `config` object: identical keys for the B200 arm and the reference arm"""
    wl = wls[0]
    return {"workload": wl.name if len(wls) == 1 else name + ":" + "+".join(w.name for w in wls),
            "voices": int(sum(w.V for w in wls)), "samples": int(wl.T), "sample_rate": 48000, "group": int(wl.group),
            "layout": "voice-major [V/G][T] f32", "note": wl.note if len(wls) == 1 else "; ".join(f"{w.name}: {w.note}" for w in wls)}


def cpu_baseline(wls, seconds_target=12.0, threads=None):
    """The oracle (port of the reference's tick loop) on a bounded sample of the same workload.  Multi-bank workloads
    (c5: four voice archetypes) are sampled bank by bank with equal shares of the time target and combined as total
    units / total extrapolated time."""
    from quartz_b200.graphs import build
    from tests.oracle_ffi import ONet, render_bank
    threads = threads or len(os.sched_getaffinity(0)) or 1    # the cores this rank may use (a multi-rank run binds ranks to NUMA nodes)
    per_bank = []
    for wl in wls:
        per_bank.append(_cpu_sample(wl, seconds_target / len(wls), threads, build, ONet, render_bank))
    units = sum(w.V * w.T for w in wls)
    full_s = sum(w.V * w.T / pb["value"] for w, pb in zip(wls, per_bank))
    one_s = sum(w.V * w.T / pb["one_core_value"] for w, pb in zip(wls, per_bank))
    return {"value": units / full_s, "unit": UNIT, "cores": threads, "kind": "port", "one_core_value": units / one_s,
            "full_job_s": full_s, "sample": " | ".join(pb["sample"] for pb in per_bank)}


def _cpu_sample(wl, seconds_target, threads, build, ONet, render_bank):
    quantum = wl.group * threads if wl.V >= wl.group * threads else wl.group
    # calibrate on a short run, then size the sample for ~seconds_target of CPU work at the workload's full T
    nv0 = min(wl.V, quantum)
    cal = [build(wl.voice_expr(v), ONet).set_salt(int(wl.salts[v])) for v in range(nv0)]
    Tc = min(wl.T, 20000)
    render_bank(cal, min(Tc, 2000), group=wl.group, threads=threads)     # thread start-up, page faults
    for _ in range(3):                                                     # a calibration run of at least ~0.3 s
        t0 = time.perf_counter()
        render_bank(cal, Tc, group=wl.group, threads=threads)
        el = time.perf_counter() - t0
        rate = nv0 * Tc / el
        if el >= 0.3 or Tc >= wl.T:
            break
        Tc = int(min(wl.T, Tc * max(2.0, 0.5 / el)))
    T = wl.T
    nv = int(rate * seconds_target / T) // quantum * quantum
    nv = max(nv0, min(nv, wl.V, int(8e9 / (4 * T)) * wl.group // quantum * quantum))   # <= 8 GB of oracle output
    if nv <= nv0:
        nv = nv0
        T = int(min(wl.T, max(Tc, rate * seconds_target / nv)))
    onets = [build(wl.voice_expr(v), ONet).set_salt(int(wl.salts[v])) for v in range(nv)]
    out = np.empty((nv // wl.group, T), dtype=np.float32)
    out.fill(0.0)                       # first touch outside the timed region: the baseline is not charged for page faults
    # the output buffer caps the sample's size: repeat the render (the graphs simply keep running) up to the time target
    reps = int(max(1, min(8, round(seconds_target * rate / (nv * T)))))
    t0 = time.perf_counter()
    for _ in range(reps):
        render_bank(onets, T, group=wl.group, threads=threads, out=out)
    dt = (time.perf_counter() - t0) / reps
    del out
    # the reference itself evaluates one graph on ONE thread (audio.rs:95-100, process.rs:1347-1351): the faithful per-core figure
    n1 = max(wl.group, min(nv, 2 * wl.group))
    T1 = int(min(T, max(2000, rate / max(1, threads) * 1.0 / n1)))
    one = [build(wl.voice_expr(v), ONet).set_salt(int(wl.salts[v])) for v in range(n1)]
    out = np.empty((n1 // wl.group, T1), dtype=np.float32)
    out.fill(0.0)
    t1 = time.perf_counter()
    render_bank(one, T1, group=wl.group, threads=1, out=out)
    one_core = n1 * T1 / (time.perf_counter() - t1)
    return {"value": nv * T / dt, "one_core_value": one_core,
            "sample": f"{nv} voices x {T} samples x {reps} passes of {wl.name} (oracle/, {threads} threads, {dt * reps:.1f} s); "
                      f"one core: {n1} voices x {T1} samples"}


def make_workloads(name, rank, world=1, strong=False):
    from quartz_b200 import shard, workloads
    fake = int(os.environ.get("QG_BENCH_FAKE_WORLD", "0"))    # experiment hook: rank 0's share of a `fake`-rank strong split, on one GPU
    if fake > 1:
        world, rank, strong = fake, 0, True
    w = shard.shard_workload(workloads.WORKLOADS[name], world, rank, strong=strong)
    return w if isinstance(w, list) else [w]


def run_reference(args, rank):
    """the reference's CPU path (oracle port: no Rust toolchain here) on all host threads, bounded sample per step"""
    if rank != 0:
        return
    wls = make_workloads(args.workload, 0)
    vals, cb = [], None
    per_step = min(15.0, max(2.0, 90.0 / max(1, args.warmup + args.steps)))
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(wls, seconds_target=per_step)
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals))
    units = sum(w.V * w.T for w in wls)
    cfg = config_of(args.workload, wls)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            # one full job at the sampled rate (the sample is a bounded share of the workload's voices at its full length)
            "ms_per_step": units / v * 1e3, "ms_per_step_is": "extrapolated from the bounded sample: voices x samples / value",
            "sampled": True, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "reference_note": "the reference is Rust (no toolchain in this image): CPU oracle port (oracle/), all host threads; the "
                              "reference itself evaluates one graph on one thread (cpu_baseline.one_core_value)",
            "cpu_baseline": dict(cb, value=v),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "sample": cb["sample"]}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index):
    """Run this rank (and first-touch its pinned host buffers) on the CPUs of the GPU's NUMA node: with 8 ranks per box the
    device->host ingest is the end-to-end limiter, and a buffer on the far socket halves what a rank gets.  Returns a note."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = None
        if all(hasattr(pr, k) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bdf = f"{int(pr.pci_domain_id):04x}:{int(pr.pci_bus_id):02x}:{int(pr.pci_device_id):02x}.0"
        if bdf is None or not os.path.exists(f"/sys/bus/pci/devices/{bdf}"):
            out = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=20).stdout.strip()
            bdf = out[-12:].lower() if len(out) >= 12 else None          # 00000000:1B:00.0 -> 0000:1b:00.0
        if not bdf:
            return "numa: bus id unknown"
        node = int(open(f"/sys/bus/pci/devices/{bdf.lower()}/numa_node").read())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) < 8:        # a container that exposes only a few of the node's cores: leave the rank where it is
            return f"numa: node {node} offers {len(cpus)} usable cpus, not bound"
        os.sched_setaffinity(0, cpus)
        return f"numa: bound to node {node} ({len(cpus)} cpus)"
    except Exception as e:   # the binding is an optimisation, never a requirement
        return f"numa: not bound ({type(e).__name__})"


class Runner:
    """one GPU context + the timing protocol shared by the headline and the sub-configs"""

    def __init__(self, args, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        import quartz_b200 as qb
        self.torch, self.dist, self.qb = torch, dist, qb
        self.args, self.rank, self.local_rank, self.world = args, rank, local_rank, world
        # the library launches on the stream it is given; use a real (non-legacy) torch stream so that torch CUDA
        # events bracket exactly those launches
        self.numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "numa: one rank, not bound"
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        self.ctx = qb.Context(local_rank, stream=self.stream.cuda_stream)
        self.hbm_peak, self.hbm_src = _peaks()
        self._fp32 = None

    def fp32_peak(self):
        if self._fp32 is None:
            self._fp32 = self.ctx.measure_fp32_tflops()
        return self._fp32

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def specialise(self, b, w):
        """--path specialised: K1s for every lane-interpreter bank whose tape allows it"""
        qb = self.qb
        if b.kernel() not in ("k_interp_blk", "k_interp<uniform>"):
            return                                 # fused / time-vector banks keep their kernel
        try:
            b.set_path(qb.PATH_SPECIALISED)
        except qb.QuartzGpuError:
            pass                                   # tape not specialisable / NVRTC missing: the bank keeps its AUTO kernel

    def set_paths(self, banks, wls):
        if self.args.path == "interp":
            [b.set_path(self.qb.PATH_INTERP) for b in banks]
        if self.args.path == "specialised":
            [self.specialise(b, w) for b, w in zip(banks, wls)]

    def measure(self, name, wls, steps, warmup, do_e2e, do_cpu, clocks=False):
        """Time `steps` renders of the workload (a list of banks run back to back).  Returns the record."""
        torch, qb, ctx, world = self.torch, self.qb, self.ctx, self.world
        from quartz_b200.graphs import build as build_graph
        tmpls = [build_graph(w.expr, qb.Net) for w in wls]
        # A sharded multi-bank workload (configs[4] over N ranks) leaves each bank too few voices to fill a GPU: V/N voices =
        # V/(32 N) warps of a latency-bound lane kernel.  The banks are independent, so each gets its own context = its own
        # stream and they render concurrently.  (At N = 1 every bank fills the machine: one stream, per-bank times add up.)
        concurrent = len(wls) > 1 and (world > 1 or int(os.environ.get("QG_BENCH_FAKE_WORLD", "0")) > 1)
        if os.environ.get("QG_BENCH_CONCURRENT") in ("0", "1"):
            concurrent = len(wls) > 1 and os.environ["QG_BENCH_CONCURRENT"] == "1"
        if concurrent:
            streams = [torch.cuda.Stream() for _ in wls]
            ctxs = [qb.Context(self.local_rank, stream=st.cuda_stream) for st in streams]
        else:
            streams, ctxs = [self.stream] * len(wls), [ctx] * len(wls)
        banks = [qb.Bank(t, w.V, raw=w.raw, salts=w.salts, ctx=cx) for t, w, cx in zip(tmpls, wls, ctxs)]
        self.set_paths(banks, wls)
        rows_l = [(w.V // w.group) * t.outputs() for t, w in zip(tmpls, wls)]
        rows = sum(rows_l)
        T = wls[0].T
        out_bytes = sum(r * w.T * 4 for r, w in zip(rows_l, wls))
        d_out = torch.empty(out_bytes // 4, dtype=torch.float32, device="cuda")
        offs = np.concatenate([[0], np.cumsum([r * w.T * 4 for r, w in zip(rows_l, wls)])])[:-1]
        nb = len(banks)

        def step(evs=None):
            if concurrent:
                fork = torch.cuda.Event()
                fork.record(self.stream)
                joins = []
                for k, (b, w, o, st) in enumerate(zip(banks, wls, offs, streams)):
                    st.wait_event(fork)
                    b.reset()
                    if evs:
                        evs[2 * k].record(st)
                    b.render_device(w.T, d_out.data_ptr() + int(o), group=w.group)
                    if evs:
                        evs[2 * k + 1].record(st)
                    j = torch.cuda.Event()
                    j.record(st)
                    joins.append(j)
                for j in joins:
                    self.stream.wait_event(j)
                return
            for k, (b, w, o) in enumerate(zip(banks, wls, offs)):
                b.reset()
                if evs:
                    evs[k].record()
                b.render_device(w.T, d_out.data_ptr() + int(o), group=w.group)
            if evs:
                evs[nb].record()

        for _ in range(warmup):
            step()
        # AUTO compiles a bank's tape / plan into its kernel once the bank's accumulated work pays for it (K1s, K5s): a small
        # per-rank share (strong scaling at N = 8) reaches that point only after several renders.  Keep warming up until the
        # kernel selection has been stable for one whole step, so that the timed region measures the steady state.
        per_step = min(w.V * w.T for w in wls)
        for _ in range(max(0, min(64, int(np.ceil(1.0e10 / per_step)) + 1 - warmup))):   # 1e10: the library's default K1s threshold
            step()
        names = [b.kernel() for b in banks]
        for _ in range(8):
            step()
            now = [b.kernel() for b in banks]
            stable = now == names
            names = now
            if stable:
                break
        self.barrier()
        sampler = None
        if clocks:
            sampler = ClockSampler(self.local_rank)
            sampler.start()
            time.sleep(0.25)
        l0 = sum(cx.launch_count() for cx in set(ctxs))
        # per step: one event before each bank's render and one after the last (reset copies sit before their bank's event);
        # concurrent banks: a start / end pair per bank on its own stream
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2 * nb if concurrent else nb + 1)] for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for ev in evs:
            step(ev)
        e1.record()
        self.barrier()
        launches = sum(cx.launch_count() for cx in set(ctxs)) - l0
        total_ms = self.max_ranks(e0.elapsed_time(e1))
        if concurrent:   # each bank's own span, measured while the others run beside it
            bank_ms = [float(np.mean([ev[2 * k].elapsed_time(ev[2 * k + 1]) for ev in evs])) for k in range(nb)]
            kern_ms = total_ms / steps
        else:
            bank_ms = [float(np.mean([ev[k].elapsed_time(ev[k + 1]) for ev in evs])) for k in range(nb)]
            # a bank's share = its render + the next bank's state reset copy (a device-to-device memcpy of the state table)
            kern_ms = float(np.mean([ev[0].elapsed_time(ev[nb]) for ev in evs]))
        if sampler:
            time.sleep(0.15)
            sampler.stop()
        ms_per_step = total_ms / steps
        units_local = sum(w.V * w.T for w in wls)
        units = self.sum_ranks(units_local)
        rec = {"value": units / (ms_per_step * 1e-3), "unit": UNIT, "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
               "gpu_launches": int(launches), "config": dict(config_of(name, wls), kernel="+".join(sorted({b.kernel() for b in banks})),
                                                             l2=f"each step writes {out_bytes / 1e9:.2f} GB of output per rank (>> 126 MB L2), state re-initialised per step"),
               "roofline": self.roofline(name, wls, banks, bank_ms, kern_ms, concurrent)}
        if concurrent:
            rec["config"]["streams"] = f"{nb} banks rendered concurrently, one context and stream each (per-bank times overlap)"
        if sampler:
            rec["clocks"] = sampler.summary()
        del d_out
        torch.cuda.empty_cache()
        if do_e2e:
            rec["e2e"] = self.e2e(wls, tmpls, rows_l, out_bytes, min(steps, 3))
        del banks
        if concurrent:
            torch.cuda.synchronize()
            for cx in ctxs:
                cx.close()
        if do_cpu and self.rank == 0:
            rec["cpu_baseline"] = cpu_baseline(wls, seconds_target=self.args.cpu_seconds)
        return rec

    def roofline(self, name, wls, banks, bank_ms, kern_ms, concurrent=False):
        """Per bank: the bound is max(bytes / HBM peak, flops / FP32 peak) — the resource that binds that kernel; frac is
        that minimum time over the measured time.  The record's headline figures are those of the dominant kernel (the bank
        that takes the most time); `frac_all` weighs every bank (sum of minimum times / sum of measured times)."""
        hbm = self.hbm_peak
        per = []
        for w, b, ms in zip(wls, banks, bank_ms):
            by, fl = w.V * w.T * w.bytes_per_unit, w.V * w.T * w.flops_per_unit
            t_hbm = by / (hbm * 1e9)
            if w.bound == "latency":
                per.append({"workload": w.name, "kernel": b.kernel(), "kernel_ms": ms, "bound": "latency", "frac": None,
                            "note": "one voice: no throughput roofline applies (SURVEY.md 8d); wall time only"})
                continue
            t_fp = fl / (self.fp32_peak() * 1e12) if fl > 0 else 0.0
            bound = "hbm" if t_hbm >= t_fp else "fp32"
            r = {"workload": w.name, "kernel": b.kernel(), "kernel_ms": ms, "bound": bound,
                 "algorithmic_bytes_per_launch": by, "algorithmic_flops_per_launch": fl,
                 "hbm_gbs": by / (ms * 1e-3) / 1e9, "fp32_tflops": fl / (ms * 1e-3) / 1e12,
                 "t_min_ms": max(t_hbm, t_fp) * 1e3, "frac": max(t_hbm, t_fp) * 1e3 / ms}
            per.append(r)
        dom = max(per, key=lambda r: r["kernel_ms"])
        if dom["bound"] == "latency":
            return {"bound": "latency", "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                    "kernel": dom["kernel"], "kernel_ms": kern_ms, "note": dom["note"]}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        traffic, tsrc = None, None
        if os.path.exists(tp):
            ent = json.load(open(tp)).get(f"{dom['workload']}:{dom['kernel']}")
            if isinstance(ent, dict):
                traffic, tsrc = ent.get("bytes"), ent.get("source")
            elif ent is not None:
                traffic, tsrc = ent, "profiles/traffic.json"
        if dom["bound"] == "hbm":
            roof = {"bound": "hbm", "achieved": dom["hbm_gbs"], "peak": hbm, "unit": "GB/s", "frac": dom["hbm_gbs"] / hbm,
                    "peak_source": self.hbm_src}
        else:
            roof = {"bound": "fp32", "achieved": dom["fp32_tflops"], "peak": self.fp32_peak(), "unit": "TFLOP/s",
                    "frac": dom["fp32_tflops"] / self.fp32_peak(),
                    "peak_source": "measured in this run (FFMA probe, qg_ctx_measure_fp32_tflops)",
                    "note": "non-tensor FP32 pipe: algorithmic flops (SURVEY.md 8d) / kernel time"}
        roof.update({"traffic": traffic, "traffic_source": tsrc or "no ncu capture of this kernel on this workload committed",
                     "kernel": dom["kernel"], "kernel_ms": dom["kernel_ms"], "dominant_workload": dom["workload"],
                     "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                     "algorithmic_flops_per_launch": dom["algorithmic_flops_per_launch"]})
        if len(per) > 1:
            roof["kernels"] = per
            # concurrent banks overlap: the step time, not the sum of the per-bank spans, is what the minimum times add up against
            roof["frac_all"] = sum(r["t_min_ms"] for r in per) / (kern_ms if concurrent else sum(r["kernel_ms"] for r in per))
            roof["step_kernel_ms"] = kern_ms
        return roof

    def e2e(self, wls, tmpls, rows_l, out_bytes, n_e2e):
        """bank build from host tables + render + D2H of every output sample through the public API, host buffers"""
        torch, qb, ctx, world = self.torch, self.qb, self.ctx, self.world
        n_e2e = max(1, n_e2e)
        T = wls[0].T
        rows = sum(rows_l)
        # host buffer: the whole render when this rank's share of host memory allows it, otherwise a block of T_host samples
        # that the render streams through (state persists across calls, every output byte still crosses the link)
        T_host = T
        try:
            avail = int(next(l for l in open("/proc/meminfo") if l.startswith("MemAvailable")).split()[1]) * 1024
            share = avail // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
            if os.environ.get("QG_BENCH_HOST_GB"):   # testing hook: pretend this rank may only use that much host memory
                share = int(float(os.environ["QG_BENCH_HOST_GB"]) * 1e9)
            if rows * T * 4 > 0.55 * share:
                T_host = max(4096, int(0.45 * share / (rows * 4)) // 4096 * 4096)
        except (OSError, StopIteration, ValueError):
            pass
        T_host = min(T_host, T)
        pinned = True
        try:
            h_out = torch.empty(rows * T_host, dtype=torch.float32).pin_memory()
        except RuntimeError:   # not enough lockable host memory on this box: pageable buffer, noted below
            h_out, pinned = torch.empty(rows * T_host, dtype=torch.float32), False
        h_all = h_out.numpy()
        row_off = np.concatenate([[0], np.cumsum(rows_l)])[:-1]

        def host_view(k, n):   # bank k's [rows][outputs][n] block inside the (reused) host buffer
            o = int(row_off[k]) * T_host
            return h_all[o: o + rows_l[k] * n].reshape(wls[k].V // wls[k].group, tmpls[k].outputs(), n)

        h2d = sum((0 if w.raw is None else w.raw.nbytes) + w.salts.nbytes for w in wls)

        dbg = bool(os.environ.get("QG_BENCH_DEBUG"))

        def e2e_step():
            for k, (t, w) in enumerate(zip(tmpls, wls)):
                ta = time.perf_counter()
                b2 = qb.Bank(t, w.V, raw=w.raw, salts=w.salts, ctx=ctx)
                self.set_paths([b2], [w])          # --path specialised: the NVRTC compile is inside the end-to-end time
                tb = time.perf_counter()
                for t0 in range(0, w.T, T_host):
                    n = min(T_host, w.T - t0)
                    b2.render(n, group=w.group, out=host_view(k, n))
                tc = time.perf_counter()
                kern = b2.kernel()
                del b2
                if dbg:
                    print(f"  e2e {w.name}: build {tb - ta:.4f} s, render+copy {tc - tb:.4f} s ({kern}), free {time.perf_counter() - tc:.4f} s", file=sys.stderr)

        e2e_step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            ts = time.perf_counter()
            e2e_step()
            if os.environ.get("QG_BENCH_DEBUG"):
                print(f"e2e step {time.perf_counter() - ts:.4f} s", file=sys.stderr)
        self.barrier()
        dt = self.max_ranks((time.perf_counter() - t0) / n_e2e)
        checksum = float(host_view(0, min(T_host, wls[0].T))[0, 0, : min(T_host, 4096)].astype(np.float64).sum())
        # what the link allows: one large pinned device->host copy per rank, ALL RANKS AT ONCE (the host side of the box is
        # shared: a probe taken alone would overstate what a rank can get while its neighbours copy too)
        nb = int(min(out_bytes, 1 << 30))
        dprobe = torch.empty(nb, dtype=torch.uint8, device="cuda")
        hprobe = h_out.view(torch.uint8)[:nb]
        hprobe.copy_(dprobe, non_blocking=True)
        self.barrier()
        tp0 = time.perf_counter()
        hprobe.copy_(dprobe, non_blocking=True)
        torch.cuda.synchronize()
        d2h_peak = nb / (time.perf_counter() - tp0) / 1e9
        d2h_peak_min = -self.max_ranks(-d2h_peak)
        del dprobe, h_out
        units = self.sum_ranks(sum(w.V * w.T for w in wls))
        return {"value": units / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(out_bytes), "ms_per_step": dt * 1e3, "steps": n_e2e,
                "d2h_gbs": out_bytes / dt / 1e9, "d2h_link_gbs": d2h_peak_min,
                "d2h_link_note": f"one {nb >> 20} MiB pinned copy per rank, all {world} rank(s) copying at the same time, slowest rank",
                "host_buffer": ("pinned" if pinned else "pageable") + ("" if T_host == T else f", streamed in blocks of {T_host} samples"),
                "host_placement": self.numa,
                "note": "bank build from host tables + render + device->host copy of every output sample (pinned host buffer)",
                "checksum": checksum}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=HEADLINE, choices=["c1", "c2", "c3", "c4", "c5", "c3_saw", "c2_butterpass", "c2_lowpole"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--path", default="auto", choices=["auto", "interp", "specialised"],
                    help="interp: lane interpreters only; specialised: force K1s (NVRTC) on every lane-interpreter bank")
    ap.add_argument("--only", action="store_true", help="time the selected workload only (no `configs` sub-records)")
    ap.add_argument("--strong", action="store_true", help="with --gpus N: shard the selected workload over the ranks (strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU oracle time per cpu_baseline sample")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3 if args.steps > 2 else args.warmup   # timing rule: W >= 3 (kept lower only for 1-2 step profiling runs)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: quartz_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    R = Runner(args, rank, local_rank, world)

    strong = args.strong and world > 1
    wls = make_workloads(args.workload, rank, world, strong=strong)
    head = R.measure(args.workload, wls, args.steps, args.warmup, not args.no_e2e, not args.no_cpu_baseline, clocks=True)
    line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": head["config"], "roofline": head["roofline"],
            "gpu_launches": head["gpu_launches"], "clocks": head.get("clocks")}
    for k in ("e2e", "cpu_baseline"):
        if k in head:
            line[k] = head[k]
    if not args.only and args.workload == HEADLINE:
        # the other BASELINE configs, each timed in this run with the same protocol (fewer steps for the long ones).
        # N = 1: all of them.  N > 1: configs[4] SHARDED over the ranks (strong scaling of the 1M-voice graph).
        subs = SUB_CONFIGS if world == 1 else ["c5"]
        line["configs"] = {}
        for name in subs:
            s_strong = world > 1
            sw = make_workloads(name, rank, world, strong=s_strong)
            steps = max(2, min(args.steps, 3)) if name in ("c4", "c5", "c2") else args.steps
            rec = R.measure(name, sw, steps, 3, not args.no_e2e, not args.no_cpu_baseline and world == 1)
            rec["scaling"] = "strong" if s_strong else "weak"
            line["configs"][name] = rec
        line["scaling_detail"] = {"headline": "weak: every rank renders a full-size configs[2] bank",
                                  "configs.c5": "strong: the 1,048,576 voices of configs[4] are split over the ranks" if world > 1
                                  else "single GPU"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
