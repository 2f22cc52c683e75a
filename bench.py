#!/usr/bin/env python3
"""bench.py — voice-samples/s @ 48 kHz for quartz's audio-graph hot path on N B200s (one process per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c1] [--impl reference]

A "step" is one offline render of the workload (all voices x all samples).  Default workload = BASELINE.json
configs[1] (4,096 noise->lowpass voices, 60 s, per-voice outputs kept: the HBM-write-bound configuration).
Prints ONE JSON line (see the contract in the task statement): `value` = kernel throughput with every input
resident in HBM; `e2e` = the same job through the public API with host buffers (bank build from host parameter
tables + render + device->host copy of every output sample); `roofline` for the dominant kernel; `cpu_baseline` =
the CPU oracle (a restatement of the reference: it is Rust and cannot be built here) on this box's host cores.
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


def cpu_baseline(wl, seconds_target=12.0, threads=None):
    """The oracle (port of the reference's tick loop) on a bounded sample of the same workload."""
    from tests.graphs import build
    from tests.oracle_ffi import ONet, render_bank
    threads = threads or os.cpu_count() or 1
    quantum = wl.group * threads if wl.V >= wl.group * threads else wl.group
    # calibrate on a short run, then size the sample for ~seconds_target of CPU work at the workload's full T
    nv0 = min(wl.V, quantum)
    cal = [build(wl.voice_expr(v), ONet).set_salt(int(wl.salts[v])) for v in range(nv0)]
    Tc = min(wl.T, 20000)
    t0 = time.perf_counter()
    render_bank(cal, Tc, group=wl.group, threads=threads)
    rate = nv0 * Tc / (time.perf_counter() - t0)
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
    T1 = int(min(T, max(2000, rate / max(1, threads) * 2.0 / n1)))
    one = [build(wl.voice_expr(v), ONet).set_salt(int(wl.salts[v])) for v in range(n1)]
    out = np.empty((n1 // wl.group, T1), dtype=np.float32)
    out.fill(0.0)
    t1 = time.perf_counter()
    render_bank(one, T1, group=wl.group, threads=1, out=out)
    one_core = n1 * T1 / (time.perf_counter() - t1)
    return {"value": nv * T / dt, "unit": "voice-samples/s", "cores": threads, "kind": "port", "one_core_value": one_core,
            "sample": f"{nv} voices x {T} samples x {reps} passes of {wl.name} (oracle/, {threads} threads, {dt * reps:.1f} s); one_core_value: "
                      f"{n1} voices x {T1} samples on 1 thread"}


def make_workload(name, rank, world=1):
    from quartz_b200 import shard, workloads
    # weak scaling: every rank renders a full-size, differently seeded bank
    return shard.shard_workload(workloads.WORKLOADS[name], world, rank)


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = make_workload(args.workload, 0)
    wl = wl[-1] if isinstance(wl, list) else wl
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(wl, seconds_target=min(15.0, max(2.0, 90.0 / max(1, args.warmup + args.steps))))
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals))
    nv_t = cb["sample"]
    line = {"impl": "reference", "metric": "voice-samples/s @48 kHz", "value": v, "unit": "voice-samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "voices": wl.V, "samples": wl.T, "sample_rate": 48000, "group": wl.group,
                       "note": "reference is Rust (no toolchain here): CPU oracle port, all host threads, bounded sample per step"},
            "cpu_baseline": dict(cb, value=v),
            "e2e": {"value": v, "unit": "voice-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "sample": nv_t}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--path", default="auto", choices=["auto", "interp", "specialised"],
                    help="specialised: K1s (NVRTC, opt-in) for every bank whose tape allows it and has no delay line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3 if args.steps > 2 else args.warmup   # timing rule: W >= 3 (kept lower only for 1-2 step profiling runs)

    import torch
    import torch.distributed as dist

    import quartz_b200 as qb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: quartz_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wls = make_workload(args.workload, rank, world)
    wls = wls if isinstance(wls, list) else [wls]
    wl = wls[0]
    # the library launches on the stream it is given; use a real (non-legacy) torch stream so that torch CUDA
    # events bracket exactly those launches
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = qb.Context(local_rank, stream=stream.cuda_stream)
    build_graph = __import__("tests.graphs", fromlist=["build"]).build
    tmpls = [build_graph(w.expr, qb.Net) for w in wls]
    tmpl = tmpls[0]
    banks = [qb.Bank(t, w.V, raw=w.raw, salts=w.salts, ctx=ctx) for t, w in zip(tmpls, wls)]
    bank = banks[0]
    if args.path == "interp":
        [b.set_path(qb.PATH_INTERP) for b in banks]

    def specialise(b, w):
        """K1s where it pays today: uniform tapes without a delay line (a ring load per sample is still dependent there)"""
        if "delay" in w.name or b.kernel() not in ("k_interp_blk", "k_interp<uniform>"):
            return                                 # fused / time-vector banks keep their kernel
        try:
            b.set_path(qb.PATH_SPECIALISED)
        except qb.QuartzGpuError:
            pass                                   # tape not specialisable / NVRTC missing: the bank keeps its AUTO kernel
    if args.path == "specialised":
        [specialise(b, w) for b, w in zip(banks, wls)]
    rows_l = [(w.V // w.group) * t.outputs() for t, w in zip(tmpls, wls)]
    rows = sum(rows_l)
    T = wl.T
    out_bytes = rows * T * 4
    d_out = torch.empty(rows * T, dtype=torch.float32, device="cuda")
    offs = np.concatenate([[0], np.cumsum(rows_l)])[:-1] * T * 4

    def step():
        for b, w, o in zip(banks, wls, offs):
            b.reset()
            b.render_device(w.T, d_out.data_ptr() + int(o), group=w.group)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    l0 = ctx.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for a, b in evs:
        a.record()
        step()
        b.record()
    e1.record()
    barrier()
    launches = ctx.launch_count() - l0
    total_ms = e0.elapsed_time(e1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    time.sleep(0.15)
    sampler.stop()
    if world > 1:
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    units = sum(w.V * w.T for w in wls)
    value = world * units / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host buffers (bank build from host tables + render + D2H)
    e2e = None
    if not args.no_e2e:
        n_e2e = max(1, min(args.steps, 3))
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

        h_np = host_view(0, T_host)
        del d_out
        torch.cuda.empty_cache()
        h2d = sum((0 if w.raw is None else w.raw.nbytes) + w.salts.nbytes for w in wls)

        def e2e_step():
            for k, (t, w) in enumerate(zip(tmpls, wls)):
                b2 = qb.Bank(t, w.V, raw=w.raw, salts=w.salts, ctx=ctx)
                if args.path == "interp":
                    b2.set_path(qb.PATH_INTERP)
                if args.path == "specialised":
                    specialise(b2, w)              # the NVRTC compile is inside the end-to-end time
                for t0 in range(0, w.T, T_host):
                    n = min(T_host, w.T - t0)
                    b2.render(n, group=w.group, out=host_view(k, n))
                del b2

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            ts = time.perf_counter()
            e2e_step()
            if os.environ.get("QG_BENCH_DEBUG"):
                print(f"e2e step {time.perf_counter() - ts:.4f} s", file=sys.stderr)
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        checksum = float(h_np[0, 0, : min(T_host, 4096)].astype(np.float64).sum())   # before the probe reuses the buffer
        # what the link allows: one large pinned device->host copy, timed alone (explains the e2e number, not part of it)
        nb = min(out_bytes, 1 << 30)
        dprobe = torch.empty(nb, dtype=torch.uint8, device="cuda")
        hprobe = h_out.view(torch.uint8)[:nb]
        hprobe.copy_(dprobe, non_blocking=True)
        torch.cuda.synchronize()
        tp0 = time.perf_counter()
        hprobe.copy_(dprobe, non_blocking=True)
        torch.cuda.synchronize()
        d2h_peak = nb / (time.perf_counter() - tp0) / 1e9
        del dprobe
        e2e = {"value": world * units / dt, "unit": "voice-samples/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(out_bytes), "ms_per_step": dt * 1e3, "steps": n_e2e,
               "d2h_gbs": out_bytes / dt / 1e9, "d2h_link_gbs": d2h_peak,
               "host_buffer": ("pinned" if pinned else "pageable") + ("" if T_host == T else f", streamed in blocks of {T_host} samples"),
               "note": "bank build from host tables + render + device->host copy of every output sample (pinned host buffer); "
                       "bounded by the PCIe link: d2h_gbs vs d2h_link_gbs (one large pinned copy timed alone)",
               "checksum": checksum}

    if rank == 0:
        peak, peak_src = _peaks()
        alg_bytes = sum(w.V * w.T * w.bytes_per_unit for w in wls)
        alg_flops = sum(w.V * w.T * w.flops_per_unit for w in wls)
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(f"{wl.name}:{bank.kernel()}")
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": bank.kernel(), "kernel_ms": kern_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "binding_resource": wl.bound}
        if wl.bound != "hbm" and alg_flops > 0:
            # compute-bound workload: the binding roofline is the FP32 (non-tensor) pipe, measured here with an FFMA probe
            fp32_peak = ctx.measure_fp32_tflops()
            tf = alg_flops / (kern_ms * 1e-3) / 1e12
            roof = {"bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak, "traffic": traffic,
                    "peak_source": "measured here (FFMA probe, qg_ctx_measure_fp32_tflops)", "kernel": bank.kernel(),
                    "kernel_ms": kern_ms, "algorithmic_flops_per_launch": alg_flops, "algorithmic_bytes_per_launch": alg_bytes,
                    "hbm_frac": achieved / peak, "binding_resource": wl.bound,
                    "note": "algorithmic flops (SURVEY.md 8d) / kernel time; issue-slot utilisation is in profiles/"}
        line = {
            "metric": "voice-samples/s @48 kHz", "value": value, "unit": "voice-samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name if len(wls) == 1 else args.workload + ":" + "+".join(w.name for w in wls),
                       "voices": sum(w.V for w in wls), "samples": wl.T, "sample_rate": 48000, "group": wl.group,
                       "layout": "voice-major [V/G][T] f32", "kernel": "+".join(sorted({b.kernel() for b in banks})), "note": wl.note,
                       "l2": f"each step writes {out_bytes / 1e9:.1f} GB of output (>> 126 MB L2), state re-initialised per step"},
            "roofline": roof,
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wls[0] if len(wls) == 1 else wls[-1])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
