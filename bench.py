#!/usr/bin/env python
"""POMCP simulations/second on RockSample(15,15) -- the BASELINE.json metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one move's search: a fresh belief tree over the 2000-particle root belief, ``n_sims`` simulations per
GPU (root-parallel: every rank grows its own tree) and the merge of the root statistics (one
all-reduce of 64 int64 over NCCL when N > 1).  `value` = simulations of all ranks / device time (CUDA events on
the launching stream, inputs resident in HBM, max over ranks); `e2e` = the same through the C ABI with HOST
buffers (belief copied host->device, root statistics device->host, every step).

--impl reference: the CPU arm.  The reference is pure Python and cannot travel to the GPU box, so this arm runs
the oracle's C port of the reference's sequential POMCP (oracle/pomcp_oracle.c, "ref" mode, pinned bit-for-bit
to the executed reference) on every host core, one independent search per core.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

MAP = "map-15-15"
N_SIMS = 1_000_000
SCHEDULE = (32768, 4, 1 << 19)          # generation widths: 32768, 131072, 524288 (+ the rest) simulations in flight;
                                        # return at 1e6 simulations per move does not depend on the first width
                                        # (160 paired episodes per width, 4096 ... 65536: profiles/r01_schedule_sweep.txt)
REF_SAMPLE_SIMS = 100_000               # per core and step, reference arm
METRIC = "pomcp_sims_per_sec_rocksample_15_15"


def build_inputs(n_sims):
    """World and root belief exactly as BASELINE.md 8d: np.random.seed(123); world = first draw; 2000 particles."""
    from pomdpy_b200.rock_sample import RockModel
    args = dict(map_file=MAP, discount=0.95, max_depth=100, ucb_coefficient=3.0, n_start_states=2000,
                max_particle_count=2000, n_sims=n_sims, seed=123)
    model = RockModel(args)
    np.random.seed(123)
    from pomdpy_b200 import util
    util.VERBOSITY = 0
    model.reset_for_epoch()
    bag = np.fromiter((model.pack_state(model.sample_an_init_state()) for _ in range(2000)), dtype=np.uint32, count=2000)
    actual, sampled = model.world_masks()
    return model, model.device_spec(), bag, actual, sampled


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons while the GPU is under load."""
    Q = "clocks.sm,clocks.max.sm,utilization.gpu,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                clk, cmax, util = float(r[0]), float(r[1]), float(r[2])
            except Exception:
                continue
            mx = max(mx, cmax)
            if util > 0:
                sm.append(clk)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons),
                    samples=len(self.rows))


def cpu_baseline(spec, bag, actual, n_sims):
    """The oracle's port of the reference's sequential POMCP on ONE host core (the reference is single-threaded)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    om = oracle.Model(spec["rows"], spec["cols"], spec["cell"], spec["start_cell"], spec["thr53"], spec["rewards"])
    sample = min(n_sims, 1_000_000)
    ref = oracle.RefPOMCP(om, actual, 0, (bag >> 8).astype(np.uint32), seed=123)
    t0 = time.perf_counter()
    ref.search(sample)
    dt = time.perf_counter() - t0
    return dict(value=sample / dt, unit="sims/s", cores=1, kind="port",
                sample="one search of %d simulations (reference algorithm, quirks on) from the step's root belief; "
                       "%.1f s on 1 core" % (sample, dt))


def _ref_worker(a):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    spec, bag_rocks, actual, sims, seed = a
    om = oracle.Model(spec["rows"], spec["cols"], spec["cell"], spec["start_cell"], spec["thr53"], spec["rewards"])
    ref = oracle.RefPOMCP(om, actual, 0, bag_rocks, seed=seed)
    ref.search(sims)
    return int(ref.root_stats()["N"])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.build()
    _, spec, bag, actual, _ = build_inputs(N_SIMS)
    cores = os.cpu_count() or 1
    spec_small = {k: spec[k] for k in ("rows", "cols", "cell", "start_cell", "thr53", "rewards")}
    bag_rocks = (bag >> 8).astype(np.uint32)
    times = []
    with mp.get_context("fork").Pool(cores) as pool:
        for it in range(args.warmup + args.steps):
            jobs = [(spec_small, bag_rocks, actual, REF_SAMPLE_SIMS, 1000 * it + c) for c in range(cores)]
            t0 = time.perf_counter()
            done = pool.map(_ref_worker, jobs)
            dt = time.perf_counter() - t0
            assert sum(done) == cores * REF_SAMPLE_SIMS
            if it >= args.warmup:
                times.append(dt)
    value = cores * REF_SAMPLE_SIMS * len(times) / sum(times)
    sample = "%d independent searches (one per host core) of %d simulations each per step, C port of the " \
             "reference's sequential POMCP" % (cores, REF_SAMPLE_SIMS)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "sims/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/f64", "data": "synthetic",
        "config": {"workload": "RockSample(15,15) POMCP, n_sims=1M per move per GPU, root-parallel", "map": MAP,
                   "sample_sims_per_core": REF_SAMPLE_SIMS},
        "cpu_baseline": {"value": value, "unit": "sims/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))
    return 0


def run_b200(args):
    import torch
    import torch.distributed as dist
    from pomdpy_b200 import native
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_sims = args.n_sims
    model, spec, bag, actual, sampled = build_inputs(n_sims)
    w0, growth, wmax = args.w0, args.growth, args.wmax
    pl = native.Planner(max_depth=100, tree_depth_cap=64, max_particle_count=2000, ucb_coefficient=3.0, discount=0.95,
                        seed=123, node_capacity=n_sims + 4096, gen_w0=w0, gen_growth=growth, gen_wmax=wmax,
                        device=local)
    pl.set_model_rocksample(spec["rows"], spec["cols"], spec["cell"], spec["start_cell"], spec["thr53"], spec["rewards"])
    pl.set_world(actual, sampled)
    host_bag = torch.from_numpy(bag.view(np.int32).copy()).pin_memory()     # pinned host belief
    host_bag_np = host_bag.numpy().view(np.uint32)
    pl.set_root_particles(host_bag_np, spec["start_cell"])
    stream = torch.cuda.current_stream().cuda_stream
    fused = None
    if world > 1 and os.environ.get("POMCP_FUSED_MERGE", "1") != "0":
        from pomdpy_b200 import dist as dist_mod
        fused = dist_mod.try_fused_merge(pl)             # merge inside the search kernel (NVLink peer stores)
    stats = torch.zeros(64, dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")         # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(i, ev=None):
        pl.reset_tree(stream)
        if ev:
            ev[0].record()
        pl.search(n_sims, move=i, sim_offset=rank * n_sims, stream=stream)
        if ev:
            ev[1].record()
        if fused is None:
            pl.root_stats_device(stats.data_ptr(), stream)
            if world > 1:
                dist.all_reduce(stats, op=dist.ReduceOp.SUM)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for i in range(args.warmup):
        flush.zero_()
        device_step(i)
    barrier()
    c0 = pl.counters()
    ev_step = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev_kern = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()                                   # L2 flushed between timed steps (not timed)
        ev_step[i][0].record()
        device_step(args.warmup + i, ev_kern[i])
        ev_step[i][1].record()
    barrier()
    c1 = pl.counters()
    step_ms = sum(a.elapsed_time(b) for a, b in ev_step)
    kern_ms = sum(a.elapsed_time(b) for a, b in ev_kern) / args.steps
    A = spec["n_actions"]
    merged = stats.cpu().numpy()[:A] if fused is None else pl.merged_root_stats()["n"]
    assert int(merged.sum()) == n_sims * world, "root visit counts do not add up to the simulations run"

    # ---- end to end through the C ABI with host buffers --------------------------------------------------
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        pl.set_root_particles(host_bag_np, spec["start_cell"])                     # H2D from pinned memory
        pl.search(n_sims, move=1000 + i, sim_offset=rank * n_sims, stream=stream)
        if fused is not None:
            res = pl.merged_root_stats()                                           # D2H of the in-kernel merge
        elif world > 1:
            pl.root_stats_device(stats.data_ptr(), stream)
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
            res = stats.cpu()                                                      # D2H
        else:
            res = pl.root_stats()                                                  # D2H
    barrier()
    e2e_s = time.perf_counter() - t0
    del res
    clocks = sampler.stop() if sampler else None

    t = torch.tensor([step_ms, e2e_s * 1e3, kern_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, e2e_ms, kern_ms = (float(x) for x in t.cpu())

    if rank == 0:
        sims = n_sims * world * args.steps
        d = {k: c1[k] - c0[k] for k in ("levels", "rollout_steps", "created", "generations", "level_steps", "launches")}
        L = d["levels"] / (n_sims * args.steps)
        R = d["rollout_steps"] / (n_sims * args.steps)
        bytes_per_sim = L * (8 * A + 56) + 8 * A + 20
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = bytes_per_sim * n_sims / (kern_ms * 1e-3) / 1e9
        prof = {}
        try:
            with open(os.path.join(ROOT, "profiles", "search_kernel_summary.json")) as f:
                prof = json.load(f)
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": sims / (step_ms * 1e-3), "unit": "sims/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32/f64", "data": "synthetic",
            "config": {"workload": "RockSample(15,15) POMCP, n_sims=1M per move per GPU, root-parallel", "map": MAP,
                       "n_sims_per_gpu": n_sims, "root_particles": 2000, "max_depth": 100, "ucb_coefficient": 3.0,
                       "discount": 0.95, "generation_schedule": [w0, growth, wmax],
                       "l2": "flushed between timed steps (256 MiB device write, untimed)",
                       "step": "fresh tree + search + root-statistics merge",
                       "merge": ("none (1 GPU)" if world == 1 else
                                 "fused into the search kernel: peer stores over NVLink, no collective call" if fused
                                 else "NCCL all-reduce of 64 int64"),
                       "tree_levels_per_sim": L, "rollout_steps_per_sim": R,
                       "generate_steps_per_s": (d["levels"] + d["rollout_steps"]) * world / (step_ms * 1e-3)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": prof.get("dram_bytes_per_launch"),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "kernel": "pomcp_search_kernel", "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_sim": bytes_per_sim,
                         "note": "not HBM-bound by construction (SURVEY 8d): the binding resource is instruction issue; "
                                 "see issue_slot_utilisation (from the committed ncu capture) and profiles/",
                         "issue_slot_utilisation_pct": prof.get("issue_active_pct"),
                         "thread_instructions_per_generate_step": prof.get("thread_inst_per_generate_step")},
            "e2e": {"value": sims / (e2e_ms * 1e-3), "unit": "sims/s", "h2d_bytes_per_step": int(bag.nbytes),
                    "d2h_bytes_per_step": int(c1["root_stats_d2h_bytes"] if (world == 1 or fused is not None) else 64 * 8)},
            "gpu_launches": int(d["launches"]),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(spec, bag, actual, n_sims)
        print(json.dumps(out))
    pl.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n_sims", type=int, default=N_SIMS)
    ap.add_argument("--w0", type=int, default=SCHEDULE[0])
    ap.add_argument("--growth", type=int, default=SCHEDULE[1])
    ap.add_argument("--wmax", type=int, default=SCHEDULE[2])
    ap.add_argument("--no_cpu_baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
