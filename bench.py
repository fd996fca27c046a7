"""bench.py -- auction opportunities/s of the B200 engine on the SP_Truthful_TS-shaped synthetic workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[4]): 4096 runs sharded over the GPUs of one box (4096 / world
each: STRONG scaling, the whole job is fixed), T = 10 000 rounds per iteration, A = 64 agents x I = 64 items, D = 5, Do = 4,
P = 2, SecondPrice + TruthfulBidder + learnt Thompson-sampling allocator.  `--runs-per-gpu R` switches to weak scaling.
One step = one iteration of every resident run = T rounds of the fused round-loop kernel (K1-K5) + the per-iteration
allocator fits (K6, adam_ref: the reference's Adam + plateau scheduler + early-stop state machine) + the per-iteration
metric read-out to the host; at N > 1 the timed region ends with ONE NCCL all-gather of its iterations' metric blocks (K8:
main.py:186-222 assembles its per-run rows once the runs have finished).  value = runs * T / step time (max over ranks).
Each GPU processes its runs as 4 independent sub-shards on 4 CUDA streams (the tail of one sub-shard's fit grid, and end to
end its PCIe copies, are covered by the others' kernels; runs are independent, results do not depend on the split).
The timed steps are iterations W .. W+K-1 of ONE learning trajectory that starts from m ~ N(0, 1), q = 1: the fit gets
cheaper as the allocators learn, so the iteration range is part of the configuration and is printed.

The JSON line also carries: e2e (same iterations through the public API with the learnt state and metrics crossing PCIe
from / to pinned host memory every step; a sub-shard's next step waits on the host for its own read-back), fit_epochs_mean,
full_workload (the whole N = 100-iteration trajectory), opt_in_newton_mode (the same under AGYM_FIT_NEWTON: a different algorithm),
roofline (dominant kernel) + roofline_kernels (every kernel, incl. the staged resolution kernel K4 the north star puts
the HBM bar on), cpu_baseline, clocks, gpu_launches (counted by the library).

`--impl reference` times the UNMODIFIED reference (oracle/_ref, a verbatim copy shipped by gpurun) on all host cores on
the SAME iterations: each step fits the dumped fit inputs of iteration W + i of this trajectory
(tests/golden/bench_fit_inputs.npz, written by tools/dump_bench_fit_inputs.py) and times the reference's round loop at the
bench shape; fit_epochs_mean is printed by both arms.  The numpy port is timed beside it as a second, labelled number.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = dict(A=64, I=64, D=5, Do=4, P=2, T=10000, runs_total=4096, iterations_full=100)
METRIC = "auction opportunities/sec"
UNIT = "opportunities/s"
SEED = 0            # Philox seed of the round loop (per-run key = (SEED, global run index))
INIT_SEED = 1000    # initial allocator state: m[r] ~ N(0, 1) from default_rng([INIT_SEED, global run index])


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--runs-total", type=int, default=WORKLOAD["runs_total"], help="strong scaling: runs of the whole job (default 4096)")
    ap.add_argument("--runs-per-gpu", type=int, default=0, help="weak scaling: runs per GPU (overrides --runs-total)")
    ap.add_argument("--rounds", type=int, default=WORKLOAD["T"])
    ap.add_argument("--allocator", default="ts", choices=["ts", "oracle"], help="ts = SP_Truthful_TS shape (headline); oracle = SP_Oracle shape")
    ap.add_argument("--fit-mode", default="adam_ref", choices=["adam_ref", "adam_fast", "newton"],
                    help="adam_ref: the reference's algorithm (the headline); newton: opt-in, a different algorithm, never comparable with the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the staged-kernel roofline measurements and the shipped config")
    ap.add_argument("--no-full", action="store_true", help="skip the whole-trajectory (100 iterations) measurement")
    ap.add_argument("--full-iterations", type=int, default=WORKLOAD["iterations_full"])
    ap.add_argument("--subshards", type=int, default=0, help="process this GPU's runs as K independent sub-shards on K CUDA streams "
                    "(0 = auto: 2 when a GPU holds at most 2048 runs, else 1)")
    return ap.parse_args(argv)


def shard(args, world, rank):
    """(first global run, runs on this rank, runs of the whole job, scaling)."""
    if args.runs_per_gpu > 0:
        return rank * args.runs_per_gpu, args.runs_per_gpu, args.runs_per_gpu * world, "weak"
    base, extra = divmod(args.runs_total, world)
    return rank * base + min(rank, extra), base + (1 if rank < extra else 0), args.runs_total, "strong"


def config_dict(args, world, runs_job, scaling, first_it, n_it):
    w = WORKLOAD
    return {"workload": f"synthetic SP_Truthful_TS shape: {runs_job} runs x {args.rounds} rounds/iteration x {w['A']} agents x {w['I']} items, "
                        f"D={w['D']} Do={w['Do']} P={w['P']}, SecondPrice, TruthfulBidder, "
                        f"{'learnt Thompson-sampling allocator + ' + args.fit_mode + ' fits' if args.allocator == 'ts' else 'OracleAllocator (no fits)'}; "
                        f"timed steps = iterations {first_it}..{first_it + n_it - 1} of the learning trajectory from m ~ N(0,1), q = 1",
            "runs": runs_job, "runs_per_gpu": runs_job // world, "rounds_per_step": args.rounds, "iterations": [first_it, first_it + n_it],
            "agents": w["A"], "items": w["I"], "embedding_size": w["D"], "obs_embedding_size": w["Do"],
            "participants": w["P"], "allocation": "SecondPrice", "allocator": args.allocator, "fit_mode": args.fit_mode,
            "step": "one iteration: T rounds (fused K1-K5) + allocator fits (K6) + metric read-out to the host" + ("; the timed region ends with ONE NCCL all-gather of the metric blocks of all its iterations (K8, agym_gather_block_nccl), as main.py:186-222 assembles its per-run rows once the runs have finished" if world > 1 else ""),
            "parallelism": f"runs sharded over {world} GPU(s) ({scaling} scaling), no data-path collective" +
                           (f"; each GPU's runs go as {args.subshards} independent sub-shards on {args.subshards} CUDA streams, so the tail of one "
                            f"sub-shard's fit grid overlaps the other's kernels (runs are independent, main.py:186-189; results are bit-identical)"
                            if args.subshards > 1 else ""),
            "l2": "no flush needed: per-step working set (learnt state + winner log + fit workspace, > 400 MB per 512 runs) exceeds the 126 MB L2"}


def initial_m(first_run, count):
    """Models.py:22 m ~ N(0, 1), one independent stream per GLOBAL run (the same run starts from the same state on any shard)."""
    import torch

    w = WORKLOAD
    out = np.empty((count, w["A"], w["I"], w["Do"] + 1), np.float32)
    for r in range(count):
        out[r] = np.random.default_rng([INIT_SEED, first_run + r]).standard_normal(out.shape[1:], dtype=np.float32)
    return torch.from_numpy(out)


def make_engine(ag, _lib, R, T, learnt, device, run_offset):
    from oracle import auction_oracle as ao  # catalog sampler only (main.py:60-72); never on the timed path

    w = WORKLOAD
    E, V = ao.make_catalog(np.random.default_rng(0), w["A"], w["I"], w["D"])  # shared by every run and rank
    return ag.Engine(R=R, A=w["A"], I=w["I"], D=w["D"], Do=w["Do"], P=w["P"], mechanism=_lib.SECOND_PRICE, E=E, V=V, n_items=[w["I"]] * w["A"],
                     alloc_kind=[_lib.ALLOC_TS if learnt else _lib.ALLOC_ORACLE] * w["A"], bidder_kind=[_lib.BID_TRUTHFUL] * w["A"],
                     precision=_lib.FP32, device=device, run_offset=run_offset, rounds_capacity=T)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_numbers(kernel_pattern, grid=None):
    """Counters of the newest committed ncu summary (profiles/r*_ncu_full.txt, written by tools/ncu_summary.py) whose kernel
    name matches; nothing here is a constant of this file.  `traffic` only when the capture's grid equals this run's."""
    best, best_grid = None, None  # newest capture of the kernel; newest one whose grid equals this run's, if any

    def consider(path, block):
        nonlocal best, best_grid
        if block.get("name") and re.search(kernel_pattern, block["name"]):
            best = (path, block)
            if grid is not None and block.get("launch__grid_size") == grid:
                best_grid = (path, block)

    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full.txt"))):
        block = {}
        for ln in open(path):
            if ln.startswith("Kernel Name"):
                consider(path, block)
                block = {"name": ln[len("Kernel Name"):].strip()}
                continue
            f = ln.split()
            if len(f) >= 2 and not ln.startswith("-"):
                try:
                    block[f[0]] = float(f[1]) * ({"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(f[2], 1.0) if len(f) > 2 else 1.0)
                except ValueError:
                    pass
        consider(path, block)
    if best is None:
        return None
    path, b = best_grid or best
    out = {"source": os.path.relpath(path, ROOT), "kernel": b["name"],
           "warp_instructions_per_sm_cycle": b.get("sm__inst_executed.avg.per_cycle_elapsed"),
           "issue_slots_active_pct": b.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "warps_active_pct": b.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
           "registers_per_thread": b.get("launch__registers_per_thread"), "grid": b.get("launch__grid_size"),
           "dram_bytes": (b.get("dram__bytes_read.sum", 0.0) + b.get("dram__bytes_write.sum", 0.0)) * 1e6}
    out["traffic_matches_this_grid"] = bool(grid is not None and out["grid"] == grid)
    return out


# ------------------------------------------------------------------------------------------------
def cpu_kind():
    from oracle import ref_bench

    return "reference" if ref_bench.reference_available() else "port"


def run_reference_arm(args, rank, world):
    """CPU arm (rank 0 only): the unmodified reference on all host cores, on the SAME iterations as the GPU arm."""
    if rank != 0:
        return
    from oracle import ref_bench

    w = WORKLOAD
    learnt = args.allocator == "ts"
    kind = cpu_kind()
    workers = max(1, min(os.cpu_count() or 1, 64))
    _, _, runs_job, scaling = shard(args, world, 0)
    fits_per_worker = max(1, 16 // workers) if learnt else 0  # >= 16 fits per step: the mean epoch count over K = 20 steps is good to ~2 %
    n_rounds = 400
    t_all = time.perf_counter()
    rows, ports = [], []
    with ref_bench.make_pool(workers) as pool:
        for i in range(args.warmup + args.steps):
            warm = i < args.warmup
            r = ref_bench.run(kind, w["A"], w["I"], w["D"], w["Do"], w["P"], args.rounds, iteration=i, n_rounds=100 if warm else n_rounds,
                              fits_per_worker=0 if warm else fits_per_worker, workers=workers, learnt=learnt, seed=i, pool=pool)
            if not warm:
                rows.append(r)
                if kind == "reference":  # the numpy port on the same fits, as a second, labelled number
                    ports.append(ref_bench.run("port", w["A"], w["I"], w["D"], w["Do"], w["P"], args.rounds, iteration=i, n_rounds=n_rounds,
                                               fits_per_worker=fits_per_worker, workers=workers, learnt=learnt, seed=i, pool=pool))

    def agg(rs):
        # whole-job throughput, as the GPU arm's `value`: K steps' opportunities / the time K steps take on `workers` cores
        # (NOT the mean of per-step rates, which would weight the cheap late iterations like the expensive early ones)
        if learnt:
            secs = [args.rounds / r["round_only_per_core"] + w["A"] * r["fit_seconds_mean"] for r in rs]  # one run's iteration on one core
            return workers * args.rounds * len(rs) / float(np.sum(secs))
        return workers * len(rs) / float(np.sum([1.0 / r["round_only_per_core"] for r in rs]))

    value = agg(rows)
    sample = (f"per step i (iteration {args.warmup} + i of the GPU arm's trajectory) and per core: {n_rounds} rounds of the {'unmodified reference' if kind == 'reference' else 'numpy port'}'s "
              f"simulate_opportunity at the bench shape + {fits_per_worker} allocator fit(s) on the dumped inputs of that iteration "
              f"(tests/golden/bench_fit_inputs.npz); {workers} processes, one torch thread each; opportunities/s = cores * K * T / sum over the K steps of (T / round_rate + A * fit_seconds)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean([r["wall_seconds"] for r in rows])), "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32/f64 (the reference's own mix)", "data": "synthetic",
            "config": config_dict(args, world, runs_job, scaling, args.warmup, args.steps),
            "fit_epochs_mean": float(np.mean([r["fit_epochs_mean"] for r in rows])) if learnt else None,
            "fit_epochs_per_step": [r["fit_epochs_mean"] for r in rows] if learnt else None,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample,
                             "per_core": value / workers,
                             "round_loop_only_per_core": float(np.mean([r["round_only_per_core"] for r in rows])),
                             "fit_seconds": float(np.mean([r["fit_seconds_mean"] for r in rows])),
                             "fit_epochs": float(np.mean([r["fit_epochs_mean"] for r in rows])),
                             "fits_timed": int(sum(r["fits_timed"] for r in rows))},
            "port": ({"value": agg(ports), "unit": UNIT, "kind": "port", "cores": workers,
                      "round_loop_only_per_core": float(np.mean([r["round_only_per_core"] for r in ports])),
                      "fit_seconds": float(np.mean([r["fit_seconds_mean"] for r in ports])),
                      "fit_epochs": float(np.mean([r["fit_epochs_mean"] for r in ports])),
                      "note": "oracle/auction_oracle.py + oracle/fit_oracle.py (the numpy restatement) on the same sample"} if ports else None),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "total_seconds": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = WORKLOAD
    A, I, D, Do, P, T = w["A"], w["I"], w["D"], w["Do"], w["P"], args.rounds
    first_run, R, runs_job, scaling = shard(args, world, rank)
    K = Do + 1
    learnt = args.allocator == "ts"
    fit_mode = {"adam_ref": _lib.FIT_ADAM_REF, "adam_fast": _lib.FIT_ADAM_FAST, "newton": _lib.FIT_NEWTON}[args.fit_mode]
    if args.subshards <= 0:
        # measured on B200, whole trajectory / e2e step: 512 runs per GPU, 1 / 2 / 4 sub-shards -> 4.60 / 4.20 / 4.16 s;
        # 4096 runs per GPU -> 32.9 / 32.3 / 32.4 s and 555.9 / 532.0 / 524.5 ms (the copies of one sub-shard overlap the fits of another)
        args.subshards = (4 if R % 4 == 0 else 2 if R % 2 == 0 else 1) if learnt else 1
    NS = args.subshards
    assert R % NS == 0, "--subshards must divide the runs per GPU"
    Rs = R // NS

    class Sub:  # one sub-shard: its engine, its stream, its pinned host buffers
        pass

    subs = []
    for k in range(NS):
        sb = Sub()
        sb.first = first_run + k * Rs
        sb.eng = make_engine(ag, _lib, Rs, T, learnt, local_rank, sb.first)
        sb.stream = torch.cuda.current_stream(dev) if NS == 1 else torch.cuda.Stream(dev)
        if learnt:
            sb.m0_host = initial_m(sb.first, Rs)
            sb.m_host = sb.m0_host.clone().pin_memory()
            sb.q_host = torch.ones((Rs, A, I, K)).pin_memory()
            sb.mp_host = sb.m0_host.clone().pin_memory()
            sb.eng.set_allocator_state(sb.m_host, sb.q_host, sb.mp_host)
        sb.acc_host = torch.empty((Rs, A, _lib.NUM_METRICS), dtype=torch.float64).pin_memory()
        sb.rev_host = torch.empty((Rs,), dtype=torch.float64).pin_memory()
        # K8: every iteration's metric block of every rank (main.py:186-222 keeps per-run rows, so gather, not reduce)
        if world > 1:
            sb.eng.comm_init(rank, world)  # this sub-shard's NCCL communicator (agym_comm_init; the id is broadcast by torch.distributed)
        sb.epochs_sum = torch.zeros(2, dtype=torch.float64, device=dev)  # {sum of epochs, fits} over the steps that ask for it
        if world > 1:  # the metric blocks of a region's iterations, kept for the one all-gather at its end
            n_hist = max(args.steps, args.full_iterations)
            sb.hist_acc = torch.zeros((n_hist, Rs, A, _lib.NUM_METRICS), dtype=torch.float64, device=dev)
            sb.hist_rev = torch.zeros((n_hist, Rs), dtype=torch.float64, device=dev)
        subs.append(sb)
    eng = subs[0].eng
    stream = torch.cuda.current_stream(dev)
    torch.cuda.synchronize(dev)
    ev_pairs = {"rounds": [], "fit": []}

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])

    def read_out(sb, slot=None):
        """Per-iteration metric read-out: D2H of this rank's block.  `slot`: the block is also kept on the device for the job's
        one collective (gather_history: main.py:186-222 assembles its per-run rows once, when the runs have finished)."""
        e = sb.eng
        sb.acc_host.copy_(e.acc, non_blocking=True)
        sb.rev_host.copy_(e.revenue, non_blocking=True)
        if slot is not None and world > 1:
            sb.hist_acc[slot].copy_(e.acc)
            sb.hist_rev[slot].copy_(e.revenue)

    def gather_history(n_slots):
        """K8 inside the timed region: ONE NCCL all-gather per sub-shard of the metric blocks of the region's n_slots iterations
        ([n, Rs, A, 12] and [n, Rs] doubles, agym_gather_block_nccl), on the sub-shard's stream."""
        if world == 1:
            return
        for sb in subs:
            with torch.cuda.stream(sb.stream):
                sb.gathered = (sb.eng.gather_block(sb.hist_acc[:n_slots]), sb.eng.gather_block(sb.hist_rev[:n_slots]))

    def step_device(it, timed, count_epochs=False):
        """One iteration with everything resident in HBM (every sub-shard on its own stream)."""
        for sb in subs:
            with torch.cuda.stream(sb.stream):
                e = sb.eng
                e.clear_iteration()
                if timed:
                    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                    e0.record(sb.stream)
                e.simulate(SEED, it, T)
                if timed:
                    e1.record(sb.stream)
                if learnt:
                    info = e.update_allocators(want_info=count_epochs, fit_mode=fit_mode)
                    if count_epochs:
                        ran = info[..., 1]
                        sb.epochs_sum.add_(torch.stack([ran.sum(dtype=torch.float64), (ran > 0).sum().to(torch.float64)]))
                if timed:
                    e2.record(sb.stream)
                    ev_pairs["rounds"].append((e0, e1))
                    ev_pairs["fit"].append((e1, e2))
                read_out(sb, slot=(it - args.warmup) if timed else None)

    def step_e2e(it):
        """The same iteration through the public API with HOST buffers: learnt state up, metrics + state down.  A sub-shard's runs
        start their next step only when their previous step's state and metrics ARE on the host (host-side synchronize of that
        sub-shard's stream); the other sub-shards, whose runs are independent, keep the GPU busy meanwhile."""
        for sb in subs:
            sb.stream.synchronize()
            with torch.cuda.stream(sb.stream):
                e = sb.eng
                if learnt:
                    e.set_allocator_state(sb.m_host, sb.q_host, sb.mp_host, non_blocking=True)
                e.clear_iteration()
                e.simulate(SEED, it, T)
                if learnt:
                    e.update_allocators(want_info=False, fit_mode=fit_mode)
                    sb.m_host.copy_(e.m, non_blocking=True)
                    sb.q_host.copy_(e.q, non_blocking=True)
                    sb.mp_host.copy_(e.m_prev, non_blocking=True)
                read_out(sb, slot=(it - args.warmup) if it >= args.warmup else None)

    def join_streams():
        for sb in subs:
            if sb.stream is not stream:
                stream.wait_stream(sb.stream)

    def fork_streams():
        for sb in subs:
            if sb.stream is not stream:
                sb.stream.wait_stream(stream)

    def timed_region(fn, steps, it0):
        barrier()
        torch.cuda.synchronize(dev)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        fork_streams()
        for i in range(steps):
            fn(it0 + i)
        gather_history(steps)  # K8: the region's one collective, inside the region
        join_streams()
        e.record(stream)
        torch.cuda.synchronize(dev)
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        rank_ms.append(per_rank(ms))
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    rank_ms = []  # per timed region: every rank's own time (the reported time is their maximum)

    def per_rank(ms):
        if world == 1:
            return [round(float(ms.item()), 3)]
        allr = [torch.zeros_like(ms) for _ in range(world)]
        dist.all_gather(allr, ms)
        return [round(float(x.item()), 3) for x in allr]

    def reset_state():
        torch.cuda.synchronize(dev)
        if learnt:
            for sb in subs:
                sb.m_host.copy_(sb.m0_host); sb.q_host.fill_(1.0); sb.mp_host.copy_(sb.m0_host)
                sb.eng.set_allocator_state(sb.m_host, sb.q_host, sb.mp_host)
        torch.cuda.synchronize(dev)

    # ---- warm-up, then the device-resident timed region: iterations W .. W+K-1 ----
    for it in range(args.warmup):
        step_device(it, False, count_epochs=True)  # same code path as the timed steps (no first-call costs inside the timed region)
    gather_history(1)  # ... including the collective's first call
    torch.cuda.synchronize(dev)
    for sb in subs:
        sb.epochs_sum.zero_()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = sum(sb.eng.launch_count() for sb in subs)
    ms_total = timed_region(lambda i: step_device(i, True, count_epochs=True), args.steps, args.warmup)
    launches = sum(sb.eng.launch_count() for sb in subs) - launches0
    clk = clocks.stop()
    opp_per_step = runs_job * T
    value = opp_per_step * args.steps / (ms_total * 1e-3)
    # per-kernel intervals: with sub-shards the intervals of different streams overlap, so the per-step figure is the SUM over the
    # sub-shards of each one's interval (an upper bound of the kernel's share, exact when NS == 1)
    per_step_ms = {k: [sum(a.elapsed_time(b) for a, b in v[i * NS:(i + 1) * NS]) for i in range(len(v) // NS)] for k, v in ev_pairs.items()}
    k_ms = {k: float(np.mean(v)) if v else 0.0 for k, v in per_step_ms.items()}
    es = torch.stack([sb.epochs_sum for sb in subs]).sum(dim=0)
    if world > 1:
        dist.all_reduce(es)
    fit_epochs_mean = float(es[0] / es[1]) if learnt and float(es[1]) > 0 else None

    # ---- end to end through host buffers: the SAME iterations of the same learning trajectory ----
    reset_state()
    for i in range(args.warmup):
        step_e2e(i)
    ms_e2e = timed_region(step_e2e, args.steps, args.warmup)
    e2e_value = opp_per_step * args.steps / (ms_e2e * 1e-3)
    state_bytes = 3 * R * A * I * K * 4 if learnt else 0
    h2d = state_bytes
    d2h = state_bytes + R * A * _lib.NUM_METRICS * 8 + R * 8

    # ---- the whole trajectory: N = 100 iterations from the initial state ----
    def trajectory(n_full, mode):
        """n_full iterations from the initial state with allocator fit mode `mode`, timed as one region (max over ranks)."""
        reset_state()
        per_it = []
        info_keep = []
        barrier()
        torch.cuda.synchronize(dev)
        s_all, e_all = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_all.record(stream)
        fork_streams()
        for it in range(n_full):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(subs[0].stream)
            acc_ep = []
            for sb in subs:
                with torch.cuda.stream(sb.stream):
                    sb.eng.clear_iteration()
                    sb.eng.simulate(SEED, it, T)
                    if learnt:
                        info = sb.eng.update_allocators(want_info=True, fit_mode=mode)
                        ran = info[..., 1]
                        acc_ep.append(torch.stack([ran.sum(dtype=torch.float64), (ran > 0).sum().to(torch.float64)]))
                    read_out(sb, slot=it)
            if learnt:
                info_keep.append(acc_ep)
            b.record(subs[0].stream)
            per_it.append((a, b))
        gather_history(n_full)
        join_streams()
        e_all.record(stream)
        torch.cuda.synchronize(dev)
        barrier()
        ms_full = torch.tensor([s_all.elapsed_time(e_all)], dtype=torch.float64, device=dev)
        ms_ranks = per_rank(ms_full)
        if world > 1:
            dist.all_reduce(ms_full, op=dist.ReduceOp.MAX)
        ms_full = float(ms_full.item())
        ep = None
        if learnt:
            ep_t = torch.stack([torch.stack(x).sum(dim=0) for x in info_keep])
            if world > 1:
                dist.all_reduce(ep_t)
            ep = [float(x[0] / max(float(x[1]), 1.0)) for x in ep_t.cpu()]
        welfare = float(sum(float(sb.acc_host[..., _lib.M_GROSS].sum()) for sb in subs) / R)  # last iteration, mean over this rank's runs
        return {"iterations": n_full, "opportunities": runs_job * T * n_full, "seconds": ms_full * 1e-3,
                "value": runs_job * T * n_full / (ms_full * 1e-3), "unit": UNIT,
                "ms_per_iteration": [round(a.elapsed_time(b), 3) for a, b in per_it],  # on sub-shard 0's stream
                "fit_epochs_mean_per_iteration": [round(x, 1) for x in ep] if ep else None,
                "welfare_last_iteration_per_run": welfare, "ms_per_rank": ms_ranks}

    # ---- the whole trajectory: N = 100 iterations from the initial state ----
    full, newton = None, None
    if not args.no_full:
        full = trajectory(args.full_iterations, fit_mode)
        full.update(target={"value": 1e9, "n_gpus": 8, "source": "BASELINE.json north_star"},
                    note="whole job, max over ranks, device-resident state, per-iteration metric read-out and (N > 1) the one all-gather of all iterations' metric blocks included")
        if learnt and fit_mode != _lib.FIT_NEWTON:
            newton = trajectory(args.full_iterations, _lib.FIT_NEWTON)
            newton.update(fit_mode="newton", fit_passes_mean_per_iteration=newton.pop("fit_epochs_mean_per_iteration"),
                          note="OPT-IN mode AGYM_FIT_NEWTON, a DIFFERENT ALGORITHM from the reference's (regularised Newton solve per item instead "
                               "of the Adam trajectory; csrc/agym_fit_newton.cu): its own learning trajectory, no parity claim, not comparable "
                               "with `value`, `e2e` or the reference arm; reported because BASELINE.json's north star names Newton / Laplace fits")

    # ---- roofline bookkeeping ----
    peak, peak_src = measured_peaks()
    rows_per_fit = T / A
    bytes_rounds = R * T * (4 * Do + 4) if learnt else 0          # fused loop: winner record only (20 B/opportunity)
    bytes_fit = R * T * (4 * Do + 4) + R * A * I * K * 4 * 7     # K6: winner records once + m,q,m_prev read, m,q,m_prev,sigma written
    kernels = {}
    raw_ms = dict(k_ms)
    if NS > 1 and (k_ms["rounds"] + k_ms["fit"]) > 0:
        # The sub-shards' streams overlap, so the event intervals of a step add up to more than the step.  Each kernel is charged its
        # SHARE (of the two sums) of the measured step; the raw sums stay in the line as ms_summed_over_subshards.
        tot = k_ms["rounds"] + k_ms["fit"]
        k_ms = {k: v / tot * (ms_total / args.steps) for k, v in k_ms.items()}
    if k_ms["rounds"] > 0:
        kernels["sim_kernel (fused K1-K5)"] = {"ms": k_ms["rounds"], "share": k_ms["rounds"] / (k_ms["rounds"] + k_ms["fit"]),
                                                "algorithmic_bytes": bytes_rounds, "achieved_gbs": bytes_rounds / k_ms["rounds"] / 1e6,
                                                "opportunities_per_s": R * T / k_ms["rounds"] * 1e3,
                                                # production mode draws the Thompson noise in logit space: one normal per (participant, item)
                                                "normals_per_s": (R * T * P * I / k_ms["rounds"] * 1e3) if learnt else 0.0,
                                                "sigmoids_per_s": R * T * P * (2 * I + 1) / k_ms["rounds"] * 1e3,
                                                "bound": "issue (exp, FMA, Philox + Box-Muller) and L1/L2 reads of the learnt state; HBM traffic is the 20 B/opportunity winner record",
                                                "from_profile": profile_numbers(r"sim_kernel", None)}
    if learnt and k_ms["fit"] > 0:
        fit_epochs_total = float(es[0]) / max(world, 1)
        kernels["fit kernels (K6)"] = {"ms": k_ms["fit"], "share": k_ms["fit"] / (k_ms["rounds"] + k_ms["fit"]), "ms_per_step": per_step_ms["fit"],
                                       "algorithmic_bytes": bytes_fit, "achieved_gbs": bytes_fit / k_ms["fit"] / 1e6,
                                       "fits_per_s": R * A / k_ms["fit"] * 1e3, "rows_per_fit": rows_per_fit,
                                       "fit_epochs_per_s": fit_epochs_total / (k_ms["fit"] * args.steps) * 1e3,
                                       "bound": "instruction issue: thousands of sequential Adam epochs per fit on register / shared-memory resident state",
                                       "from_profile": profile_numbers(r"fit_warp_kernel<3", Rs * A),
                                       "subshards": NS}
    for name, v in kernels.items():
        v["ms_note"] = ("CUDA events on the launching stream around the kernel(s), mean over the timed steps" +
                        ("; the sub-shards' streams overlap, so `ms` = this kernel's share of the summed intervals x the measured step "
                         "(ms_per_step entries and ms_summed_over_subshards are the raw sums)" if NS > 1 else ""))
        if NS > 1:
            v["ms_summed_over_subshards"] = raw_ms["rounds" if name.startswith("sim_kernel") else "fit"]
    aux = {}
    if not args.no_aux:
        # staged resolution kernel K4(+K5) on this GPU's resident runs: HBM-bound.  With the per-agent accumulation it moves
        # 13P + 10 = 36 B/opportunity; resolution + click alone never reads the values and agent ids: 8P + 10 = 26 B
        # (all of them in ONE launch, whatever the sub-shard split of the timed steps: its own engine when NS > 1)
        Rk = R
        engk = eng
        if NS > 1:
            engk = make_engine(ag, _lib, R, T, learnt, local_rank, first_run)
            if learnt:
                engk.set_allocator_state(initial_m(first_run, R))
        engk.clear_iteration()
        b = engk.staged_round(SEED, 0, T)
        flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # 4x the 126 MB L2
        for accumulate in (False, True):
            for _ in range(3):
                engk.k4_resolve(SEED, 0, T, b, accumulate)
            torch.cuda.synchronize(dev)
            ts = []
            for _ in range(10):
                flush.fill_(1)  # evict K4's inputs from L2 so every timed launch reads HBM
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(stream); engk.k4_resolve(SEED, 0, T, b, accumulate); e.record(stream)
                torch.cuda.synchronize(dev)
                ts.append(s.elapsed_time(e))
            ms4 = float(np.mean(ts))
            by = Rk * T * ((13 * P + 10) if accumulate else (8 * P + 10))
            k4_grid = Rk * ((T + 2047) // 2048) if accumulate else (Rk * T // 4 + 255) // 256  # csrc/agym_staged.cu launch_k4
            prof = profile_numbers(r"k4_kernel_p2_acc" if accumulate else r"k4_kernel_p2\b", k4_grid)
            aux["k4_resolve" + ("+accumulate" if accumulate else "")] = {
                "bound": "hbm", "achieved": by / ms4 / 1e6, "peak": peak, "unit": "GB/s", "frac": by / ms4 / 1e6 / peak,
                "ms": ms4, "algorithmic_bytes": by, "opportunities_per_s": Rk * T / ms4 * 1e3,
                "traffic": prof["dram_bytes"] if prof and prof.get("traffic_matches_this_grid") else None, "from_profile": prof,
                "bytes_per_opportunity": (13 * P + 10) if accumulate else (8 * P + 10),
                "note": f"{Rk * T} opportunities, {by / 1e6:.0f} MB algorithmic; L2 flushed (512 MB fill) before every timed launch; outside the timed step"}
        del b, flush
        # the fused round loop alone, all resident runs in one launch, nothing else on the GPU (inside the timed steps the
        # sub-shards' streams overlap, so the per-kernel intervals there are stretched by the other streams' kernels)
        ts = []
        for i in range(5):
            engk.clear_iteration()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream); engk.simulate(SEED, args.warmup + i, T); e.record(stream)
            torch.cuda.synchronize(dev)
            ts.append(s.elapsed_time(e))
        ms_r = float(np.mean(ts[2:]))
        aux["round_loop_alone"] = {"ms": ms_r, "opportunities_per_s": Rk * T / ms_r * 1e3, "runs": Rk,
                                   "bound": "issue / L1-L2 latency (see sim_kernel)", "from_profile": profile_numbers(r"sim_kernel", None),
                                   "note": "sim_kernel on all resident runs, one launch, alone on the GPU, outside the timed step"}
        if engk is not eng:
            engk.close()
    dominant = max(kernels.items(), key=lambda kv: kv[1]["ms"])
    dprof = dominant[1].get("from_profile")
    roofline = {"kernel": dominant[0], "bound": "hbm", "achieved": dominant[1]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": dominant[1]["achieved_gbs"] / peak,
                "traffic": dprof["dram_bytes"] if dprof and dprof.get("traffic_matches_this_grid") else None, "peak_source": peak_src,
                "from_profile": dprof,
                "note": "the dominant kernel is not HBM-bound (" + dominant[1]["bound"] + "); its algorithmic HBM bytes are tiny by design. "
                        "The HBM-bound kernel of the path is the staged resolution kernel: see roofline_kernels.k4_resolve"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import ref_bench

        kind = cpu_kind()
        # a bounded sample (about 20 s on one core) spread over the timed iterations: 4 fits of 4 iterations + 500 rounds
        its = sorted({args.warmup + int(round(f * (args.steps - 1))) for f in (0.0, 0.33, 0.67, 1.0)})
        rs = [ref_bench.run(kind, A, I, D, Do, P, T, iteration=i, n_rounds=500 // len(its), fits_per_worker=1 if learnt else 0, workers=1, learnt=learnt, seed=i)
              for i in its]
        round_rate = float(np.mean([r["round_only_per_core"] for r in rs]))
        fit_s = float(np.mean([r["fit_seconds_mean"] for r in rs]))
        cpu_val = T / (T / round_rate + A * fit_s) if learnt else round_rate
        cpu = {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"iterations {its} of this trajectory: {500 // len(its)} rounds of simulate_opportunity at the bench shape + 1 allocator fit on the dumped inputs "
                         f"(tests/golden/bench_fit_inputs.npz) each, one core, one torch thread; value = T / (T / round_rate + A * fit_seconds)",
               "round_loop_only": round_rate, "fit_seconds": fit_s, "fit_epochs": float(np.mean([r["fit_epochs_mean"] for r in rs])),
               "host_cores": os.cpu_count(), "ideal_all_cores": cpu_val * (os.cpu_count() or 1)}

    shipped = None
    if rank == 0 and world == 1 and not args.no_aux:
        # BASELINE.json configs[1] through the reference-facing driver (parse_config -> Auction / Agent -> five CSV tables'
        # worth of metrics): 3 runs x 20 iterations x 10 000 rounds, 6 agents x 12 items, 360 allocator fits.  Wall clock,
        # everything included (engine construction, catalog upload, fits, per-iteration metric read-out); second of two runs.
        cfg_path = os.path.join(ROOT, "config", "SP_Truthful_TS.json")
        walls = []
        for _ in range(2):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            res = ag.run_experiment(cfg_path, device=local_rank)
            torch.cuda.synchronize(dev)
            walls.append(time.perf_counter() - t0)
        c = res["config"]
        n_opp = int(c.get("num_runs", res["metrics"].shape[0])) * c["num_iter"] * c["rounds_per_iter"]
        shipped = {"config": "config/SP_Truthful_TS.json", "opportunities": n_opp, "wall_s": walls[-1], "value": n_opp / walls[-1], "unit": UNIT,
                   "reference_surveyed": {"wall_s": 1201, "value": 500, "where": "BASELINE.md section 2: python src/main.py config/SP_Truthful_TS.json on 8 host cores (surveyor-measured, not published)"}}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_dict(args, world, runs_job, scaling, args.warmup, args.steps),
                "fit_epochs_mean": fit_epochs_mean,
                "ms_per_rank": {"value_region": rank_ms[0] if rank_ms else None, "e2e_region": rank_ms[1] if len(rank_ms) > 1 else None,
                                "note": "each rank's own time for the timed region, its closing all-gather included (the collective aligns the ranks, so "
                                        "the spread that is left is what happens after it); the reported time is the maximum"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                        "note": "same iterations of the same trajectory as `value` (restart from the initial host state, same warm-up); every step "
                                "uploads the learnt state from pinned host memory and reads state + metrics back; a sub-shard's runs wait on the host "
                                "for their own previous step's read-back, not for the other sub-shards'"},
                "gpu_launches": int(launches), "gpu_launches_note": "agym_launch_count difference over the timed region on rank 0 (sim_kernel, bucket_kernel, "
                                                                     "fit_classify_kernel, fit_order_kernel, two fit_warp_kernel instantiations, pack_state_kernel per step and sub-shard)",
                "subshards": NS,
                "round_loop": ({"value": aux["round_loop_alone"]["opportunities_per_s"] * world, "unit": UNIT, "ms": aux["round_loop_alone"]["ms"],
                                "note": "the fused round loop alone (roofline_kernels.round_loop_alone), per GPU x n_gpus"} if "round_loop_alone" in aux else
                               {"value": runs_job * T / k_ms["rounds"] * 1e3 if k_ms["rounds"] else None, "unit": UNIT, "ms": k_ms["rounds"],
                                "note": "CUDA-event interval inside the timed steps" + ("; stretched by the other sub-shards' kernels" if NS > 1 else "")}),
                "full_workload": full, "opt_in_newton_mode": newton,
                "roofline": roofline, "roofline_kernels": {**kernels, **aux}, "cpu_baseline": cpu, "shipped_config": shipped, "clocks": clk}
        print(json.dumps(line), flush=True)
    for sb in subs:
        sb.eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
