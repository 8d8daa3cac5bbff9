#!/usr/bin/env python
"""Benchmark of the consensus-selection hot path (BASELINE.json metric: genome bins/s, score+solve+budget).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, all host cores

Workload (config.workload): synthetic hg38, 50 bp bins (61,765,409 bins, 24 chromosomes) x 100 samples,
float64, per-chromosome (budget, gamma) from the reference's hg_params.csv, chromosomes LPT-packed
over the ranks (strong scaling: the genome is fixed).  One step = every chromosome of the genome
scored (log2p1 -> baseline -> WLS), budget-searched, solved and emitted as merged BED text.
`value` times that with the count matrices resident in HBM; `e2e` times the reference-facing host
API (NumPy in / NumPy + BED files out) with host<->device copies inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from rocco_b200.synth import HG38_SIZES, HG_PARAMS, chrom_bins, chrom_matrix_numpy, chrom_seed  # noqa: E402

METRIC = "genome_bins_per_sec"
UNIT = "bins/s"
PRIOR_DF = 6.0          # the reference CLI default (rocco.py:569-574)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--step-bp", type=int, default=50)
    ap.add_argument("--chroms", default="all", help="comma list (debug); default = all 24 hg38 chromosomes")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="dtype of the resident count matrices")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="-1: min(steps, 2); 0 disables the e2e leg")
    ap.add_argument("--score-streams", type=int, default=1, help="host threads / CUDA streams scoring chromosomes concurrently")
    ap.add_argument("--e2e-threads", type=int, default=4, help="host threads driving chromosomes through the public API")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-extra-legs", dest="e2e_extra_legs", action="store_false",
                    help="skip the pinned-input and float32-input end-to-end legs (the pageable float64 leg is the headline)")
    ap.add_argument("--cpu-sample-bins", type=int, default=250_000)
    ap.add_argument("--levels", type=int, default=0, help="bisection levels per launch (0 = library default)")
    ap.add_argument("--config", type=int, default=4, choices=[4, 5],
                    help="BASELINE.json config: 4 = hg38 @ 50 bp x 100 samples, chromosome-sharded (default, the metric's config); "
                         "5 = hg38 @ 20 bp x 1000 samples float32, SAMPLE-sharded scoring + 256-multiplier sweep")
    ap.add_argument("--multipliers", type=int, default=256)
    return ap.parse_args()


def workload(args):
    names = list(HG38_SIZES) if args.chroms == "all" else args.chroms.split(",")
    bins = [chrom_bins(c, args.step_bp) for c in names]
    return names, bins


def workload_name(args, names):
    g = "hg38" if len(names) == 24 else "+".join(names)
    return f"synthetic {g} @ {args.step_bp} bp x {args.samples} samples, {args.dtype}, hg_params budgets/gammas"


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        self.gpus = ",".join(str(g) for g in gpu_indices)
        self.proc = None
        self.lines: list[tuple[float, str]] = []
        self.window = None

    def start(self):
        # ONE nvidia-smi for all GPUs of the job, started before the warm-up steps: its start-up (NVML initialisation takes
        # driver-wide locks for ~100 ms) must not fall into the timed region, least of all once per rank
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", self.gpus, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self, t0: float, t1: float):
        """host-clock window of the timed region: only samples that arrived inside it are reported"""
        self.window = (t0, t1)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        lines = [ln for t, ln in self.lines if self.window is None or self.window[0] <= t <= self.window[1] + 0.1]
        scope = "timed region"
        if not lines:
            lines, scope = [ln for _, ln in self.lines], "warm-up + timed region (timed region shorter than the sampling period)"
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "gpus": self.gpus, "sampled_over": scope}


# ---------------------------------------------------------------------------------------------- reference arm
def _ref_worker(task, keep=False):
    """One worker = the reference's own kernels (oracle/_ref, else the oracle port) on one slice."""
    m, n, seed, budget, gamma, kind = task
    from oracle import oracle as orc
    x = chrom_matrix_numpy(m, n, seed=seed)
    t0 = time.perf_counter()
    scores = orc.score_loci_wls(x, prior_df=PRIOR_DF, kind=kind)
    sol, obj, det = orc.solve_chrom_exact(scores, budget=budget, gamma=gamma, kind=kind, return_details=True)
    recs = orc.solution_to_records("chr21", np.arange(0, 50 * n, 50), sol)
    dt = time.perf_counter() - t0
    if keep:
        return n, dt, len(recs), {"x": x, "scores": scores, "solution": sol, "objective": obj, "details": det, "records": recs}
    return n, dt, len(recs)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    names, bins = workload(args)
    kind = "reference" if orc.reference_available() else "port"
    orc.native(kind)             # mapped in THIS process (and inherited by the forked workers): the run's loaded-library record shows it
    cores = os.cpu_count() or 1
    import multiprocessing as mp
    # per-step sample: one slice per core, sized for ~4-6 s of single-core work each
    n_slice = max(2_000, int(4.5 / (4.5e-7 * args.samples)))
    budget, gamma = HG_PARAMS["chr21"]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        def step(k):
            tasks = [(args.samples, n_slice, 1000 * k + w, budget, gamma, kind) for w in range(cores)]
            res = pool.map(_ref_worker, tasks)
            # all slices run concurrently, one per core: the step takes as long as the slowest worker's compute
            # (synthetic-data generation inside the workers is not part of the path and is excluded)
            return sum(r[0] for r in res), max(r[1] for r in res)
        for k in range(args.warmup):
            step(k)
        tot_bins, tot_t = 0, 0.0
        for k in range(args.steps):
            b, t = step(100 + k)
            tot_bins += b
            tot_t += t
    value = tot_bins / tot_t
    sample = (f"{cores} slices/step of {n_slice} bins x {args.samples} samples (chr21 generator, budget {budget}, gamma {gamma}): "
              f"score_loci_wls + solve_chrom_exact + BED records, one process per core")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, names), "genome_bins": int(sum(bins)), "bounded_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)



# ---------------------------------------------------------------------------------------------- config 5
def run_config5(args):
    """BASELINE.json config 5: hg38 @ 20 bp x 1000 samples stored float32, the SAMPLES sharded over the ranks.

    Per chromosome: every rank runs the per-sample stages on its rows (-> four per-bin accumulators), ONE all-reduce
    (sum, float64) of [4, bins] over NVLink, the per-bin finalisation, and on the rank that owns the chromosome (LPT) a
    sort for the multiplier grid plus a 256-multiplier sweep in one launch set.  The all-reduce / finalise / sweep chain
    of chromosome c runs on a second stream from a helper thread while the main thread scores chromosome c+1.
    The matrices stay resident, so at fewer than 8 GPUs the run covers the chromosomes that fit (named in the workload)."""
    import queue
    import torch
    import torch.distributed as dist
    from rocco_b200 import _lib, pipeline
    from rocco_b200.synth import chrom_matrix_torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    step_bp = 20 if args.step_bp == 50 else args.step_bp
    samples = 1000 if args.samples == 100 else args.samples
    assert samples % world == 0, "samples must divide over the ranks"
    m_local = samples // world
    all_names = list(HG38_SIZES) if args.chroms == "all" else args.chroms.split(",")
    all_bins = {c: chrom_bins(c, step_bp) for c in all_names}
    # resident inputs <= 80 GB per rank, scratch (centred matrix + variance track, float64) of the largest chromosome <= 80 GB
    n_max = int(80e9 / (16.0 * m_local))
    names, used = [], 0.0
    for c in sorted(all_names, key=lambda c: all_bins[c]):
        need = all_bins[c] * m_local * 4.0
        if all_bins[c] <= n_max and used + need <= 80e9:
            names.append(c)
            used += need
    names = [c for c in all_names if c in names]
    bins = [all_bins[c] for c in names]
    owner = {}
    for r, part in enumerate(pipeline.lpt_partition(bins, world)):
        for k in part:
            owner[k] = r
    prm = pipeline.score_params(prior_df=PRIOR_DF)
    mats = [chrom_matrix_torch(m_local, n, chrom_seed(c), dev, torch.float32, sample_stream=rank + 1) for c, n in zip(names, bins)]
    comm = torch.cuda.Stream(device=dev)
    total_bins = int(sum(bins))
    sweeps = {}

    def tail_worker(q):
        torch.cuda.set_device(local_rank)
        with torch.cuda.stream(comm):
            while True:
                item = q.get()
                if item is None:
                    return
                k, acc, ready = item
                comm.wait_event(ready)
                if world > 1:
                    dist.all_reduce(acc)
                scores = pipeline.score_finalize_device(acc, samples, prm)
                if owner[k] == rank:
                    n = scores.shape[0]
                    srt = torch.sort(scores).values
                    lo, hi = float(srt[int(0.50 * (n - 1))]), float(srt[int(0.999 * (n - 1))])
                    lam = np.linspace(max(lo, 0.0), hi, args.multipliers)
                    counts, pen, obj = pipeline.sweep_multipliers(scores, HG_PARAMS[names[k]][1], lam)
                    sweeps[names[k]] = (int(counts[0]), int(counts[-1]))

    def step():
        q = queue.Queue()
        th = threading.Thread(target=tail_worker, args=(q,))
        th.start()
        for k, x in enumerate(mats):
            acc = pipeline.score_partial_device(x, prm)
            ready = torch.cuda.Event()
            ready.record()
            acc.record_stream(comm)
            q.put((k, acc, ready))
        q.put(None)
        th.join()
        torch.cuda.current_stream().wait_stream(comm)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(range(world)) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = _lib.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    t_host1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.kernel_launches() - launches0
    clocks = None
    if sampler:
        sampler.mark(t_host0, t_host1)
        clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt)
    ms = float(t.item())
    # the collective on its own: [4, bins] float64 of the largest resident chromosome, ranks aligned by a barrier
    ar = None
    if world > 1:
        buf = torch.zeros((4, max(bins)), dtype=torch.float64, device=dev)
        for _ in range(2):
            dist.all_reduce(buf)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            dist.all_reduce(buf)
        a1.record()
        torch.cuda.synchronize()
        at = torch.tensor([a0.elapsed_time(a1) / 5.0], dtype=torch.float64, device=dev)
        dist.all_reduce(at, op=dist.ReduceOp.MAX)
        nbytes = buf.numel() * 8
        ar = {"bytes": nbytes, "ms": float(at.item()), "algbw_GBps": nbytes / 1e9 / (float(at.item()) / 1e3),
              "busbw_GBps": 2.0 * (world - 1) / world * nbytes / 1e9 / (float(at.item()) / 1e3),
              "reference_busbw_GBps": 725.0, "per_step_bytes": total_bins * 32}
    if rank == 0:
        subset = "hg38" if len(names) == 24 else "+".join(names)
        line = {
            "metric": METRIC, "value": total_bins * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"config 5: synthetic {subset} @ {step_bp} bp x {samples} samples stored float32, scoring "
                                   f"sample-sharded {m_local} per rank, {args.multipliers}-multiplier sweep per chromosome",
                       "genome_bins": total_bins, "samples": samples, "chromosomes": len(names),
                       "resident_input_GB_per_rank": used / 1e9,
                       "sample_bins_per_sec": total_bins * samples * args.steps / (ms / 1e3),
                       "l2_policy": "inputs larger than L2", "collective": "one NCCL all-reduce(sum, float64) of [4, bins] per chromosome, "
                       "overlapped with the scoring of the next chromosome", "all_reduce": ar,
                       "sweep_counts_first_last": sweeps},
            "clocks": clocks, "e2e": None, "gpu_launches": int(lt.item()), "roofline": None, "cpu_baseline": None,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- our arm
def cpu_baseline_single_core(args):
    from oracle import oracle as orc
    kind = "reference" if orc.reference_available() else "port"
    n = int(args.cpu_sample_bins * min(1.0, 100.0 / max(args.samples, 1)))
    budget, gamma = HG_PARAMS["chr21"]
    n_done, dt, _, ref = _ref_worker((args.samples, n, chrom_seed("chr21"), budget, gamma, kind), keep=True)
    cpu = {"value": n_done / dt, "unit": UNIT, "cores": 1, "kind": kind,
           "sample": f"first {n} bins of synthetic chr21 x {args.samples} samples: score_loci_wls + solve_chrom_exact "
                     f"(budget {budget}, gamma {gamma}) + BED records, {dt:.1f} s on one core"}
    return cpu, parity_against(ref, budget, gamma, kind)


def parity_against(ref, budget, gamma, kind):
    """The GPU path on the SAME bytes the CPU baseline just processed, compared output by output (SURVEY.md 8d gates)."""
    import rocco_b200
    from rocco_b200.rocco import solution_runs
    x = ref["x"]
    scores = rocco_b200.score_loci_wls(x, prior_df=PRIOR_DF)
    sol, obj, det = rocco_b200.solve_chrom_exact(scores, budget=budget, gamma=gamma, return_details=True)
    starts, ends = solution_runs(sol)
    n = x.shape[1]
    recs = [("chr21", int(50 * a), int(50 * b)) for a, b in zip(starts.tolist(), ends.tolist())]
    want = ref["scores"]
    rel = float(np.max(np.abs(scores - want) / np.maximum(np.abs(want), 1e-3)))
    wdet = ref["details"]
    # the multiplier search on identical scores: the searched double itself
    sol2, obj2, det2 = rocco_b200.solve_chrom_exact(want, budget=budget, gamma=gamma, return_details=True)
    # ... and with the opt-in exact search (decisions replayed in the reference's own operation order near ties)
    from rocco_b200 import _lib
    prev = _lib.load().rocco_b200_chain_set_exact_search(1)
    try:
        sol3, obj3, det3 = rocco_b200.solve_chrom_exact(want, budget=budget, gamma=gamma, return_details=True)
    finally:
        _lib.load().rocco_b200_chain_set_exact_search(prev)
    return {
        "against": f"oracle kind={kind}, same input bytes ({x.shape[0]} x {n})",
        "scores_max_rel_err": rel, "scores_tolerance": 1e-5,
        "mask_identical": bool(np.array_equal(sol, ref["solution"])), "mask_differing_bins": int(np.sum(sol != ref["solution"])),
        "selected_count": [int(det["selected_count"]), int(wdet["selected_count"])],
        "intervals_identical": [tuple(r) for r in recs] == [tuple(r) for r in ref["records"]],
        "objective_rel_err": float(abs(obj - ref["objective"]) / max(abs(ref["objective"]), 1e-300)),
        "lambda_abs_diff": float(abs(det["selection_penalty"] - wdet["selection_penalty"])),
        "same_scores_search": {"mask_identical": bool(np.array_equal(sol2, ref["solution"])),
                               "lambda_identical": bool(det2["selection_penalty"] == wdet["selection_penalty"]),
                               "lambda_abs_diff": float(abs(det2["selection_penalty"] - wdet["selection_penalty"]))},
        "same_scores_exact_search": {"mask_identical": bool(np.array_equal(sol3, ref["solution"])),
                                     "lambda_identical": bool(det3["selection_penalty"] == wdet["selection_penalty"]),
                                     "objective_identical": bool(obj3 == ref["objective"])},
    }


_RESULT_FD = None


def emit(line: dict) -> None:
    """The one JSON line goes to the process's ORIGINAL stdout; see main() for why fd 1 is redirected meanwhile."""
    try:                                       # the literal metric string of BASELINE.json, next to this line's short name for it
        line.setdefault("baseline_metric", json.load(open(os.path.join(REPO, "BASELINE.json")))["metric"])
    except (OSError, KeyError, ValueError):
        pass
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    args = parse_args()
    # Libraries print to stdout behind Python's back (NCCL's "NCCL version ..." banner at communicator creation): keep a
    # private handle on the real stdout for the JSON line and point fd 1 at stderr for everything else.
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return
    if args.config == 5:
        run_config5(args)
        return

    import torch
    import torch.distributed as dist

    import rocco_b200
    from rocco_b200 import _lib, pipeline
    from rocco_b200.synth import chrom_matrix_torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    names, bins = workload(args)
    parts = pipeline.lpt_partition(bins, world)
    mine = parts[rank]
    my_names = [names[i] for i in mine]
    my_bins = [bins[i] for i in mine]
    budgets = [HG_PARAMS[c][0] for c in my_names]
    gammas = [HG_PARAMS[c][1] for c in my_names]
    tdtype = torch.float64 if args.dtype == "f64" else torch.float32
    esz = 8 if args.dtype == "f64" else 4
    d_mats = [chrom_matrix_torch(args.samples, n, chrom_seed(c), dev, tdtype) for c, n in zip(my_names, my_bins)]
    params = pipeline.score_params(prior_df=PRIOR_DF)
    genome_bins = int(sum(bins))
    from rocco_b200 import distributed as rdist
    tmp_holder = [tempfile.mkdtemp(prefix="rocco_b200_bench_") if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(tmp_holder, src=0)          # one node: every rank sees rank 0's directory
    tmpdir = tmp_holder[0]
    count_buf = torch.zeros(2, dtype=torch.int64, device=dev)
    # record order of the reference's combined BED (rocco.py:74-95 sorts by the chromosome STRING: chr1 < chr10 < chr2)
    lex_names = sorted(names)
    lex_order = sorted(range(len(my_names)), key=lambda k: my_names[k])      # shard-local chromosome indices in that order

    host_phase = {}

    def step():
        t_a = time.perf_counter()
        shard = pipeline.run_shard(d_mats, budgets, gammas, params=params, levels_per_round=args.levels,
                                   score_streams=args.score_streams) if mine else None
        host_phase["run_shard_ms"] = 1e3 * (time.perf_counter() - t_a)     # returns once the runs are on the host
        t_a = time.perf_counter()
        # ONE merged BED for the genome in the reference's record order.  world == 1: the runs are regrouped and written in
        # one call.  world > 1: every rank announces the text size of each of its chromosomes in the step's one collective
        # (an all-gather of [selected, bins, 24 sizes] per rank), derives every chromosome's byte offset in the genome file,
        # and writes its own chromosomes there with positioned writes -- no part files, no gather pass on rank 0.
        genome_bed = os.path.join(tmpdir, "genome.bed")
        sel_mine = sum(r["selected_count"] for r in shard["results"]) if mine else 0
        if world == 1:
            if mine:
                pipeline.runs_to_bed_file(genome_bed, lex_names, pipeline.reorder_runs(shard["runs"], lex_order), args.step_bp)
            host_phase["bed_ms"] = 1e3 * (time.perf_counter() - t_a)
            count_buf[0] = sel_mine
            count_buf[1] = sum(my_bins)
        else:
            empty = (np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int64))
            tot = rdist.write_genome_bed(genome_bed, lex_names, my_names, shard["runs"] if mine else empty, args.step_bp,
                                         extras=(sel_mine, sum(my_bins)), device=dev)
            count_buf[0] = tot[0]
            count_buf[1] = tot[1]
            host_phase["bed_ms"] = 1e3 * (time.perf_counter() - t_a)
        return shard

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(range(world)) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    _lib.profile_enable(True)
    step()                                  # one profiled step outside the timed region fills the event pool
    _lib.profile_report()
    launches0 = _lib.kernel_launches()
    fb_rows0 = int(_lib.load().rocco_b200_trend_fallback_rows())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    ev0.record()
    last = None
    for _ in range(args.steps):
        last = step()
    ev1.record()
    barrier()
    t_host1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.kernel_launches() - launches0
    fb_rows = int(_lib.load().rocco_b200_trend_fallback_rows()) - fb_rows0
    reasons = (ctypes.c_longlong * 8)()
    _lib.load().rocco_b200_trend_fallback_reasons(reasons)
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    clocks = None
    if sampler:
        sampler.mark(t_host0, t_host1)
        clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt)
    ms = float(t.item())
    value = genome_bins * args.steps / (ms / 1e3)
    selected_total = int(count_buf[0].item())
    by_chrom = {c: (r["selected_count"], r["selection_penalty"]) for c, r in zip(my_names, last["results"])} if last else {}
    if world > 1:
        parts_ = [None] * world
        dist.all_gather_object(parts_, by_chrom)
        by_chrom = {k: v for p_ in parts_ for k, v in p_.items()}

    # ---- the column statistics kernel (not on the default CLI path, measured on its own)
    colstat = None
    if rank == 0 and mine:
        k_small = min(range(len(d_mats)), key=lambda k: my_bins[k])
        xm = d_mats[k_small]
        out_cs = torch.empty(xm.shape[1], dtype=torch.float64, device=dev)
        lib = _lib.load()

        def colstat_once(stat):
            st_ = lib.rocco_b200_column_stat_dev(ctypes.c_void_p(xm.data_ptr()), 0 if xm.dtype == torch.float64 else 1, xm.shape[0], xm.shape[1],
                                                 stat, 0.0, 0.0, 1.0, ctypes.c_void_p(out_cs.data_ptr()),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(st_, "column_stat")

        colstat = {"workload": f"{my_names[k_small]} x {xm.shape[0]} samples ({xm.shape[1]} bins), {args.dtype}", "kernels": {}}
        for nm, code in (("median", 0), ("mad", 4)):
            for _ in range(3):
                colstat_once(code)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(5):
                colstat_once(code)
            c1.record()
            torch.cuda.synchronize()
            ms_cs = c0.elapsed_time(c1) / 5.0
            nbytes = xm.numel() * xm.element_size() + 8 * xm.shape[1]
            colstat["kernels"][nm] = {"ms": ms_cs, "algorithmic_bytes": nbytes, "GBps": nbytes / 1e9 / (ms_cs / 1e3),
                                      "bins_per_sec": xm.shape[1] / (ms_cs / 1e3)}

    # ---- end-to-end through the reference-facing host API (NumPy in, BED files out)
    e2e = None
    e2e_steps = min(args.steps, 2) if args.e2e_steps < 0 else args.e2e_steps
    if e2e_steps > 0:
        from concurrent.futures import ThreadPoolExecutor
        cwd = os.getcwd()
        os.chdir(tmpdir)

        def one_chrom(job):
            c, n, x, b, g = job
            scores = rocco_b200.score_loci_wls(x, prior_df=PRIOR_DF)
            sol, obj = rocco_b200.solve_chrom_exact(scores, budget=b, gamma=g)
            return rocco_b200.chrom_solution_to_bed(c, np.arange(0, args.step_bp * n, args.step_bp), sol, ID="bench")

        def e2e_step(host):
            # the reference solves chromosomes in a pool of <= 4 workers (rocco.py:1146-1184); the host threads here keep
            # the PCIe link busy: one chromosome's upload overlaps the kernels / solve / BED writing of the others (ctypes
            # drops the GIL)
            jobs = sorted(zip(my_names, my_bins, host, budgets, gammas), key=lambda j: -j[1])   # longest first: short tail
            with ThreadPoolExecutor(max_workers=args.e2e_threads, initializer=torch.cuda.set_device, initargs=(local_rank,)) as pool:   # the CUDA current device is per host thread
                files = list(pool.map(one_chrom, jobs))
            # ONE combined BED for the genome, as the reference's parent process builds it from its workers' files
            # (rocco.py:194-240): the ranks share the node's file system, rank 0 combines all 24 per-chromosome files
            if world > 1:
                dist.barrier()
            if rank == 0:
                allf = [os.path.join(tmpdir, f"rocco_bench_{c}.bed") for c in names]
                rocco_b200.combine_chrom_results([f for f in allf if os.path.exists(f)], "combined.bed")
            return files

        def e2e_leg(host, warm, steps, label):
            for _ in range(warm):                           # warm-up: the per-lease scratch pools reach their steady-state sizes
                e2e_step(host)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                t_step = time.perf_counter()
                e2e_step(host)
                print(f"[bench] rank {rank} e2e ({label}) step: {1e3 * (time.perf_counter() - t_step):.0f} ms", file=sys.stderr)
            barrier()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return genome_bins * steps / float(tt.item())

        def host_bytes(esz_):
            h2d = sum(args.samples * n * esz_ + n * 8 + n for n in my_bins)     # matrix + scores (solve) + mask (BED)
            d2h = sum(n * 8 + n for n in my_bins)                               # scores + mask
            bb = torch.tensor([h2d, d2h], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(bb)
            return int(bb[0].item()), int(bb[1].item())

        # leg 1 (headline): the caller's arrays are PAGEABLE NumPy memory, which is what score_loci_wls is handed in practice
        host = []
        for x in d_mats:
            h = np.empty(tuple(x.shape), dtype=np.float64 if args.dtype == "f64" else np.float32)
            torch.from_numpy(h).copy_(x)
            host.append(h)
        del d_mats
        torch.cuda.empty_cache()
        pageable = e2e_leg(host, 2, e2e_steps, "pageable")
        h2d, d2h = host_bytes(esz)
        e2e = {"value": pageable, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "pinned_host": False, "host_threads": args.e2e_threads,
               "api": "score_loci_wls + solve_chrom_exact + chrom_solution_to_bed + combine_chrom_results (pageable NumPy in, ONE combined BED out)"}
        if args.e2e_extra_legs:
            # leg 2: the same arrays in pinned memory (the best the link can do; round 1's headline)
            pinned_ok = True
            for k in range(len(host)):
                try:
                    t = torch.empty(host[k].shape, dtype=torch.from_numpy(host[k]).dtype, pin_memory=True)
                except RuntimeError:
                    pinned_ok = False
                    break
                t.copy_(torch.from_numpy(host[k]))
                host[k] = t.numpy()
            if pinned_ok:
                e2e["pinned_input"] = {"value": e2e_leg(host, 1, 1, "pinned"), "unit": UNIT, "steps": 1}
                e2e["pageable_over_pinned"] = e2e["value"] / e2e["pinned_input"]["value"]
            # leg 3: float32 counts, what the reference's --low_memory path hands over (readtracks.py:621)
            if args.dtype == "f64":
                for k in range(len(host)):
                    host[k] = np.ascontiguousarray(host[k], dtype=np.float32)
                h2d32, d2h32 = host_bytes(4)
                e2e["float32_input"] = {"value": e2e_leg(host, 1, 1, "float32 pageable"), "unit": UNIT, "steps": 1,
                                        "h2d_bytes_per_step": h2d32, "d2h_bytes_per_step": d2h32, "pinned_host": False}
        os.chdir(cwd)

    if rank == 0:
        peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # dominant kernel = the scope with the largest share of device time on rank 0
        roof = None
        if prof:
            total_ms = sum(v[0] for v in prof.values())
            dom = max(prof, key=lambda k: prof[k][0])
            d_ms, d_cnt, d_bytes = prof[dom]
            achieved = (d_bytes / 1e9) / (d_ms / 1e3) if d_ms > 0 else 0.0
            traffic, traffic_src = None, None
            tpath = os.path.join(REPO, "profiles", "r2_traffic.json")
            if os.path.exists(tpath):
                tj = json.load(open(tpath))
                if dom in tj["kernels"]:
                    # DRAM bytes per launch = this run's algorithmic bytes per launch x the ncu-measured DRAM/algorithmic ratio
                    traffic = (d_bytes / max(d_cnt, 1)) * tj["kernels"][dom]["traffic_per_algorithmic_byte"]
                    traffic_src = "dram__bytes_read+write per algorithmic byte from " + tj["source"]
            roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": d_bytes / max(d_cnt, 1), "peak_source": peak_src,
                    "launches": d_cnt, "avg_launch_ms": d_ms / max(d_cnt, 1), "share_of_profiled_time": d_ms / total_ms,
                    "all_scopes": {k: {"ms": round(v[0], 3), "launch_sets": v[1],
                                       "GBps_algorithmic": round((v[2] / 1e9) / (v[0] / 1e3), 1) if v[0] > 0 else 0.0}
                                   for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}
        if colstat:
            for v in colstat["kernels"].values():
                v["frac_of_peak"] = v["GBps"] / peak
        cpu, parity = None, None
        if world == 1 and not args.no_cpu_baseline:
            cpu, parity = cpu_baseline_single_core(args)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_name(args, names), "genome_bins": genome_bins, "samples": args.samples,
                       "chromosomes": len(names), "sharding": f"chromosomes LPT-packed over {world} rank(s)", "score_streams": args.score_streams,
                       "l2_policy": "inputs larger than L2 (per-chromosome matrices 0.7-4 GB vs 126 MB L2)",
                       "selected_bins": selected_total,
                       "selected_by_chrom": {c: by_chrom[c][0] for c in names if c in by_chrom},
                       "lambda_by_chrom": {c: by_chrom[c][1] for c in names if c in by_chrom}, "trend_sort_fallback_rows": fb_rows,
                       "trend_fallback_reason_counts": list(reasons)[:5], "collective": "one NCCL all-gather of [selected, bins, per-chromosome BED text sizes] per step",
                       "host_phases_last_step_ms": {k: round(v, 2) for k, v in host_phase.items()}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(lt.item()), "roofline": roof, "cpu_baseline": cpu, "parity": parity, "column_stat": colstat,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
