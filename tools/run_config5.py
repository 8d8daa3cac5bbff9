"""BASELINE.json config 5 at full size: hg38 @ 20 bp (154.4 M bins) x 1000 samples stored float32 (617.7 GB), SAMPLES sharded
over the ranks (125 per GPU at 8), one all-reduce of the four per-bin accumulators per chromosome, then a 256-multiplier sweep
per chromosome on the rank that owns it (LPT).  Chromosomes are generated, scored and dropped one at a time, so a rank never
holds more than one chromosome's matrix + scratch (~35 GB).

    torchrun --nproc-per-node 8 tools/run_config5.py [--samples 1000] [--step-bp 20] [--chroms chr1,chr2]
"""
import argparse, json, math, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocco_b200 import pipeline, _lib
from rocco_b200.synth import HG38_SIZES, HG_PARAMS, chrom_matrix_torch, chrom_seed

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=1000)
ap.add_argument("--step-bp", type=int, default=20)
ap.add_argument("--chroms", default="")
ap.add_argument("--multipliers", type=int, default=256)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
names = [c for c in HG38_SIZES if not args.chroms or c in args.chroms.split(",")]
bins = [int(math.ceil(HG38_SIZES[c] / args.step_bp)) for c in names]
m_local = args.samples // world
owner = {}
for r, part in enumerate(pipeline.lpt_partition(bins, world)):
    for k in part:
        owner[k] = r
prm = pipeline.score_params(prior_df=6.0)
ev = lambda: torch.cuda.Event(enable_timing=True)
t_score = t_reduce = t_final = t_sweep = 0.0
gen = 0.0
selected = {}
for k, (c, n) in enumerate(zip(names, bins)):
    g0 = time.perf_counter()
    x = chrom_matrix_torch(m_local, n, chrom_seed(c), dev, torch.float32, sample_stream=rank + 1)
    torch.cuda.synchronize(); gen += time.perf_counter() - g0
    if world > 1:
        dist.barrier()
    e = [ev() for _ in range(5)]
    e[0].record()
    acc = pipeline.score_partial_device(x, prm)
    e[1].record()
    if world > 1:
        dist.all_reduce(acc)
    e[2].record()
    scores = pipeline.score_finalize_device(acc, m_local * world, prm)
    e[3].record()
    if owner[k] == rank:
        srt = torch.sort(scores).values
        lo, hi = float(srt[int(0.50 * (n - 1))]), float(srt[int(0.999 * (n - 1))])
        lam = np.linspace(max(lo, 0.0), hi, args.multipliers)
        counts, pen, obj = pipeline.sweep_multipliers(scores, HG_PARAMS[c][1], lam)
        selected[c] = (int(counts[0]), int(counts[-1]))
    e[4].record()
    torch.cuda.synchronize()
    t_score += e[0].elapsed_time(e[1]); t_reduce += e[1].elapsed_time(e[2]); t_final += e[2].elapsed_time(e[3])
    t_sweep += e[3].elapsed_time(e[4])
    del x, acc, scores
t = torch.tensor([t_score, t_reduce, t_final, t_sweep, t_score + t_reduce + t_final + t_sweep, gen * 1e3], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    tot_bins = sum(bins)
    ms = t.tolist()
    print(json.dumps({"config": f"hg38 @ {args.step_bp} bp x {m_local * world} samples f32, sample-sharded over {world} GPUs, "
                                f"{args.multipliers}-multiplier sweep", "bins": tot_bins, "input_GB": tot_bins * m_local * world * 4 / 1e9,
                      "ms_max_over_ranks": {"score_partial": ms[0], "all_reduce_4xn_f64": ms[1], "finalize": ms[2],
                                            "sort+sweep (owner rank)": ms[3], "total": ms[4], "synthetic_generation(untimed)": ms[5]},
                      "bins_per_s": tot_bins / (ms[4] / 1e3), "sample_bins_per_s": tot_bins * m_local * world / (ms[4] / 1e3),
                      "allreduce_bytes": tot_bins * 32, "sweep_counts_first_last(rank0 chroms)": selected}))
if world > 1:
    dist.destroy_process_group()
