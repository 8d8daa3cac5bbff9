import sys, time, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
from rocco_b200 import pipeline, _lib
from rocco_b200.synth import chrom_matrix_torch, chrom_bins, HG38_SIZES, HG_PARAMS, chrom_seed
dev = torch.device('cuda', 0)
names = list(HG38_SIZES)
mats = [chrom_matrix_torch(100, chrom_bins(c), chrom_seed(c), dev, torch.float64) for c in names]
budgets = [HG_PARAMS[c][0] for c in names]; gammas = [HG_PARAMS[c][1] for c in names]
prm = pipeline.score_params(prior_df=6.0)
lengths = [int(x.shape[1]) for x in mats]
offsets, total = pipeline.layout_offsets(lengths)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d_scores = torch.zeros(total, dtype=torch.float64, device=dev)
    tsc = []
    for x, off, n in zip(mats, offsets, lengths):
        a = time.perf_counter()
        pipeline.score_loci_wls_device(x, out_scores=d_scores[off:off + n], params=prm)
        tsc.append(time.perf_counter() - a)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    d_masks, results = pipeline.solve_packed(d_scores, offsets, lengths, budgets, gammas)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    runs = pipeline.masks_to_runs(d_masks, offsets, lengths)
    t3 = time.perf_counter()
    pipeline.runs_to_bed_file("/tmp/x.bed", names, runs, 50)
    t4 = time.perf_counter()
    print(f"score {1e3*(t1-t0):.1f} ms (sum of calls {1e3*sum(tsc):.1f}) | solve {1e3*(t2-t1):.1f} | runs {1e3*(t3-t2):.1f} | bed {1e3*(t4-t3):.1f} | total {1e3*(t4-t0):.1f}")
_lib.profile_enable(True); _lib.profile_report()
torch.cuda.synchronize(); t0 = time.perf_counter()
for x, off, n in zip(mats, offsets, lengths):
    pipeline.score_loci_wls_device(x, out_scores=d_scores[off:off + n], params=prm)
torch.cuda.synchronize(); t1 = time.perf_counter()
rep = _lib.profile_report()
print("score wall", 1e3*(t1-t0), "sum of profiled scopes", sum(v[0] for v in rep.values()))
