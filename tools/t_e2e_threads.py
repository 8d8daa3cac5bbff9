import sys, time, os, numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, '/root/repo')
import rocco_b200
from rocco_b200.synth import chrom_matrix_torch, chrom_bins, HG_PARAMS
dev = torch.device('cuda', 0)
names = ["chr1", "chr5", "chr9", "chr13", "chr17", "chr21"]
host = []
for i, c in enumerate(names):
    x = chrom_matrix_torch(100, chrom_bins(c), i, dev, torch.float64)
    h = torch.empty(x.shape, dtype=x.dtype, pin_memory=True); h.copy_(x); host.append(h); del x
torch.cuda.synchronize()
nbytes = sum(h.numel() * 8 for h in host)
d = [torch.empty(h.shape, dtype=h.dtype, device=dev) for h in host]
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for a, b in zip(d, host): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"pure H2D {nbytes/1e9:.1f} GB: {dt*1e3:.0f} ms = {nbytes/dt/1e9:.1f} GB/s")
del d; torch.cuda.empty_cache()
arrs = [h.numpy() for h in host]
os.chdir("/tmp")
def score_only(k): return rocco_b200.score_loci_wls(arrs[k], prior_df=6.0)
def full(k):
    s = rocco_b200.score_loci_wls(arrs[k], prior_df=6.0)
    b, g = HG_PARAMS[names[k]]
    sol, obj = rocco_b200.solve_chrom_exact(s, budget=b, gamma=g)
    return rocco_b200.chrom_solution_to_bed(names[k], np.arange(0, 50 * len(sol), 50), sol, ID="t")
for fn in (score_only, full):
    for T in (1, 2, 3):
        for rep in range(3):
            t0 = time.perf_counter()
            with ThreadPoolExecutor(T) as pool: list(pool.map(fn, range(len(names))))
            dt = time.perf_counter() - t0
        print(f"{fn.__name__} threads={T}: {dt*1e3:.0f} ms")
