import sys, time, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
import rocco_b200
from rocco_b200.synth import chrom_matrix_torch, chrom_bins
dev = torch.device('cuda', 0)
m, n = 100, chrom_bins("chr1")
x = chrom_matrix_torch(m, n, 5, dev, torch.float64)
h = torch.empty(x.shape, dtype=x.dtype, pin_memory=True); h.copy_(x); torch.cuda.synchronize(); del x
arr = h.numpy()
os.chdir("/tmp")
for rep in range(3):
    t0 = time.perf_counter(); scores = rocco_b200.score_loci_wls(arr, prior_df=6.0); t1 = time.perf_counter()
    sol, obj = rocco_b200.solve_chrom_exact(scores, budget=0.03, gamma=1.0); t2 = time.perf_counter()
    iv = np.arange(0, 50 * n, 50); t3 = time.perf_counter()
    f = rocco_b200.chrom_solution_to_bed("chr1", iv, sol, ID="t"); t4 = time.perf_counter()
    print(f"score {1e3*(t1-t0):.0f} ms | solve {1e3*(t2-t1):.0f} ms | arange {1e3*(t3-t2):.0f} ms | bed {1e3*(t4-t3):.0f} ms")
