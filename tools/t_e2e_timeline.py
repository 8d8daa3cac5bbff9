"""Timeline of the e2e leg: per-thread start/end of each public-API call (scratch tool, not part of the product)."""
import sys, time, os, numpy as np, torch, threading
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rocco_b200
from rocco_b200.synth import chrom_matrix_torch, chrom_bins, HG38_SIZES, HG_PARAMS, chrom_seed
dev = torch.device('cuda', 0)
names = list(HG38_SIZES)
host = []
for c in names:
    x = chrom_matrix_torch(100, chrom_bins(c), chrom_seed(c), dev, torch.float64)
    h = torch.empty(x.shape, dtype=x.dtype, pin_memory=True); h.copy_(x); host.append(h.numpy()); del x
torch.cuda.synchronize(); torch.cuda.empty_cache()
os.chdir("/tmp")
T0 = [0.0]
log = []
def one(k):
    c = names[k]; b, g = HG_PARAMS[c]
    t = [time.perf_counter() - T0[0]]
    s = rocco_b200.score_loci_wls(host[k], prior_df=6.0); t.append(time.perf_counter() - T0[0])
    sol, obj = rocco_b200.solve_chrom_exact(s, budget=b, gamma=g); t.append(time.perf_counter() - T0[0])
    f = rocco_b200.chrom_solution_to_bed(c, np.arange(0, 50 * len(sol), 50), sol, ID="t"); t.append(time.perf_counter() - T0[0])
    log.append((c, threading.get_ident() % 1000, t))
    return f
def pure_h2d():
    h = torch.from_numpy(host[0]); d = torch.empty(h.shape, dtype=h.dtype, device=dev)
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    return h.numel() * 8 / (time.perf_counter() - t0) / 1e9
print("pure H2D GB/s before:", [round(pure_h2d(), 1) for _ in range(3)])
order = sorted(range(len(names)), key=lambda k: -host[k].shape[1])
for T in (6,):
    for rep in range(3):
        log.clear()
        T0[0] = time.perf_counter()
        with ThreadPoolExecutor(T) as pool: files = list(pool.map(one, order))
        t1 = time.perf_counter() - T0[0]
        rocco_b200.combine_chrom_results(files, "comb.bed")
        t2 = time.perf_counter() - T0[0]
        print(f"threads={T} rep {rep}: pool {t1*1e3:.0f} ms, combine {1e3*(t2-t1):.0f} ms, total {t2*1e3:.0f} ms")
        if rep:
            for c, th, t in sorted(log, key=lambda r: r[2][0]):
                print(f"  {c:6s} th{th:03d} start {t[0]*1e3:7.1f} score-end {1e3*t[1]:7.1f} ({1e3*(t[1]-t[0]):6.1f}) solve {1e3*(t[2]-t[1]):6.1f} bed {1e3*(t[3]-t[2]):6.1f}  GB {host[names.index(c)].nbytes/1e9:.2f}")
print("pure H2D GB/s after:", [round(pure_h2d(), 1) for _ in range(3)])
