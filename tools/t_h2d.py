import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import rocco_b200
from rocco_b200 import pipeline
from rocco_b200.synth import chrom_matrix_torch
dev = torch.device('cuda', 0)
m, n = 100, 2_000_000
x = chrom_matrix_torch(m, n, 5, dev, torch.float64)
h = torch.empty(x.shape, dtype=x.dtype, pin_memory=True); h.copy_(x); torch.cuda.synchronize()
hp = h.numpy().copy()    # pageable copy
d = torch.empty_like(x)
for name, src in (("pinned", h), ("pageable", torch.from_numpy(hp))):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"torch H2D {name}: {x.numel()*8/dt/1e9:.1f} GB/s")
prm = pipeline.score_params(prior_df=6.0)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); s = pipeline.score_loci_wls_device(x, params=prm); torch.cuda.synchronize(); dt_dev = time.perf_counter() - t0
print(f"device-resident score: {dt_dev*1e3:.1f} ms")
for name, arr in (("pinned", h.numpy()), ("pageable", hp)):
    for _ in range(2):
        t0 = time.perf_counter(); s = rocco_b200.score_loci_wls(arr, prior_df=6.0); dt = time.perf_counter() - t0
    print(f"host API score_loci_wls {name}: {dt*1e3:.1f} ms  -> implied H2D {x.numel()*8/(dt-dt_dev)/1e9:.1f} GB/s")
