"""H2D rate from pinned memory alone vs while scoring kernels run on another stream (scratch tool)."""
import os, sys, time, threading, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocco_b200 import pipeline
from rocco_b200.synth import chrom_matrix_torch, chrom_bins, chrom_seed
dev = torch.device('cuda', 0)
x = chrom_matrix_torch(100, chrom_bins("chr1"), chrom_seed("chr1"), dev, torch.float64)
h = torch.empty(x.shape, dtype=x.dtype, pin_memory=True); h.copy_(x)
y = chrom_matrix_torch(100, chrom_bins("chr5"), chrom_seed("chr5"), dev, torch.float64)
d = torch.empty_like(x)
cs = torch.cuda.Stream()
def h2d(reps=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(cs):
        for _ in range(reps): d.copy_(h, non_blocking=True)
    cs.synchronize(); dt = time.perf_counter() - t0
    return reps * h.numel() * 8 / dt / 1e9
print("pure H2D GB/s:", [round(h2d(), 1) for _ in range(3)])
stop = False
def compute():
    ks = torch.cuda.Stream()
    with torch.cuda.stream(ks):
        while not stop:
            pipeline.score_loci_wls_device(y)
th = threading.Thread(target=compute); th.start(); time.sleep(0.5)
print("H2D GB/s with scoring kernels running:", [round(h2d(), 1) for _ in range(3)])
stop = True; th.join()
# D2D copy (HBM heavy) concurrently
z = torch.empty_like(y)
def hbm():
    ks = torch.cuda.Stream()
    with torch.cuda.stream(ks):
        while not stop: z.copy_(y)
stop = False
th = threading.Thread(target=hbm); th.start(); time.sleep(0.3)
print("H2D GB/s with a device-to-device copy loop running:", [round(h2d(), 1) for _ in range(3)])
stop = True; th.join()
