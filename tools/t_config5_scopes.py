"""Where does sample-sharded scoring spend its time at 20 bp (chr1: 12.4 M bins x 125 samples f32)? (scratch tool)"""
import sys, time, torch
sys.path.insert(0, '/root/repo')
from rocco_b200 import pipeline, _lib
from rocco_b200.synth import chrom_matrix_torch, chrom_seed
import math
dev = torch.device('cuda', 0)
n = math.ceil(248956422 / 20)
x = chrom_matrix_torch(125, n, chrom_seed("chr1"), dev, torch.float32, sample_stream=1)
prm = pipeline.score_params(prior_df=6.0)
for rep in range(2):
    if rep: _lib.profile_enable(True); _lib.profile_report()
    fb0 = int(_lib.load().rocco_b200_trend_fallback_rows())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    acc = pipeline.score_partial_device(x, prm)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"rep {rep}: {dt*1e3:.0f} ms, fallback rows {int(_lib.load().rocco_b200_trend_fallback_rows()) - fb0} of 125")
rep_ = _lib.profile_report()
import ctypes
reasons = (ctypes.c_longlong * 8)(); _lib.load().rocco_b200_trend_fallback_reasons(reasons); print("reason counts", list(reasons))
for k, v in sorted(rep_.items(), key=lambda kv: -kv[1][0])[:10]:
    print(f"   {k:24s} {v[0]:8.2f} ms  x{v[1]}")
