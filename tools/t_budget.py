"""Timing of the device-resident budget null at BASELINE config sizes (scratch tool)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocco_b200 import pipeline, _lib
from rocco_b200.synth import chrom_matrix_torch, chrom_bins, chrom_seed
dev = torch.device('cuda', 0)
for chrom, m in (("chr21", 100), ("chr1", 100)):
    x = chrom_matrix_torch(m, chrom_bins(chrom), chrom_seed(chrom), dev, torch.float64)
    prm = pipeline.score_params(prior_df=6.0)
    sc, det = pipeline.score_loci_wls_device(x, params=prm, details=True)
    del x
    for draws, mind in ((25, None), (25, 25)):
        for rep in range(2):
            if rep == 1: _lib.profile_enable(True); _lib.profile_report()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            frac, meta = pipeline.budget_null_device(det["centered_matrix"], sc, params=prm, dependence_lag_hint=det["local_baseline_window"],
                                                     num_null_draws=draws, min_null_draws=mind)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        rep_ = _lib.profile_report(); _lib.profile_enable(False)
        nd = int(meta["num_null_draws"])
        print(f"{chrom} x {m}: {nd} draws {dt*1e3:.0f} ms ({dt*1e3/(nd+2):.1f} ms per scoring pass), frac {frac:.5f}, tau {meta['autocorrelation_time']:.2f}, null_tail {meta['null_tail_occupancy']:.5f}")
        tot = sum(v[0] for v in rep_.values())
        for k, v in sorted(rep_.items(), key=lambda kv: -kv[1][0])[:8]:
            print(f"     {k:24s} {v[0]:8.2f} ms {100*v[0]/tot:5.1f}%  x{v[1]}")
    del det, sc
    torch.cuda.empty_cache()
