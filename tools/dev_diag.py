import os, sys, numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import rocco_b200
from rocco_b200 import _lib
from oracle import oracle as orc
g = np.load(os.path.join(REPO, "tests/golden/reference_assembly_v1_11_0.npz"))
x = g["single_end_matrix"]
print(x.shape, x.min(), x.max(), np.median(x))
ws, wd = orc.score_loci_wls(x, prior_df=6.0, return_details=True, kind="reference" if orc.reference_available() else "port")
for mode in (0, 1):
    _lib.load().rocco_b200_whittaker_set_mode(mode)
    for tm in (0, 1):
        _lib.load().rocco_b200_trend_set_mode(tm)
        gs, gd = rocco_b200.score_loci_wls(x, prior_df=6.0, return_details=True)
        rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3)))
        print("whit mode", mode, "trend mode", tm, "scores", rel(gs, ws), {k: rel(gd[k], wd[k]) for k in ("mean", "raw_variance", "prior_variance", "moderated_variance", "standard_error")},
              "centered", float(np.max(np.abs(gd["centered_matrix"] - wd["centered_matrix"]))))
        j = int(np.argmax(np.abs(gs - ws) / np.maximum(np.abs(ws), 1e-3)))
        print("   worst bin", j, gs[j], ws[j], "prior", gd["prior_variance"][j], wd["prior_variance"][j], "raw", gd["raw_variance"][j], wd["raw_variance"][j])
_lib.load().rocco_b200_trend_set_mode(0); _lib.load().rocco_b200_whittaker_set_mode(0)
