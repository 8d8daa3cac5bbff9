"""Round-2 development probes (run on the GPU box): A/B checks and timings of the reworked kernels.

    python tools/dev_r2.py whit        # Whittaker steady-tile kernels: pair modes vs the round-1 kernel vs the oracle
    python tools/dev_r2.py scopes      # per-scope times of one chromosome set (profile scopes)

Results go to stdout and gpurun_out/dev_<cmd>.json.
"""
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from rocco_b200 import _lib, pipeline  # noqa: E402
from rocco_b200.synth import HG38_SIZES, HG_PARAMS, chrom_bins, chrom_matrix_numpy, chrom_matrix_torch, chrom_seed  # noqa: E402

DEV = torch.device("cuda", 0)
OUT = os.path.join(REPO, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def scopes_of(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    _lib.profile_report()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    rep = _lib.profile_report()
    _lib.profile_enable(False)
    return {k: {"ms": v[0] / reps, "GBps": (v[2] / 1e9) / (v[0] / 1e3) if v[0] > 0 else 0.0} for k, v in rep.items()}


def cmd_whit():
    lib = _lib.load()
    res = {"parity": [], "timing": []}
    prm = pipeline.score_params(prior_df=6.0)
    from oracle import oracle as orc
    shapes = [(3, 60000, torch.float64, True), (3, 60001, torch.float64, True), (4, 934200, torch.float64, False),
              (5, 250001, torch.float32, False), (5, 250002, torch.float32, False), (6, 250003, torch.float32, False),
              (3, 1000001, torch.float64, False)]
    for m, n, dt, use_oracle in shapes:
        x = chrom_matrix_torch(m, n, 11 + n % 7, DEV, dt)
        out = {}
        for mode in (1, 2, 3):
            prev = lib.rocco_b200_whittaker_set_mode(mode)
            s, d = pipeline.score_loci_wls_device(x, params=prm, details=True)
            torch.cuda.synchronize()
            out[mode] = (s.clone(), d["centered_matrix"].clone())
            lib.rocco_b200_whittaker_set_mode(prev)
        row = {"m": m, "n": n, "dtype": str(dt)}
        for mode in (2, 3):
            row[f"centered_maxabs_mode{mode}_vs_old"] = float((out[mode][1] - out[1][1]).abs().max())
            row[f"scores_maxabs_mode{mode}_vs_old"] = float((out[mode][0] - out[1][0]).abs().max())
        if use_oracle:
            want_s, want_d = orc.score_loci_wls(x.cpu().numpy(), prior_df=6.0, return_details=True,
                                                kind="reference" if orc.reference_available() else "port")
            for mode in (1, 2, 3):
                row[f"centered_maxabs_mode{mode}_vs_oracle"] = float(np.max(np.abs(out[mode][1].cpu().numpy() - want_d["centered_matrix"])))
                ref = np.maximum(np.abs(want_s), 1e-3)
                row[f"scores_maxrel_mode{mode}_vs_oracle"] = float(np.max(np.abs(out[mode][0].cpu().numpy() - want_s) / ref))
        print(json.dumps(row), flush=True)
        res["parity"].append(row)
        del x, out
    # non-finite input is still reported
    x = chrom_matrix_torch(3, 60000, 5, DEV, torch.float64)
    x[1, 30000] = float("nan")
    try:
        pipeline.score_loci_wls_device(x, params=prm)
        res["nonfinite_raises"] = False
    except ValueError:
        res["nonfinite_raises"] = True
    print("nonfinite raises:", res["nonfinite_raises"], flush=True)
    # timings
    for name, m, dt in (("chr21", 100, torch.float64), ("chr21", 100, torch.float32), ("chr1", 100, torch.float64)):
        n = chrom_bins(name)
        x = chrom_matrix_torch(m, n, chrom_seed(name), DEV, dt)
        sc = torch.empty(n, dtype=torch.float64, device=DEV)
        for mode in (1, 2, 3):
            prev = lib.rocco_b200_whittaker_set_mode(mode)
            sp = scopes_of(lambda: pipeline.score_loci_wls_device(x, out_scores=sc, params=prm))
            lib.rocco_b200_whittaker_set_mode(prev)
            row = {"chrom": name, "m": m, "dtype": str(dt), "mode": mode,
                   "steady_ms": sp.get("k_whittaker_steady", {}).get("ms"), "steady_GBps": sp.get("k_whittaker_steady", {}).get("GBps"),
                   "edge_ms": sp.get("k_whittaker_edge", {}).get("ms"), "all": {k: round(v["ms"], 3) for k, v in sp.items()}}
            print(json.dumps(row), flush=True)
            res["timing"].append(row)
        del x
    json.dump(res, open(os.path.join(OUT, "dev_whit.json"), "w"), indent=1)


def cmd_scopes():
    names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["chr21"]
    m = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    mats = [chrom_matrix_torch(m, chrom_bins(c), chrom_seed(c), DEV, torch.float64) for c in names]
    budgets = [HG_PARAMS[c][0] for c in names]
    gammas = [HG_PARAMS[c][1] for c in names]
    prm = pipeline.score_params(prior_df=6.0)
    sp = scopes_of(lambda: pipeline.run_shard(mats, budgets, gammas, params=prm))
    tot = sum(v["ms"] for v in sp.values())
    for k, v in sorted(sp.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"{k:28s} {v['ms']:9.3f} ms  {v['GBps']:8.1f} GB/s")
    print("total profiled", tot)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        pipeline.run_shard(mats, budgets, gammas, params=prm)
    torch.cuda.synchronize()
    print("wall per step ms", 1e3 * (time.perf_counter() - t0) / 3)
    json.dump(sp, open(os.path.join(OUT, "dev_scopes.json"), "w"), indent=1)




def cmd_fallback():
    """which rows of a chromosome take the sort fallback, and why"""
    import ctypes
    name = sys.argv[2] if len(sys.argv) > 2 else "chr1"
    m = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    x = chrom_matrix_torch(m, chrom_bins(name), chrom_seed(name), DEV, torch.float64)
    lib = _lib.load()
    r0 = (ctypes.c_longlong * 8)(); r1 = (ctypes.c_longlong * 8)()
    lib.rocco_b200_trend_fallback_reasons(r0)
    f0 = lib.rocco_b200_trend_fallback_rows()
    pipeline.score_loci_wls_device(x, params=pipeline.score_params(prior_df=6.0))
    lib.rocco_b200_trend_fallback_reasons(r1)
    print(name, "fallback rows", lib.rocco_b200_trend_fallback_rows() - f0, "reasons [xslot,xcount,ytotal,yslot,ycount]", [r1[k] - r0[k] for k in range(5)])


def cmd_timeline():
    """device-side gaps between the profile scopes of one run_shard call"""
    names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["chr21"]
    m = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    mats = [chrom_matrix_torch(m, chrom_bins(c), chrom_seed(c), DEV, torch.float64) for c in names]
    budgets = [HG_PARAMS[c][0] for c in names]
    gammas = [HG_PARAMS[c][1] for c in names]
    prm = pipeline.score_params(prior_df=6.0)
    for _ in range(2):
        pipeline.run_shard(mats, budgets, gammas, params=prm)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    _lib.profile_report()
    pipeline.run_shard(mats, budgets, gammas, params=prm)
    torch.cuda.synchronize()
    tl = _lib.profile_timeline()
    _lib.profile_enable(False)
    _lib.profile_report()
    end_prev, gaps = None, {}
    busy = 0.0
    for name, t0, ms in tl:
        if end_prev is not None:
            key = f"{prev_name}->{name}"
            g = gaps.setdefault(key, [0, 0.0])
            g[0] += 1; g[1] += t0 - end_prev
        end_prev, prev_name = t0 + ms, name
        busy += ms
    span = tl[-1][1] + tl[-1][2] - tl[0][1]
    print("search rounds ms:", " ".join(f"{ms:.3f}" for name, t0, ms in tl if name == "chain_search_round"))
    print(f"span {span:.3f} ms, inside scopes {busy:.3f} ms, gaps {span - busy:.3f} ms over {len(tl)} scopes")
    for k, (c, g) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{k:55s} n={c:4d} gap {g:8.3f} ms ({1e3 * g / c:7.1f} us each)")


if __name__ == "__main__":
    {"whit": cmd_whit, "scopes": cmd_scopes, "fallback": cmd_fallback, "timeline": cmd_timeline}[sys.argv[1]]()
