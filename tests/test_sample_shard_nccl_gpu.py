"""sample-sharded scoring over NCCL (BASELINE.json config 5's data-plane collective).

Two ranks, one GPU each, launched with torch.distributed.run from inside the test (skipped on a box with fewer than
two GPUs -- the single-GPU emulation of the same path is tests/test_score_gpu.py::test_sample_sharded_scoring_matches_unsharded).
Each rank scores its half of the samples, the [4, bins] float64 accumulators are all-reduced over NVLink, and every rank
finalises; the result must match the unsharded device path and the oracle.
"""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ROCCO_B200_REPO"])
from rocco_b200 import pipeline
from rocco_b200.synth import chrom_matrix_numpy
from oracle import oracle as orc

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
m, n = 16, 300_001
x = chrom_matrix_numpy(m, n, seed=5)
prm = pipeline.score_params(prior_df=6.0)
lo, hi = rank * m // world, (rank + 1) * m // world
d = torch.from_numpy(x[lo:hi]).to(dev)
sharded = pipeline.score_loci_wls_sample_sharded(d, m, prm)
whole = pipeline.score_loci_wls_device(torch.from_numpy(x).to(dev), params=prm)
err_dev = float((sharded - whole).abs().max())
# every rank holds the same result
ref = sharded.clone()
dist.broadcast(ref, src=0)
same = bool(torch.equal(ref, sharded))
out = {"rank": rank, "rows": [lo, hi], "max_abs_vs_unsharded": err_dev, "identical_across_ranks": same}
if rank == 0:
    want = orc.score_loci_wls(x, prior_df=6.0, kind="reference" if orc.reference_available() else "port")
    got = sharded.cpu().numpy()
    out["max_rel_vs_oracle"] = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)))
    # float32 shards, the storage type of config 5
    x32 = x.astype(np.float32)
d32 = torch.from_numpy(x.astype(np.float32)[lo:hi]).to(dev)
s32 = pipeline.score_loci_wls_sample_sharded(d32, m, prm)
w32 = pipeline.score_loci_wls_device(torch.from_numpy(x.astype(np.float32)).to(dev), params=prm)
out["max_abs_f32_vs_unsharded"] = float((s32 - w32).abs().max())
print("RESULT " + json.dumps(out), flush=True)
dist.destroy_process_group()
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_nccl_sample_sharded_scoring(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, ROCCO_B200_REPO=REPO)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    rows = [json.loads(ln.split("RESULT ", 1)[1]) for ln in res.stdout.splitlines() if "RESULT " in ln]
    assert len(rows) == 2
    for r in rows:
        assert r["max_abs_vs_unsharded"] <= 1e-11, r          # the sample order of the sum differs: ~1e-16 relative
        assert r["max_abs_f32_vs_unsharded"] <= 1e-11, r
        assert r["identical_across_ranks"], r
    assert [r for r in rows if r["rank"] == 0][0]["max_rel_vs_oracle"] <= 1e-5
