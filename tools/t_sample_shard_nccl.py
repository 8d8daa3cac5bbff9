"""2-rank NCCL check of sample-sharded scoring: torchrun --nproc-per-node 2 tools/t_sample_shard_nccl.py"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '/root/repo')
from rocco_b200 import pipeline
from rocco_b200.synth import chrom_matrix_numpy
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
m, n = 16, 300_000
x = chrom_matrix_numpy(m, n, seed=5)
prm = pipeline.score_params(prior_df=6.0)
lo, hi = rank * m // world, (rank + 1) * m // world
d = torch.from_numpy(x[lo:hi]).cuda()
s = pipeline.score_loci_wls_sample_sharded(d, m, prm)
whole = pipeline.score_loci_wls_device(torch.from_numpy(x).cuda(), params=prm)
err = float((s - whole).abs().max())
print(f"rank {rank}: rows [{lo},{hi}) max abs diff vs unsharded {err:.3e}")
assert err <= 1e-11
dist.destroy_process_group()
