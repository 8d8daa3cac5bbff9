import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from rocco_b200 import pipeline
from rocco_b200.synth import chrom_matrix_torch
dev = torch.device('cuda', 0)
mats = [chrom_matrix_torch(20, n, 100 + i, dev, torch.float64) for i, n in enumerate([934200, 1172353, 400000])]
prm = pipeline.score_params(prior_df=6.0)
ref = None
for rep in range(4):
    sh = pipeline.run_shard(mats, [0.02, 0.045, 0.03], [1.0, 1.0, 1.0], params=prm)
    torch.cuda.synchronize()
    sc = sh['d_scores'].cpu().numpy(); mk = sh['d_masks'].cpu().numpy()
    lam = [r['selection_penalty'] for r in sh['results']]; cnt = [r['selected_count'] for r in sh['results']]
    if ref is None: ref = (sc, mk, lam, cnt)
    print(rep, 'scores equal', np.array_equal(sc, ref[0]), 'max abs diff', float(np.max(np.abs(sc - ref[0]))),
          'mask diff', int(np.sum(mk != ref[1])), 'lam', lam, 'cnt', cnt)
