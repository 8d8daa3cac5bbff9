import torch, ctypes, sys, os
sys.path.insert(0, os.getcwd())
from rocco_b200 import _lib
from rocco_b200.synth import chrom_matrix_torch, chrom_bins, chrom_seed
dev = torch.device("cuda", 0)
x = chrom_matrix_torch(100, chrom_bins("chr21"), chrom_seed("chr21"), dev, torch.float64)
out = torch.empty(x.shape[1], dtype=torch.float64, device=dev)
lib = _lib.load()
for stat in (0, 4):
    for _ in range(2):
        st = lib.rocco_b200_column_stat_dev(ctypes.c_void_p(x.data_ptr()), 0, x.shape[0], x.shape[1], stat, 0.0, 0.0, 1.0,
                                            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(st, "column_stat")
torch.cuda.synchronize()
print("ok", float(out.sum()))
