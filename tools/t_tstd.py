import numpy as np, sys
sys.path.insert(0,'/root/repo')
import rocco_b200
from rocco_b200.synth import chrom_matrix_numpy
x = chrom_matrix_numpy(100, 2000, seed=4)
print(rocco_b200.score_dispersion_chrom(x, method="tstd", tprop=0.1)[:5])
