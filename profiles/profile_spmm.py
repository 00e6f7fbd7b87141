"""Driver for `ncu --set full` captures of the SpMM kernel on the C3 batch."""
import sys
from os.path import dirname, realpath
sys.path.insert(0, dirname(dirname(realpath(__file__))))
import numpy as np
import torch
import springcraft_b200 as sc
from springcraft_b200._engine import DeviceModel
from bench import make_ensemble, N_CONF

B = int(sys.argv[1]) if len(sys.argv) > 1 else N_CONF
base, coords, seq = make_ensemble(0, B)
ff = sc.TabulatedForceField.e_anm(sc.AtomArray(base, *seq))
model = DeviceModel(coords, ff, 3)
X = torch.randn((B, 900, 32), dtype=torch.float64, device="cuda")
for _ in range(5):
    Y = model.spmm_paired(X)
torch.cuda.synchronize()
print("ok", float(Y.abs().sum()))
