"""One warm-up + one timed pass of the C3 hot path (device-resident inputs); the command ncu wraps for launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.realpath(__file__))))
import numpy as np, torch
import springcraft_b200 as sc
from springcraft_b200 import ensemble
from bench import make_ensemble, N_CONF

n_conf = int(os.environ.get("STEP_CONF", N_CONF))
base, coords, seq = make_ensemble(0, n_conf)
ff = sc.TabulatedForceField.e_anm(sc.AtomArray(base, *seq))
xyz = torch.from_numpy(np.ascontiguousarray(coords.transpose(0, 2, 1))).cuda()
reps = int(os.environ.get("STEP_REPS", 1))
ensemble.enm_ensemble_device(xyz, ff, k=20)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    eig, msf, iters, npairs, conv = ensemble.enm_ensemble_device(xyz, ff, k=20)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"step {ms:.1f} ms  {n_conf / ms * 1e3:.0f} structures/s  outer {iters.abs().float().mean().item():.2f} converged {conv}")
