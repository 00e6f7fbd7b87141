"""2-GPU checks of the partitioned paths (run with `gpurun --gpus 2`; skipped on one GPU)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from .conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["SCB_ROOT"])
import numpy as np, torch, torch.distributed as dist
import springcraft_b200 as sc
from springcraft_b200 import _engine, parallel
from oracle import enm_oracle as orc
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
n, m = 600, 50
coord = orc.synthetic_chain(n, seed=3)
anm = sc.ANM(coord, sc.InvariantForceField(13.0))
lam, modes = anm.eigen(k=6 + m)
lam_d = torch.from_numpy(lam[6:]).cuda(); modes_d = torch.from_numpy(modes[6:]).cuda()
full = parallel.dcc_row_partitioned(3, lam_d, modes_d, norm=True, gather=True).cpu().numpy()
row0, row1, slab = parallel.dcc_row_partitioned(3, lam_d, modes_d, norm=False)
want = orc.dcc(lam, modes, 3, mode_subset=np.arange(6, 6 + m))
want_abs = orc.dcc(lam, modes, 3, mode_subset=np.arange(6, 6 + m), norm=False)
assert np.allclose(full, want, atol=1e-10), "gathered DCC"
assert np.allclose(slab.cpu().numpy(), want_abs[row0:row1], rtol=1e-10, atol=1e-14), "row slab"
# ensemble shards: each rank solves its conformations, results gathered
B = 6
a, b = parallel.shard_range(B, rank, dist.get_world_size())
confs = np.stack([orc.perturbed_conformation(coord, c) for c in range(B)])
res = sc.enm_ensemble(confs[a:b], sc.InvariantForceField(13.0), k=20)
ev = parallel.gather_results(torch.from_numpy(res.eigenvalues).cuda(), B).cpu().numpy()
ref = sc.enm_ensemble(confs, sc.InvariantForceField(13.0), k=20).eigenvalues
assert np.allclose(ev, ref, rtol=1e-9)
# C4: dense all-pairs Hessian, row slabs + all-gather of the block per operator application
from tests.conftest import golden
g = golden("ref_c4_cloud400.npz")
lam4, modes4, it4 = sc.allpairs_lowest_modes(g["coord"], sc.ParameterFreeForceField(), 56)
assert np.allclose(lam4[6:56], g["eigval"][6:56], rtol=1e-8, atol=0), "C4 eigenvalues"
A4 = g["modes_6_106"][:50]; Q, _ = np.linalg.qr(modes4[6:56].T)
Qa, _ = np.linalg.qr(A4.T)
assert np.linalg.norm(Q - Qa @ (Qa.T @ Q), 2) < 1e-6, "C4 subspace"
# the same solve with the plain NCCL all-gather instead of the fused peer-memory epilogue
lam4n, modes4n, it4n = sc.allpairs_lowest_modes(g["coord"], sc.ParameterFreeForceField(), 56, exchange="nccl")
assert it4n == it4 and np.allclose(lam4n, lam4, rtol=1e-11, atol=1e-12), "peer vs nccl exchange"
# one operator application, both exchanges, bit for bit
from springcraft_b200.dense_solver import DenseRowOperator
Xb = torch.randn((3 * len(g["coord"]), 64), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(7))
outs = []
for ex in ("peer", "nccl"):
    op = DenseRowOperator(g["coord"], sc.ParameterFreeForceField(), 3, exchange=ex)
    outs.append(op.apply(Xb).clone()); outs.append(op.apply(Xb, Xb, (0.5, 0.1, 0.25)).clone())
    op.close()
assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[3]), "fused all-gather differs from NCCL all-gather"
# the oracle-free checks that bench.py --gpus N also runs (C4 peer/nccl, borderline tolerance, C5 slabs, C3 shards)
from tests.multigpu_checks import run_checks
run_checks()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_gpu_partitioned_paths(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SCB_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ok") == 2
