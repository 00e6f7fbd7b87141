"""One short run that launches every kernel of the hot path at its natural size; the command that
`ncu --set full -k regex:<kernel>` wraps (profiles/README.md).  Sizes: C3 batch (STEP_CONF structures x 300 residues),
C5-sized DCC (n = 4,000, m = 200), C4-sized dense slab product (n = 6,000, b = 128), C2 full spectrum (N = 3,000)."""
import os
import sys
from os.path import dirname, realpath

sys.path.insert(0, dirname(dirname(realpath(__file__))))
import numpy as np
import torch

import springcraft_b200 as sc
from bench import make_ensemble, jittered_grid
from springcraft_b200 import _engine
from springcraft_b200.dense_solver import DenseRowOperator
from springcraft_b200.ensemble import enm_ensemble_device

what = set(sys.argv[1:]) or {"c3", "dcc", "slab", "tf32", "full"}
if "c3" in what:
    B = int(os.environ.get("STEP_CONF", 4096))
    base, coords, seq = make_ensemble(0, B)
    ff = sc.TabulatedForceField.e_anm(sc.AtomArray(base, *seq))
    xyz = torch.from_numpy(np.ascontiguousarray(coords.transpose(0, 2, 1))).cuda()
    for _ in range(int(os.environ.get("STEP_PASSES", 2))):
        eig, msf, iters, npairs, conv = enm_ensemble_device(xyz, ff, k=20)
    torch.cuda.synchronize()
    print("c3", B, "structures, converged", conv, "outer", float(iters.abs().float().mean()))
if "dcc" in what:
    n, m = 4000, 200
    gen = torch.Generator("cuda").manual_seed(0)
    V = torch.linalg.qr(torch.randn((3 * n, m), dtype=torch.float64, device="cuda", generator=gen)).Q
    modes = V.T.contiguous()
    lam = torch.sort(torch.rand(m, dtype=torch.float64, device="cuda", generator=gen) * 9.9 + 0.1).values
    for _ in range(2):
        d = _engine.modes_dcc(3, lam, modes, norm=True)
    torch.cuda.synchronize()
    print("dcc", n, m, float(d[0, 0]))
if "slab" in what:
    n, b = 6000, 128
    op = DenseRowOperator(jittered_grid(n, seed=0), sc.ParameterFreeForceField(), 3)
    X = torch.randn((3 * n, b), dtype=torch.float64, device="cuda")
    for _ in range(2):
        Y = op.apply(X, X, (0.5, 0.1, 0.25))
    torch.cuda.synchronize()
    print("slab", n, b, float(Y.abs().sum()))
if "tf32" in what:
    from springcraft_b200 import _lib
    h = _lib.require_device()
    n, b = 8000, 128
    op = DenseRowOperator(jittered_grid(n, seed=0), sc.ParameterFreeForceField(), 3)
    N = op.N
    ld = int(h.scb_tf32_ld(N))
    hi, lo = op.slab32(True)
    zc = torch.randn((2 * b, ld), dtype=torch.float32, device="cuda")
    zp = torch.randn_like(zc)
    rh = torch.randn((b, ld), dtype=torch.float32, device="cuda")
    ca = torch.rand(b, device="cuda") * 0.1
    cb = torch.rand(b, device="cuda")
    for _ in range(3):
        _lib.check(h.scb_dense_slab_tf32_apply(N, 0, N, _lib.ptr(hi), _lib.ptr(lo), b, _lib.ptr(zc), _lib.ptr(zp),
                                               _lib.ptr(rh), _lib.ptr(zp), _lib.ptr(ca), _lib.ptr(cb), 0.7, 1,
                                               _lib.stream_ptr()))
    torch.cuda.synchronize()
    print("tf32 filter step", n, b, float(zp.abs().max()))
if "full" in what:
    from synthetic_inputs import synthetic_chain
    anm = sc.ANM(synthetic_chain(1000, seed=0), sc.HinsenForceField())
    dense = anm._model_device().dense()
    for _ in range(2):
        lam, modes = _engine.eig_full_dense(dense.clone())
    torch.cuda.synchronize()
    print("full spectrum N", dense.shape[-1], float(lam[0, -1]))
