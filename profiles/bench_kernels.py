"""Secondary measurements: achieved rates of the non-dominant kernels at their natural sizes
(SURVEY 8d rooflines).  CUDA events on the launching stream, 3 warm-ups, best of 5."""
import ctypes as C
import json
import sys
import time
from os.path import dirname, realpath

sys.path.insert(0, dirname(dirname(realpath(__file__))))
import numpy as np
import torch

import springcraft_b200 as sc
from oracle import enm_oracle as orc
from springcraft_b200 import _engine, _lib
from springcraft_b200._engine import DeviceModel

h = _lib.require_device()
HBM = json.load(open(dirname(dirname(realpath(__file__))) + "/MEASURED_PEAKS.json"))["hbm_gbs"] \
    if True else 6534.8


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {}
st = _lib.stream_ptr

# ---- K1 / K2 on the C3 batch (4096 x 300) and on one 20k structure
from bench import make_ensemble
base, coords, seq = make_ensemble(0, 4096)
ff = sc.TabulatedForceField.e_anm(sc.AtomArray(base, *seq))
model = DeviceModel(coords, ff, 3)
B, n, P = model.B, model.n, model.P
rowcount = torch.empty(B * n, dtype=torch.int32, device="cuda")
ms = timed(lambda: _lib.check(h.scb_contacts_count(_lib.ptr(model.xyz), B, n, model.cutoff_sq, None, 0, _lib.ptr(rowcount), st())))
ms2 = timed(lambda: _lib.check(h.scb_contacts_fill(_lib.ptr(model.xyz), B, n, model.cutoff_sq, None, 0, _lib.ptr(model.rowptr), _lib.ptr(model.col), st())))
k1_bytes = 24.0 * B * n + 8.0 * (B * n + 1) + 4.0 * P
out["K1 contacts (tiled), C3 batch"] = {"count_ms": ms, "fill_ms": ms2, "pair_tests_per_s": B * n * n / (ms2 * 1e-3),
                                        "alg_GBps_fill": k1_bytes / (ms2 * 1e-3) / 1e9}
flag = torch.zeros(1, dtype=torch.int32, device="cuda")
ms = timed(lambda: _lib.check(h.scb_assemble(3, _lib.ptr(model.xyz), B, n, C.byref(model.desc), _lib.ptr(model.rowptr), _lib.ptr(model.col),
                                             None, _lib.ptr(model.offdiag), _lib.ptr(model.diag), None, _lib.ptr(flag), st())))
k2_bytes = 24.0 * B * n + 4.0 * P + 72.0 * P + 72.0 * B * n
out["K2 fused e_anm force constants + Hessian blocks, C3 batch"] = {"ms": ms, "alg_GBps": k2_bytes / (ms * 1e-3) / 1e9,
                                                                   "frac_of_measured_hbm": k2_bytes / (ms * 1e-3) / 1e9 / HBM}
big = orc.synthetic_chain(20000, seed=0)
for use_cell in (0, 1):
    xyz = torch.from_numpy(np.ascontiguousarray(big.T)).cuda()[None].contiguous()
    rc = torch.empty(20000, dtype=torch.int32, device="cuda")
    ms = timed(lambda: _lib.check(h.scb_contacts_count(_lib.ptr(xyz), 1, 20000, 169.0, None, use_cell, _lib.ptr(rc), st())))
    out[f"K1 contacts n=20000 {'cell list' if use_cell else 'tiled all-pairs'} (count pass)"] = {"ms": ms}

# ---- K4: DCC / covariance DMMA GEMM at C5 size (n=10,000, m=500)
n5, m5 = 10000, 500
rng = np.random.default_rng(0)
for D in (3, 1):
    N5 = D * n5
    V = torch.linalg.qr(torch.randn(N5, m5, dtype=torch.float64, device="cuda"))[0].T.contiguous()
    lam = torch.sort(torch.rand(m5, dtype=torch.float64, device="cuda") * 9.9 + 0.1)[0]
    ws_bytes = h.scb_dcc_workspace_bytes(D, n5, m5)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    o = torch.empty((n5, n5), dtype=torch.float64, device="cuda")
    ms = timed(lambda: _lib.check(h.scb_dcc(D, n5, m5, _lib.ptr(lam), _lib.ptr(V), 1, 1.0, 0, n5, _lib.ptr(o), _lib.ptr(ws), ws_bytes, st())))
    flops = 2.0 * n5 * n5 * D * m5
    out[f"K4 DCC n=10000 m=500 D={D} (prepare + DMMA GEMM + fused normalisation)"] = {
        "ms": ms, "fp64_TFLOPs": flops / (ms * 1e-3) / 1e12}
    # parity of one tile against torch
    W = (V / lam[:, None]).T.reshape(n5, D, m5)[:64].reshape(64, -1) if False else None
# DGEMM reference rate for the roofline denominator (cuBLAS through torch)
A = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
ms = timed(lambda: torch.matmul(A, A))
out["cuBLAS DGEMM 8192^3 (FP64 denominator)"] = {"ms": ms, "fp64_TFLOPs": 2 * 8192.0 ** 3 / (ms * 1e-3) / 1e12}

# ---- K3c: full spectrum block Jacobi, C2 size
coord = orc.synthetic_chain(1000, seed=0)
anm = sc.ANM(coord, sc.HinsenForceField())
dense = anm._model_device().dense()[0]
t0 = time.perf_counter()
lam_f, modes_f = _engine.eig_full_dense(dense.clone())
torch.cuda.synchronize()
t1 = time.perf_counter()
lam_f, modes_f = _engine.eig_full_dense(dense.clone())
torch.cuda.synchronize()
t2 = time.perf_counter()
out["K3c block-Jacobi full eigh N=3000 (C2)"] = {"seconds": t2 - t1, "first_call_seconds": t1 - t0}
t0 = time.perf_counter()
np.linalg.eigh(dense.cpu().numpy())
out["numpy eigh N=3000 on the host (reference path)"] = {"seconds": time.perf_counter() - t0}
print(json.dumps(out, indent=1))
