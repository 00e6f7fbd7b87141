"""Full-size runs of the BASELINE.json configurations on the GPU, checked through size-independent
properties (the oracle cannot finish these sizes in seconds) plus spot checks against the oracle:

C3  4,096 conformations x 300 residues, eANM, 20 lowest non-trivial modes + MSF
C4  20,000-residue cloud, all-pairs ParameterFree (dense 60,000 x 60,000 fp64), 100 lowest modes
C5  DCC of 10,000 residues from 500 modes
(C1 and C2 run at full size in test_gpu_parity.py.)
"""
import numpy as np
import pytest

import springcraft_b200 as sc
from oracle import enm_oracle as orc
from .conftest import golden

pytestmark = pytest.mark.gpu


def test_c3_full_batch_properties():
    import torch
    from springcraft_b200 import _lib
    ref = golden("ref_c3_chain300.npz")
    B, n, k = 4096, 300, 20
    rng_members = [0, 1, 2, 4095]
    coords = np.stack([orc.perturbed_conformation(ref["base"], c) for c in range(B)])
    atoms = sc.AtomArray(ref["base"], ref["res_name"], ref["chain_id"], ref["res_id"])
    res = sc.enm_ensemble(coords, sc.TabulatedForceField.e_anm(atoms), k=k, return_modes=True)
    assert res.converged
    lam, msf, modes = res.eigenvalues, res.msf, res.modes
    # ascending positive spectra, positive MSF
    assert np.all(lam > 0) and np.all(np.diff(lam, axis=1) >= 0) and np.all(msf > 0)
    # sum_i msf_i = sum_k 1/lambda_k (unit-norm modes), for every conformation
    assert np.allclose(msf.sum(axis=1), (1.0 / lam).sum(axis=1), rtol=1e-10)
    # orthonormal modes, orthogonal to the rigid-body motions, for every conformation
    gram = np.einsum("bkn,bln->bkl", modes, modes)
    assert np.abs(gram - np.eye(k)).max() < 1e-9
    trans = np.tile(np.eye(3), (n, 1)).T / np.sqrt(n)          # (3, 3n) translations
    assert np.abs(np.einsum("bkn,tn->bkt", modes, trans)).max() < 1e-9
    # golden members (outputs of the unmodified reference)
    for c in rng_members:
        assert np.allclose(lam[c], ref[f"c{c}/e_anm/eigval"][6:26], rtol=1e-8, atol=0)
        assert np.allclose(msf[c], ref[f"c{c}/e_anm/msf_6_26"], rtol=1e-8, atol=0)
    # eigen-residuals of a random sample against oracle Hessians
    spec = orc.preset_spec("e_anm", ref["res_name"], ref["chain_id"], ref["res_id"])
    for c in np.random.default_rng(0).choice(B, 6, replace=False):
        H, _ = orc.compute_hessian(coords[c], spec)
        r = H @ modes[c].T - modes[c].T * lam[c]
        assert np.linalg.norm(r, axis=0).max() <= 1e-8 * lam[c][-1]
    del torch, _lib


@pytest.mark.parametrize("filt", ["tf32", "fp64"])
def test_c4_full_size_properties(filt):
    """60,000 x 60,000 dense Hessian on one GPU (28.8 GB): residuals, orthonormality, null space -- with the
    residual-form filter on the TF32 tensor cores (3-term split product, + 28.8 GB of single-precision slabs) and
    with the FP64 filter."""
    import torch
    from springcraft_b200 import _lib
    from springcraft_b200.dense_solver import DenseRowOperator, eig_lowest_dense
    free, _ = torch.cuda.mem_get_info()
    if free < 80e9:
        pytest.skip("needs ~75 GB of free device memory")
    n, k = 20000, 100
    rng = np.random.default_rng(0)
    side = (n / 0.008) ** (1 / 3)
    g = int(np.ceil(n ** (1 / 3)))
    pts = np.stack(np.meshgrid(*[np.arange(g)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n] * (side / g)
    coord = pts + rng.uniform(-0.25, 0.25, pts.shape) * (side / g) * 0.5      # min distance > 3 A
    op = DenseRowOperator(coord, sc.ParameterFreeForceField(), 3)
    Z = op.rigid_basis()
    theta, X, resid, iters = eig_lowest_dense(op, k, Z=Z, filter=filt)
    th = theta.cpu().numpy()
    assert np.all(th[:k] > 0) and np.all(np.diff(th[:k]) >= 0)
    assert float(resid[:k].max()) <= 3e-9 * th[k - 1]
    # independent residual check with one more operator application
    HX = op.apply(X)
    r = (HX[:, :k] - X[:, :k] * theta[:k]).norm(dim=0)
    assert float(r.max()) <= 1e-8 * th[k - 1]
    G = X[:, :k].T @ X[:, :k]
    assert float((G - torch.eye(k, dtype=torch.float64, device="cuda")).abs().max()) < 1e-10
    assert float((Z.T @ X[:, :k]).abs().max()) < 1e-10
    # rigid-body motions are in the null space: ||H Z|| tiny next to the spectrum bound
    Zp = torch.zeros((3 * n, 64), dtype=torch.float64, device="cuda")
    Zp[:, :6] = Z
    assert float(op.apply(Zp).abs().max()) <= 1e-9 * op.spectrum_bound()
    # spot check of slab rows against the reference formula H_ij = -d d^T / |d|^4
    i, j = 1234, 15678
    d = coord[j] - coord[i]
    want = -np.outer(d, d) / (d @ d) ** 2
    got = op.slab[3 * i:3 * i + 3, 3 * j:3 * j + 3].cpu().numpy()
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    op.close()
    del _lib


def test_c5_full_size_dcc_properties():
    """DCC of 10,000 residues from 500 modes (30,000-dimensional): unit diagonal, symmetry, sampled entries."""
    import torch
    from springcraft_b200 import _engine
    n, m, D = 10000, 500, 3
    g = torch.Generator("cuda").manual_seed(5)
    Q, _ = torch.linalg.qr(torch.randn((D * n, m), dtype=torch.float64, device="cuda", generator=g))
    modes = Q.T.contiguous()                                    # orthonormal rows
    lam = torch.linspace(0.5, 40.0, m, dtype=torch.float64, device="cuda")
    dcc = _engine.modes_dcc(D, lam, modes, norm=True)
    assert dcc.shape == (n, n)
    assert float((dcc.diagonal() - 1.0).abs().max()) < 1e-12
    assert float((dcc - dcc.T).abs().max()) < 1e-12
    assert float(dcc.abs().max()) <= 1.0 + 1e-12
    raw = _engine.modes_dcc(D, lam, modes, norm=False)
    mo, la = modes.cpu().numpy().reshape(m, n, D), lam.cpu().numpy()
    rng = np.random.default_rng(1)
    for i, j in rng.integers(0, n, size=(40, 2)):
        want = np.einsum("ka,ka,k->", mo[:, i], mo[:, j], 1.0 / la)   # nma.py:346-347 summed over the subset
        assert abs(float(raw[i, j]) - want) <= 1e-12 * abs(want) + 1e-18
        dii = np.einsum("ka,ka,k->", mo[:, i], mo[:, i], 1.0 / la)
        djj = np.einsum("ka,ka,k->", mo[:, j], mo[:, j], 1.0 / la)
        assert abs(float(dcc[i, j]) - want / np.sqrt(dii * djj)) <= 1e-10
    # row-slab form (what each rank computes in the partitioned path) equals the full matrix
    slab = _engine.modes_dcc(D, lam, modes, norm=True, rows=(2500, 5000))
    assert torch.equal(slab, dcc[2500:5000])
