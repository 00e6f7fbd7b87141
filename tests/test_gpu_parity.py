"""GPU parity tests: the CUDA path (through the C ABI / the springcraft-style
API) against (i) golden vectors produced by the unmodified reference and
(ii) the NumPy oracle on seeded inputs.

Tolerances (BASELINE.json north_star): contact sets and Invariant Kirchhoff
bit-exact; Hessians <= 1e-12 relative; eigenvalues <= 1e-8 relative;
eigenvector subspaces: sine of the largest principal angle < 1e-6;
MSF / DCC <= 1e-8.
"""
import numpy as np
import pytest

import springcraft_b200 as sc
from oracle import enm_oracle as orc
from .conftest import golden

pytestmark = pytest.mark.gpu

HESS_RTOL = 1e-12
EIG_RTOL = 1e-8
ANGLE_TOL = 1e-6
PROD_RTOL = 1e-8


def atoms_of(st, name):
    return sc.AtomArray(st[f"{name}_coord"], st[f"{name}_res_name"], st[f"{name}_chain_id"], st[f"{name}_res_id"])


FF = {
    "invariant7": lambda a: sc.InvariantForceField(7.0),
    "invariant13": lambda a: sc.InvariantForceField(13.0),
    "hinsen": lambda a: sc.HinsenForceField(),
    "hinsen_cut12": lambda a: sc.HinsenForceField(12.0),
    "pfree": lambda a: sc.ParameterFreeForceField(),
    "pfree_cut10": lambda a: sc.ParameterFreeForceField(10.0),
    "e_anm": lambda a: sc.TabulatedForceField.e_anm(a),
    "e_anm_mean": lambda a: sc.TabulatedForceField.e_anm(a, nonbonded_mean=True),
    "e_anm_mj": lambda a: sc.TabulatedForceField.e_anm_mj(a),
    "e_anm_ke": lambda a: sc.TabulatedForceField.e_anm_ke(a),
    "sd_enm": lambda a: sc.TabulatedForceField.sd_enm(a),
    "d_enm": lambda a: sc.TabulatedForceField.d_enm(a),
    "s_enm_10": lambda a: sc.TabulatedForceField.s_enm_10(a),
    "s_enm_13": lambda a: sc.TabulatedForceField.s_enm_13(a),
}


def rel_err(a, b):
    scale = np.max(np.abs(b))
    return np.max(np.abs(a - b)) / (scale if scale > 0 else 1.0)


def subspace_sin(A, B):
    """sine of the largest principal angle between the row spaces of A and B."""
    Qa, _ = np.linalg.qr(A.T)
    Qb, _ = np.linalg.qr(B.T)
    return np.linalg.norm(Qb - Qa @ (Qa.T @ Qb), 2)


# --------------------------------------------------------------------------- K1 + K2
@pytest.mark.parametrize("cutoff", [5, 10, 15])
@pytest.mark.parametrize("use_cell_list", [False, True])
def test_kirchhoff_random500_bit_exact(cutoff, use_cell_list, monkeypatch):
    """test_interaction.py:11-40 -- ProDy Kirchhoff, both contact kernels."""
    from springcraft_b200 import _engine
    monkeypatch.setattr(_engine, "CELL_LIST_MIN_N", 0)  # force the cell-list kernels when asked
    tp = golden("thirdparty_random500.npz")
    ref = golden("ref_random500.npz")
    K, pairs = sc.compute_kirchhoff(tp["coord"], sc.InvariantForceField(cutoff), use_cell_list)
    assert pairs.dtype == np.int64
    assert np.array_equal(pairs, ref[f"pairs_{cutoff}_1"])
    assert np.array_equal(K, tp[f"prody_kirchhoff_{cutoff}"].astype(float))


@pytest.mark.parametrize("cutoff", [5, 10, 15])
def test_hessian_random500(cutoff):
    ref = golden("ref_random500.npz")
    tp = golden("thirdparty_random500.npz")
    H, pairs = sc.compute_hessian(tp["coord"], sc.InvariantForceField(cutoff))
    b = H.reshape(500, 3, 500, 3).transpose(0, 2, 1, 3)
    assert np.array_equal(b[pairs[:, 0], pairs[:, 1]], ref[f"hessian_offdiag_{cutoff}"])
    assert np.array_equal(b[np.arange(500), np.arange(500)], ref[f"hessian_diag_{cutoff}"])
    assert np.count_nonzero(b) == np.count_nonzero(ref[f"hessian_offdiag_{cutoff}"]) + \
        np.count_nonzero(ref[f"hessian_diag_{cutoff}"])


@pytest.mark.parametrize("key", sorted(FF))
def test_1l2y_assembly(structures, key):
    ref = golden("ref_1l2y.npz")
    atoms = atoms_of(structures, "1l2y")
    ff = FF[key](atoms)
    H, pairs = sc.compute_hessian(atoms.coord, ff)
    K, pairs_k = sc.compute_kirchhoff(atoms.coord, ff)
    assert np.array_equal(pairs, ref[f"{key}/pairs"])
    assert np.array_equal(pairs_k, ref[f"{key}/pairs"])
    if key.startswith("hinsen"):   # pow(): libm vs CUDA, <= 1e-12 relative
        assert rel_err(K, ref[f"{key}/kirchhoff"]) <= HESS_RTOL
        assert rel_err(H, ref[f"{key}/hessian"]) <= HESS_RTOL
    else:
        assert np.array_equal(K, ref[f"{key}/kirchhoff"])
        assert np.array_equal(H, ref[f"{key}/hessian"])
    # mass weighting (anm.py:89-113)
    anm = sc.ANM(atoms, ff, masses=ref["masses"])
    assert rel_err(anm.hessian, ref[f"{key}/mw_hessian"]) <= HESS_RTOL
    gnm = sc.GNM(atoms, ff, masses=ref["masses"])
    assert rel_err(gnm.kirchhoff, ref[f"{key}/gnm_mw_kirchhoff"]) <= HESS_RTOL
    if not key.startswith("hinsen"):
        assert np.array_equal(anm.hessian, ref[f"{key}/mw_hessian"])


def test_two_chain_patched():
    """test_forcefield.py:39-114 fixture: overlapping chains + PatchedForceField."""
    ref = golden("ref_two_chain.npz")
    atoms = sc.AtomArray(ref["coord"], ref["res_name"], ref["chain_id"], ref["res_id"])
    base = sc.InvariantForceField(7.0)
    K, pairs = sc.compute_kirchhoff(atoms.coord, base)
    assert np.array_equal(pairs, ref["invariant7/pairs"])
    assert np.array_equal(K, ref["invariant7/kirchhoff"])
    patches = {
        "shutdown": dict(contact_shutdown=ref["shutdown"]),
        "pair_off": dict(contact_pair_off=ref["pair_off"]),
        "pair_on": dict(contact_pair_on=ref["pair_on"], force_constants=ref["pair_on_fc"]),
        "all": dict(contact_shutdown=ref["shutdown"], contact_pair_off=ref["pair_off"],
                    contact_pair_on=ref["pair_on"], force_constants=ref["pair_on_fc"]),
    }
    for tag, kw in patches.items():
        K, pairs = sc.compute_kirchhoff(atoms.coord, sc.PatchedForceField(base, **kw))
        assert np.array_equal(pairs, ref[f"patched_{tag}/pairs"]), tag
        assert np.array_equal(K, ref[f"patched_{tag}/kirchhoff"]), tag
    for key in ("e_anm", "sd_enm", "d_enm", "s_enm_13"):
        ff = FF[key](atoms)
        K, pairs = sc.compute_kirchhoff(atoms.coord, ff)
        assert np.array_equal(pairs, ref[f"{key}/pairs"])
        assert np.array_equal(K, ref[f"{key}/kirchhoff"]), key
        pf = sc.PatchedForceField(ff, contact_pair_off=ref["pair_off"], contact_pair_on=ref["pair_on"],
                                  force_constants=ref["pair_on_fc"])
        K, _ = sc.compute_kirchhoff(atoms.coord, pf)
        assert np.array_equal(K, ref[f"{key}_patched/kirchhoff"]), key
    shifted = sc.AtomArray(ref["shifted_coord"], ref["res_name"], ref["chain_id"], ref["res_id"])
    for key in ("e_anm", "sd_enm", "hinsen", "invariant13"):
        ff = FF[key](shifted)
        H, pairs = sc.compute_hessian(shifted.coord, ff)
        assert np.array_equal(pairs, ref[f"shifted_{key}/pairs"])
        assert rel_err(H, ref[f"shifted_{key}/hessian"]) <= HESS_RTOL, key
        pf = sc.PatchedForceField(ff, contact_shutdown=ref["shutdown"], contact_pair_off=ref["pair_off"],
                                  contact_pair_on=ref["pair_on"], force_constants=ref["pair_on_fc"])
        H, pairs = sc.compute_hessian(shifted.coord, pf)
        assert np.array_equal(pairs, ref[f"shifted_{key}_patched/pairs"])
        assert rel_err(H, ref[f"shifted_{key}_patched/hessian"]) <= HESS_RTOL, key


@pytest.mark.parametrize("seed", [0, 7])
@pytest.mark.parametrize("cutoff", [5, 10, 15])
@pytest.mark.parametrize("use_cell_list", [False, True])
def test_cloud1000_vs_oracle(seed, cutoff, use_cell_list, monkeypatch):
    """test_interaction.py:71-89 generator (n=1000, box 50), checked against the oracle."""
    from springcraft_b200 import _engine
    monkeypatch.setattr(_engine, "CELL_LIST_MIN_N", 0)
    np.random.seed(seed)
    coord = np.random.rand(1000, 3) * 50
    H, pairs = sc.compute_hessian(coord, sc.InvariantForceField(cutoff), use_cell_list)
    Ho, pairs_o = orc.compute_hessian(coord, orc.FFSpec("invariant", float(cutoff)))
    assert np.array_equal(pairs, pairs_o)
    assert np.array_equal(H, Ho)
    assert np.allclose(H, H.T)


def test_cartesian_index_product_user_forcefield():
    """test_interaction.py:92-116: user-defined ForceField without cutoff."""
    class AllConnectedForceField(sc.ForceField):
        def force_constant(self, atom_i, atom_j, sq_distance):
            return np.ones(len(atom_i))

    np.random.seed(0)
    coord = np.random.rand(10, 3) * 50
    _, pairs = sc.compute_hessian(coord, AllConnectedForceField())
    m = np.zeros((10, 10), dtype=bool)
    m[tuple(pairs.T)] = True
    assert (m == ~np.identity(10).astype(bool)).all()


def test_force_constant_method(structures):
    atoms = atoms_of(structures, "1l2y")
    rng = np.random.default_rng(0)
    i = rng.integers(0, 20, 200)
    j = rng.integers(0, 20, 200)
    sq = rng.uniform(4.0, 16.4, 200) ** 2
    pairs = np.stack([i, j], 1)
    for key, spec in (("hinsen", orc.FFSpec("hinsen")), ("pfree", orc.FFSpec("pfree")),
                      ("sd_enm", orc.preset_spec("sd_enm", atoms.res_name, atoms.chain_id, atoms.res_id)),
                      ("e_anm", orc.preset_spec("e_anm", atoms.res_name, atoms.chain_id, atoms.res_id))):
        got = FF[key](atoms).force_constant(i, j, sq)
        want = orc.force_constants(spec, pairs, sq)
        assert got.dtype == want.dtype, key
        assert np.allclose(got, want, rtol=1e-13, atol=0), key
    with pytest.raises(ValueError):
        FF["sd_enm"](atoms).force_constant(np.array([0]), np.array([5]), np.array([17.0 ** 2]))


# --------------------------------------------------------------------------- K3
@pytest.mark.parametrize("D", [1, 3])
@pytest.mark.parametrize("shape", [(1, 257, 32), (3, 300, 32), (70, 120, 32), (2, 301, 64), (1, 2500, 32)])
def test_spmm_kernels_vs_dense(D, shape):
    """Both SpMM kernels (row-per-warp and row-paired) against the dense product."""
    import torch
    from springcraft_b200._engine import DeviceModel
    B, n, b = shape
    rng = np.random.default_rng(B * 1000 + n)
    base = orc.synthetic_chain(n, seed=2)
    coords = base[None] + rng.normal(0, 0.4, size=(B, n, 3))
    model = DeviceModel(coords, sc.HinsenForceField(11.0), D, masses=rng.uniform(50, 200, n))
    dense = model.dense()
    X = torch.from_numpy(rng.normal(size=(B, D * n, b))).cuda()
    want = torch.bmm(dense, X).cpu().numpy()
    scale = np.abs(want).max()
    assert np.abs(model.spmm(X).cpu().numpy() - want).max() <= 1e-13 * scale
    assert np.abs(model.spmm_paired(X).cpu().numpy() - want).max() <= 1e-13 * scale


@pytest.mark.parametrize("key", ["invariant7", "invariant13", "hinsen", "e_anm", "sd_enm", "pfree"])
def test_1l2y_full_spectrum(structures, key):
    ref = golden("ref_1l2y.npz")
    atoms = atoms_of(structures, "1l2y")
    anm = sc.ANM(atoms, FF[key](atoms))
    lam, vec = anm.eigen()
    want = ref[f"{key}/anm_eigval"]
    scale = np.abs(want).max()
    assert np.allclose(lam[6:], want[6:], rtol=EIG_RTOL, atol=1e-12 * scale)
    assert np.abs(lam[:6]).max() <= 1e-9 * scale
    assert np.allclose(vec @ vec.T, np.eye(60), atol=1e-12)
    H = ref[f"{key}/hessian"]
    assert np.abs(H @ vec.T - vec.T * lam).max() <= 1e-11 * scale
    gnm = sc.GNM(atoms, FF[key](atoms))
    lam, vec = gnm.eigen()
    want = ref[f"{key}/gnm_eigval"]
    assert np.allclose(lam[1:], want[1:], rtol=EIG_RTOL, atol=1e-12 * np.abs(want).max())


def test_c2_chain1000_full_spectrum():
    """Config C2: 1,000-residue chain, Hinsen (all pairs), full spectrum + MSF + B-factor."""
    ref = golden("ref_c2_chain1000.npz")
    anm = sc.ANM(ref["coord"], sc.HinsenForceField())
    assert rel_err(np.diagonal(anm.hessian), ref["hessian_diag"]) <= HESS_RTOL
    lam, modes = anm.eigen()
    want = ref["eigval"]
    assert np.allclose(lam[6:], want[6:], rtol=EIG_RTOL, atol=0)
    assert np.abs(lam[:6]).max() <= 1e-9 * want.max()
    assert subspace_sin(modes[6:26], ref["modes_6_26"]) < ANGLE_TOL
    assert np.allclose(anm.mean_square_fluctuation(), ref["msf"], rtol=PROD_RTOL)
    assert np.allclose(anm.bfactor(), ref["bfactor"], rtol=PROD_RTOL)


@pytest.mark.parametrize("N", [65, 128, 200, 331])
def test_eig_full_block_random(N):
    """Full eigensolver (N > 64: tridiagonalisation + divide and conquer) vs LAPACK on random symmetric matrices
    of small order (including a rank-deficient PSD one), lower triangle referenced."""
    import torch
    from springcraft_b200 import _engine
    rng = np.random.default_rng(N)
    A = rng.normal(size=(N, N))
    A = A + A.T
    if N == 200:
        G = rng.normal(size=(N, N - 7))
        A = G @ G.T
    Al = np.tril(A) + 1e3 * np.triu(rng.normal(size=(N, N)), 1)   # garbage in the strict upper triangle
    lam, modes = _engine.eig_full_dense(torch.from_numpy(Al).cuda())
    lam, modes = lam[0].cpu().numpy(), modes[0].cpu().numpy()
    want = np.linalg.eigvalsh(A)
    scale = np.abs(want).max()
    assert np.allclose(lam, want, rtol=0, atol=1e-12 * scale)
    assert np.allclose(modes @ modes.T, np.eye(N), atol=1e-12)
    assert np.abs(A @ modes.T - modes.T * lam).max() <= 1e-11 * scale


def _check_full(A, lam, modes, tol=1e-12):
    want = np.linalg.eigvalsh(A)
    scale = max(np.abs(want).max(), 1e-300)
    N = len(want)
    assert np.isfinite(modes).all()
    assert np.allclose(lam, want, rtol=0, atol=tol * scale)
    assert np.allclose(modes @ modes.T, np.eye(N), atol=tol)
    assert np.abs(A @ modes.T - modes.T * lam).max() <= 10 * tol * scale


@pytest.mark.parametrize("N", [257, 300, 515, 1001])
def test_eig_full_tridiag_random(N):
    """N > 64: Householder tridiagonalisation + divide and conquer + back-transformation (eig_full_tridiag.cu)
    vs LAPACK (np.linalg.eigh, reference nma.py:61): random symmetric matrices of even and odd order, lower
    triangle referenced, 1e-13-class eigenvalues, orthogonality and residuals."""
    import torch
    from springcraft_b200 import _engine
    rng = np.random.default_rng(N)
    A = rng.normal(size=(N, N))
    A = A + A.T
    Al = np.tril(A) + 1e3 * np.triu(rng.normal(size=(N, N)), 1)   # garbage in the strict upper triangle
    lam, modes = _engine.eig_full_dense(torch.from_numpy(Al).cuda())
    _check_full(A, lam[0].cpu().numpy(), modes[0].cpu().numpy(), tol=2e-13)


@pytest.mark.parametrize("kind", ["rank_deficient", "diagonal", "clustered", "decoupled", "toeplitz", "glued", "scaled"])
def test_eig_full_tridiag_special(kind):
    """Spectra that stress the deflation and the secular-equation solver of the divide and conquer: many zero
    eigenvalues, an already diagonal matrix (every reflector is the identity), a tight cluster, two decoupled
    blocks (zero off-diagonal of the tridiagonal form), the 1-2-1 Toeplitz matrix (deflation by rotation at
    every level), glued Wilkinson blocks (pairs of eigenvalues agreeing to 1e-9) and a badly scaled matrix."""
    import torch
    from springcraft_b200 import _engine
    N = 420
    rng = np.random.default_rng(11)
    A = rng.normal(size=(N, N)); A = A + A.T
    tol = 2e-13
    if kind == "rank_deficient":
        G = rng.normal(size=(N, N - 60)); A = G @ G.T
    elif kind == "diagonal":
        A = np.diag(rng.normal(size=N))
    elif kind == "clustered":
        A = np.eye(N) + 1e-7 * A
    elif kind == "decoupled":
        A[:200, 200:] = 0.0; A[200:, :200] = 0.0
    elif kind == "toeplitz":
        A = 2.0 * np.eye(N) - np.eye(N, k=1) - np.eye(N, k=-1)
    elif kind == "glued":
        W = np.abs(np.arange(-10, 11)).astype(float)
        A = np.zeros((N, N))
        for b in range(N // 21):
            s = 21 * b
            A[s:s + 21, s:s + 21] = np.diag(W) + np.eye(21, k=1) + np.eye(21, k=-1)
            if b: A[s - 1, s] = A[s, s - 1] = 1e-9
    elif kind == "scaled":
        sc_ = 10.0 ** rng.uniform(-6, 6, size=N)
        A = A * np.sqrt(np.outer(sc_, sc_))
    lam, modes = _engine.eig_full_dense(torch.from_numpy(A.copy()).cuda())
    _check_full(A, lam[0].cpu().numpy(), modes[0].cpu().numpy(), tol=tol)


def test_eig_full_tridiag_batched():
    """Batches share every launch of the tridiagonal solver: the CTA groups of the reduction work on different
    matrices concurrently (more matrices than groups, so groups also take several in turn), the merges of all
    matrices share the launches of a level; matrices with different deflation patterns must each match LAPACK."""
    import torch
    from springcraft_b200 import _engine
    N = 300
    rng = np.random.default_rng(6)
    mats = []
    for kind in range(11):
        A = rng.normal(size=(N, N)); A = A + A.T
        if kind % 4 == 1:
            G = rng.normal(size=(N, N - 40)); A = G @ G.T
        elif kind % 4 == 2:
            A = np.diag(rng.normal(size=N))
        elif kind % 4 == 3:
            A = np.eye(N) + 1e-6 * A
        mats.append(A)
    lam, modes = _engine.eig_full_dense(torch.from_numpy(np.stack(mats)).cuda())
    lam, modes = lam.cpu().numpy(), modes.cpu().numpy()
    for A, l, m in zip(mats, lam, modes):
        _check_full(A, l, m, tol=2e-13)


def test_eig_full_tridiag_fuzz():
    """Seeded fuzz of the full-spectrum solver: random orders 65..500 (odd and even: both alignment paths of the
    grouped merge products), batches of 1..4, eight kinds of spectra (indefinite, rank deficient, two clusters 1e-10
    apart, decoupled blocks, banded, almost diagonal, integer eigenvalues with high multiplicity, negative definite)."""
    import torch
    from springcraft_b200 import _engine
    rng = np.random.default_rng(2024)
    kinds = ["rand", "psd_def", "clustered", "blocks", "band", "sparse_diag", "repeat", "neg"]
    for it in range(16):
        N = int(rng.integers(65, 500))
        B = int(rng.integers(1, 5))
        kind = kinds[it % len(kinds)]
        mats = []
        for _ in range(B):
            M = rng.standard_normal((N, N)); A = M + M.T
            if kind == "psd_def":
                G = rng.standard_normal((N, int(rng.integers(1, N)))); A = G @ G.T
            elif kind == "clustered":
                Q, _ = np.linalg.qr(M)
                lam = np.concatenate([np.ones(N // 2), 1 + 1e-10 * rng.standard_normal(N - N // 2)])
                A = (Q * lam) @ Q.T; A = (A + A.T) / 2
            elif kind == "blocks":
                c = int(rng.integers(1, N - 1)); A[:c, c:] = 0; A[c:, :c] = 0
            elif kind == "band":
                bw = int(rng.integers(1, 6)); A = np.triu(np.tril(A, bw), -bw)
            elif kind == "sparse_diag":
                A = np.diag(rng.standard_normal(N)); i, j = rng.integers(0, N, 2); A[i, j] = A[j, i] = 0.5
            elif kind == "repeat":
                Q, _ = np.linalg.qr(M); A = (Q * rng.integers(0, 4, N).astype(float)) @ Q.T; A = (A + A.T) / 2
            elif kind == "neg":
                A = -np.abs(A) - 5 * np.eye(N)
            mats.append(A)
        lam, modes = _engine.eig_full_dense(torch.from_numpy(np.stack(mats)).cuda())
        lam, modes = lam.cpu().numpy(), modes.cpu().numpy()
        for A, l, m in zip(mats, lam, modes):
            _check_full(A, l, m, tol=5e-13)


def test_eig_full_block_jacobi_fallback(monkeypatch):
    """The two-sided block Jacobi solver (round 1) stays as the fallback for orders the tridiagonal solver does not
    take (N > 9,200, devices without cooperative launch); SCB_EIG_FULL=jacobi selects it."""
    import torch
    from springcraft_b200 import _engine
    monkeypatch.setenv("SCB_EIG_FULL", "jacobi")
    rng = np.random.default_rng(9)
    N = 200
    A = rng.normal(size=(N, N)); A = A + A.T
    lam, modes = _engine.eig_full_dense(torch.from_numpy(A.copy()).cuda())
    _check_full(A, lam[0].cpu().numpy(), modes[0].cpu().numpy(), tol=1e-12)
    # the explicit selector of the ABI (scb_eig_full_ex): what the dense row-slab solver asks for on several ranks
    monkeypatch.delenv("SCB_EIG_FULL")
    for solver in ("jacobi", "tridiag", "auto"):
        lam, modes = _engine.eig_full_dense(torch.from_numpy(A.copy()).cuda(), solver=solver)
        _check_full(A, lam[0].cpu().numpy(), modes[0].cpu().numpy(), tol=1e-12)


def test_eig_full_block_batched():
    """A batch of matrices shares every launch of the block-Jacobi solver (grid.y = matrix): matrices that
    converge after different numbers of sweeps (random, rank-deficient, already diagonal, tightly clustered)
    must each match LAPACK."""
    import torch
    from springcraft_b200 import _engine
    N = 130
    rng = np.random.default_rng(5)
    mats = []
    for kind in range(5):
        A = rng.normal(size=(N, N)); A = A + A.T
        if kind == 1:
            G = rng.normal(size=(N, N - 11)); A = G @ G.T
        elif kind == 2:
            A = np.diag(rng.normal(size=N))
        elif kind == 3:
            A = np.eye(N) + 1e-6 * A
        mats.append(A)
    lam, modes = _engine.eig_full_dense(torch.from_numpy(np.stack(mats)).cuda())
    lam, modes = lam.cpu().numpy(), modes.cpu().numpy()
    for A, l, m in zip(mats, lam, modes):
        want = np.linalg.eigvalsh(A)
        scale = np.abs(want).max()
        assert np.allclose(l, want, rtol=0, atol=1e-12 * scale)
        assert np.allclose(m @ m.T, np.eye(N), atol=1e-12)
        assert np.abs(A @ m.T - m.T * l).max() <= 1e-11 * scale


@pytest.mark.parametrize("key", ["e_anm", "sd_enm"])
@pytest.mark.parametrize("c", [0, 1, 2, 4095])
def test_c3_lowest_modes(key, c):
    """Config C3 member: 300-residue chain, lowest 20 non-trivial modes + MSF."""
    ref = golden("ref_c3_chain300.npz")
    coord = orc.perturbed_conformation(ref["base"], c)
    atoms = sc.AtomArray(coord, ref["res_name"], ref["chain_id"], ref["res_id"])
    ff = FF[key](atoms)
    anm = sc.ANM(coord, ff)
    lam, modes = anm.eigen(k=26)
    want = ref[f"c{c}/{key}/eigval"]
    assert np.allclose(lam[6:26], want[6:26], rtol=EIG_RTOL, atol=0)
    assert subspace_sin(modes[6:26], ref[f"c{c}/{key}/modes_6_26"]) < ANGLE_TOL
    # trivial modes: analytic rigid-body basis, orthonormal and in the null space
    assert np.allclose(modes[:6] @ modes[:6].T, np.eye(6), atol=1e-12)
    msf = anm.mean_square_fluctuation(mode_subset=np.arange(6, 26))
    assert np.allclose(msf, ref[f"c{c}/{key}/msf_6_26"], rtol=PROD_RTOL, atol=0)
    if c == 0:
        assert rel_err(anm.hessian, ref[f"c0/{key}/hessian"]) <= HESS_RTOL


def test_c3_ensemble_api():
    ref = golden("ref_c3_chain300.npz")
    members = [0, 1, 2, 4095]
    coords = np.stack([orc.perturbed_conformation(ref["base"], c) for c in members])
    atoms = sc.AtomArray(ref["base"], ref["res_name"], ref["chain_id"], ref["res_id"])
    for key in ("e_anm", "sd_enm"):
        res = sc.enm_ensemble(coords, FF[key](atoms), k=20, return_modes=True)
        assert res.converged
        for q, c in enumerate(members):
            assert np.allclose(res.eigenvalues[q], ref[f"c{c}/{key}/eigval"][6:26], rtol=EIG_RTOL, atol=0)
            assert np.allclose(res.msf[q], ref[f"c{c}/{key}/msf_6_26"], rtol=PROD_RTOL, atol=0)
            assert subspace_sin(res.modes[q], ref[f"c{c}/{key}/modes_6_26"]) < ANGLE_TOL
        assert res.n_pairs == sum(int(ref[f"c{c}/{key}/n_pairs"]) for c in members)


def test_chain1500_lowest_modes_vs_oracle():
    """Sparse lowest-k path on a mid-size system (cell-list contacts are used from n=2048 on, so
    both a 1,500- and a 2,500-residue chain are checked) against the oracle's dense eigh."""
    for n, use_eigh in ((1500, True), (2500, False)):
        coord = orc.synthetic_chain(n, seed=5)
        anm = sc.ANM(coord, sc.InvariantForceField(13.0))
        lam, modes = anm.eigen(k=26)
        Ho, pairs_o = orc.compute_hessian(coord, orc.FFSpec("invariant", 13.0))
        assert anm._model_device().P == len(pairs_o)
        scale = np.abs(Ho).sum(1).max()
        # residuals against the ORACLE's matrix: || H u - lam u || small, modes orthonormal
        R = Ho @ modes[6:].T - modes[6:].T * lam[6:]
        assert np.abs(R).max() <= 1e-9 * scale
        assert np.allclose(modes @ modes.T, np.eye(26), atol=1e-10)
        assert np.abs(Ho @ modes[:6].T).max() <= 1e-9 * scale        # analytic rigid-body modes
        if use_eigh:
            want, vec = np.linalg.eigh(Ho)
            assert np.allclose(lam[6:26], want[6:26], rtol=EIG_RTOL, atol=0)
            assert subspace_sin(modes[6:26], vec.T[6:26]) < ANGLE_TOL
            msf = anm.mean_square_fluctuation(mode_subset=np.arange(6, 26))
            want_msf = orc.mean_square_fluctuation(want, vec.T, 3, mode_subset=np.arange(6, 26))
            assert np.allclose(msf, want_msf, rtol=PROD_RTOL, atol=0)


def test_ensemble_gnm_and_tiny_systems(structures):
    """Batched path for GNM (D=1) and for structures smaller than the solver block (dense fallback)."""
    ref = golden("ref_c3_chain300.npz")
    coords = np.stack([orc.perturbed_conformation(ref["base"], c) for c in (0, 1, 2)])
    res = sc.enm_ensemble(coords, sc.InvariantForceField(10.0), k=20, kind="gnm", return_modes=True)
    for q in range(3):
        K, _ = orc.compute_kirchhoff(coords[q], orc.FFSpec("invariant", 10.0))
        lam, vec = np.linalg.eigh(K)
        assert np.allclose(res.eigenvalues[q], lam[1:21], rtol=EIG_RTOL, atol=0)
        assert subspace_sin(res.modes[q], vec.T[1:21]) < ANGLE_TOL
        want = orc.mean_square_fluctuation(lam, vec.T, 1, mode_subset=np.arange(1, 21))
        assert np.allclose(res.msf[q], want, rtol=PROD_RTOL, atol=0)
    # 20-residue Trp-cage: N = 60 < block width -> shared-memory Jacobi path, same API
    g = golden("ref_1l2y.npz")
    c = structures["1l2y_coord"].astype(np.float64)
    rng = np.random.default_rng(3)
    tiny = np.stack([c, c + rng.normal(0, 0.05, c.shape)])
    res = sc.enm_ensemble(tiny, sc.InvariantForceField(13.0), k=20, return_modes=True)
    assert np.allclose(res.eigenvalues[0], g["invariant13/anm_eigval"][6:26], rtol=EIG_RTOL)
    assert np.allclose(res.msf[0], g["invariant13/anm_msf_sub"], rtol=PROD_RTOL)
    H1, _ = orc.compute_hessian(tiny[1], orc.FFSpec("invariant", 13.0))
    assert np.allclose(res.eigenvalues[1], np.linalg.eigvalsh(H1)[6:26], rtol=EIG_RTOL)


def test_ensemble_variants_vs_oracle():
    """Ensemble path with mass weighting, a patched force field and a distance-dependent one."""
    n = 200
    base = orc.synthetic_chain(n, seed=9)
    res_name, chain_id, res_id = orc.synthetic_sequence(n, seed=9)
    rng = np.random.default_rng(9)
    coords = base[None] + rng.normal(0, 0.3, size=(3, n, 3))
    masses = rng.uniform(60, 190, n)
    atoms = sc.AtomArray(base, res_name, chain_id, res_id)
    pair_off = np.array([[0, 1], [50, 51]])
    pair_on = np.array([[3, 150], [10, 190]])
    fcs = np.array([5.0, 2.5])
    cases = [
        (sc.HinsenForceField(12.0), orc.FFSpec("hinsen", 12.0), masses),
        (sc.PatchedForceField(sc.TabulatedForceField.e_anm(atoms), contact_pair_off=pair_off, contact_pair_on=pair_on,
                              force_constants=fcs),
         None, None),
        (sc.TabulatedForceField.d_enm(atoms), orc.preset_spec("d_enm", res_name, chain_id, res_id), masses),
    ]
    spec_p = orc.preset_spec("e_anm", res_name, chain_id, res_id)
    spec_p.patched, spec_p.pair_off, spec_p.pair_on, spec_p.pair_on_fc = True, pair_off, pair_on, fcs
    cases[1] = (cases[1][0], spec_p, None)
    for ff, spec, m in cases:
        res = sc.enm_ensemble(coords, ff, k=20, masses=m, return_modes=True)
        assert res.converged
        for q in range(len(coords)):
            H, _ = orc.compute_hessian(coords[q], spec, masses=m)
            lam, vec = np.linalg.eigh(H)
            assert np.allclose(res.eigenvalues[q], lam[6:26], rtol=EIG_RTOL, atol=0)
            assert subspace_sin(res.modes[q], vec.T[6:26]) < ANGLE_TOL
            want = orc.mean_square_fluctuation(lam, vec.T, 3, mode_subset=np.arange(6, 26))
            assert np.allclose(res.msf[q], want, rtol=PROD_RTOL, atol=0)


def test_ensemble_all_modes_vs_oracle():
    """``enm_ensemble(k=None)``: every non-trivial mode per conformation through the batched full-spectrum
    solver (the reference's default mode set for MSF, nma.py:145-151), ANM with masses and GNM."""
    n = 100
    base = orc.synthetic_chain(n, seed=4)
    res_name, chain_id, res_id = orc.synthetic_sequence(n, seed=4)
    rng = np.random.default_rng(4)
    coords = base[None] + rng.normal(0, 0.3, size=(5, n, 3))
    masses = rng.uniform(60, 190, n)
    atoms = sc.AtomArray(base, res_name, chain_id, res_id)
    spec = orc.preset_spec("e_anm", res_name, chain_id, res_id)
    res = sc.enm_ensemble(coords, sc.TabulatedForceField.e_anm(atoms), k=None, masses=masses, return_modes=True)
    assert res.eigenvalues.shape == (5, 3 * n - 6) and res.msf.shape == (5, n)
    for q in range(len(coords)):
        H, _ = orc.compute_hessian(coords[q], spec, masses=masses)
        lam, vec = np.linalg.eigh(H)
        assert np.allclose(res.eigenvalues[q], lam[6:], rtol=EIG_RTOL, atol=1e-10 * lam[-1])
        assert np.allclose(res.msf[q], orc.mean_square_fluctuation(lam, vec.T, 3), rtol=PROD_RTOL, atol=0)
        assert np.abs(H @ res.modes[q].T - res.modes[q].T * res.eigenvalues[q]).max() <= 1e-10 * lam[-1]
    gres = sc.enm_ensemble(coords, sc.InvariantForceField(10.0), k=None, kind="gnm")
    for q in range(len(coords)):
        K, _ = orc.compute_kirchhoff(coords[q], orc.FFSpec("invariant", 10.0))
        lam, vec = np.linalg.eigh(K)
        assert np.allclose(gres.eigenvalues[q], lam[1:], rtol=EIG_RTOL, atol=1e-10 * lam[-1])
        assert np.allclose(gres.msf[q], orc.mean_square_fluctuation(lam, vec.T, 1), rtol=PROD_RTOL, atol=0)


def test_ensemble_chunked_equals_single_call(monkeypatch):
    """`enm_ensemble` feeds the library in chunks (grid limits, device memory): forcing 3-structure chunks
    must reproduce the single-call result for every conformation, uneven tail included."""
    ref = golden("ref_c3_chain300.npz")
    coords = np.stack([orc.perturbed_conformation(ref["base"], c) for c in range(8)])
    atoms = sc.AtomArray(ref["base"], ref["res_name"], ref["chain_id"], ref["res_id"])
    ff = sc.TabulatedForceField.e_anm(atoms)
    whole = sc.enm_ensemble(coords, ff, k=20, return_modes=True)
    monkeypatch.setenv("SCB_ENSEMBLE_CHUNK", "3")
    parts = sc.enm_ensemble(coords, ff, k=20, return_modes=True)
    assert parts.converged and whole.converged and parts.n_pairs == whole.n_pairs
    assert np.allclose(parts.eigenvalues, whole.eigenvalues, rtol=1e-10, atol=0)
    assert np.allclose(parts.msf, whole.msf, rtol=1e-8, atol=0)
    for q in range(8):
        assert subspace_sin(parts.modes[q], whole.modes[q]) < ANGLE_TOL


def _fuzz_force_field(ref, i, atoms):
    pre = f"case{i}/"
    kind, cutoff = str(ref[pre + "kind"]), float(ref[pre + "cutoff"])
    if kind == "invariant":
        ff = sc.InvariantForceField(cutoff)
    elif kind == "hinsen":
        ff = sc.HinsenForceField(cutoff)
    elif kind == "hinsen_nocut":
        ff = sc.HinsenForceField()
    elif kind == "pfree":
        ff = sc.ParameterFreeForceField(cutoff)
    elif kind == "pfree_nocut":
        ff = sc.ParameterFreeForceField()
    elif kind == "e_anm_mean":
        ff = sc.TabulatedForceField.e_anm(atoms, nonbonded_mean=True)
    else:
        ff = getattr(sc.TabulatedForceField, kind)(atoms)
    if bool(ref[pre + "patched"]):
        sd = ref[pre + "shutdown"]
        ff = sc.PatchedForceField(ff, contact_shutdown=sd if len(sd) else None, contact_pair_off=ref[pre + "pair_off"],
                                  contact_pair_on=ref[pre + "pair_on"], force_constants=ref[pre + "pair_on_fc"])
    return ff


@pytest.mark.parametrize("i", range(26))
def test_fuzz_cases_vs_reference(i):
    """26 seeded random cases generated by the unmodified reference (chain breaks, numbering gaps, random
    sequences, every force-field kind, random patches, masses): contacts exact, matrices <= 1e-12, spectra
    <= 1e-8 of the spectral radius, MSF <= 1e-8 where the matrix has no extra (near-)null directions."""
    ref = golden("ref_fuzz.npz")
    pre = f"case{i}/"
    atoms = sc.AtomArray(ref[pre + "coord"], ref[pre + "res_name"], ref[pre + "chain_id"], ref[pre + "res_id"])
    ff = _fuzz_force_field(ref, i, atoms)
    masses = ref[pre + "masses"] if pre + "masses" in ref else None
    H, pairs = sc.compute_hessian(atoms.coord, ff)
    K, _ = sc.compute_kirchhoff(atoms.coord, ff)
    assert np.array_equal(pairs, ref[pre + "pairs"])
    assert rel_err(K, ref[pre + "kirchhoff"]) <= HESS_RTOL and rel_err(H, ref[pre + "hessian"]) <= HESS_RTOL
    if str(ref[pre + "kind"]) == "invariant":
        assert np.array_equal(K, ref[pre + "kirchhoff"])
    for cls, key, ntriv in ((sc.ANM, "anm", 6), (sc.GNM, "gnm", 1)):
        enm = cls(atoms, ff, masses=masses)
        M = enm.hessian if key == "anm" else enm.kirchhoff
        assert rel_err(M, ref[pre + f"{key}_matrix"]) <= HESS_RTOL
        want = ref[pre + f"{key}_eigval"]
        lam, modes = enm.eigen()
        assert np.allclose(lam, want, rtol=0, atol=EIG_RTOL * np.abs(want).max())
        assert np.abs(M @ modes.T - modes.T * lam).max() <= 1e-10 * np.abs(want).max()
        if want[ntriv] > 1e-3 * want[-1]:
            assert np.allclose(enm.mean_square_fluctuation(), ref[pre + f"{key}_msf"], rtol=PROD_RTOL, atol=0)


def test_degenerate_inputs():
    """No contacts at all, and structures of 2-3 nodes (edge cases of the contact / assembly kernels)."""
    rng = np.random.default_rng(1)
    coord = rng.random((12, 3)) * 100.0
    K, pairs = sc.compute_kirchhoff(coord, sc.InvariantForceField(0.5))
    assert pairs.shape == (0, 2) and not K.any()
    H, pairs = sc.compute_hessian(coord, sc.InvariantForceField(0.5))
    assert pairs.shape == (0, 2) and H.shape == (36, 36) and not H.any()
    for n in (2, 3):
        c = rng.random((n, 3)) * 5.0
        H, pairs = sc.compute_hessian(c, sc.HinsenForceField())
        Ho, po = orc.compute_hessian(c, orc.FFSpec("hinsen"))
        assert np.array_equal(pairs, po) and rel_err(H, Ho) <= HESS_RTOL
        lam, _ = sc.ANM(c, sc.HinsenForceField()).eigen()
        assert np.allclose(lam, np.linalg.eigvalsh(Ho), atol=1e-9 * np.abs(Ho).max())
    with pytest.raises(ValueError):
        sc.compute_hessian(np.zeros((4, 2)), sc.InvariantForceField(5.0))
    ca = sc.AtomArray(rng.random((5, 3)), res_name=["ALA"] * 5)
    with pytest.raises(ValueError):   # natoms mismatch (interaction.py:143-147)
        sc.compute_hessian(rng.random((6, 3)), sc.TabulatedForceField.e_anm(ca))
    with pytest.raises(ValueError):   # contact beyond the last bin edge via a patched pair (forcefield.py:526-533)
        far = rng.random((5, 3)) * 3.0
        far[4] += 100.0
        ff = sc.TabulatedForceField.sd_enm(ca)
        sc.compute_hessian(far, sc.PatchedForceField(sc.InvariantForceField(8.0), contact_pair_on=[[0, 4]],
                                                     force_constants=[1.0]))
        ff.force_constant(np.array([0]), np.array([4]), np.array([100.0 ** 2]))


def test_smoke_entry():
    import __graft_entry__
    __graft_entry__.smoke()


@pytest.mark.parametrize("key", ["invariant13", "e_anm"])
def test_7cal_lowest_modes(structures, key):
    """1,776-residue tetramer (test_anm.py:60-84 structure), lowest modes."""
    ref = golden("ref_7cal.npz")
    atoms = atoms_of(structures, "7cal")
    anm = sc.ANM(atoms, FF[key](atoms))
    lam, modes = anm.eigen(k=26)
    want = ref[f"{key}/eigval"]
    assert np.allclose(lam[6:26], want[6:26], rtol=EIG_RTOL, atol=0)
    assert subspace_sin(modes[6:26], ref[f"{key}/modes_6_26"]) < ANGLE_TOL
    msf = anm.mean_square_fluctuation(mode_subset=np.arange(6, 26))
    assert np.allclose(msf, ref[f"{key}/msf_6_26"], rtol=PROD_RTOL, atol=0)


def test_c4_cloud400_allpairs_100_modes():
    """Config C4 scaled down: ParameterFree all-pairs Hessian, lowest 100 non-trivial modes + MSF."""
    ref = golden("ref_c4_cloud400.npz")
    anm = sc.ANM(ref["coord"], sc.ParameterFreeForceField())
    assert rel_err(np.diagonal(anm.hessian), ref["hessian_diag"]) <= HESS_RTOL
    lam, modes = anm.eigen(k=106)
    assert np.allclose(lam[6:106], ref["eigval"][6:106], rtol=EIG_RTOL, atol=0)
    assert subspace_sin(modes[6:106], ref["modes_6_106"]) < ANGLE_TOL
    msf = anm.mean_square_fluctuation(mode_subset=np.arange(6, 106))
    assert np.allclose(msf, ref["msf_6_106"], rtol=PROD_RTOL, atol=0)
    # the dense row-slab assembly (C4 partitioning, scb_assemble_dense_allpairs) agrees with the BSR path
    import ctypes, torch
    from springcraft_b200 import _lib
    model = anm._model_device()
    n = model.n
    slab = torch.empty((3 * (300 - 100), 3 * n), dtype=torch.float64, device="cuda")
    _lib.check(model.handle.scb_assemble_dense_allpairs(3, _lib.ptr(model.xyz), n, ctypes.byref(model.desc), None,
                                                        100, 300, _lib.ptr(slab), _lib.stream_ptr()))
    assert rel_err(slab.cpu().numpy(), anm.hessian[300:900]) <= HESS_RTOL


@pytest.mark.parametrize("k", [56, 106])
def test_c4_dense_row_operator(k):
    """Dense slab operator + Python-driven subspace iteration (the multi-GPU C4 path, here on one rank):
    block width 64 (k=56) and 128 (k=106)."""
    import torch
    from springcraft_b200.dense_solver import DenseRowOperator
    ref = golden("ref_c4_cloud400.npz")
    op = DenseRowOperator(ref["coord"], sc.ParameterFreeForceField(), 3)
    H = sc.ANM(ref["coord"], sc.ParameterFreeForceField()).hessian
    assert rel_err(op.slab.cpu().numpy(), H) <= HESS_RTOL
    X = torch.from_numpy(np.random.default_rng(0).normal(size=(1200, 64))).cuda()
    assert rel_err(op.apply(X).cpu().numpy(), H @ X.cpu().numpy()) <= 1e-13
    lam, modes, iters = sc.allpairs_lowest_modes(ref["coord"], sc.ParameterFreeForceField(), k)
    assert np.allclose(lam[6:k], ref["eigval"][6:k], rtol=EIG_RTOL, atol=0)
    assert subspace_sin(modes[6:k], ref["modes_6_106"][: k - 6]) < ANGLE_TOL
    assert np.abs(H @ modes[:6].T).max() <= 1e-10 * np.abs(H).max()


@pytest.mark.parametrize("splits", [1, 2, 4, 8])
def test_dense_slab_split_k(splits, monkeypatch):
    """Split-K variant of the slab product (used when a rank's slab has too few tiles to fill the GPU): same
    result as torch's fp64 matmul, fused Chebyshev epilogue included, ragged row/column counts, and
    bit-reproducible from call to call (deterministic reduction order)."""
    import torch
    from springcraft_b200 import _lib
    monkeypatch.setenv("SCB_SLAB_SPLITS", str(splits))
    h = _lib.require_device()
    rows, N, b, row0 = 1000, 5003, 128, 37
    g = torch.Generator("cuda").manual_seed(3)
    slab = torch.randn((rows, N), dtype=torch.float64, device="cuda", generator=g)
    X = torch.randn((N, b), dtype=torch.float64, device="cuda", generator=g)
    W = torch.randn((N, b), dtype=torch.float64, device="cuda", generator=g)
    ws = torch.zeros(h.scb_dense_slab_workspace_bytes(N, row0, row0 + rows, b), dtype=torch.uint8, device="cuda")
    outs = []
    for _ in range(2):
        Y = torch.empty((rows, b), dtype=torch.float64, device="cuda")
        _lib.check(h.scb_dense_slab_apply(N, row0, row0 + rows, _lib.ptr(slab), _lib.ptr(X), _lib.ptr(W), _lib.ptr(Y),
                                          b, 1, 0.7, 0.3, 0.2, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        outs.append(Y)
    ref = 0.7 * (slab @ X - 0.3 * X[row0:row0 + rows]) - 0.2 * W[row0:row0 + rows]
    assert float((outs[0] - ref).abs().max() / ref.abs().max()) <= 1e-13
    assert torch.equal(outs[0], outs[1])


# --------------------------------------------------------------------------- K4
@pytest.mark.parametrize("key", ["invariant13", "hinsen", "e_anm", "sd_enm", "pfree"])
def test_1l2y_nma_products(structures, key):
    ref = golden("ref_1l2y.npz")
    atoms = atoms_of(structures, "1l2y")
    anm = sc.ANM(atoms, FF[key](atoms))
    tol = dict(rtol=1e-8, atol=1e-10 * np.abs(ref[f"{key}/anm_cov"]).max())
    assert np.allclose(anm.frequencies()[6:], ref[f"{key}/anm_freq"][6:], rtol=EIG_RTOL)
    assert np.allclose(anm.mean_square_fluctuation(), ref[f"{key}/anm_msf"], rtol=PROD_RTOL)
    assert np.allclose(anm.mean_square_fluctuation(mode_subset=np.arange(6, 26)), ref[f"{key}/anm_msf_sub"],
                       rtol=PROD_RTOL)
    assert np.allclose(anm.mean_square_fluctuation(tem=300, tem_factors=orc.K_B * orc.N_A),
                       ref[f"{key}/anm_msf_tem"], rtol=PROD_RTOL)
    assert np.allclose(anm.bfactor(), ref[f"{key}/anm_bfactor"], rtol=PROD_RTOL)
    assert np.allclose(anm.covariance, ref[f"{key}/anm_cov"], **tol)
    assert np.allclose(anm.dcc(), ref[f"{key}/anm_dcc"], atol=1e-8)
    assert np.allclose(anm.dcc(norm=False), ref[f"{key}/anm_dcc_abs"], **tol)
    assert np.allclose(anm.dcc(mode_subset=np.arange(6, 36)), ref[f"{key}/anm_dcc_sub"], atol=1e-8)
    assert np.allclose(anm.dcc(mode_subset=np.arange(6, 36), norm=False, tem=300), ref[f"{key}/anm_dcc_sub_tem"],
                       rtol=1e-8, atol=1e-10 * np.abs(ref[f"{key}/anm_dcc_sub_tem"]).max())
    assert np.allclose(anm.linear_response(ref["force_unit"]), ref[f"{key}/anm_lr_unit"], **tol)
    assert np.allclose(anm.linear_response(ref["force_rand"].flatten()), ref[f"{key}/anm_lr_rand"],
                       rtol=1e-8, atol=1e-9 * np.abs(ref[f"{key}/anm_lr_rand"]).max())
    prs, eff, sens = anm.prs_effector_sensor()
    assert np.allclose(prs, ref[f"{key}/anm_prs"], rtol=1e-6)
    assert np.allclose(eff, ref[f"{key}/anm_eff"], rtol=1e-6)
    assert np.allclose(sens, ref[f"{key}/anm_sens"], rtol=1e-6)
    nm = anm.normal_mode(6, 5.0, 8)
    want = ref[f"{key}/anm_normal_mode"]
    sign = np.sign(np.sum(nm[1] * want[1]))          # eigenvector sign is arbitrary
    assert nm.shape == (8, 20, 3) and np.allclose(sign * nm, want, atol=1e-7)
    nm = anm.normal_mode(7, 2.0, 6, movement="triangle")
    want = ref[f"{key}/anm_normal_mode_tri"]
    sign = np.sign(np.sum(nm[0] * want[0]))
    assert np.allclose(sign * nm, want, atol=1e-7)
    with pytest.raises(ValueError):
        anm.normal_mode(6, 1.0, 4, movement="square")
    with pytest.raises(ValueError):
        anm.mean_square_fluctuation(mode_subset=np.array([5, 6, 7]))
    with pytest.raises(ValueError):
        anm.linear_response(np.zeros((19, 3)))
    gnm = sc.GNM(atoms, FF[key](atoms))
    assert np.allclose(gnm.mean_square_fluctuation(), ref[f"{key}/gnm_msf"], rtol=PROD_RTOL)
    assert np.allclose(gnm.bfactor(tem=300), ref[f"{key}/gnm_bfactor"], rtol=PROD_RTOL)
    assert np.allclose(gnm.covariance, ref[f"{key}/gnm_cov"], rtol=1e-8,
                       atol=1e-10 * np.abs(ref[f"{key}/gnm_cov"]).max())
    assert np.allclose(gnm.dcc(), ref[f"{key}/gnm_dcc"], atol=1e-8)
    # nodes that do not move in the selected modes have d_ii ~ 1e-32: their normalised rows are
    # rounding noise in the reference itself (0/0), so compare the well-defined nodes only
    dsub = gnm.dcc(mode_subset=np.arange(1, 17), norm=False)
    live = np.diagonal(dsub) > 1e-10 * np.diagonal(dsub).max()
    got = gnm.dcc(mode_subset=np.arange(1, 17))
    assert np.allclose(got[np.ix_(live, live)], ref[f"{key}/gnm_dcc_sub"][np.ix_(live, live)], atol=1e-8)
    mw = sc.ANM(atoms, FF[key](atoms), masses=ref["masses"])
    assert np.allclose(mw.eigen()[0][6:], ref[f"{key}/mw_eigval"][6:], rtol=EIG_RTOL)
    assert np.allclose(mw.dcc(), ref[f"{key}/mw_dcc"], atol=1e-8)


def test_setters_and_roundtrip(structures):
    """anm.py:114-148: covariance <-> hessian through the setters."""
    ref = golden("ref_1l2y.npz")
    atoms = atoms_of(structures, "1l2y")
    anm = sc.ANM(atoms, sc.InvariantForceField(13.0))
    H = anm.hessian
    assert anm.hessian is H
    C = anm.covariance
    assert np.allclose(H, H @ C @ H, atol=1e-9)
    assert np.allclose(C, C @ H @ C, atol=1e-12)
    other = sc.ANM(atoms, sc.InvariantForceField(13.0))
    other.covariance = C.copy()
    assert np.allclose(other.hessian, ref["invariant13/hessian"], atol=1e-8)
    with pytest.raises(IndexError):
        other.hessian = np.zeros((3, 3))
    g = sc.GNM(atoms, sc.InvariantForceField(7.0))
    with pytest.raises(ValueError):
        g.kirchhoff = np.zeros((3, 3))


def test_matrix_edits_and_assigned_covariance(structures):
    """anm.py:53-57: `hessian` is returned by reference, so the reference sees in-place edits at the next call;
    anm.py:138-148 + nma.py:324-336, 473: an assigned covariance is used as is by dcc() and linear_response()."""
    atoms = atoms_of(structures, "1l2y")
    anm = sc.ANM(atoms, sc.InvariantForceField(13.0))
    lam0, _ = anm.eigen()
    msf0 = anm.mean_square_fluctuation()
    H = anm.hessian
    H *= 2.0                                   # in place: no setter involved
    lam1, _ = anm.eigen()
    assert np.allclose(lam1[6:], 2.0 * lam0[6:], rtol=1e-10)
    assert np.allclose(anm.mean_square_fluctuation(), 0.5 * msf0, rtol=1e-8)
    # a covariance assigned by the caller
    n = len(atoms.coord)
    C = 3.0 * sc.ANM(atoms, sc.InvariantForceField(13.0)).covariance
    other = sc.ANM(atoms, sc.InvariantForceField(13.0))
    other.covariance = C
    tr = C.reshape(n, 3, n, 3).trace(axis1=1, axis2=3)
    d = np.diagonal(tr)
    assert np.allclose(other.dcc(norm=False), tr, rtol=1e-12)
    assert np.allclose(other.dcc(), tr / np.sqrt(np.outer(d, d)), rtol=1e-12)
    assert np.allclose(other.dcc(tem=300), tr / np.sqrt(np.outer(d, d)) * 300 * sc.anm.K_B, rtol=1e-12)
    f = np.random.default_rng(3).normal(size=(n, 3))
    assert np.allclose(other.linear_response(f), (C @ f.ravel()).reshape(n, 3), rtol=1e-11, atol=1e-13)
    g = sc.GNM(atoms, sc.InvariantForceField(7.0))
    Cg = 2.0 * sc.GNM(atoms, sc.InvariantForceField(7.0)).covariance
    g.covariance = Cg
    assert np.allclose(g.dcc(norm=False), Cg, rtol=1e-12)


def test_c5_dcc_linear_response():
    """Config C5 scaled down: DMMA W W^T contraction from a mode subset."""
    ref = golden("ref_c5_chain400.npz")
    coord = ref["coord"]
    anm = sc.ANM(coord, sc.InvariantForceField(13.0))
    gnm = sc.GNM(coord, sc.InvariantForceField(10.0))
    assert np.allclose(anm.dcc(mode_subset=np.arange(6, 56)), ref["anm_dcc_sub"], atol=1e-8)
    assert np.allclose(anm.dcc(mode_subset=np.arange(6, 56), norm=False), ref["anm_dcc_sub_abs"],
                       rtol=1e-8, atol=1e-10 * np.abs(ref["anm_dcc_sub_abs"]).max())
    assert np.allclose(gnm.dcc(mode_subset=np.arange(1, 51)), ref["gnm_dcc_sub"], atol=1e-8)
    # full covariance products (nma.py:324-336, 473) and the default mode set of the MSF (nma.py:145-151)
    assert np.allclose(anm.dcc(), ref["anm_dcc_all"], atol=1e-8)
    assert np.allclose(gnm.dcc(norm=False), ref["gnm_dcc_all_abs"], rtol=1e-8,
                       atol=1e-10 * np.abs(ref["gnm_dcc_all_abs"]).max())
    assert np.allclose(anm.mean_square_fluctuation(), ref["anm_msf"], rtol=PROD_RTOL)
    assert np.allclose(gnm.mean_square_fluctuation(), ref["gnm_msf"], rtol=PROD_RTOL)
    scale = np.abs(ref["anm_lr"]).max()
    assert np.allclose(anm.linear_response(ref["force"]), ref["anm_lr"], rtol=1e-8, atol=1e-10 * scale)
    assert np.allclose(anm.linear_response(ref["force"].flatten()), ref["anm_lr"], rtol=1e-8, atol=1e-10 * scale)
    unit = np.zeros((len(coord), 3))
    unit[42, 0] = 1.0
    assert np.allclose(anm.linear_response(unit), ref["anm_lr_unit42"], rtol=1e-8,
                       atol=1e-10 * np.abs(ref["anm_lr_unit42"]).max())
    # "from m modes": low-rank response V_S^T L_S^-1 V_S f through the lowest-k solver (fresh object: no cached spectrum)
    low = sc.ANM(coord, sc.InvariantForceField(13.0))
    got = low.linear_response(ref["force"], mode_subset=np.arange(6, 56))
    assert low._spectrum_cache.get("full") is None, "the subset must not trigger a full decomposition"
    assert np.allclose(got, ref["anm_lr_sub"], rtol=1e-7, atol=1e-9 * np.abs(ref["anm_lr_sub"]).max())
    with pytest.raises(ValueError):
        low.linear_response(ref["force"], mode_subset=np.arange(5, 20))


def test_tabulated_interaction_matrix_edited_in_place(structures):
    """forcefield.py:429-434: interaction_matrix is returned by reference; edits by the caller are what
    force_constant() uses (device side: SCB_FF_TABULATED_DENSE reads the explicit (n,n,k) float32 table)."""
    ref = golden("ref_dense_table.npz")
    atoms = atoms_of(structures, "1l2y")
    for key in ("e_anm", "sd_enm"):
        ff = FF[key](atoms)
        M = ff.interaction_matrix
        assert M.dtype == np.float32 and ff.interaction_matrix is M
        M[...] = ref[f"{key}/table"]
        H, pairs = sc.compute_hessian(atoms.coord, ff)
        assert np.array_equal(pairs, ref[f"{key}/pairs"])
        assert np.array_equal(H, ref[f"{key}/hessian"])
        K, _ = sc.compute_kirchhoff(atoms.coord, ff)
        assert np.array_equal(K, ref[f"{key}/kirchhoff"])
        # the pristine force field gives a different matrix: the edit really went through the dense table
        H0, _ = sc.compute_hessian(atoms.coord, FF[key](atoms))
        assert not np.array_equal(H0, H)


def test_dense_tf32_filter_kernel():
    """tcgen05 TF32 slab product (dense_tf32.cu) against FP64 matmul: single product ~1e-3 (TF32), 3-term split
    product at FP32 level; fused recurrence epilogue; ragged sizes (rows and K not multiples of the tile)."""
    import torch
    from springcraft_b200 import _lib
    from springcraft_b200.dense_solver import DenseRowOperator
    from synthetic_inputs import synthetic_cloud
    h = _lib.require_device()
    n, b = 250, 128
    op = DenseRowOperator(synthetic_cloud(n, seed=2), sc.ParameterFreeForceField(), 3)
    N = op.N
    ld = int(h.scb_tf32_ld(N))
    gen = torch.Generator("cuda").manual_seed(3)
    Z = torch.zeros((b, ld), dtype=torch.float32, device="cuda")
    Z[:, :N] = torch.randn((b, N), generator=gen, device="cuda")
    ref = (op.slab @ Z[:, :N].double().T).T
    scale = ref.abs().max().item()
    s32, _ = op.slab32(False)
    out = torch.full((b, ld), 7.0, dtype=torch.float32, device="cuda")
    _lib.check(h.scb_dense_slab_tf32_apply(N, 0, N, _lib.ptr(s32), None, b, _lib.ptr(Z), None, None, _lib.ptr(out),
                                           None, None, 0.0, 0, _lib.stream_ptr()))
    assert (out[:, :N].double() - ref).abs().max().item() < 5e-3 * scale
    assert bool((out[:, N:] == 7.0).all())                       # padding columns are never written
    hi, lo = op.slab32(True)
    assert torch.equal(hi.double()[:, :N] + lo.double()[:, :N], (hi.double() + lo.double())[:, :N])
    assert (hi.double()[:, :N] + lo.double()[:, :N] - op.slab).abs().max().item() < 1e-9 * op.slab.abs().max().item()
    Z2 = torch.zeros((2 * b, ld), dtype=torch.float32, device="cuda")
    zh = (Z.view(torch.int32) & -8192).view(torch.float32)
    Z2[:b], Z2[b:] = zh, Z - zh
    Zp = torch.randn((2 * b, ld), generator=gen, device="cuda")
    Rh = torch.randn((b, ld), generator=gen, device="cuda")
    cA = torch.rand(b, generator=gen, device="cuda") * 0.1
    cB = torch.rand(b, generator=gen, device="cuda")
    zp = Zp[:b, :N].double() + Zp[b:, :N].double()
    want = cA[:, None].double() * (ref - 0.7 * Z[:, :N].double() + Rh[:, :N].double()) - cB[:, None].double() * zp
    _lib.check(h.scb_dense_slab_tf32_apply(N, 0, N, _lib.ptr(hi), _lib.ptr(lo), b, _lib.ptr(Z2), _lib.ptr(Zp),
                                           _lib.ptr(Rh), _lib.ptr(Zp), _lib.ptr(cA), _lib.ptr(cB), 0.7, 1,
                                           _lib.stream_ptr()))
    got = Zp[:b, :N].double() + Zp[b:, :N].double()
    assert (got - want).abs().max().item() < 5e-5 * want.abs().max().item()
    op.close()


@pytest.mark.parametrize("filt", ["tf32", "tf32x1", "fp64"])
def test_c4_cloud400_filters(filt):
    """C4 scaled down (reference golden): all-pairs pfENM, lowest 50 non-trivial modes with every filter variant."""
    from springcraft_b200.dense_solver import DenseRowOperator, eig_lowest_dense
    g = golden("ref_c4_cloud400.npz")
    op = DenseRowOperator(g["coord"], sc.ParameterFreeForceField(), 3)
    theta, X, res, it = eig_lowest_dense(op, 50, Z=op.rigid_basis(), b=128, filter=filt)
    assert np.allclose(theta[:50].cpu().numpy(), g["eigval"][6:56], rtol=1e-8, atol=0)
    Q, _ = np.linalg.qr(X[:, :50].cpu().numpy())
    Qa, _ = np.linalg.qr(g["modes_6_106"][:50].T)
    assert np.linalg.norm(Q - Qa @ (Qa.T @ Q), 2) < 1e-6
    op.close()
