"""CPU oracle for the springcraft elastic-network hot path (TEST INFRASTRUCTURE).

Independent NumPy restatement of the reference algorithm.  Every function
cites the reference lines (relative to /root/reference/src/springcraft) whose
arithmetic -- including evaluation order -- it reproduces.  Parity is pinned by
tests/test_oracle_golden.py (see oracle/__init__.py).

Conventions
-----------
n      number of nodes (CA atoms)
P      number of ORDERED contact pairs (both (i,j) and (j,i))
rowptr CSR row pointer (n+1,), col CSR column indices (P,), sorted inside a row
       => walking the CSR row by row yields the reference's ``np.where(adj)``
       order (interaction.py:177-178).
"""

from dataclasses import dataclass, field
from os.path import dirname, join, realpath

import numpy as np

K_B = 1.380649e-23  # nma.py:24
N_A = 6.02214076e23  # nma.py:25

DATA_DIR = join(dirname(dirname(realpath(__file__))), "springcraft_b200", "data")
# biotite ProteinSequence alphabet order, first 20 symbols (forcefield.py:28-34)
AA_ORDER = ["ALA", "CYS", "ASP", "GLU", "PHE", "GLY", "HIS", "ILE", "LYS", "LEU",
            "MET", "ASN", "PRO", "GLN", "ARG", "SER", "THR", "VAL", "TRP", "TYR"]
AA_INDEX = {aa: k for k, aa in enumerate(AA_ORDER)}


# --------------------------------------------------------------------------
# force-field description (the oracle's own POD; mirrors include/scb200.h)
# --------------------------------------------------------------------------
@dataclass
class FFSpec:
    kind: str                      # "invariant" | "hinsen" | "pfree" | "tabulated"
    cutoff: float | None = None    # None => all pairs (interaction.py:151-153)
    # tabulated only (forcefield.py:437-513)
    bonded: np.ndarray | None = None   # (20,20,k) float32
    intra: np.ndarray | None = None    # (20,20,k) float32
    inter: np.ndarray | None = None    # (20,20,k) float32
    edges: np.ndarray | None = None    # (k,) float64 or None
    res_type: np.ndarray | None = None     # (n,) int  index into AA_ORDER
    chain: np.ndarray | None = None        # (n,) int  chain label id
    bonded_next: np.ndarray | None = None  # (n,) bool: atom a and a+1 are peptide bonded
    # patches (forcefield.py:117-261)
    shutdown: np.ndarray | None = None     # (s,) int
    pair_off: np.ndarray | None = None     # (q,2) int
    pair_on: np.ndarray | None = None      # (q,2) int
    pair_on_fc: np.ndarray | None = None   # (q,) float64
    patched: bool = False
    extra: dict = field(default_factory=dict)


def load_table(name):
    """forcefield.py:940-950 -- plain comma separated float table."""
    return np.loadtxt(join(DATA_DIR, name), delimiter=",")


def as_table(value, n_bins):
    """forcefield.py:879-923 -- broadcast scalar/1D/2D/3D input to (20,20,k) f32."""
    if np.isnan(value).any():
        raise IndexError("Array contains NaN elements")
    if np.isscalar(value):
        return np.full((20, 20, n_bins), value, dtype=np.float32)
    arr = np.asarray(value, dtype=np.float32)
    if arr.ndim == 1:
        if len(arr) != n_bins:
            raise IndexError("bin count mismatch")
        return np.broadcast_to(arr, (20, 20, n_bins)).copy()
    if arr.ndim == 2:
        return np.repeat(arr[:, :, None], n_bins, axis=2)
    if arr.ndim == 3:
        if arr.shape[-1] != n_bins:
            raise IndexError("bin count mismatch")
        return arr
    raise IndexError("too many dimensions")


def tabulated_spec(res_name, chain_id, res_id, bonded, intra, inter, cutoff):
    """forcefield.py:437-513 -- per-atom attributes instead of the (n,n,k) table."""
    res_name = np.asarray(res_name)
    chain_id = np.asarray(chain_id)
    res_id = np.asarray(res_id)
    if cutoff is None:
        edges, k = None, 1
    elif np.isscalar(cutoff):
        edges, k = np.array([float(cutoff)]), 1
    else:
        edges = np.asarray(cutoff, dtype=np.float64)
        k = len(edges)
    _, chain_num = np.unique(chain_id, return_inverse=True)
    n = len(res_name)
    nxt = np.zeros(n, dtype=bool)
    if n > 1:
        # forcefield.py:471-473
        nxt[:-1] = (np.diff(res_id) == 1) & (chain_id[:-1] == chain_id[1:])
    return FFSpec(
        kind="tabulated",
        cutoff=None if edges is None else float(edges[-1]),
        bonded=as_table(bonded, k), intra=as_table(intra, k), inter=as_table(inter, k),
        edges=edges,
        res_type=np.array([AA_INDEX[r] for r in res_name], dtype=np.int64),
        chain=chain_num.astype(np.int64), bonded_next=nxt,
    )


def preset_spec(name, res_name, chain_id, res_id, nonbonded_mean=False):
    """forcefield.py:547-876 -- the seven preset constructors."""
    args = (res_name, chain_id, res_id)
    if name == "s_enm_10":
        fc = load_table("s_enm_10.csv")
        return tabulated_spec(*args, 10.0, fc, fc, 10.0)
    if name == "s_enm_13":
        fc = load_table("s_enm_13.csv")
        return tabulated_spec(*args, 10.0, fc, fc, 13.0)
    if name == "d_enm":
        fc = load_table("d_enm.csv")
        return tabulated_spec(*args, 46.83, fc, fc, load_table("d_enm_edges.csv"))
    if name == "sd_enm":
        # forcefield.py:693-699 (file is 26 stacked 20x20 blocks)
        fc = load_table("sd_enm.csv").reshape(-1, 20, 20).T
        fc = fc * 0.0083144621 * 300 * 10
        bonded = 43.52 * 0.0083144621 * 300 * 10
        return tabulated_spec(*args, bonded, fc, fc, load_table("d_enm_edges.csv"))
    if name in ("e_anm", "e_anm_mj", "e_anm_ke"):
        intra = load_table("keskin.csv" if name == "e_anm_ke" else "miyazawa.csv")
        inter = load_table("miyazawa.csv" if name == "e_anm_mj" else "keskin.csv")
        if nonbonded_mean:
            intra = np.average(intra) * np.ones((20, 20))
            inter = np.average(inter) * np.ones((20, 20))
        return tabulated_spec(*args, 82.0, intra, inter, 13.0)
    raise KeyError(name)


# --------------------------------------------------------------------------
# a1-a3: contact search, patches, pair list, geometry
# --------------------------------------------------------------------------
def squared_distance_rows(coord, i):
    """sq[i, :] with the reference's rounding: (dx*dx + dy*dy) + dz*dz,
    every product and sum rounded separately (interaction.py:162-166, 183-184)."""
    d = coord - coord[i]
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]


def contact_csr(coord, spec):
    """Adjacency in CSR form.  interaction.py:149-174, 193-213.

    Criterion: fp64 ``sq <= cutoff**2`` (the reference's brute-force branch,
    interaction.py:165-166; pinned bit-for-bit by the ProDy Kirchhoff goldens).
    Patches applied in the order shutdown, pair_off, pair_on.
    """
    coord = np.asarray(coord, dtype=np.float64)
    if coord.ndim != 2 or coord.shape[1] != 3:
        raise ValueError(f"Expected coordinates with shape (n,3), got {coord.shape}")
    n = len(coord)
    off = set()
    on = {}
    dead = np.zeros(n, dtype=bool)
    if spec.shutdown is not None:
        dead[np.asarray(spec.shutdown, dtype=np.int64)] = True
    if spec.pair_off is not None:
        for a, b in np.asarray(spec.pair_off, dtype=np.int64).reshape(-1, 2):
            off.add((int(a), int(b)))
            off.add((int(b), int(a)))
    if spec.pair_on is not None:
        for a, b in np.asarray(spec.pair_on, dtype=np.int64).reshape(-1, 2):
            if a == b:
                raise ValueError("Cannot turn on interaction of an atom with itself")
            on.setdefault(int(a), set()).add(int(b))
            on.setdefault(int(b), set()).add(int(a))
    cut2 = None if spec.cutoff is None else spec.cutoff ** 2
    rowptr = np.zeros(n + 1, dtype=np.int64)
    cols = []
    for i in range(n):
        if cut2 is None:
            row = np.ones(n, dtype=bool)
        else:
            row = squared_distance_rows(coord, i) <= cut2
        row[i] = False
        if dead[i]:
            row[:] = False
        row[dead] = False
        for (a, b) in off:
            if a == i:
                row[b] = False
        for b in on.get(i, ()):
            row[b] = True
        c = np.nonzero(row)[0]
        cols.append(c)
        rowptr[i + 1] = rowptr[i] + len(c)
    col = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    return rowptr, col.astype(np.int64)


def pairs_from_csr(rowptr, col):
    """interaction.py:177-178 -- (P,2) int64, lexicographic."""
    n = len(rowptr) - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    return np.stack([rows, col.astype(np.int64)], axis=1)


def pair_geometry(coord, pairs):
    """interaction.py:182-184 -- disp = x_j - x_i ; sq = (d0^2 + d1^2) + d2^2."""
    coord = np.asarray(coord, dtype=np.float64)
    disp = coord[pairs[:, 1]] - coord[pairs[:, 0]]
    sq = (disp[:, 0] * disp[:, 0] + disp[:, 1] * disp[:, 1]) + disp[:, 2] * disp[:, 2]
    return disp, sq


# --------------------------------------------------------------------------
# a6-a11: force constants
# --------------------------------------------------------------------------
def _fc_base(spec, pairs, sq):
    i, j = pairs[:, 0], pairs[:, 1]
    if spec.kind == "invariant":          # forcefield.py:283-284
        return np.ones(len(sq))
    if spec.kind == "hinsen":             # forcefield.py:318-326
        r = np.maximum(np.sqrt(sq), 2.9)
        return np.where(r < 4.0, r * 860.0 - 2390.0, r ** (-6) * 1280000.0)
    if spec.kind == "pfree":              # forcefield.py:361-362
        return 1 / sq
    if spec.kind == "tabulated":          # forcefield.py:497-533
        dense = spec.extra.get("dense_table")
        if dense is not None:
            # interaction_matrix edited by the caller (forcefield.py:429-434): the explicit (n,n,k) float32 table
            if dense.shape[-1] == 1:      # forcefield.py:516-518
                return dense[i, j, 0]
            b = np.searchsorted(spec.edges ** 2, sq, side="left")
            if (b >= len(spec.edges)).any():
                raise ValueError("Atom interactions above cutoff distance are "
                                 "not allowed in TabulatedForceField")
            return dense[i, j, b]
        if spec.edges is None or len(spec.edges) == 1:
            b = np.zeros(len(sq), dtype=np.int64)
        else:
            # left rule: number of squared edges strictly below sq
            b = np.searchsorted(spec.edges ** 2, sq, side="left")
            if (b >= len(spec.edges)).any():
                raise ValueError("Atom interactions above cutoff distance are "
                                 "not allowed in TabulatedForceField")
        ti, tj = spec.res_type[i], spec.res_type[j]
        same = spec.chain[i] == spec.chain[j]
        out = np.where(same, spec.intra[ti, tj, b], spec.inter[ti, tj, b]).astype(np.float32)
        lo = np.minimum(i, j)
        bonded = (np.abs(i - j) == 1) & spec.bonded_next[lo]
        hi = np.minimum(lo + 1, len(spec.res_type) - 1)
        # forcefield.py:504-509: both orientations take bonded[type_lo, type_hi]
        out = np.where(bonded, spec.bonded[spec.res_type[lo], spec.res_type[hi], b], out)
        out = np.where(i == j, np.float32(0), out)   # forcefield.py:512-513
        return out.astype(np.float32)
    raise KeyError(spec.kind)


def force_constants(spec, pairs, sq):
    """ForceField.force_constant for the built-in kinds (+ Patched wrapper,
    forcefield.py:183-226)."""
    if not spec.patched:
        return _fc_base(spec, pairs, sq)
    if spec.cutoff is None:
        fc = np.asarray(_fc_base(spec, pairs, sq), dtype=np.float64)
    else:
        fc = np.zeros(len(sq))
        m = sq <= spec.cutoff ** 2
        fc[m] = _fc_base(spec, pairs[m], sq[m])
    if spec.pair_on is not None:
        lut = {}
        for (a, b), v in zip(np.asarray(spec.pair_on).reshape(-1, 2), spec.pair_on_fc):
            lut[(int(a), int(b))] = float(v)
            lut[(int(b), int(a))] = float(v)
        for p, (a, b) in enumerate(pairs):
            v = lut.get((int(a), int(b)))
            if v is not None and v != -1:      # sentinel, forcefield.py:214-223
                fc[p] = v
    return fc


# --------------------------------------------------------------------------
# a4-a5: dense assembly
# --------------------------------------------------------------------------
def kirchhoff_dense(n, pairs, fc):
    """interaction.py:47-52.  Diagonal = -(column sum), accumulated in
    ascending row order (np.sum(axis=0) is sequential)."""
    K = np.zeros((n, n))
    K[pairs[:, 0], pairs[:, 1]] = -np.asarray(fc, dtype=np.float64)
    diag = np.zeros(n)
    np.add.at(diag, pairs[:, 1], K[pairs[:, 0], pairs[:, 1]])
    K[np.arange(n), np.arange(n)] = -diag
    return K


def hessian_blocks(disp, sq, fc):
    """interaction.py:96-101: ((-fc / sq) * d_a) * d_b."""
    t = -np.asarray(fc)[:, None, None] / sq[:, None, None]
    return (t * disp[:, :, None]) * disp[:, None, :]


def hessian_dense(n, pairs, disp, sq, fc):
    """interaction.py:93-109."""
    blocks = hessian_blocks(disp, sq, fc)
    H4 = np.zeros((n, 3, n, 3))
    H4[pairs[:, 0], :, pairs[:, 1], :] = blocks
    diag = np.zeros((n, 3, 3))
    np.add.at(diag, pairs[:, 1], blocks)       # ascending i for each j
    idx = np.arange(n)
    H4[idx, :, idx, :] = -diag
    return H4.reshape(3 * n, 3 * n)


def mass_weight(M, masses, dim):
    """anm.py:89-96,112-113 / gnm.py:85-89,105-106: M *= outer(w, w)."""
    w = 1 / np.sqrt(np.asarray(masses, dtype=float))
    if dim == 3:
        w = np.repeat(w, 3)
    return M * np.outer(w, w)


def compute_kirchhoff(coord, spec, masses=None):
    coord = np.asarray(coord).astype(np.float64, copy=False)
    rowptr, col = contact_csr(coord, spec)
    pairs = pairs_from_csr(rowptr, col)
    _, sq = pair_geometry(coord, pairs)
    K = kirchhoff_dense(len(coord), pairs, force_constants(spec, pairs, sq))
    if masses is not None:
        K = mass_weight(K, masses, 1)
    return K, pairs


def compute_hessian(coord, spec, masses=None):
    coord = np.asarray(coord).astype(np.float64, copy=False)
    rowptr, col = contact_csr(coord, spec)
    pairs = pairs_from_csr(rowptr, col)
    disp, sq = pair_geometry(coord, pairs)
    H = hessian_dense(len(coord), pairs, disp, sq, force_constants(spec, pairs, sq))
    if masses is not None:
        H = mass_weight(H, masses, 3)
    return H, pairs


# --------------------------------------------------------------------------
# a13-a19: NMA
# --------------------------------------------------------------------------
def eigen(M):
    """nma.py:61-63: full eigh (lower triangle, ascending); modes as ROWS."""
    lam, vec = np.linalg.eigh(M)
    return lam, vec.T


def frequencies(lam, ntriv):
    """nma.py:97-104."""
    lam = np.array(lam, dtype=float)
    lam[:ntriv] = np.abs(lam[:ntriv])
    with np.errstate(invalid="ignore"):
        return 1 / (2 * np.pi) * np.sqrt(lam)


def mean_square_fluctuation(lam, modes, dim, mode_subset=None, tem=None, tem_factors=K_B):
    """nma.py:145-183.  ``modes`` rows are modes; dim = 3 (ANM) or 1 (GNM)."""
    ntriv = 6 if dim == 3 else 1
    sqv = np.square(modes)
    if dim == 3:
        sqv = (sqv[:, 0::3] + sqv[:, 1::3]) + sqv[:, 2::3]
    if mode_subset is None:
        mode_subset = np.arange(ntriv, len(lam))
    elif np.any(np.asarray(mode_subset) <= ntriv - 1):
        raise ValueError("Trivial modes are included in the current selection.")
    msf = np.sum(sqv[mode_subset] / lam[mode_subset][:, None], axis=0)
    return msf * (1 if tem is None else tem * tem_factors)


def bfactor(*args, **kw):
    """nma.py:227-229."""
    return 8 * np.pi ** 2 * mean_square_fluctuation(*args, **kw) / 3


def covariance(M):
    """anm.py:132-136 / gnm.py:125-131: pinv(hermitian=True, rcond=1e-6)
    == sum over |lam_k| > 1e-6*max|lam| of u_k u_k^T / lam_k."""
    lam, U = np.linalg.eigh(M)
    keep = np.abs(lam) > 1e-6 * np.max(np.abs(lam))
    return (U[:, keep] / lam[keep]) @ U[:, keep].T


def dcc(lam, modes, dim, cov=None, mode_subset=None, norm=True, tem=None, tem_factors=K_B):
    """nma.py:296-358."""
    ntriv = 6 if dim == 3 else 1
    n = modes.shape[1] // dim
    if mode_subset is None:
        if cov is None:
            raise ValueError("all-mode DCC needs the covariance")
        d = cov if dim == 1 else np.einsum("iaja->ij", cov.reshape(n, 3, n, 3))
    else:
        mode_subset = np.asarray(mode_subset)
        if np.any(mode_subset <= ntriv - 1):
            raise ValueError("Trivial modes are included in the current selection.")
        d = np.zeros((n, n))
        for k in mode_subset:          # rank-dim updates, nma.py:346-347
            u = modes[k].reshape(n, dim)
            d += u @ u.T / lam[k]
    if norm:
        dii = np.diagonal(d).reshape(1, n)
        d = d / np.sqrt(dii * dii.T)
    if tem is not None:
        d = d * tem * tem_factors
    return d


def linear_response(cov, force):
    """nma.py:457-473."""
    n = cov.shape[0] // 3
    return np.dot(cov, np.asarray(force, dtype=float).reshape(-1)).reshape(n, 3)


def linear_response_modes(lam, modes, force, mode_subset):
    """Low-rank linear response "from m modes" (BASELINE config C5): the pseudo-inverse of nma.py:473 restricted
    to the given non-trivial modes, V_S^T (L_S^-1 (V_S f))."""
    sub = np.asarray(mode_subset)
    n = modes.shape[1] // 3
    f = np.asarray(force, dtype=float).reshape(-1)
    return (modes[sub].T @ ((modes[sub] @ f) / lam[sub])).reshape(n, 3)


def prs(cov, norm=True):
    """nma.py:511-531."""
    n = cov.shape[0] // 3
    m = (cov ** 2).reshape(n, 3, n, 3).sum(axis=(1, 3))
    if norm:
        m = m / np.diagonal(m)[:, None]
    return m


def effector_sensor(prs_matrix):
    """nma.py:563-568."""
    n = len(prs_matrix)
    w = 1 - np.eye(n)
    return (np.average(prs_matrix, weights=w, axis=1),
            np.average(prs_matrix, weights=w, axis=0))


# --------------------------------------------------------------------------
# synthetic inputs live in the neutral module synthetic_inputs.py (the GPU arm of
# bench.py must not import the oracle); re-exported here for the tests and the golden generator
# --------------------------------------------------------------------------
import sys  # noqa: E402

_ROOT = dirname(dirname(realpath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from synthetic_inputs import (perturbed_conformation, synthetic_chain, synthetic_cloud,  # noqa: E402,F401
                              synthetic_sequence)
