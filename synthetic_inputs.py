"""Synthetic inputs of the benchmark configurations: seeded NumPy generators shared by
bench.py, the tests and the golden-vector generator.  No arithmetic of the hot path lives here."""
import numpy as np

# biotite ProteinSequence alphabet, first 20 symbols, as 3-letter codes (forcefield.py:28-34)
AA_ORDER = ["ALA", "CYS", "ASP", "GLU", "PHE", "GLY", "HIS", "ILE", "LYS", "LEU",
            "MET", "ASN", "PRO", "GLN", "ARG", "SER", "THR", "VAL", "TRP", "TYR"]


def synthetic_chain(n, seed=0, jitter=0.25):
    """Boustrophedon CA chain: 3.8 A steps along x, 6.0 A row/layer pitch."""
    nx = int(np.ceil((n * 36.0 / 3.8 ** 2) ** (1.0 / 3.0)))
    ny = int(np.ceil(np.sqrt(n / nx)))
    pts = np.zeros((n, 3))
    for k in range(n):
        iz, rem = divmod(k, nx * ny)
        iy, ix = divmod(rem, nx)
        if iy % 2 == 1:
            ix = nx - 1 - ix
        if iz % 2 == 1:
            iy = ny - 1 - iy
        pts[k] = (3.8 * ix, 6.0 * iy, 6.0 * iz)
    rng = np.random.default_rng(seed)
    return pts + rng.normal(0.0, jitter, size=(n, 3))


def synthetic_sequence(n, seed=0):
    rng = np.random.default_rng(seed)
    res_name = np.array(AA_ORDER)[rng.integers(0, 20, size=n)]
    return res_name, np.full(n, "A"), np.arange(1, n + 1)


def synthetic_cloud(n, seed=0, density=0.008, min_dist=3.0):
    """Uniform cloud at `density` atoms/A^3 with a minimum pair distance."""
    rng = np.random.default_rng(seed)
    side = (n / density) ** (1.0 / 3.0)
    pts = np.zeros((0, 3))
    while len(pts) < n:
        cand = rng.random((n, 3)) * side
        for c in cand:
            if len(pts) == 0 or np.min(((pts - c) ** 2).sum(1)) >= min_dist ** 2:
                pts = np.vstack([pts, c])
                if len(pts) == n:
                    break
    return pts


def perturbed_conformation(base, c, sigma=0.5):
    """C3 ensemble member c: base + N(0, sigma) with seed 1000+c."""
    rng = np.random.default_rng(1000 + c)
    return base + rng.normal(0.0, sigma, size=base.shape)
