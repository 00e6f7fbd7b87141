"""Device-resident elastic-network model: thin Python plumbing over the C ABI.

Holds the SoA coordinates, the CSR neighbour list and the BSR interaction matrix
of one structure or of an ensemble in HBM, and exposes the hot-path stages
(contacts, assembly, eigen, products) as methods that launch the CUDA kernels.
No arithmetic happens here.
"""

import ctypes as C

import numpy as np

from . import _lib

__all__ = ["DeviceModel", "eig_full_dense", "modes_msf", "modes_dcc", "modes_covariance",
           "modes_linear_response", "cov_dcc", "cov_matvec"]

# structures up to this size keep an explicit CSR for all-pairs force fields
# (bit-exact diagonal); larger ones are assembled straight into dense slabs
ALLPAIRS_CSR_MAX_N = 4096
# below this size the tiled all-against-all contact kernel is used even when
# the caller asks for the cell list (both give identical results)
CELL_LIST_MIN_N = 2048


def _torch():
    import torch
    return torch


def _patch_struct(ff, n, keep):
    """scb_patch from the force field's contact_* properties (interaction.py:169-174)."""
    torch = _torch()
    shutdown, off, on = ff.contact_shutdown, ff.contact_pair_off, ff.contact_pair_on
    if shutdown is None and off is None and on is None:
        return None
    p = _lib.Patch()
    if shutdown is not None:
        dead = np.zeros(n, dtype=np.uint8)
        sd = np.asarray(shutdown)
        # NumPy indexing semantics of the reference (interaction.py:201-203): a boolean mask selects atoms,
        # integers may be negative
        dead[sd if sd.dtype == bool else sd.astype(np.int64)] = 1
        t = _lib.to_device(dead, torch.uint8)
        keep.append(t)
        p.dead = t.data_ptr()
    if off is not None:
        a = _bounds(np.asarray(off, dtype=np.int64).reshape(-1, 2), n)
        t = _lib.to_device(a.astype(np.int32), torch.int32)
        keep.append(t)
        p.pair_off = t.data_ptr()
        p.n_pair_off = len(a)
    if on is not None:
        a = _bounds(np.asarray(on, dtype=np.int64).reshape(-1, 2), n)
        if (a[:, 0] == a[:, 1]).any():  # interaction.py:210-211
            raise ValueError("Cannot turn on interaction of an atom with itself")
        t = _lib.to_device(a.astype(np.int32), torch.int32)
        keep.append(t)
        p.pair_on = t.data_ptr()
        p.n_pair_on = len(a)
    return p


def _bounds(idx, n):
    """Bounds check with NumPy's semantics; negative indices are normalised (the kernels compare node numbers)."""
    if idx.size and (idx.max() >= n or idx.min() < -n):
        bad = int(idx.max()) if idx.max() >= n else int(idx.min())
        raise IndexError(f"index {bad} is out of bounds for axis 0 with size {n}")
    return np.where(idx < 0, idx + n, idx)


class DeviceModel:
    """Contacts + BSR interaction matrix of B structures of n nodes on the device."""

    def __init__(self, coord, force_field, D, masses=None, use_cell_list=True):
        torch = _torch()
        self.handle = _lib.require_device()
        coord = np.asarray(coord)
        if coord.ndim == 2:
            coord = coord[None]
        if coord.ndim != 3 or coord.shape[2] != 3:
            raise ValueError(f"Expected coordinates with shape (n,3), got {coord.shape[1:]}")
        self.B, self.n = int(coord.shape[0]), int(coord.shape[1])
        self.D = int(D)
        self.N = self.D * self.n
        if force_field.natoms is not None and self.n != force_field.natoms:
            raise ValueError(
                f"Got coordinates for {self.n} atoms, but forcefield was built for {force_field.natoms} atoms"
            )
        self.ff = force_field
        self._keep = []
        st = _lib.stream_ptr()
        # fp64 SoA coordinates (interaction.py:43,88: coord.astype(float64))
        aos = _lib.to_device(coord.astype(np.float64, copy=False), torch.float64)
        self.xyz = torch.empty((self.B, 3, self.n), dtype=torch.float64, device="cuda")
        _lib.check(self.handle.scb_coords_to_soa(_lib.ptr(aos), self.B, self.n, _lib.ptr(self.xyz), st))
        self.masses = None if masses is None else _lib.to_device(np.asarray(masses, dtype=np.float64), torch.float64)
        cutoff = force_field.cutoff_distance
        self.cutoff_sq = -1.0 if cutoff is None else float(cutoff ** 2)
        self.patch = _patch_struct(force_field, self.n, self._keep)
        self.use_cell_list = bool(use_cell_list and cutoff is not None and self.B == 1
                                  and self.n >= CELL_LIST_MIN_N)
        self._contacts()
        self._assemble()

    # ---- K1 -----------------------------------------------------------------
    def _contacts(self):
        torch = _torch()
        h, st = self.handle, _lib.stream_ptr()
        nrows = self.B * self.n
        pp = C.byref(self.patch) if self.patch is not None else None
        rowcount = torch.empty(nrows, dtype=torch.int32, device="cuda")
        self.rowptr = torch.empty(nrows + 1, dtype=torch.int64, device="cuda")
        scratch = torch.empty(max(1, h.scb_scan_scratch_bytes(nrows)), dtype=torch.uint8, device="cuda")
        _lib.check(h.scb_contacts_count(_lib.ptr(self.xyz), self.B, self.n, self.cutoff_sq, pp,
                                        int(self.use_cell_list), _lib.ptr(rowcount), st))
        _lib.check(h.scb_contacts_scan(_lib.ptr(rowcount), nrows, _lib.ptr(self.rowptr), _lib.ptr(scratch), st))
        self.P = int(self.rowptr[-1].item())
        self.col = torch.empty(max(self.P, 1), dtype=torch.int32, device="cuda")
        if self.P:
            _lib.check(h.scb_contacts_fill(_lib.ptr(self.xyz), self.B, self.n, self.cutoff_sq, pp,
                                           int(self.use_cell_list), _lib.ptr(self.rowptr), _lib.ptr(self.col), st))

    def pairs(self):
        """(P,2) int64 host array, lexicographic, both directions (interaction.py:177-178)."""
        torch = _torch()
        out = torch.empty((max(self.P, 1), 2), dtype=torch.int64, device="cuda")
        if self.P:
            _lib.check(self.handle.scb_pairs_materialize(_lib.ptr(self.rowptr), _lib.ptr(self.col), self.B, self.n,
                                                         _lib.ptr(out), _lib.stream_ptr()))
        return out[: self.P].cpu().numpy()

    def pair_geometry(self):
        torch = _torch()
        disp = torch.empty((max(self.P, 1), 3), dtype=torch.float64, device="cuda")
        sq = torch.empty(max(self.P, 1), dtype=torch.float64, device="cuda")
        if self.P:
            _lib.check(self.handle.scb_pair_geometry(_lib.ptr(self.xyz), self.B, self.n, _lib.ptr(self.rowptr),
                                                     _lib.ptr(self.col), _lib.ptr(disp), _lib.ptr(sq),
                                                     _lib.stream_ptr()))
        return disp[: self.P], sq[: self.P]

    # ---- K2 -----------------------------------------------------------------
    def _assemble(self):
        torch = _torch()
        h, st = self.handle, _lib.stream_ptr()
        desc = self.ff._descriptor(self.n)
        if desc is None:
            # user-defined ForceField: host callback between kernel 1 and kernel 2
            # (doc/advanced.rst:23-70, test_interaction.py:99-101)
            pairs = self.pairs()
            _, sq = self.pair_geometry()
            fc = np.asarray(self.ff.force_constant(pairs[:, 0], pairs[:, 1], sq.cpu().numpy()), dtype=np.float64)
            if fc.shape != (self.P,):
                raise IndexError(f"force_constant() returned shape {fc.shape}, expected {(self.P,)}")
            d = _lib.FFDesc()
            d.kind = _lib.SCB_FF.EXTERNAL
            d.nbins = 1
            d.cutoff_sq = self.cutoff_sq
            t = _lib.to_device(fc, torch.float64)
            d.external_fc = t.data_ptr()
            desc = (d, [t])
        self.desc, keep = desc
        self._keep.extend(keep)
        DD = self.D * self.D
        self.offdiag = torch.empty((max(self.P, 1), DD), dtype=torch.float64, device="cuda")
        self.diag = torch.empty((self.B * self.n, DD), dtype=torch.float64, device="cuda")
        self.gersh = torch.empty(self.B, dtype=torch.float64, device="cuda")
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(h.scb_assemble(self.D, _lib.ptr(self.xyz), self.B, self.n, C.byref(self.desc),
                                  _lib.ptr(self.rowptr), _lib.ptr(self.col), _lib.ptr(self.masses),
                                  _lib.ptr(self.offdiag), _lib.ptr(self.diag), _lib.ptr(self.gersh),
                                  _lib.ptr(flag), st))
        status = int(flag.item())
        if status:
            _lib.check(status)

    def dense(self):
        """Dense [B][N][N] device tensor in the reference's layout (interaction.py:106-109)."""
        torch = _torch()
        out = torch.empty((self.B, self.N, self.N), dtype=torch.float64, device="cuda")
        _lib.check(self.handle.scb_densify(self.D, self.B, self.n, _lib.ptr(self.rowptr), _lib.ptr(self.col),
                                           _lib.ptr(self.offdiag), _lib.ptr(self.diag), _lib.ptr(out),
                                           _lib.stream_ptr()))
        return out

    # ---- K3 -----------------------------------------------------------------
    def spmm(self, X):
        torch = _torch()
        b = int(X.shape[-1])
        Y = torch.empty_like(X)
        _lib.check(self.handle.scb_spmm(self.D, self.B, self.n, _lib.ptr(self.rowptr), _lib.ptr(self.col),
                                        _lib.ptr(self.offdiag), _lib.ptr(self.diag), _lib.ptr(X), _lib.ptr(Y), b,
                                        _lib.stream_ptr()))
        return Y

    def paired(self):
        """Row-paired copy of the operator (built once, cached)."""
        torch = _torch()
        if getattr(self, "_paired", None) is None:
            nbytes = self.handle.scb_paired_bytes(self.D, self.B, self.n, self.P, None, None)
            buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            _lib.check(self.handle.scb_paired_build(self.D, self.B, self.n, self.P, _lib.ptr(self.rowptr),
                                                    _lib.ptr(self.col), _lib.ptr(self.offdiag), _lib.ptr(self.diag),
                                                    _lib.ptr(buf), _lib.stream_ptr()))
            self._paired = buf
        return self._paired

    def spmm_paired(self, X, Y=None):
        torch = _torch()
        b = int(X.shape[-1])
        if Y is None:
            Y = torch.empty_like(X)
        _lib.check(self.handle.scb_spmm_paired(self.D, self.B, self.n, self.P, _lib.ptr(self.rowptr),
                                               _lib.ptr(self.paired()), _lib.ptr(X), _lib.ptr(Y), b,
                                               _lib.stream_ptr()))
        return Y

    def rigid_basis(self):
        torch = _torch()
        nz = 6 if self.D == 3 else 1
        Z = torch.empty((self.B, self.N, nz), dtype=torch.float64, device="cuda")
        _lib.check(self.handle.scb_rigid_basis(self.D, _lib.ptr(self.xyz), self.B, self.n, _lib.ptr(self.masses),
                                               _lib.ptr(Z), _lib.stream_ptr()))
        return Z

    def eig_lowest(self, k, deflate=True, tol=3e-9, max_outer=300, degree=24, seed=0x5CB200, b=None):
        """The k lowest modes of the (optionally rigid-body deflated) operator.

        Returns (eigval[B][b], X[B][N][b], resid[B][b], iters[B]) device tensors;
        columns 0..k-1 are converged."""
        torch = _torch()
        h = self.handle
        nz = (6 if self.D == 3 else 1) if deflate else 0
        if b is None:
            b = 32 if k + 8 <= 32 else 64
        if k > b:
            raise NotImplementedError(f"k={k} modes need a block wider than {b}")
        Z = self.rigid_basis() if deflate else None
        eigval = torch.empty((self.B, b), dtype=torch.float64, device="cuda")
        X = torch.empty((self.B, self.N, b), dtype=torch.float64, device="cuda")
        resid = torch.empty((self.B, b), dtype=torch.float64, device="cuda")
        iters = torch.empty(self.B, dtype=torch.int32, device="cuda")
        ws_bytes = h.scb_eig_lowest_workspace_bytes(self.D, self.B, self.n, b, nz, self.P)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        status = h.scb_eig_lowest(self.D, self.B, self.n, self.P, _lib.ptr(self.rowptr), _lib.ptr(self.col),
                                  _lib.ptr(self.offdiag), _lib.ptr(self.diag), _lib.ptr(self.gersh), _lib.ptr(Z), nz,
                                  k, b, tol, max_outer, degree, seed, _lib.ptr(eigval), _lib.ptr(X), _lib.ptr(resid),
                                  _lib.ptr(iters), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
        _lib.check(status)
        return eigval, X, resid, iters, Z


_EIG_SOLVERS = {"auto": 0, "jacobi": 1, "tridiag": 2}


def eig_full_dense(A, solver="auto"):
    """Full eigendecomposition of dense symmetric device matrices A[B][N][N]
    (destroyed).  Returns (eigval[B][N], modes[B][N][N]) with rows = modes.
    ``solver``: "auto" (tridiagonalisation + divide and conquer for N > 64), "jacobi" (block Jacobi: ordinary
    launches only, for callers whose other streams keep SMs busy for an unknown time) or "tridiag"."""
    torch = _torch()
    h = _lib.require_device()
    if A.dim() == 2:
        A = A[None]
    B, N = int(A.shape[0]), int(A.shape[1])
    eigval = torch.empty((B, N), dtype=torch.float64, device="cuda")
    modes = torch.empty((B, N, N), dtype=torch.float64, device="cuda")
    code = _EIG_SOLVERS[solver]
    ws_bytes = h.scb_eig_full_workspace_bytes_ex(code, B, N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    _lib.check(h.scb_eig_full_ex(code, B, N, _lib.ptr(A), _lib.ptr(eigval), _lib.ptr(modes), _lib.ptr(ws), ws_bytes,
                                 _lib.stream_ptr()))
    return eigval, modes


def modes_msf(D, lam, modes, scale=1.0):
    """lam[B][m], modes[B][m][N] -> msf[B][n] (nma.py:145-183)."""
    torch = _torch()
    h = _lib.require_device()
    B, m, N = (int(x) for x in modes.shape)
    n = N // D
    out = torch.empty((B, n), dtype=torch.float64, device="cuda")
    _lib.check(h.scb_msf(D, B, n, m, _lib.ptr(lam), _lib.ptr(modes), float(scale), _lib.ptr(out), _lib.stream_ptr()))
    return out


def modes_dcc(D, lam, modes, norm=True, scale=1.0, rows=None):
    """lam[m], modes[m][N] -> dcc rows [row0,row1) x n (nma.py:338-357)."""
    torch = _torch()
    h = _lib.require_device()
    m, N = (int(x) for x in modes.shape)
    n = N // D
    row0, row1 = (0, n) if rows is None else rows
    out = torch.empty((row1 - row0, n), dtype=torch.float64, device="cuda")
    ws_bytes = h.scb_dcc_workspace_bytes(D, n, m)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    _lib.check(h.scb_dcc(D, n, m, _lib.ptr(lam), _lib.ptr(modes), int(bool(norm)), float(scale), row0, row1,
                         _lib.ptr(out), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
    return out


def modes_covariance(lam, modes, rows=None):
    """sum_k u_k u_k^T / lam_k over the given modes (anm.py:132-136)."""
    torch = _torch()
    h = _lib.require_device()
    m, N = (int(x) for x in modes.shape)
    row0, row1 = (0, N) if rows is None else rows
    out = torch.empty((row1 - row0, N), dtype=torch.float64, device="cuda")
    ws_bytes = h.scb_dcc_workspace_bytes(1, N, m)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    _lib.check(h.scb_covariance(N, m, _lib.ptr(lam), _lib.ptr(modes), row0, row1, _lib.ptr(out), _lib.ptr(ws),
                                ws_bytes, _lib.stream_ptr()))
    return out


def modes_linear_response(lam, modes, force):
    """sum_k u_k (u_k . f) / lam_k (nma.py:473)."""
    torch = _torch()
    h = _lib.require_device()
    m, N = (int(x) for x in modes.shape)
    out = torch.empty(N, dtype=torch.float64, device="cuda")
    ws = torch.empty(8 * m + 256, dtype=torch.uint8, device="cuda")
    _lib.check(h.scb_linear_response(N, m, _lib.ptr(lam), _lib.ptr(modes), _lib.ptr(force), _lib.ptr(out),
                                     _lib.ptr(ws), 8 * m + 256, _lib.stream_ptr()))
    return out


def cov_dcc(D, cov, norm=True, scale=1.0):
    """DCC from a covariance matrix assigned by the caller (nma.py:324-357)."""
    torch = _torch()
    h = _lib.require_device()
    c = torch.from_numpy(np.ascontiguousarray(cov, dtype=np.float64)).cuda()
    n = int(c.shape[0]) // D
    out = torch.empty((n, n), dtype=torch.float64, device="cuda")
    diag = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.check(h.scb_dcc_from_covariance(D, n, _lib.ptr(c), int(bool(norm)), float(scale), _lib.ptr(out),
                                         _lib.ptr(diag), _lib.stream_ptr()))
    return out


def cov_matvec(cov, force):
    """cov @ force for a covariance assigned by the caller (nma.py:473)."""
    torch = _torch()
    h = _lib.require_device()
    c = torch.from_numpy(np.ascontiguousarray(cov, dtype=np.float64)).cuda()
    f = _lib.to_device(force, torch.float64)
    y = torch.empty(int(c.shape[0]), dtype=torch.float64, device="cuda")
    _lib.check(h.scb_symv(int(c.shape[0]), _lib.ptr(c), _lib.ptr(f), _lib.ptr(y), _lib.stream_ptr()))
    return y
