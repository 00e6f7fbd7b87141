"""Ensemble entry point: many conformations of one structure
through contacts -> assembly -> lowest-k modes -> MSF in batched kernels.
The reference has no batched API (it rejects ndim != 2, interaction.py:141-142);
this is the drop-in for the loop ``for c in confs: ANM(c, ff)...``."""

import ctypes as C
import os

import numpy as np

from . import _lib

__all__ = ["enm_ensemble", "enm_ensemble_device", "EnsembleResult"]


class EnsembleResult:
    def __init__(self, eigenvalues, msf, modes, n_pairs, converged, iterations=None):
        self.eigenvalues = eigenvalues   # (B, k)   non-trivial modes, ascending
        self.msf = msf                   # (B, n)   MSF from those k modes
        self.modes = modes               # (B, k, N) or None; rows = modes
        self.n_pairs = n_pairs           # ordered contact pairs in the batch
        self.converged = converged       # every structure met the tolerance
        # (B,) outer iterations of the lowest-k solver per structure; NEGATIVE where the structure did not converge
        # (its rows hold the last iterate); 0 for structures solved by the dense full-spectrum path
        self.iterations = iterations

    @property
    def converged_mask(self):
        """(B,) bool: which structures met the tolerance."""
        if self.iterations is None:
            return np.full(len(self.eigenvalues), bool(self.converged))
        return self.iterations >= 0


def _chunk_limit(B):
    """Structures per library call (SCB_ENSEMBLE_CHUNK overrides, for tests)."""
    forced = os.environ.get("SCB_ENSEMBLE_CHUNK")
    if forced:
        return max(1, min(B, int(forced)))
    return min(B, 32768)


def _patch_of(ff, n, keep):
    from ._engine import _patch_struct
    return _patch_struct(ff, n, keep)


def enm_ensemble(coords, force_field, k=20, kind="anm", masses=None, tol=3e-9, return_modes=False,
                 pinned_out=None):
    """ANM/GNM of B conformations: coords (B, n, 3) host array (float64) or an object with such a
    ``coord`` attribute (biotite ``AtomArrayStack``).

    Returns the k lowest NON-trivial modes' eigenvalues (reference indices
    6..6+k-1 for ANM, 1..k for GNM) and ``mean_square_fluctuation(mode_subset=
    those modes)`` for every conformation.  Host->device and device->host copies
    happen inside the single C-ABI call ``scb_enm_ensemble_host``.

    ``k=None`` asks for ALL non-trivial modes (the reference's default mode set,
    nma.py:145-151): dense matrices go through the batched full-spectrum solver
    in chunks; two orders of magnitude slower than a small k."""
    import torch
    # a biotite AtomArrayStack (or anything with a (B,n,3) ``coord`` attribute) is accepted as well
    coords = np.ascontiguousarray(getattr(coords, "coord", coords), dtype=np.float64)
    if coords.ndim != 3 or coords.shape[2] != 3:
        raise ValueError(f"Expected coordinates with shape (B,n,3), got {coords.shape}")
    handle = _lib.require_device()
    B, n = int(coords.shape[0]), int(coords.shape[1])
    D = 3 if kind == "anm" else 1
    if k is None:
        return _ensemble_all_modes(coords, force_field, D, masses, return_modes)
    if force_field.natoms is not None and force_field.natoms != n:
        raise ValueError(f"Got coordinates for {n} atoms, but forcefield was built for {force_field.natoms} atoms")
    if force_field.cutoff_distance is None:
        # all-pairs force fields (HinsenForceField(), ParameterFreeForceField() defaults) have dense matrices: the
        # batched full-spectrum solver computes every mode, the k lowest non-trivial ones are kept
        return _ensemble_all_modes(coords, force_field, D, masses, return_modes, keep=int(k))
    built = force_field._descriptor(n)
    if built is None:
        raise NotImplementedError("user-defined ForceField subclasses are not supported by the batched path")
    desc, keep = built
    patch = _patch_of(force_field, n, keep)
    m_dev = None if masses is None else _lib.to_device(np.asarray(masses, dtype=np.float64), torch.float64)
    if pinned_out is not None:
        eig, msf, modes = pinned_out
    else:
        eig = np.empty((B, k))
        msf = np.empty((B, n))
        modes = np.empty((B, k, D * n)) if return_modes else None
    # The batch goes through the library in chunks: kernels index structures with blockIdx.y (<= 65,535) and a
    # chunk's scratch must fit the device; on a CUDA out-of-memory status the chunk is halved and retried.
    chunk = _chunk_limit(B)
    iters = np.zeros(B, dtype=np.int32)
    total_pairs, converged, c0 = 0, True, 0
    while c0 < B:
        c1 = min(B, c0 + chunk)
        npairs = C.c_int64(0)
        status = handle.scb_enm_ensemble_host(
            D, coords[c0:c1].ctypes.data_as(C.c_void_p), c1 - c0, n, C.byref(desc),
            C.byref(patch) if patch is not None else None, _lib.ptr(m_dev), k, tol,
            eig[c0:c1].ctypes.data_as(C.c_void_p), msf[c0:c1].ctypes.data_as(C.c_void_p),
            modes[c0:c1].ctypes.data_as(C.c_void_p) if modes is not None else None,
            iters[c0:c1].ctypes.data_as(C.c_void_p), C.byref(npairs), _lib.stream_ptr())
        try:
            _lib.check(status, allow=(_lib.SCB_ERR_NOT_CONVERGED,))
        except RuntimeError as err:
            if "out of memory" in str(err).lower() and chunk > 1:
                chunk = (chunk + 1) // 2
                handle.scb_trim_pool()      # the library's cached scratch
                torch.cuda.empty_cache()    # and the caller's
                continue
            raise
        total_pairs += int(npairs.value)
        converged = converged and status == 0
        c0 = c1
    del keep
    return EnsembleResult(eig, msf, modes, total_pairs, converged, iters)


def _ensemble_all_modes(coords, force_field, D, masses, return_modes, keep=None):
    """All non-trivial modes per conformation (or the `keep` lowest of them): batched assembly -> dense ->
    batched full-spectrum solver (scb_eig_full) -> MSF."""
    from . import _engine
    B, n = int(coords.shape[0]), int(coords.shape[1])
    N = D * n
    ntriv = 6 if D == 3 else 1
    if N <= ntriv:
        raise ValueError("system too small: no non-trivial modes")
    m = N - ntriv if keep is None else keep
    if m > N - ntriv:
        raise ValueError(f"{m} non-trivial modes requested, the system has {N - ntriv}")
    chunk = int(max(1, min(64, 2e9 // (40 * N * N))))
    eig = np.empty((B, m))
    msf = np.empty((B, n))
    modes_out = np.empty((B, m, N)) if return_modes else None
    npairs = 0
    for c0 in range(0, B, chunk):
        c1 = min(B, c0 + chunk)
        model = _engine.DeviceModel(coords[c0:c1], force_field, D, masses)
        lam, modes = _engine.eig_full_dense(model.dense())
        npairs += int(model.P)
        lam_nt = lam[:, ntriv:ntriv + m].contiguous()
        modes_nt = modes[:, ntriv:ntriv + m, :].contiguous()
        eig[c0:c1] = lam_nt.cpu().numpy()
        msf[c0:c1] = _engine.modes_msf(D, lam_nt, modes_nt).cpu().numpy()
        if return_modes:
            modes_out[c0:c1] = modes_nt.cpu().numpy()
    return EnsembleResult(eig, msf, modes_out, npairs, True)


def enm_ensemble_device(xyz_soa, force_field, k=20, kind="anm", masses=None, tol=3e-9, out=None):
    """Same path with DEVICE buffers: xyz_soa is a (B, 3, n) float64 cuda tensor.
    Returns (eigval (B,k), msf (B,n), iters (B,), n_pairs, converged); all tensors stay in HBM."""
    import torch
    handle = _lib.require_device()
    B, _, n = (int(x) for x in xyz_soa.shape)
    D = 3 if kind == "anm" else 1
    desc, keep = force_field._descriptor(n)
    patch = _patch_of(force_field, n, keep)
    m_dev = None if masses is None else _lib.to_device(np.asarray(masses, dtype=np.float64), torch.float64)
    if out is None:
        out = (torch.empty((B, k), dtype=torch.float64, device="cuda"),
               torch.empty((B, n), dtype=torch.float64, device="cuda"),
               torch.empty(B, dtype=torch.int32, device="cuda"))
    eig, msf, iters = out
    total_pairs, converged = 0, True
    chunk = _chunk_limit(B)
    for c0 in range(0, B, chunk):       # blockIdx.y indexes structures: at most 65,535 per library call
        c1 = min(B, c0 + chunk)
        npairs = C.c_int64(0)
        status = handle.scb_enm_ensemble(D, _lib.ptr(xyz_soa[c0:c1]), c1 - c0, n, C.byref(desc),
                                         C.byref(patch) if patch is not None else None, _lib.ptr(m_dev), k, tol,
                                         _lib.ptr(eig[c0:c1]), _lib.ptr(msf[c0:c1]), None, _lib.ptr(iters[c0:c1]),
                                         C.byref(npairs), _lib.stream_ptr())
        _lib.check(status, allow=(_lib.SCB_ERR_NOT_CONVERGED,))
        total_pairs += int(npairs.value)
        converged = converged and status == 0
    del keep
    return eig, msf, iters, total_pairs, converged
