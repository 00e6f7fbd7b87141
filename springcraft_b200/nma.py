"""Normal-mode analysis on top of the device eigensolvers: same free functions,
arguments, defaults and exceptions as springcraft/nma.py.

Keyword-only extensions: ``eigen(enm, k=...)`` returns only the k
lowest modes (trivial modes included, so mode indices keep the reference's
meaning) through the sparse lowest-k solver instead of the full spectrum.
"""

import numpy as np

from . import _engine

__all__ = ["eigen", "frequencies", "mean_square_fluctuation", "bfactor", "dcc", "normal_mode",
           "linear_response", "prs", "effector_sensor"]

K_B = 1.380649e-23
N_A = 6.02214076e23

# mode subsets reaching at most this index are served by the lowest-k solver
LOWEST_K_MAX = 58
# Below this matrix order the dense full-spectrum solver is used even for a few modes: a 32..64-wide block
# would span a large part of the space, where Chebyshev-filtered subspace iteration loses rank.
LOWEST_N_MIN = 512


def _use_lowest(enm, k_total):
    D, _ = _kind(enm)
    return (enm._has_model() and k_total <= LOWEST_K_MAX and D * len(enm._coord) > LOWEST_N_MIN
            and enm._spectrum_cache.get("full") is None)


def _kind(enm, what="GNM/ANM"):
    from .anm import ANM
    from .gnm import GNM
    if isinstance(enm, (GNM, ANM)):
        enm._sync_exposed()      # in-place edits of a matrix that was handed to the caller (anm.py:53-57)
        return (1, 1) if isinstance(enm, GNM) else (3, 6)
    raise ValueError(f"Instance of {what} class expected.")


def _full_spectrum(enm):
    """(lam[N], modes[N][N]) device tensors of the full decomposition (cached)."""
    enm._sync_exposed()
    cache = enm._spectrum_cache
    if cache.get("full") is None:
        A = enm._matrix_device().clone()
        lam, modes = _engine.eig_full_dense(A)
        cache["full"] = (lam[0], modes[0])
    return cache["full"]


def _low_spectrum(enm, k_total):
    """The k_total lowest modes (trivial ones included) via the lowest-k solver:
    analytic rigid-body modes with eigenvalue 0, then the deflated operator's."""
    import torch
    D, ntriv = _kind(enm)
    cache = enm._spectrum_cache
    have = cache.get("low")
    if have is None or have[0].shape[0] < k_total:
        model = enm._model_device()
        k = max(k_total - ntriv, 1)
        lam, X, _, _, Z = model.eig_lowest(k)
        b = lam.shape[1]
        lam_all = torch.cat([torch.zeros(ntriv, dtype=torch.float64, device="cuda"), lam[0, :k]])
        modes = torch.cat([Z[0].T.contiguous(), X[0, :, :k].T.contiguous()])
        cache["low"] = (lam_all, modes)
        have = cache["low"]
        del b
    return have[0][:k_total], have[1][:k_total]


def eigen(enm, *, k=None):
    """Eigenvalues (ascending) and eigenvectors as rows (nma.py:29-63)."""
    _kind(enm)
    if k is None or not _use_lowest(enm, int(k)):
        # every mode (or more than the lowest-k solver's block holds): dense block-Jacobi path
        lam, modes = _full_spectrum(enm)
        if k is not None:
            lam, modes = lam[:k], modes[:k]
    else:
        lam, modes = _low_spectrum(enm, int(k))
    return lam.cpu().numpy(), modes.cpu().numpy()


def frequencies(enm):
    """nma.py:66-105."""
    _, ntriv = _kind(enm)
    eig_values, _ = eigen(enm)
    eig_values[0:ntriv] = np.abs(eig_values[0:ntriv])
    with np.errstate(invalid="ignore"):
        return 1 / (2 * np.pi) * np.sqrt(eig_values)


def _select(enm, mode_subset, ntriv):
    """Device (lam, modes) rows for the requested subset + the subset itself."""
    import torch
    if mode_subset is None:
        lam, modes = _full_spectrum(enm)
        return lam[ntriv:], modes[ntriv:]
    mode_subset = np.asarray(mode_subset)
    if any(mode_subset <= (ntriv - 1)):  # nma.py:161-165, 316-320
        raise ValueError("Trivial modes are included in the current selection. Please check your input.")
    top = int(mode_subset.max()) + 1
    if _use_lowest(enm, top):
        lam, modes = _low_spectrum(enm, top)
    else:
        lam, modes = _full_spectrum(enm)
    idx = torch.as_tensor(mode_subset, dtype=torch.int64, device="cuda")
    return lam[idx].contiguous(), modes[idx].contiguous()


def mean_square_fluctuation(enm, mode_subset=None, tem=None, tem_factors=K_B):
    """nma.py:108-184."""
    D, ntriv = _kind(enm)
    lam, modes = _select(enm, mode_subset, ntriv)
    scale = 1.0 if tem is None else tem * tem_factors
    return _engine.modes_msf(D, lam[None], modes[None].contiguous(), scale)[0].cpu().numpy()


def bfactor(enm, mode_subset=None, tem=None, tem_factors=K_B):
    """nma.py:187-230."""
    _kind(enm)
    msqf = mean_square_fluctuation(enm, mode_subset, tem, tem_factors)
    return 8 * np.pi ** 2 * msqf / 3


def _pinv_modes(enm):
    """Modes kept by np.linalg.pinv(hermitian=True, rcond=1e-6) (anm.py:132-136):
    |lam_k| > 1e-6 * max|lam|."""
    import torch
    lam, modes = _full_spectrum(enm)
    keep = torch.abs(lam) > 1e-6 * torch.max(torch.abs(lam))
    return lam[keep].contiguous(), modes[keep].contiguous()


def dcc(enm, mode_subset=None, norm=True, tem=None, tem_factors=K_B):
    """nma.py:233-359."""
    D, ntriv = _kind(enm)
    scale = 1.0 if tem is None else tem * tem_factors
    if mode_subset is None and enm._user_covariance:
        # the caller assigned the covariance: the reference uses that matrix as is (nma.py:324-336)
        return _engine.cov_dcc(D, enm._covariance, norm=norm, scale=scale).cpu().numpy()
    if mode_subset is None:
        lam, modes = _pinv_modes(enm)       # == enm.covariance (nma.py:324-336)
    else:
        lam, modes = _select(enm, mode_subset, ntriv)
    return _engine.modes_dcc(D, lam, modes, norm=norm, scale=scale).cpu().numpy()


def normal_mode(anm, index, amplitude, frames, movement="sine"):
    """nma.py:363-419: oscillation of one mode, shape (frames, n, 3)."""
    from .anm import ANM
    if not isinstance(anm, ANM):
        raise ValueError("Instance of ANM class expected.")
    if movement not in ("sine", "triangle"):
        raise ValueError(f"Movement '{movement}' is unknown")
    import torch
    from . import _lib
    if index >= 0 and _use_lowest(anm, index + 1):
        _, modes = _low_spectrum(anm, index + 1)
    else:
        _, modes = _full_spectrum(anm)
    n = len(anm._coord)
    mode = modes[index].contiguous()
    out = torch.empty((frames, n, 3), dtype=torch.float64, device="cuda")
    _lib.check(_lib.require_device().scb_normal_mode(n, int(frames), _lib.ptr(mode), float(amplitude),
                                                     int(movement == "triangle"), _lib.ptr(out), _lib.stream_ptr()))
    return out.cpu().numpy()


def linear_response(anm, force, *, mode_subset=None):
    """nma.py:422-473.  ``mode_subset`` (keyword-only extension, SURVEY 8a18 / BASELINE config C5: "linear response
    from m modes") restricts the pseudo-inverse to the given non-trivial modes: V_S^T (L_S^-1 (V_S f)); small
    subsets are served by the lowest-k solver, so no full decomposition is needed."""
    from .anm import ANM
    if not isinstance(anm, ANM):
        raise ValueError("Instance of ANM class expected.")
    force = np.asarray(force)
    n = len(anm._coord)
    if force.ndim == 2:
        if force.shape != (n, 3):
            raise ValueError(f"Expected force with shape {(n, 3)}, got {force.shape}")
        force = force.flatten()
    elif force.ndim == 1:
        if len(force) != n * 3:
            raise ValueError(f"Expected force with length {n * 3}, got {len(force)}")
    else:
        raise ValueError(f"Expected 1D or 2D array, got {force.ndim} dimensions")
    import torch
    from . import _lib
    if mode_subset is None and anm._user_covariance:
        # the caller assigned the covariance: nma.py:473 multiplies with exactly that matrix
        return _engine.cov_matvec(anm._covariance, force.astype(np.float64)).cpu().numpy().reshape(n, 3)
    if mode_subset is None:
        lam, modes = _pinv_modes(anm)
    else:
        lam, modes = _select(anm, mode_subset, 6)
    f = _lib.to_device(force.astype(np.float64), torch.float64)
    return _engine.modes_linear_response(lam, modes, f).cpu().numpy().reshape(n, 3)


def prs(anm, norm=True):
    """nma.py:476-531: perturbation response scanning matrix from the covariance."""
    from .anm import ANM
    if not isinstance(anm, ANM):
        raise ValueError("Instance of ANM class expected.")
    import torch
    from . import _lib
    if anm._covariance is not None:
        cov = torch.from_numpy(np.ascontiguousarray(anm._covariance, dtype=np.float64)).cuda()
    else:
        lam, modes = _pinv_modes(anm)
        cov = _engine.modes_covariance(lam, modes)
    n = anm._coord.shape[0]
    out = torch.empty((n, n), dtype=torch.float64, device="cuda")
    _lib.check(_lib.require_device().scb_prs(n, _lib.ptr(cov), int(bool(norm)), _lib.ptr(out), _lib.stream_ptr()))
    return out.cpu().numpy()


def effector_sensor(prs_matrix):
    """nma.py:534-569."""
    n = len(prs_matrix)
    w = 1 - np.eye(n)
    return (np.average(prs_matrix, weights=w, axis=1), np.average(prs_matrix, weights=w, axis=0))
