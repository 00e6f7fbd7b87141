"""Shared façade logic of ANM and GNM: lazy caches with mutual invalidation
(anm.py:98-148, gnm.py:91-143) in front of a device-resident model."""

import numpy as np

from . import _engine, _lib
from .structure import coord as _coord_of
from .structure import is_atom_array, residue_mass

K_B = 1.380649e-23
N_A = 6.02214076e23


class ENMBase:
    _D = 3          # 3 = ANM / Hessian, 1 = GNM / Kirchhoff
    _NAME = "hessian"

    def __init__(self, atoms, force_field, masses=None, use_cell_list=True):
        self._coord = _coord_of(atoms)
        self._ff = force_field
        self._use_cell_list = use_cell_list
        n = len(self._coord)
        if masses is None or masses is False:      # anm.py:67-68
            self._masses = None
        elif masses is True:                       # anm.py:69-79
            if not is_atom_array(atoms):
                raise TypeError("An AtomArray is required to automatically infer masses")
            self._masses = np.array([residue_mass(r) for r in atoms.res_name])
        else:                                      # anm.py:80-87
            if len(masses) != n:
                raise IndexError(f"{len(masses)} masses for {n} atoms given")
            if np.any(np.asarray(masses) == 0):
                raise ValueError("Masses must not be 0")
            self._masses = np.array(masses, dtype=float)
        self._matrix = None        # host dense Hessian / Kirchhoff
        self._covariance = None    # host dense covariance
        self._user_matrix = False  # matrix (or covariance) was assigned by the caller
        self._user_covariance = False
        self._exposed_token = None  # fingerprint of the host matrix at the time it was handed to the caller
        self._model = None
        self._spectrum_cache = {}

    # ---- device side ---------------------------------------------------------
    def _has_model(self):
        self._sync_exposed()
        return not self._user_matrix

    def _matrix_token(self):
        M = self._matrix
        return (float(M.sum()), float(np.vdot(M, M)), float(M.flat[:: max(1, M.size // 97)].sum()))

    def _sync_exposed(self):
        """`enm.hessian` / `enm.kirchhoff` return the cached array itself, not a copy (anm.py:53-57); the reference
        therefore sees in-place edits by the caller at the next eigen()/MSF/... call.  The device model cannot, so
        the array's fingerprint is compared with the one taken when it was handed out: if it changed, the host array
        becomes the source of truth (dense path) and cached spectra are dropped."""
        if self._matrix is None or self._exposed_token is None:
            return
        token = self._matrix_token()
        if token != self._exposed_token:
            self._exposed_token = token
            self._user_matrix = True
            self._spectrum_cache = {}

    def _model_device(self):
        if self._model is None:
            self._model = _engine.DeviceModel(self._coord, self._ff, self._D, masses=self._masses,
                                              use_cell_list=self._use_cell_list)
        return self._model

    def _matrix_device(self):
        """Dense [N][N] device tensor of the mechanical matrix."""
        import torch
        if self._has_model() and self._matrix is None:
            return self._model_device().dense()[0]
        return torch.from_numpy(np.ascontiguousarray(self._get_matrix(), dtype=np.float64)).cuda()

    @staticmethod
    def _pinv_device(M):
        """np.linalg.pinv(M, hermitian=True, rcond=1e-6) (anm.py:115-117,135)."""
        import torch
        A = torch.from_numpy(np.ascontiguousarray(M, dtype=np.float64)).cuda()
        lam, modes = _engine.eig_full_dense(A)
        lam, modes = lam[0], modes[0]
        keep = torch.abs(lam) > 1e-6 * torch.max(torch.abs(lam))
        return _engine.modes_covariance(lam[keep].contiguous(), modes[keep].contiguous()).cpu().numpy()

    # ---- lazy caches -----------------------------------------------------------
    def _get_matrix(self):
        if self._matrix is None:
            if self._covariance is None:
                self._matrix = self._model_device().dense()[0].cpu().numpy()
            else:
                self._matrix = self._pinv_device(self._covariance)
        if self._exposed_token is None:
            self._exposed_token = self._matrix_token()
        else:
            self._sync_exposed()
        return self._matrix

    def _set_matrix(self, value, exc):
        N = len(self._coord) * self._D
        if value.shape != (N, N):
            raise exc(f"Expected shape {(N, N)}, got {value.shape}")
        self._matrix = value
        self._covariance = None
        self._user_matrix = True
        self._user_covariance = False
        self._exposed_token = self._matrix_token()
        self._spectrum_cache = {}

    @property
    def masses(self):
        return self._masses

    @property
    def covariance(self):
        if self._covariance is None:
            from . import nma
            lam, modes = nma._pinv_modes(self)
            self._covariance = _engine.modes_covariance(lam, modes).cpu().numpy()
        return self._covariance

    @covariance.setter
    def covariance(self, value):
        N = len(self._coord) * self._D
        if value.shape != (N, N):
            raise IndexError(f"Expected shape {(N, N)}, got {value.shape}")
        self._covariance = value
        self._matrix = None
        self._user_matrix = True
        self._user_covariance = True
        self._exposed_token = None
        self._spectrum_cache = {}

    # ---- NMA methods shared by both models -----------------------------------
    def eigen(self, *, k=None):
        from . import nma
        return nma.eigen(self, k=k)

    def frequencies(self):
        from . import nma
        return nma.frequencies(self)

    def mean_square_fluctuation(self, mode_subset=None, tem=None, tem_factors=K_B):
        from . import nma
        return nma.mean_square_fluctuation(self, mode_subset, tem, tem_factors)

    def bfactor(self, mode_subset=None, tem=None, tem_factors=K_B):
        from . import nma
        return nma.bfactor(self, mode_subset, tem, tem_factors)

    def dcc(self, mode_subset=None, norm=True, tem=None, tem_factors=K_B):
        from . import nma
        return nma.dcc(self, mode_subset, norm, tem, tem_factors)


del _lib
