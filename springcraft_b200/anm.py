"""Anisotropic Network Model: the springcraft.ANM surface (anm.py:20-445) on
the device-resident engine."""

from . import nma
from ._enm import K_B, N_A, ENMBase  # noqa: F401

__all__ = ["ANM"]


class ANM(ENMBase):
    """``ANM(atoms, force_field, masses=None, use_cell_list=True)`` (anm.py:62)."""

    _D = 3

    @property
    def hessian(self):
        return self._get_matrix()

    @hessian.setter
    def hessian(self, value):
        self._set_matrix(value, IndexError)  # anm.py:122-127

    def normal_mode(self, index, amplitude, frames, movement="sine"):
        return nma.normal_mode(self, index, amplitude, frames, movement)

    def linear_response(self, force, *, mode_subset=None):
        return nma.linear_response(self, force, mode_subset=mode_subset)

    def prs_effector_sensor(self, norm=True):
        prs_mat = nma.prs(self, norm)
        eff, sens = nma.effector_sensor(prs_mat)
        return prs_mat, eff, sens
