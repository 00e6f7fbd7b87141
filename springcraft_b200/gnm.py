"""Gaussian Network Model: the springcraft.GNM surface (gnm.py:20-303) on the
device-resident engine."""

from ._enm import K_B, N_A, ENMBase  # noqa: F401

__all__ = ["GNM"]


class GNM(ENMBase):
    """``GNM(atoms, force_field, masses=None, use_cell_list=True)`` (gnm.py:58)."""

    _D = 1

    @property
    def kirchhoff(self):
        return self._get_matrix()

    @kirchhoff.setter
    def kirchhoff(self, value):
        self._set_matrix(value, ValueError)  # gnm.py:115-120
