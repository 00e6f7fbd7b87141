"""Interaction-matrix assembly: same functions and return values as
springcraft/interaction.py, computed by the contact (K1) and fused
force-constant/assembly (K2) kernels."""

import numpy as np

from ._engine import DeviceModel

__all__ = ["compute_kirchhoff", "compute_hessian"]


def _compute(coord, force_field, use_cell_list, D):
    coord = np.asarray(coord)
    if coord.ndim != 2 or coord.shape[1] != 3:  # interaction.py:141-142
        raise ValueError(f"Expected coordinates with shape (n,3), got {coord.shape}")
    model = DeviceModel(coord, force_field, D, use_cell_list=use_cell_list)
    dense = model.dense()[0].cpu().numpy()
    return dense, model.pairs()


def compute_kirchhoff(coord, force_field, use_cell_list=True):
    """Kirchhoff matrix (n,n) float64 and the (k,2) contact pairs (interaction.py:14-54).

    ``use_cell_list`` selects the cell-list contact kernel for large systems;
    both contact kernels return identical contact sets."""
    return _compute(coord, force_field, use_cell_list, 1)


def compute_hessian(coord, force_field, use_cell_list=True):
    """Hessian (3n,3n) float64, ``[x1, y1, z1, ... xn, yn, zn]`` partitioning, and
    the (k,2) contact pairs (interaction.py:57-111)."""
    return _compute(coord, force_field, use_cell_list, 3)
