"""Minimal structure container for the ENM hot path.

The reference takes ``biotite.structure.AtomArray`` objects (anm.py:62-63,
forcefield.py:437-443).  biotite is optional here: any object exposing
``coord, res_name, chain_id, res_id, atom_name, element`` arrays (biotite's
AtomArray does) is accepted, and :class:`AtomArray` below is a small stand-in
with the same attribute names for environments without biotite.
"""

import numpy as np

__all__ = ["AtomArray", "BadStructureError", "coord", "is_atom_array", "residue_mass", "read_pdb_ca"]

try:  # pragma: no cover - biotite is not installed in the build container
    from biotite.structure import BadStructureError  # type: ignore
except Exception:  # noqa: BLE001
    class BadStructureError(Exception):
        """Raised for structures that are not pure CA traces (forcefield.py:440-443)."""


class AtomArray:
    """Lightweight CA-trace container (attribute-compatible with biotite's AtomArray)."""

    _FIELDS = ("coord", "chain_id", "res_id", "res_name", "atom_name", "element")

    def __init__(self, coord, res_name=None, chain_id=None, res_id=None, atom_name=None, element=None):
        c = np.asarray(coord)
        if c.ndim != 2 or c.shape[1] != 3:
            raise ValueError(f"Expected coordinates with shape (n,3), got {c.shape}")
        n = len(c)
        self.coord = c
        self.res_name = np.asarray(res_name) if res_name is not None else np.full(n, "GLY")
        self.chain_id = np.asarray(chain_id) if chain_id is not None else np.full(n, "A")
        self.res_id = np.asarray(res_id) if res_id is not None else np.arange(1, n + 1)
        self.atom_name = np.asarray(atom_name) if atom_name is not None else np.full(n, "CA")
        self.element = np.asarray(element) if element is not None else np.full(n, "C")
        for f in self._FIELDS:
            if len(getattr(self, f)) != n:
                raise IndexError(f"annotation '{f}' has length {len(getattr(self, f))}, expected {n}")

    def array_length(self):
        return len(self.coord)

    def __len__(self):
        return len(self.coord)

    def __getitem__(self, index):
        return AtomArray(*(getattr(self, f)[index] for f in
                           ("coord", "res_name", "chain_id", "res_id", "atom_name", "element")))

    def copy(self):
        return AtomArray(*(getattr(self, f).copy() for f in
                           ("coord", "res_name", "chain_id", "res_id", "atom_name", "element")))

    def __add__(self, other):
        return AtomArray(*(np.concatenate([getattr(self, f), getattr(other, f)]) for f in
                           ("coord", "res_name", "chain_id", "res_id", "atom_name", "element")))


def is_atom_array(obj):
    return all(hasattr(obj, f) for f in ("coord", "res_name", "chain_id", "res_id", "atom_name", "element"))


def coord(item):
    """biotite.structure.coord: the (n,3) coordinate array of `item` (anm.py:63)."""
    if is_atom_array(item):
        return item.coord
    return np.asarray(item)


# average residue masses (Da) used for ``masses=True`` (anm.py:74-79 -> biotite info.mass)
_RES_MASS = {
    "ALA": 89.09, "ARG": 175.21, "ASN": 132.12, "ASP": 133.10, "CYS": 121.16,
    "GLN": 146.15, "GLU": 147.13, "GLY": 75.07, "HIS": 155.16, "ILE": 131.17,
    "LEU": 131.17, "LYS": 147.20, "MET": 149.21, "PHE": 165.19, "PRO": 115.13,
    "SER": 105.09, "THR": 119.12, "TRP": 204.23, "TYR": 181.19, "VAL": 117.15,
}


def residue_mass(res_name):
    try:  # pragma: no cover
        from biotite.structure import info  # type: ignore
        return info.mass(res_name, is_residue=True)
    except Exception:  # noqa: BLE001
        return _RES_MASS[res_name]


def read_pdb_ca(path, model=1):
    """CA trace of the given model of a PDB file (first altloc only)."""
    rows = []
    current = 1
    with open(path) as fh:
        for line in fh:
            rec = line[:6]
            if rec.startswith("MODEL"):
                current = int(line[10:14])
            elif rec.startswith("ENDMDL"):
                if current == model:
                    break
            elif rec in ("ATOM  ", "HETATM") and current == model:
                if line[16] not in (" ", "A"):
                    continue
                if line[12:16].strip() == "CA" and line[76:78].strip().upper() in ("C", ""):
                    rows.append(line)
    return AtomArray(
        np.array([[float(l[30:38]), float(l[38:46]), float(l[46:54])] for l in rows], dtype=np.float32),
        res_name=np.array([l[17:20].strip() for l in rows]),
        chain_id=np.array([l[21].strip() for l in rows]),
        res_id=np.array([int(l[22:26]) for l in rows]),
    )
