"""Stub of biotite.structure (test infrastructure; see ../__init__.py)."""
import numpy as np
from . import info  # noqa: F401


class BadStructureError(Exception):
    pass


class AtomArray:
    _FIELDS = ("coord", "chain_id", "res_id", "res_name", "atom_name", "element", "hetero")

    def __init__(self, length):
        self.coord = np.zeros((length, 3), dtype=np.float32)
        self.chain_id = np.zeros(length, dtype="U4")
        self.res_id = np.zeros(length, dtype=int)
        self.res_name = np.zeros(length, dtype="U5")
        self.atom_name = np.zeros(length, dtype="U6")
        self.element = np.zeros(length, dtype="U2")
        self.hetero = np.zeros(length, dtype=bool)

    def array_length(self):
        return len(self.coord)

    def __len__(self):
        return len(self.coord)

    def _take(self, index):
        if isinstance(index, (int, np.integer)):
            raise NotImplementedError("single Atom access not needed by the reference")
        new = AtomArray.__new__(AtomArray)
        for f in self._FIELDS:
            setattr(new, f, getattr(self, f)[index])
        return new

    def __getitem__(self, index):
        return self._take(index)

    def copy(self):
        new = AtomArray.__new__(AtomArray)
        for f in self._FIELDS:
            setattr(new, f, getattr(self, f).copy())
        return new

    def __add__(self, other):
        new = AtomArray.__new__(AtomArray)
        for f in self._FIELDS:
            setattr(new, f, np.concatenate([getattr(self, f), getattr(other, f)]))
        return new


def coord(item):
    if isinstance(item, AtomArray):
        return item.coord
    return np.asarray(item)


def displacement(atoms1, atoms2, box=None):
    return coord(atoms2) - coord(atoms1)


def index_displacement(atoms, indices, periodic=False, box=None):
    c = coord(atoms)
    return c[indices[:, 1]] - c[indices[:, 0]]


def distance(atoms1, atoms2, box=None):
    d = displacement(atoms1, atoms2)
    return np.sqrt((d * d).sum(axis=-1))


class CellList:
    """Brute-force stand-in (fp32 like biotite's Cython cell list is believed to be)."""

    def __init__(self, atom_array, cell_size, periodic=False, box=None, selection=None):
        self._coord = coord(atom_array).astype(np.float32)

    def create_adjacency_matrix(self, threshold_distance):
        c = self._coord
        d = c[np.newaxis, :, :] - c[:, np.newaxis, :]
        sq = (d * d).sum(axis=-1)
        return sq <= np.float32(threshold_distance) ** 2


def get_chain_count(array):
    starts = np.where(array.chain_id[1:] != array.chain_id[:-1])[0]
    return len(starts) + 1


def check_res_id_continuity(array):
    diff = np.diff(array.res_id)
    return np.where((diff != 0) & (diff != 1))[0] + 1
