"""Force fields of the elastic-network hot path (host-side mirror).

Same classes, constructor arguments, properties and exceptions as
springcraft/forcefield.py (reference lines cited per item); the arithmetic runs
in the CUDA library: every built-in force field compiles itself into the POD
``scb_ff_desc`` (include/scb200.h) that the fused assembly kernel consumes, and
``force_constant()`` itself is a kernel launch (``scb_force_constant``).
User-defined :class:`ForceField` subclasses keep working through the
``SCB_FF_EXTERNAL`` path (their own ``force_constant`` is called on the host
between the contact kernel and the assembly kernel, doc/advanced.rst:23-70).
"""

import abc
import functools
import numbers
from os.path import dirname, join, realpath

import numpy as np

from . import _lib
from .structure import BadStructureError, is_atom_array

__all__ = ["ForceField", "PatchedForceField", "InvariantForceField", "HinsenForceField",
           "ParameterFreeForceField", "TabulatedForceField"]

DATA_DIR = join(dirname(realpath(__file__)), "data")
N_AMINO_ACIDS = 20
# biotite ProteinSequence alphabet, first 20 symbols, as 3-letter codes (forcefield.py:28-34)
AA_LIST = ["ALA", "CYS", "ASP", "GLU", "PHE", "GLY", "HIS", "ILE", "LYS", "LEU",
           "MET", "ASN", "PRO", "GLN", "ARG", "SER", "THR", "VAL", "TRP", "TYR"]
AA_TO_INDEX = {aa: i for i, aa in enumerate(AA_LIST)}


class ForceField(metaclass=abc.ABCMeta):
    """Interface of springcraft.ForceField (forcefield.py:37-114)."""

    @abc.abstractmethod
    def force_constant(self, atom_i, atom_j, sq_distance):
        pass

    @property
    def cutoff_distance(self):
        return None

    @property
    def contact_shutdown(self):
        return None

    @property
    def contact_pair_off(self):
        return None

    @property
    def contact_pair_on(self):
        return None

    @property
    def natoms(self):
        return None

    # ---- device side -------------------------------------------------------
    def _descriptor(self, n):
        """(FFDesc, keepalive) for built-in kinds; None for user subclasses."""
        return None


def _cutoff_sq(cutoff):
    return -1.0 if cutoff is None else float(cutoff ** 2)


def _eval_on_device(ff, atom_i, atom_j, sq_distance, n=None):
    """Run the force-constant kernel for explicit (i, j, sq) triples."""
    import torch
    handle = _lib.require_device()
    atom_i = np.asarray(atom_i)
    atom_j = np.asarray(atom_j)
    sq = np.asarray(sq_distance, dtype=np.float64)
    if atom_i.shape != atom_j.shape or atom_i.shape != sq.shape or atom_i.ndim != 1:
        raise IndexError("atom_i, atom_j and sq_distance must be 1D arrays of equal length")
    if n is None:
        n = int(max(atom_i.max(initial=0), atom_j.max(initial=0))) + 1
    desc, keep = ff._descriptor(n)
    P = len(sq)
    out = torch.empty(P, dtype=torch.float64, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    di = _lib.to_device(atom_i.astype(np.int32), torch.int32)
    dj = _lib.to_device(atom_j.astype(np.int32), torch.int32)
    ds = _lib.to_device(sq, torch.float64)
    _lib.check(handle.scb_force_constant(desc, n, _lib.ptr(di), _lib.ptr(dj), _lib.ptr(ds), P,
                                         _lib.ptr(out), _lib.ptr(flag), _lib.stream_ptr()))
    status = int(flag.item())
    if status != 0:
        _lib.check(status)
    del keep
    return out.cpu().numpy()


class PatchedForceField(ForceField):
    """forcefield.py:117-261."""

    def __init__(self, force_field, contact_shutdown=None, contact_pair_off=None,
                 contact_pair_on=None, force_constants=None):
        self._force_field = force_field
        given = {"_contact_shutdown": contact_shutdown, "_contact_pair_off": contact_pair_off,
                 "_contact_pair_on": contact_pair_on, "_force_constants": force_constants}
        for attr, value in given.items():            # any array-like is accepted
            setattr(self, attr, None if value is None else np.asarray(value))
        for attr in ("_contact_shutdown", "_contact_pair_off", "_contact_pair_on"):
            _require_in_range(getattr(self, attr), force_field.natoms)
        on, fcs = self._contact_pair_on, self._force_constants
        if on is not None and fcs is None:           # forcefield.py:171-175
            raise TypeError("Individual force constants must be given, if contacts are turned on")
        if on is not None and len(fcs) != len(on):   # forcefield.py:176-181
            raise IndexError(f"{len(fcs)} force constants were given for {len(on)} switched on contact_pairs")

    def force_constant(self, atom_i, atom_j, sq_distance):
        if self._force_field._descriptor(1) is None:
            # user-defined base force field: evaluate it on the host, patch on top
            return _patch_on_host(self, atom_i, atom_j, sq_distance)
        return _eval_on_device(self, atom_i, atom_j, sq_distance, self.natoms)

    @property
    def cutoff_distance(self):
        return self._force_field.cutoff_distance

    def _stacked(self, name):
        """Own patch entries followed by those of the wrapped force field (forcefield.py:232-257)."""
        mine, inner = getattr(self, "_" + name), getattr(self._force_field, name)
        return mine if inner is None else np.concatenate([mine, inner])

    contact_shutdown = property(lambda self: self._stacked("contact_shutdown"))
    contact_pair_off = property(lambda self: self._stacked("contact_pair_off"))
    contact_pair_on = property(lambda self: self._stacked("contact_pair_on"))

    @property
    def natoms(self):
        return self._force_field.natoms

    def _descriptor(self, n):
        import torch
        base = self._force_field._descriptor(n)
        if base is None:
            return None
        desc, keep = base
        desc.patched = 1
        if self._contact_pair_on is not None:
            # an inner PatchedForceField's own on-pairs are already in desc; ours
            # are applied after them (outermost wins, forcefield.py:197-223)
            pairs = self._contact_pair_on.reshape(-1, 2).astype(np.int32)
            fcs = self._force_constants.astype(np.float64)
            if desc.n_pair_on:
                prev_pairs, prev_fc = keep[-2], keep[-1]
                pairs_t = torch.cat([prev_pairs, _lib.to_device(pairs, torch.int32)])
                fcs_t = torch.cat([prev_fc, _lib.to_device(fcs, torch.float64)])
            else:
                pairs_t = _lib.to_device(pairs, torch.int32)
                fcs_t = _lib.to_device(fcs, torch.float64)
            desc.n_pair_on = int(pairs_t.shape[0])
            desc.pair_on = pairs_t.data_ptr()
            desc.pair_on_fc = fcs_t.data_ptr()
            keep = keep + [pairs_t, fcs_t]
        return desc, keep


def _patch_on_host(pff, atom_i, atom_j, sq_distance):
    """PatchedForceField around a USER force field: the user's callback decides
    the base constants (forcefield.py:183-226 applied to its output)."""
    base = pff._force_field
    sq_distance = np.asarray(sq_distance)
    if base.cutoff_distance is None:
        fc = np.asarray(base.force_constant(atom_i, atom_j, sq_distance), dtype=float)
    else:
        fc = np.zeros(len(sq_distance))
        mask = sq_distance <= base.cutoff_distance ** 2
        fc[mask] = base.force_constant(atom_i[mask], atom_j[mask], sq_distance[mask])
    if pff._contact_pair_on is not None:
        lut = {}
        for (a, b), v in zip(pff._contact_pair_on.reshape(-1, 2), pff._force_constants):
            lut[(int(a), int(b))] = lut[(int(b), int(a))] = float(v)
        for p, (a, b) in enumerate(zip(atom_i, atom_j)):
            v = lut.get((int(a), int(b)))
            if v is not None and v != -1:
                fc[p] = v
    return fc


class InvariantForceField(ForceField):
    """forcefield.py:264-289."""

    def __init__(self, cutoff_distance):
        if cutoff_distance is None:
            raise ValueError("Cutoff distance must be a float")
        self._cutoff_distance = cutoff_distance

    def force_constant(self, atom_i, atom_j, sq_distance):
        return _eval_on_device(self, atom_i, atom_j, sq_distance)

    @property
    def cutoff_distance(self):
        return self._cutoff_distance

    def _descriptor(self, n):
        d = _lib.FFDesc()
        d.kind = _lib.SCB_FF.INVARIANT
        d.nbins = 1
        d.cutoff_sq = _cutoff_sq(self._cutoff_distance)
        return d, []


class HinsenForceField(ForceField):
    """forcefield.py:292-330."""

    def __init__(self, cutoff_distance=None):
        self._cutoff_distance = cutoff_distance

    def force_constant(self, atom_i, atom_j, sq_distance):
        return _eval_on_device(self, atom_i, atom_j, sq_distance)

    @property
    def cutoff_distance(self):
        return self._cutoff_distance

    def _descriptor(self, n):
        d = _lib.FFDesc()
        d.kind = _lib.SCB_FF.HINSEN
        d.nbins = 1
        d.cutoff_sq = _cutoff_sq(self._cutoff_distance)
        return d, []


class ParameterFreeForceField(ForceField):
    """forcefield.py:333-366."""

    def __init__(self, cutoff_distance=None):
        self._cutoff_distance = cutoff_distance

    def force_constant(self, atom_i, atom_j, sq_distance):
        return _eval_on_device(self, atom_i, atom_j, sq_distance)

    @property
    def cutoff_distance(self):
        return self._cutoff_distance

    def _descriptor(self, n):
        d = _lib.FFDesc()
        d.kind = _lib.SCB_FF.PFREE
        d.nbins = 1
        d.cutoff_sq = _cutoff_sq(self._cutoff_distance)
        return d, []


class TabulatedForceField(ForceField):
    """forcefield.py:369-545.

    The reference materialises an (n, n, k) float32 table in the constructor
    (O(n^2 k) memory).  Here the 20x20xk residue-pair tables plus three per-atom
    attributes {residue type, chain, bonded-to-next} are kept instead; the dense
    table is only built if :attr:`interaction_matrix` is read.
    """

    def __init__(self, atoms, bonded, intra_chain, inter_chain, cutoff_distance):
        if not is_atom_array(atoms):
            raise TypeError(f"Expected 'AtomArray', not {type(atoms).__name__}")
        if not np.all((np.asarray(atoms.atom_name) == "CA") & (np.asarray(atoms.element) == "C")):
            raise BadStructureError("AtomArray does not contain exclusively CA atoms")
        self._natoms = atoms.array_length() if hasattr(atoms, "array_length") else len(atoms.coord)
        if cutoff_distance is None:
            self._edges = None
            n_bins = 1
        elif isinstance(cutoff_distance, numbers.Real):
            self._edges = np.array([cutoff_distance])
            n_bins = 1
        else:
            self._edges = np.asarray(cutoff_distance)
            if not np.all(np.diff(self._edges) >= 0):
                raise ValueError("Distance bin edges are not sorted in increasing order")
            n_bins = len(self._edges)
        self._bonded = _residue_table(bonded, n_bins)
        self._intra_chain = _residue_table(intra_chain, n_bins)
        self._inter_chain = _residue_table(inter_chain, n_bins)
        res_name = np.asarray(atoms.res_name)
        chain_id = np.asarray(atoms.chain_id)
        res_id = np.asarray(atoms.res_id)
        self._res_type = np.array([AA_TO_INDEX[aa] for aa in res_name], dtype=np.uint8)
        _, chain_num = np.unique(chain_id, return_inverse=True)
        self._chain = chain_num.astype(np.int32)
        nxt = np.zeros(self._natoms, dtype=np.uint8)
        if self._natoms > 1:  # forcefield.py:471-473
            nxt[:-1] = (np.diff(res_id) == 1) & (chain_id[:-1] == chain_id[1:])
        self._bonded_next = nxt
        self._interaction_matrix = None

    def force_constant(self, atom_i, atom_j, sq_distance):
        return _eval_on_device(self, atom_i, atom_j, sq_distance, self._natoms).astype(np.float32)

    @property
    def cutoff_distance(self):
        return None if self._edges is None else self._edges[-1]

    @property
    def natoms(self):
        return self._natoms

    @property
    def interaction_matrix(self):
        """(n, n, k) float32 table, returned by reference (forcefield.py:429-434).
        Built lazily; once read, the kernels index THIS array so that in-place
        edits by the caller take effect."""
        if self._interaction_matrix is None:
            n = self._natoms
            t = self._res_type.astype(np.int64)
            same = self._chain[:, None] == self._chain[None, :]
            m = np.where(same[:, :, None], self._intra_chain[t[:, None], t[None, :]],
                         self._inter_chain[t[:, None], t[None, :]]).astype(np.float32)
            lo = np.where(self._bonded_next[:-1])[0] if n > 1 else np.zeros(0, dtype=int)
            const = self._bonded[t[lo], t[lo + 1]]
            m[lo, lo + 1] = const
            m[lo + 1, lo] = const
            m[np.arange(n), np.arange(n)] = 0
            self._interaction_matrix = m
        return self._interaction_matrix

    def _descriptor(self, n):
        import torch
        d = _lib.FFDesc()
        k = self._bonded.shape[-1]
        d.nbins = k
        d.cutoff_sq = _cutoff_sq(self.cutoff_distance)
        keep = []
        if self._edges is not None and k > 1:
            e = _lib.to_device(np.asarray(self._edges, dtype=np.float64) ** 2, torch.float64)
            d.edges_sq = e.data_ptr()
            keep.append(e)
        if self._interaction_matrix is not None:
            d.kind = _lib.SCB_FF.TABULATED_DENSE
            t = _lib.to_device(self._interaction_matrix, torch.float32)
            d.dense_table = t.data_ptr()
            keep.append(t)
            return d, keep
        d.kind = _lib.SCB_FF.TABULATED
        for name, arr, dt in (("bonded", self._bonded, torch.float32), ("intra", self._intra_chain, torch.float32),
                              ("inter", self._inter_chain, torch.float32), ("res_type", self._res_type, torch.uint8),
                              ("chain", self._chain, torch.int32), ("bonded_next", self._bonded_next, torch.uint8)):
            t = _lib.to_device(arr, dt)
            setattr(d, name, t.data_ptr())
            keep.append(t)
        return d, keep

    # ---- presets (forcefield.py:547-876) ------------------------------------
    @staticmethod
    def s_enm_10(atoms):
        fc = _load_matrix("s_enm_10.csv")
        return TabulatedForceField(atoms, 10.0, fc, fc, 10.0)

    @staticmethod
    def s_enm_13(atoms):
        fc = _load_matrix("s_enm_13.csv")
        return TabulatedForceField(atoms, 10.0, fc, fc, 13.0)

    @staticmethod
    def d_enm(atoms):
        fc = _load_matrix("d_enm.csv")
        return TabulatedForceField(atoms, 46.83, fc, fc, _load_matrix("d_enm_edges.csv"))

    @staticmethod
    def sd_enm(atoms):
        # the file stacks 26 (20,20) blocks; kJ/(mol*A^2) scaling as in forcefield.py:693-699
        fc = _load_matrix("sd_enm.csv").reshape(-1, 20, 20).T
        fc = fc * 0.0083144621 * 300 * 10
        bonded = 43.52 * 0.0083144621 * 300 * 10
        return TabulatedForceField(atoms, bonded, fc, fc, _load_matrix("d_enm_edges.csv"))

    @staticmethod
    def e_anm(atoms, nonbonded_mean=False):
        return _e_anm(atoms, "miyazawa.csv", "keskin.csv", nonbonded_mean)

    def e_anm_mj(atoms, nonbonded_mean=False):
        return _e_anm(atoms, "miyazawa.csv", "miyazawa.csv", nonbonded_mean)

    def e_anm_ke(atoms, nonbonded_mean=False):
        return _e_anm(atoms, "keskin.csv", "keskin.csv", nonbonded_mean)


def _e_anm(atoms, intra_file, inter_file, nonbonded_mean):
    intra = _load_matrix(intra_file)
    inter = _load_matrix(inter_file)
    if nonbonded_mean:
        intra = np.average(intra) * np.ones(shape=(20, 20))
        inter = np.average(inter) * np.ones(shape=(20, 20))
    return TabulatedForceField(atoms, 82.0, intra, inter, 13.0)


def _residue_table(value, n_bins):
    """Residue-pair constants as a (20, 20, n_bins) float32 table (what forcefield.py:879-937 accepts):
    a scalar, one value per distance bin, a symmetric 20x20 matrix or a full symmetric 20x20xk table."""
    if np.isnan(value).any():
        raise IndexError("Array contains NaN elements")
    full = (N_AMINO_ACIDS, N_AMINO_ACIDS, n_bins)
    if isinstance(value, numbers.Number):
        return np.full(full, value, dtype=np.float32)
    table = np.asarray(value, dtype=np.float32)
    if table.ndim not in (1, 2, 3):
        raise IndexError(f"Expected array with at most 3 dimensions, {table.ndim} given")
    if table.ndim != 2 and table.shape[-1] != n_bins:
        raise IndexError(f"Array contains {len(table)} elements for {n_bins} distance bins")
    if table.ndim >= 2:
        if table.shape[:2] != full[:2]:
            raise IndexError(f"Expected matrix of shape {full[:2]}, got {table.shape[:2]}")
        if not np.allclose(table, table.swapaxes(0, 1)):
            raise ValueError("Input matrix is not symmetric")
    if table.ndim == 2:
        table = table[:, :, None]
    return np.ascontiguousarray(np.broadcast_to(table, full))


@functools.lru_cache(maxsize=None)
def _load_matrix(fname):
    """Comma separated table shipped in data/ (read once per process)."""
    return np.loadtxt(join(DATA_DIR, fname), delimiter=",")


def _require_in_range(indices, length):
    """Patch indices must address atoms of the structure the force field was built for (forcefield.py:953-962)."""
    if indices is None or length is None:
        return
    beyond = indices[indices >= length]
    if beyond.size:
        raise IndexError(f"Index {beyond.flat[0]} is out of bounds for a structure of length {length}")
