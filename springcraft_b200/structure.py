"""Minimal structure container for the ENM hot path.

The reference takes ``biotite.structure.AtomArray`` objects (anm.py:62-63,
forcefield.py:437-443).  biotite is optional here: any object exposing
``coord, res_name, chain_id, res_id, atom_name, element`` arrays (biotite's
AtomArray does) is accepted, and :class:`AtomArray` below is a small stand-in
with the same attribute names for environments without biotite.
"""

import numpy as np

__all__ = ["AtomArray", "BadStructureError", "coord", "is_atom_array", "residue_mass", "read_pdb_ca",
           "read_pdb_ca_models", "read_cif_ca"]

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


def _is_alpha_carbon(line):
    """ATOM/HETATM line of a C-alpha: atom name CA and element C.  When the element columns are blank the column
    alignment of the name decides: one-letter elements start in column 14 (" CA "), calcium is written "CA  "."""
    if line[12:16].strip() != "CA":
        return False
    element = line[76:78].strip().upper()
    if element:
        return element == "C"
    return line[12] == " "


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
                if _is_alpha_carbon(line):
                    rows.append(line)
    return AtomArray(
        np.array([[float(l[30:38]), float(l[38:46]), float(l[46:54])] for l in rows], dtype=np.float32),
        res_name=np.array([l[17:20].strip() for l in rows]),
        chain_id=np.array([l[21].strip() for l in rows]),
        res_id=np.array([int(l[22:26]) for l in rows]),
    )


def _pdb_ca_rows(fh):
    """CA ATOM/HETATM lines (carbon, first alternate location) grouped by MODEL record, in file order."""
    models, cur = [], None
    for line in fh:
        rec = line[:6]
        if rec.startswith("MODEL"):
            cur = []
            models.append(cur)
        elif rec in ("ATOM  ", "HETATM"):
            if cur is None:          # file without MODEL records
                cur = []
                models.append(cur)
            if line[16] in (" ", "A") and _is_alpha_carbon(line):
                cur.append(line)
        elif rec.startswith("ENDMDL"):
            cur = None
    return [m for m in models if m]


def read_pdb_ca_models(path):
    """Every model of a multi-model PDB file (NMR bundle, trajectory dump) as ``(atoms, coords)``:
    ``atoms`` is the CA trace of the first model, ``coords`` a float64 array (models, n, 3) ready for
    :func:`springcraft_b200.enm_ensemble`.  All models must hold the same CA atoms in the same order."""
    with open(path) as fh:
        models = _pdb_ca_rows(fh)
    if not models:
        raise BadStructureError("no CA atoms found")
    ident = [(l[17:20], l[21], l[22:26]) for l in models[0]]
    for k, rows in enumerate(models[1:], start=2):
        if [(l[17:20], l[21], l[22:26]) for l in rows] != ident:
            raise BadStructureError(f"model {k} does not hold the same CA atoms as model 1")
    first = models[0]
    atoms = AtomArray(
        np.array([[float(l[30:38]), float(l[38:46]), float(l[46:54])] for l in first], dtype=np.float32),
        res_name=np.array([l[17:20].strip() for l in first]),
        chain_id=np.array([l[21].strip() for l in first]),
        res_id=np.array([int(l[22:26]) for l in first]),
    )
    # rounded through float32 like AtomArray.coord (biotite stores float32; the reference widens that to float64,
    # interaction.py:43,88), so that model k of the bundle and read_pdb_ca(path, model=k) give identical matrices
    coords = np.array([[[float(l[30:38]), float(l[38:46]), float(l[46:54])] for l in rows] for rows in models],
                      dtype=np.float32).astype(np.float64)
    return atoms, coords


def _cif_tokens(line):
    """Whitespace-separated mmCIF values with single/double quoting (enough for ``_atom_site`` rows)."""
    out, i, n = [], 0, len(line)
    while i < n:
        ch = line[i]
        if ch.isspace():
            i += 1
        elif ch in "'\"":
            j = i + 1
            while j < n and not (line[j] == ch and (j + 1 == n or line[j + 1].isspace())):
                j += 1
            out.append(line[i + 1:j])
            i = j + 1
        else:
            j = i
            while j < n and not line[j].isspace():
                j += 1
            out.append(line[i:j])
            i = j
    return out


def read_cif_ca(path, model=1):
    """CA trace of one model of an mmCIF (text) file: rows of the ``_atom_site`` loop with group ATOM/HETATM,
    atom id CA, element C, first alternate location; author chain / residue numbering when present."""
    cols, rows, in_loop, in_site = [], [], False, False
    with open(path) as fh:
        for raw in fh:
            line = raw.strip()
            if not line or line.startswith("#"):
                if in_site and rows:
                    break
                continue
            if line == "loop_":
                if in_site and rows:
                    break
                in_loop, in_site, cols = True, False, []
                continue
            if in_loop and line.startswith("_"):
                if line.startswith("_atom_site."):
                    in_site = True
                    cols.append(line.split()[0][len("_atom_site."):])
                elif in_site and rows:
                    break
                else:
                    in_site = False
                continue
            if in_site:
                tok = _cif_tokens(line)
                if len(tok) == len(cols):
                    rows.append(tok)
    if not rows:
        raise BadStructureError("no _atom_site loop found")
    ix = {c: i for i, c in enumerate(cols)}

    def col(*names):
        for nm in names:
            if nm in ix:
                return ix[nm]
        raise BadStructureError(f"_atom_site loop lacks {names[0]}")

    c_atom, c_comp = col("label_atom_id", "auth_atom_id"), col("label_comp_id", "auth_comp_id")
    c_chain, c_seq = col("auth_asym_id", "label_asym_id"), col("auth_seq_id", "label_seq_id")
    c_x, c_y, c_z = col("Cartn_x"), col("Cartn_y"), col("Cartn_z")
    c_grp, c_alt = ix.get("group_PDB"), ix.get("label_alt_id")
    c_el, c_model = ix.get("type_symbol"), ix.get("pdbx_PDB_model_num")
    keep = []
    for r in rows:
        if c_model is not None and int(r[c_model]) != model:
            continue
        if c_grp is not None and r[c_grp] not in ("ATOM", "HETATM"):
            continue
        if r[c_atom] != "CA" or (c_el is not None and r[c_el].upper() != "C"):
            continue
        if c_alt is not None and r[c_alt] not in (".", "?", "A"):
            continue
        keep.append(r)
    return AtomArray(
        np.array([[float(r[c_x]), float(r[c_y]), float(r[c_z])] for r in keep], dtype=np.float32).reshape(-1, 3),
        res_name=np.array([r[c_comp] for r in keep]),
        chain_id=np.array([r[c_chain] for r in keep]),
        res_id=np.array([int(r[c_seq]) for r in keep], dtype=int),
    )
