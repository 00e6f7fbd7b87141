"""Stub of biotite.structure.io.pdb: fixed-column ATOM/HETATM reader, model 1."""
import numpy as np
from ... import AtomArray


class PDBFile:
    def __init__(self):
        self.lines = []

    @classmethod
    def read(cls, path):
        f = cls()
        with open(path) as fh:
            f.lines = fh.read().splitlines()
        return f


def get_structure(pdb_file, model=1):
    rows = []
    current = 1
    for line in pdb_file.lines:
        rec = line[:6]
        if rec.startswith("MODEL"):
            current = int(line[10:14])
        elif rec.startswith("ENDMDL"):
            if current == model:
                break
        elif rec in ("ATOM  ", "HETATM") and current == model:
            altloc = line[16]
            if altloc not in (" ", "A"):
                continue
            rows.append(line)
    arr = AtomArray(len(rows))
    for k, line in enumerate(rows):
        arr.hetero[k] = line.startswith("HETATM")
        arr.atom_name[k] = line[12:16].strip()
        arr.res_name[k] = line[17:20].strip()
        arr.chain_id[k] = line[21].strip()
        arr.res_id[k] = int(line[22:26])
        arr.coord[k] = (float(line[30:38]), float(line[38:46]), float(line[46:54]))
        el = line[76:78].strip() if len(line) >= 78 else ""
        arr.element[k] = el.upper()
    return arr
