"""Stub of biotite.structure.info: average residue masses (only `masses=True`)."""
_RES_MASS = {
    "ALA": 89.09, "ARG": 175.21, "ASN": 132.12, "ASP": 133.10, "CYS": 121.16,
    "GLN": 146.15, "GLU": 147.13, "GLY": 75.07, "HIS": 155.16, "ILE": 131.17,
    "LEU": 131.17, "LYS": 147.20, "MET": 149.21, "PHE": 165.19, "PRO": 115.13,
    "SER": 105.09, "THR": 119.12, "TRP": 204.23, "TYR": 181.19, "VAL": 117.15,
}


def mass(item, is_residue=None):
    return _RES_MASS[item]
