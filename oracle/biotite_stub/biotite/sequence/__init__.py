"""Stub of biotite.sequence: only the ProteinSequence alphabet order."""


class _Alphabet:
    _SYMBOLS = list("ACDEFGHIKLMNPQRSTVWYBZX*")

    def get_symbols(self):
        return list(self._SYMBOLS)


class ProteinSequence:
    alphabet = _Alphabet()
    _1TO3 = {
        "A": "ALA", "C": "CYS", "D": "ASP", "E": "GLU", "F": "PHE", "G": "GLY",
        "H": "HIS", "I": "ILE", "K": "LYS", "L": "LEU", "M": "MET", "N": "ASN",
        "P": "PRO", "Q": "GLN", "R": "ARG", "S": "SER", "T": "THR", "V": "VAL",
        "W": "TRP", "Y": "TYR", "B": "ASX", "Z": "GLX", "X": "UNK", "*": "TER",
    }

    @staticmethod
    def convert_letter_1to3(symbol):
        return ProteinSequence._1TO3[symbol.upper()]
