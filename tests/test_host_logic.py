"""Host-side mirror of the reference interface: constructors, properties and
error behaviour (no GPU needed).  Mirrors tests/test_forcefield.py of the
reference where the logic lives on the host."""
import numpy as np
import pytest

import springcraft_b200 as sc
from springcraft_b200 import forcefield as ffm
from .conftest import golden


@pytest.fixture
def atoms():
    """The reference's fixture (test_forcefield.py:14-30): two overlapping chains."""
    ref = golden("ref_two_chain.npz")
    return sc.AtomArray(ref["coord"], ref["res_name"], ref["chain_id"], ref["res_id"])


def test_flat_namespace():
    for name in ("ANM", "GNM", "ForceField", "PatchedForceField", "InvariantForceField", "HinsenForceField",
                 "ParameterFreeForceField", "TabulatedForceField", "compute_kirchhoff", "compute_hessian"):
        assert hasattr(sc, name)
    for name in ("eigen", "frequencies", "mean_square_fluctuation", "bfactor", "dcc", "normal_mode",
                 "linear_response", "prs", "effector_sensor"):
        assert hasattr(sc.nma, name)
    assert sc.__reference_version__ == "0.3.0"


def test_forcefield_interface_defaults():
    class Mine(sc.ForceField):
        def force_constant(self, atom_i, atom_j, sq_distance):
            return np.ones(len(atom_i))
    ff = Mine()
    assert ff.cutoff_distance is None and ff.contact_shutdown is None and ff.contact_pair_off is None
    assert ff.contact_pair_on is None and ff.natoms is None
    with pytest.raises(TypeError):
        sc.ForceField()          # abstract


def test_invariant_needs_cutoff():
    with pytest.raises(ValueError):
        sc.InvariantForceField(None)
    assert sc.InvariantForceField(7).cutoff_distance == 7
    assert sc.HinsenForceField().cutoff_distance is None
    assert sc.ParameterFreeForceField(9.5).cutoff_distance == 9.5


def test_tabulated_homogeneous_interaction_matrix(atoms):
    """test_forcefield.py:117-151."""
    ff = sc.TabulatedForceField(atoms, 1, 2, 3, None)
    assert ff.interaction_matrix.shape[2] == 1 and ff.interaction_matrix.dtype == np.float32
    m = ff.interaction_matrix[:, :, 0]
    assert np.allclose(m, m.T)
    n = len(atoms)
    for i in range(n):
        for j in range(i, n):
            if i == j:
                assert m[i, j] == 0
            elif j == i + 1 and atoms.chain_id[i] == atoms.chain_id[j]:
                assert m[i, j] == 1
            elif atoms.chain_id[i] == atoms.chain_id[j]:
                assert m[i, j] == 2
            else:
                assert m[i, j] == 3
    assert ff.natoms == 40 and ff.cutoff_distance is None
    assert ff.interaction_matrix is ff.interaction_matrix  # returned by reference


def test_tabulated_inhomogeneous_interaction_matrix(atoms):
    """test_forcefield.py:154-208."""
    mapping = np.array([ffm.AA_TO_INDEX[aa] for aa in atoms.res_name])
    np.random.seed(0)
    triu = np.triu(np.random.rand(3, 20, 20))
    bonded, intra, inter = triu + np.transpose(triu, (0, 2, 1))
    m = sc.TabulatedForceField(atoms, bonded, intra, inter, None).interaction_matrix[:, :, 0]
    n = len(atoms)
    for i in range(n):
        for j in range(i, n):
            if i == j:
                want = 0
            elif j == i + 1 and atoms.chain_id[i] == atoms.chain_id[j]:
                want = bonded[mapping[i], mapping[j]]
            elif atoms.chain_id[i] == atoms.chain_id[j]:
                want = intra[mapping[i], mapping[j]]
            else:
                want = inter[mapping[i], mapping[j]]
            assert m[i, j] == pytest.approx(want)


@pytest.mark.parametrize("key", ["e_anm", "sd_enm", "s_enm_10"])
def test_preset_interaction_matrix_equals_reference(atoms, key):
    """The lazily built (n,n,k) table is bit-identical to the reference's constructor output."""
    ref = golden("ref_two_chain.npz")
    ff = getattr(sc.TabulatedForceField, key)(atoms)
    assert np.array_equal(ff.interaction_matrix, ref[f"{key}/interaction_matrix"])
    assert ff.cutoff_distance == {"e_anm": 13.0, "sd_enm": 16.5, "s_enm_10": 10.0}[key]


@pytest.mark.parametrize(
    "shape, n_edges, is_valid",
    [[(), None, True], [(), 1, True], [(), 10, True], [(10,), None, False], [(10,), 1, False],
     [(9,), 10, False], [(10,), 10, True], [(1,), None, True], [(20, 1), 1, False], [(20, 30), 1, False],
     [(1, 20), 1, False], [(30, 20), 1, False], [(20, 20), 1, True], [(20, 20), None, True],
     [(20, 20), 10, True], [(20, 1, 10), 10, False], [(20, 30, 10), 10, False], [(1, 20, 10), 10, False],
     [(30, 20, 10), 10, False], [(20, 20, 10), 10, True], [(20, 20, 1), 1, True], [(20, 20, 1), None, True],
     [(20, 20, 10), 9, False]],
)
def test_tabulated_input_shapes(atoms, shape, n_edges, is_valid):
    """test_forcefield.py:277-320."""
    fc = np.ones(shape) if shape != () else 1
    edges = np.arange(n_edges) if n_edges is not None else None
    if is_valid:
        ff = sc.TabulatedForceField(atoms, fc, fc, fc, edges)
        assert ff.interaction_matrix.shape == (40, 40, n_edges if n_edges is not None else 1)
    else:
        with pytest.raises(IndexError):
            sc.TabulatedForceField(atoms, fc, fc, fc, edges)


@pytest.mark.parametrize("name", ["s_enm_10", "s_enm_13", "d_enm", "sd_enm", "e_anm", "e_anm_mj", "e_anm_ke"])
def test_presets_instantiate(atoms, name):
    ff = getattr(sc.TabulatedForceField, name)(atoms)
    assert ff.natoms == 40 and ff._bonded.dtype == np.float32


def test_tabulated_errors(atoms):
    with pytest.raises(TypeError):
        sc.TabulatedForceField(atoms.coord, 1, 1, 1, 7.0)
    bad = atoms.copy()
    bad.atom_name = np.array(["CB"] * 40)
    with pytest.raises(sc.BadStructureError):
        sc.TabulatedForceField(bad, 1, 1, 1, 7.0)
    with pytest.raises(ValueError):   # unsorted edges
        sc.TabulatedForceField(atoms, 1, 1, 1, np.array([5.0, 4.0]))
    asym = np.ones((20, 20))
    asym[0, 1] = 2
    with pytest.raises(ValueError):
        sc.TabulatedForceField(atoms, asym, 1, 1, 7.0)
    with pytest.raises(IndexError):
        sc.TabulatedForceField(atoms, np.nan, 1, 1, 7.0)


def test_patched_forcefield_checks(atoms):
    base = sc.TabulatedForceField(atoms, 1, 1, 1, 7.0)
    with pytest.raises(TypeError):
        sc.PatchedForceField(base, contact_pair_on=np.array([[0, 5]]))
    with pytest.raises(IndexError):
        sc.PatchedForceField(base, contact_pair_on=np.array([[0, 5]]), force_constants=np.array([1.0, 2.0]))
    with pytest.raises(IndexError):
        sc.PatchedForceField(base, contact_shutdown=np.array([40]))
    p1 = sc.PatchedForceField(base, contact_shutdown=np.array([1, 2]), contact_pair_off=np.array([[0, 1]]),
                              contact_pair_on=np.array([[0, 30]]), force_constants=np.array([4.0]))
    p2 = sc.PatchedForceField(p1, contact_shutdown=np.array([3]), contact_pair_off=np.array([[5, 6]]),
                              contact_pair_on=np.array([[7, 8]]), force_constants=np.array([2.0]))
    assert p2.natoms == 40 and p2.cutoff_distance == 7.0
    assert p2.contact_shutdown.tolist() == [3, 1, 2]
    assert p2.contact_pair_off.tolist() == [[5, 6], [0, 1]]
    assert p2.contact_pair_on.tolist() == [[7, 8], [0, 30]]


def test_enm_constructor_errors(atoms):
    ff = sc.InvariantForceField(7.0)
    with pytest.raises(IndexError):
        sc.ANM(atoms, ff, masses=np.ones(3))
    with pytest.raises(ValueError):
        sc.GNM(atoms, ff, masses=np.zeros(40))
    with pytest.raises(TypeError):
        sc.ANM(atoms.coord, ff, masses=True)
    anm = sc.ANM(atoms, ff, masses=True)
    assert anm.masses.shape == (40,) and np.all(anm.masses > 50)
    assert sc.ANM(atoms, ff).masses is None
    assert sc.GNM(atoms, ff, masses=np.arange(1, 41))._masses.dtype == float
    with pytest.raises(IndexError):
        anm.hessian = np.zeros((5, 5))
    with pytest.raises(IndexError):
        anm.covariance = np.zeros((5, 5))
    g = sc.GNM(atoms, ff)
    with pytest.raises(ValueError):
        g.kirchhoff = np.zeros((5, 5))
    with pytest.raises(ValueError):
        sc.nma.eigen("not an enm")
    with pytest.raises(ValueError):
        sc.nma.linear_response(g, np.zeros((40, 3)))
    with pytest.raises(ValueError):
        sc.nma.normal_mode(g, 6, 1.0, 4)
    with pytest.raises(ValueError):
        sc.nma.prs(g)


def test_user_set_matrix_is_returned_by_reference(atoms):
    anm = sc.ANM(atoms, sc.InvariantForceField(7.0))
    H = np.eye(120)
    anm.hessian = H
    assert anm.hessian is H and anm._covariance is None and not anm._has_model()
    C = np.eye(120) * 2
    anm.covariance = C
    assert anm.covariance is C and anm._matrix is None


def test_effector_sensor_matches_reference_formula():
    rng = np.random.default_rng(0)
    m = rng.random((7, 7))
    eff, sens = sc.nma.effector_sensor(m)
    w = 1 - np.eye(7)
    assert np.allclose(eff, (m * w).sum(1) / 6) and np.allclose(sens, (m * w).sum(0) / 6)


def test_atom_array_container():
    a = sc.AtomArray(np.zeros((3, 3)), res_name=["ALA", "GLY", "TRP"], chain_id=["A", "A", "B"], res_id=[1, 2, 1])
    assert a.array_length() == 3 and len(a[1:]) == 2 and len(a + a) == 6
    assert (a[a.chain_id == "A"].res_name == ["ALA", "GLY"]).all()
    with pytest.raises(ValueError):
        sc.AtomArray(np.zeros((3, 2)))
    with pytest.raises(IndexError):
        sc.AtomArray(np.zeros((3, 3)), res_name=["ALA"])


def test_read_pdb_ca(tmp_path):
    """Minimal CA reader (structure input is the step before the path)."""
    lines = [
        "MODEL        1",
        "ATOM      1  N   ASN A   1      -8.901   4.127  -0.555  1.00  0.00           N  ",
        "ATOM      2  CA  ASN A   1      -8.608   3.135  -1.618  1.00  0.00           C  ",
        "ATOM      3  CA  LEU A   2      -4.923   4.002  -2.452  1.00  0.00           C  ",
        "HETATM    4 CA    CA A 101       0.000   0.000   0.000  1.00  0.00          CA  ",
        "ATOM      5  CA BTYR B   3      -3.690   2.738  -5.833  1.00  0.00           C  ",
        "ATOM      6  CA  TYR B   3      -3.690   2.738  -5.833  1.00  0.00           C  ",
        "ENDMDL",
        "MODEL        2",
        "ATOM      1  CA  ASN A   1       0.000   0.000   0.000  1.00  0.00           C  ",
        "ENDMDL",
    ]
    path = tmp_path / "mini.pdb"
    path.write_text("\n".join(lines) + "\n")
    ca = sc.read_pdb_ca(str(path))
    assert len(ca) == 3                      # calcium ion and altloc B are skipped, model 1 only
    assert ca.res_name.tolist() == ["ASN", "LEU", "TYR"] and ca.chain_id.tolist() == ["A", "A", "B"]
    assert ca.res_id.tolist() == [1, 2, 3] and ca.coord.dtype == np.float32
    assert np.allclose(ca.coord[1], [-4.923, 4.002, -2.452])
    ff = sc.TabulatedForceField.e_anm(ca)
    assert ff.natoms == 3 and ff._bonded_next.tolist() == [1, 0, 0]
    # blank element columns: the alignment of the atom name tells C-alpha (" CA ") from calcium ("CA  ")
    blank = tmp_path / "blank.pdb"
    blank.write_text("\n".join(l[:76].rstrip() if l.startswith(("ATOM", "HETATM")) else l for l in lines) + "\n")
    assert len(sc.read_pdb_ca(str(blank))) == 3


def test_read_pdb_ca_models(tmp_path):
    """Multi-model PDB (NMR bundle / trajectory dump) -> (atoms, coords[models][n][3]) for enm_ensemble."""
    def model(k, shift, extra=""):
        return [f"MODEL     {k:4d}",
                f"ATOM      1  N   ASN A   1      -8.901   4.127  -0.555  1.00  0.00           N  ",
                f"ATOM      2  CA  ASN A   1    {-8.608 + shift:8.3f}   3.135  -1.618  1.00  0.00           C  ",
                f"ATOM      3  CA  LEU A   2    {-4.923 + shift:8.3f}   4.002  -2.452  1.00  0.00           C  ",
                f"ATOM      4  CA  TYR B   3    {-3.690 + shift:8.3f}   2.738  -5.833  1.00  0.00           C  "] + \
               ([extra] if extra else []) + ["ENDMDL"]
    path = tmp_path / "bundle.pdb"
    path.write_text("\n".join(model(1, 0.0) + model(2, 0.5) + model(3, 1.0)) + "\n")
    atoms, coords = sc.read_pdb_ca_models(str(path))
    assert len(atoms) == 3 and coords.shape == (3, 3, 3) and coords.dtype == np.float64
    assert np.allclose(coords[:, 0, 0], [-8.608, -8.108, -7.608]) and np.array_equal(coords[0], atoms.coord)
    # every model is rounded through float32 exactly like the single-model reader
    assert np.array_equal(coords[1], sc.read_pdb_ca(str(path), model=2).coord.astype(np.float64))
    assert atoms.chain_id.tolist() == ["A", "A", "B"]
    bad = tmp_path / "bad.pdb"
    extra = "ATOM      5  CA  GLY B   4       0.000   0.000   0.000  1.00  0.00           C  "
    bad.write_text("\n".join(model(1, 0.0) + model(2, 0.5, extra)) + "\n")
    with pytest.raises(sc.BadStructureError):
        sc.read_pdb_ca_models(str(bad))
    single = tmp_path / "single.pdb"
    single.write_text("\n".join(model(1, 0.0)[1:-1]) + "\n")          # no MODEL/ENDMDL records
    atoms1, coords1 = sc.read_pdb_ca_models(str(single))
    assert coords1.shape == (1, 3, 3) and len(atoms1) == 3


def test_read_cif_ca(tmp_path):
    """mmCIF (text) CA reader: quoted atom names, alternate locations, hetero calcium, second model."""
    cif = """data_MINI
#
loop_
_entity.id
_entity.type
1 polymer
#
loop_
_atom_site.group_PDB
_atom_site.id
_atom_site.type_symbol
_atom_site.label_atom_id
_atom_site.label_alt_id
_atom_site.label_comp_id
_atom_site.label_asym_id
_atom_site.label_seq_id
_atom_site.Cartn_x
_atom_site.Cartn_y
_atom_site.Cartn_z
_atom_site.auth_seq_id
_atom_site.auth_asym_id
_atom_site.pdbx_PDB_model_num
ATOM   1 N  N     . ASN A 1 -8.901 4.127 -0.555 11 X 1
ATOM   2 C  CA    . ASN A 1 -8.608 3.135 -1.618 11 X 1
ATOM   3 O  "O5'" . ASN A 1 -7.000 3.000 -1.000 11 X 1
ATOM   4 C  CA    A LEU A 2 -4.923 4.002 -2.452 12 X 1
ATOM   5 C  CA    B LEU A 2 -4.900 4.000 -2.400 12 X 1
HETATM 6 CA CA    . CA  B . 0.000 0.000 0.000 101 X 1
ATOM   7 C  CA    . TYR C 3 -3.690 2.738 -5.833 13 Y 1
ATOM   8 C  CA    . ASN A 1 0.000 0.000 0.000 11 X 2
#
loop_
_other.id
1
"""
    path = tmp_path / "mini.cif"
    path.write_text(cif)
    ca = sc.read_cif_ca(str(path))
    assert len(ca) == 3
    assert ca.res_name.tolist() == ["ASN", "LEU", "TYR"] and ca.chain_id.tolist() == ["X", "X", "Y"]
    assert ca.res_id.tolist() == [11, 12, 13] and ca.coord.dtype == np.float32
    assert np.allclose(ca.coord[1], [-4.923, 4.002, -2.452])
    assert len(sc.read_cif_ca(str(path), model=2)) == 1
    ff = sc.TabulatedForceField.e_anm(ca)
    assert ff.natoms == 3 and ff._bonded_next.tolist() == [1, 0, 0]
    empty = tmp_path / "empty.cif"
    empty.write_text("data_X\n#\n")
    with pytest.raises(sc.BadStructureError):
        sc.read_cif_ca(str(empty))


def test_ensemble_argument_checks_come_before_the_device():
    """Shape errors are the reference's ValueError (interaction.py:141-142) even without a GPU; a valid call
    without a device fails loudly (no CPU fallback)."""
    with pytest.raises(ValueError):
        sc.enm_ensemble(np.zeros((4, 10, 2)), sc.InvariantForceField(7.0))
    with pytest.raises(ValueError):
        sc.enm_ensemble(np.zeros((10, 3)), sc.InvariantForceField(7.0))

    class Stack:                       # stands in for biotite's AtomArrayStack
        coord = np.zeros((2, 10, 4), dtype=np.float32)
    with pytest.raises(ValueError):
        sc.enm_ensemble(Stack(), sc.InvariantForceField(7.0))
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            sc.enm_ensemble(np.zeros((2, 10, 3)), sc.InvariantForceField(7.0))


# --------------------------------------------------------------------------------------------------------------
# numerical core of the full-spectrum solver (springcraft_b200/csrc/stedc_core.cuh), run on the CPU
# --------------------------------------------------------------------------------------------------------------
def _stedc_host():
    """Compile tests/native/stedc_host.cpp (a single-threaded driver around the SAME deflation scan and secular
    root finder the CUDA kernels call) with g++.  Test infrastructure only: the product never loads it."""
    import ctypes, os, subprocess, tempfile
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(tempfile.gettempdir(), "scb_stedc_host_%d.so" % os.getuid())
    src = os.path.join(here, "native", "stedc_host.cpp")
    hdr = os.path.join(here, "..", "springcraft_b200", "csrc", "stedc_core.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    P = ctypes.POINTER(ctypes.c_double)
    lib.stedc_host.argtypes = [ctypes.c_int, P, P, P, P, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]

    def run(d, e, leaf=64):
        d = np.ascontiguousarray(d, float); e = np.ascontiguousarray(e, float)
        N = len(d)
        lam = np.zeros(N); Zt = np.zeros((N, N)); stats = (ctypes.c_longlong * 3)()
        lib.stedc_host(N, d.ctypes.data_as(P), e.ctypes.data_as(P), lam.ctypes.data_as(P), Zt.ctypes.data_as(P), leaf, stats)
        return lam, Zt, list(stats)
    return run


def _tridiag_cases():
    rng = np.random.default_rng(0)
    yield "random", rng.standard_normal(257), rng.standard_normal(256), 64
    yield "toeplitz", 2 * np.ones(300), -np.ones(299), 64
    W = np.abs(np.arange(-10, 11)).astype(float)
    yield "wilkinson21", W, np.ones(20), 8
    dg = np.concatenate([W] * 10)
    eg = np.concatenate([np.concatenate([np.ones(20), [1e-9]]) for _ in range(10)])[:-1]
    yield "glued_wilkinson", dg, eg, 16
    yield "zero_couplings", rng.standard_normal(300), np.where(rng.random(299) < 0.3, 0.0, rng.standard_normal(299)), 64
    yield "graded", 10.0 ** (-np.arange(200) / 10), 0.5 * 10.0 ** (-np.arange(199) / 10), 64
    yield "identity", np.ones(130), np.zeros(129), 64
    yield "negative_couplings", rng.standard_normal(150), -np.abs(rng.standard_normal(149)), 32


@pytest.mark.parametrize("case", list(_tridiag_cases()), ids=lambda c: c[0])
def test_stedc_core_matches_lapack(case):
    """Divide and conquer on symmetric tridiagonal matrices (Cuppen / Gu-Eisenstat; the method LAPACK dsyevd uses
    behind the reference's np.linalg.eigh, nma.py:61): eigenvalues, orthogonality and residuals at the 1e-13 level
    on random, Toeplitz, Wilkinson, glued-Wilkinson (eigenvalue pairs agreeing to 1e-9), decoupled, graded and
    identity matrices -- the deflation and root-finder code is the code the kernels run."""
    from scipy.linalg import eigh_tridiagonal
    name, d, e, leaf = case
    run = _stedc_host()
    lam, Zt, stats = run(d, e, leaf)
    N = len(d)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    want = eigh_tridiagonal(d, e, eigvals_only=True) if np.any(e) else np.sort(d)
    scale = max(np.abs(want).max(), 1e-300)
    assert np.abs(lam - want).max() <= 1e-13 * scale
    assert np.abs(Zt @ Zt.T - np.eye(N)).max() <= 1e-12
    assert np.abs(T @ Zt.T - Zt.T * lam).max() <= 1e-13 * scale
    assert stats[2] >= stats[0] >= 0


def test_stedc_core_anm_spectrum():
    """The tridiagonal form of an ANM Hessian (six-fold zero eigenvalue): D&C against the oracle's eigh."""
    from scipy.linalg import hessenberg
    from oracle import enm_oracle as orc
    from synthetic_inputs import synthetic_chain
    H, _ = orc.compute_hessian(synthetic_chain(90, seed=3), orc.FFSpec(kind="invariant", cutoff=13.0))
    Hh = hessenberg(H)
    lam, Zt, _ = _stedc_host()(np.diag(Hh).copy(), np.diag(Hh, -1).copy())
    want = np.linalg.eigvalsh(H)
    assert np.abs(lam - want).max() <= 1e-13 * want.max()
    assert np.abs(lam[:6]).max() <= 1e-12 * want.max()
    assert np.abs(Zt @ Zt.T - np.eye(len(lam))).max() <= 1e-12
