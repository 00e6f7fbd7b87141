"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONPATH=oracle/biotite_stub:/root/reference/src python tests/golden/make_golden.py

Two kinds of vectors are written:
  ref_*.npz         outputs of the reference itself (springcraft 0.3.0 from
                    /root/reference/src, imported through the test-only biotite
                    stub) on inputs stored next to them;
  thirdparty_*.npz  the ProDy / Bio3D / BioPhysConnectoR golden vectors the
                    reference's own tests compare against (read in place from
                    /root/reference/tests/data, re-packed, values unchanged).
The GPU box has no /root/reference, so everything parity needs lives here.
"""
import glob
import gzip
import os
import sys
from os.path import basename, dirname, join, realpath

import numpy as np

HERE = dirname(realpath(__file__))
ROOT = dirname(dirname(HERE))
sys.path.insert(0, ROOT)
REF_DATA = "/root/reference/tests/data"

import biotite.structure as struc  # noqa: E402  (stub)
import biotite.structure.io.pdb as pdb  # noqa: E402
import springcraft  # noqa: E402  (the reference)

from oracle import enm_oracle as orc  # noqa: E402

assert springcraft.__file__.startswith("/root/reference"), springcraft.__file__


def load_ca(name):
    atoms = pdb.get_structure(pdb.PDBFile.read(join(REF_DATA, name)), model=1)
    return atoms[(atoms.atom_name == "CA") & (atoms.element == "C")]


def make_atoms(coord, res_name, chain_id, res_id):
    a = struc.AtomArray(len(coord))
    a.coord = np.asarray(coord, dtype=np.float32)
    a.res_name = np.asarray(res_name)
    a.chain_id = np.asarray(chain_id)
    a.res_id = np.asarray(res_id)
    a.atom_name[:] = "CA"
    a.element[:] = "C"
    return a


def read_csv(name, **kw):
    path = join(REF_DATA, name)
    if not os.path.exists(path) and os.path.exists(path + ".gz"):
        path += ".gz"
    return np.genfromtxt(path, delimiter=",", **kw)


def save(name, **arrays):
    path = join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(arrays)} arrays")


FF_BUILDERS = {
    "invariant7": lambda ca: springcraft.InvariantForceField(7.0),
    "invariant13": lambda ca: springcraft.InvariantForceField(13.0),
    "hinsen": lambda ca: springcraft.HinsenForceField(),
    "hinsen_cut12": lambda ca: springcraft.HinsenForceField(12.0),
    "pfree": lambda ca: springcraft.ParameterFreeForceField(),
    "pfree_cut10": lambda ca: springcraft.ParameterFreeForceField(10.0),
    "e_anm": lambda ca: springcraft.TabulatedForceField.e_anm(ca),
    "e_anm_mean": lambda ca: springcraft.TabulatedForceField.e_anm(ca, nonbonded_mean=True),
    "e_anm_mj": lambda ca: springcraft.TabulatedForceField.e_anm_mj(ca),
    "e_anm_ke": lambda ca: springcraft.TabulatedForceField.e_anm_ke(ca),
    "sd_enm": lambda ca: springcraft.TabulatedForceField.sd_enm(ca),
    "d_enm": lambda ca: springcraft.TabulatedForceField.d_enm(ca),
    "s_enm_10": lambda ca: springcraft.TabulatedForceField.s_enm_10(ca),
    "s_enm_13": lambda ca: springcraft.TabulatedForceField.s_enm_13(ca),
}


def structures():
    out = {}
    for name in ("1l2y", "7cal"):
        ca = load_ca(name + ".pdb")
        out[f"{name}_coord"] = ca.coord
        out[f"{name}_res_name"] = ca.res_name
        out[f"{name}_chain_id"] = ca.chain_id
        out[f"{name}_res_id"] = ca.res_id
    save("structures.npz", **out)


def ref_1l2y():
    ca = load_ca("1l2y.pdb")
    masses = read_csv("bio3d_mass_1l2y.csv.gz")
    rng = np.random.default_rng(7)
    force_rand = rng.normal(size=(20, 3))
    force_unit = np.zeros((20, 3))
    force_unit[3, 0] = 1.0
    out = {"masses": masses, "force_rand": force_rand, "force_unit": force_unit}
    for key, build in FF_BUILDERS.items():
        ff = build(ca)
        H, pairs = springcraft.compute_hessian(ca.coord, ff)
        K, _ = springcraft.compute_kirchhoff(ca.coord, ff)
        out[f"{key}/pairs"] = pairs
        out[f"{key}/hessian"] = H
        out[f"{key}/kirchhoff"] = K
        anm = springcraft.ANM(ca, ff)
        lam, vec = anm.eigen()
        out[f"{key}/anm_eigval"] = lam
        out[f"{key}/anm_eigvec"] = vec
        out[f"{key}/anm_freq"] = anm.frequencies()
        out[f"{key}/anm_msf"] = anm.mean_square_fluctuation()
        out[f"{key}/anm_msf_sub"] = anm.mean_square_fluctuation(mode_subset=np.arange(6, 26))
        out[f"{key}/anm_msf_tem"] = anm.mean_square_fluctuation(tem=300, tem_factors=orc.K_B * orc.N_A)
        out[f"{key}/anm_bfactor"] = anm.bfactor()
        out[f"{key}/anm_cov"] = anm.covariance
        out[f"{key}/anm_dcc"] = anm.dcc()
        out[f"{key}/anm_dcc_abs"] = anm.dcc(norm=False)
        out[f"{key}/anm_dcc_sub"] = anm.dcc(mode_subset=np.arange(6, 36))
        out[f"{key}/anm_dcc_sub_tem"] = anm.dcc(mode_subset=np.arange(6, 36), norm=False, tem=300)
        out[f"{key}/anm_lr_unit"] = anm.linear_response(force_unit)
        out[f"{key}/anm_lr_rand"] = anm.linear_response(force_rand.flatten())
        prs, eff, sens = anm.prs_effector_sensor()
        out[f"{key}/anm_prs"] = prs
        out[f"{key}/anm_eff"] = eff
        out[f"{key}/anm_sens"] = sens
        out[f"{key}/anm_normal_mode"] = anm.normal_mode(6, 5.0, 8)
        out[f"{key}/anm_normal_mode_tri"] = anm.normal_mode(7, 2.0, 6, movement="triangle")
        mw = springcraft.ANM(ca, ff, masses=masses)
        out[f"{key}/mw_hessian"] = mw.hessian
        lam, _ = mw.eigen()
        out[f"{key}/mw_eigval"] = lam
        out[f"{key}/mw_freq"] = mw.frequencies()
        out[f"{key}/mw_msf"] = mw.mean_square_fluctuation(tem=300, tem_factors=orc.K_B * orc.N_A)
        out[f"{key}/mw_dcc"] = mw.dcc()
        gnm = springcraft.GNM(ca, ff)
        lam, vec = gnm.eigen()
        out[f"{key}/gnm_eigval"] = lam
        out[f"{key}/gnm_eigvec"] = vec
        out[f"{key}/gnm_freq"] = gnm.frequencies()
        out[f"{key}/gnm_msf"] = gnm.mean_square_fluctuation()
        out[f"{key}/gnm_bfactor"] = gnm.bfactor(tem=300)
        out[f"{key}/gnm_cov"] = gnm.covariance
        out[f"{key}/gnm_dcc"] = gnm.dcc()
        out[f"{key}/gnm_dcc_abs"] = gnm.dcc(norm=False)
        out[f"{key}/gnm_dcc_sub"] = gnm.dcc(mode_subset=np.arange(1, 17))
        gmw = springcraft.GNM(ca, ff, masses=masses)
        out[f"{key}/gnm_mw_kirchhoff"] = gmw.kirchhoff
    save("ref_1l2y.npz", **out)


def ref_two_chain():
    """The reference's `atoms` fixture (test_forcefield.py:14-30): 1L2Y CA
    duplicated into two perfectly overlapping chains A/B + PatchedForceField."""
    ca = load_ca("1l2y.pdb")
    cb = ca.copy()
    ca.chain_id[:] = "A"
    cb.chain_id[:] = "B"
    atoms = ca + cb
    out = {"coord": atoms.coord, "res_name": atoms.res_name,
           "chain_id": atoms.chain_id, "res_id": atoms.res_id}
    base = springcraft.InvariantForceField(7.0)
    out["invariant7/kirchhoff"], out["invariant7/pairs"] = springcraft.compute_kirchhoff(atoms.coord, base)
    np.random.seed(0)
    shutdown = np.random.choice(np.arange(len(atoms)), size=5, replace=False)
    pair_off = np.array([[0, 1], [5, 9], [3, 23], [10, 30], [17, 16]])
    pair_on = np.array([[0, 19], [2, 39], [7, 27], [21, 35]])
    fcs = np.array([3.5, 0.25, 82.0, 1.0])
    out["shutdown"], out["pair_off"], out["pair_on"], out["pair_on_fc"] = shutdown, pair_off, pair_on, fcs
    for tag, kw in {
        "shutdown": dict(contact_shutdown=shutdown),
        "pair_off": dict(contact_pair_off=pair_off),
        "pair_on": dict(contact_pair_on=pair_on, force_constants=fcs),
        "all": dict(contact_shutdown=shutdown, contact_pair_off=pair_off,
                    contact_pair_on=pair_on, force_constants=fcs),
    }.items():
        ff = springcraft.PatchedForceField(base, **kw)
        out[f"patched_{tag}/kirchhoff"], out[f"patched_{tag}/pairs"] = springcraft.compute_kirchhoff(atoms.coord, ff)
    for key in ("e_anm", "sd_enm", "d_enm", "s_enm_13"):
        ff = FF_BUILDERS[key](atoms)
        out[f"{key}/kirchhoff"], out[f"{key}/pairs"] = springcraft.compute_kirchhoff(atoms.coord, ff)
        pf = springcraft.PatchedForceField(ff, contact_pair_off=pair_off, contact_pair_on=pair_on,
                                           force_constants=fcs)
        out[f"{key}_patched/kirchhoff"], _ = springcraft.compute_kirchhoff(atoms.coord, pf)
    for key in ("e_anm", "sd_enm", "s_enm_10"):
        out[f"{key}/interaction_matrix"] = FF_BUILDERS[key](atoms).interaction_matrix
    # shifted second chain so that Hessians are finite (no zero distances)
    shifted = atoms.copy()
    shifted.coord[20:] += np.array([4.0, 3.0, -2.5], dtype=np.float32)
    out["shifted_coord"] = shifted.coord
    for key in ("e_anm", "sd_enm", "hinsen", "invariant13"):
        ff = FF_BUILDERS[key](shifted)
        out[f"shifted_{key}/hessian"], out[f"shifted_{key}/pairs"] = springcraft.compute_hessian(shifted.coord, ff)
        pf = springcraft.PatchedForceField(ff, contact_shutdown=shutdown, contact_pair_off=pair_off,
                                           contact_pair_on=pair_on, force_constants=fcs)
        out[f"shifted_{key}_patched/hessian"], out[f"shifted_{key}_patched/pairs"] = \
            springcraft.compute_hessian(shifted.coord, pf)
    save("ref_two_chain.npz", **out)


def thirdparty():
    out = {}
    for path in sorted(glob.glob(join(REF_DATA, "*1l2y.csv.gz"))):
        key = basename(path)[: -len(".csv.gz")]
        skip = 1 if key.startswith("biophysconnector") else 0
        out[key] = np.genfromtxt(path, delimiter=",", skip_header=skip)
    save("thirdparty_1l2y.npz", **out)
    out = {}
    for path in sorted(glob.glob(join(REF_DATA, "*7cal.csv.gz"))):
        key = basename(path)[: -len(".csv.gz")]
        skip = 1 if key.startswith("biophysconnector") else 0
        arr = np.genfromtxt(path, delimiter=",", skip_header=skip)
        if arr.size > 200_000:
            continue
        out[key] = arr
    save("thirdparty_7cal.npz", **out)
    # 500-point random cloud (the four "seeds" hold identical coordinates)
    coord = read_csv("random_coord_seed_1.csv.gz")
    for s in (323, 777, 999):
        assert np.array_equal(coord, read_csv(f"random_coord_seed_{s}.csv.gz"))
    out = {"coord": coord}
    for cutoff in (5, 10, 15):
        K = read_csv(f"prody_gnm_{cutoff}_ang_cutoff_kirchhoff_random_coords_seed_1.csv.gz")
        assert np.array_equal(K, np.round(K))
        out[f"prody_kirchhoff_{cutoff}"] = K.astype(np.int16)
    H = read_csv("prody_anm_10_ang_cutoff_hessian_random_coords_seed_1.csv.gz")
    blocks = H.reshape(500, 3, 500, 3).transpose(0, 2, 1, 3)
    nz = np.abs(blocks).max(axis=(2, 3)) > 0
    bi, bj = np.nonzero(nz)
    out["prody_hessian_10_bi"] = bi.astype(np.int32)
    out["prody_hessian_10_bj"] = bj.astype(np.int32)
    out["prody_hessian_10_blocks"] = blocks[bi, bj]
    save("thirdparty_random500.npz", **out)


def ref_random500():
    coord = read_csv("random_coord_seed_1.csv.gz")
    out = {}
    for cutoff in (5, 10, 15):
        ff = springcraft.InvariantForceField(cutoff)
        for ucl in (False, True):
            K, pairs = springcraft.compute_kirchhoff(coord, ff, ucl)
            out[f"pairs_{cutoff}_{int(ucl)}"] = pairs.astype(np.int32)
        H, _ = springcraft.compute_hessian(coord, ff, False)
        b = H.reshape(500, 3, 500, 3).transpose(0, 2, 1, 3)
        out[f"hessian_offdiag_{cutoff}"] = b[pairs[:, 0], pairs[:, 1]]
        out[f"hessian_diag_{cutoff}"] = b[np.arange(500), np.arange(500)]
    save("ref_random500.npz", **out)


def ref_7cal():
    ca = load_ca("7cal.pdb")
    masses = read_csv("bio3d_mass_7cal.csv.gz")
    out = {"masses": masses}
    for key in ("invariant13", "e_anm", "sd_enm"):
        ff = FF_BUILDERS[key](ca)
        anm = springcraft.ANM(ca, ff)
        H = anm.hessian
        pairs = springcraft.compute_hessian(ca.coord, ff)[1]
        out[f"{key}/n_pairs"] = np.array(len(pairs))
        out[f"{key}/hessian_rowsum_abs"] = np.abs(H).sum(axis=1)
        out[f"{key}/hessian_diag"] = np.diagonal(H).copy()
        lam, vec = np.linalg.eigh(H)
        out[f"{key}/eigval"] = lam
        out[f"{key}/modes_6_26"] = vec.T[6:26]
        sub = np.arange(6, 26)
        sq = np.square(vec.T[sub]).reshape(20, -1, 3).sum(-1)
        out[f"{key}/msf_6_26"] = (sq / lam[sub][:, None]).sum(0)
        sq = np.square(vec.T[6:]).reshape(len(lam) - 6, -1, 3).sum(-1)
        out[f"{key}/msf_all"] = (sq / lam[6:][:, None]).sum(0)
    save("ref_7cal.npz", **out)


def ref_synthetic():
    # C2: chain n=1000, Hinsen all pairs, full spectrum + MSF + B-factor
    n = 1000
    coord = orc.synthetic_chain(n, seed=0)
    anm = springcraft.ANM(coord, springcraft.HinsenForceField())
    lam, vec = np.linalg.eigh(anm.hessian)
    sq = np.square(vec.T[6:]).reshape(3 * n - 6, n, 3).sum(-1)
    msf = (sq / lam[6:][:, None]).sum(0)
    save("ref_c2_chain1000.npz", coord=coord, eigval=lam, msf=msf,
         bfactor=8 * np.pi ** 2 * msf / 3, hessian_diag=np.diagonal(anm.hessian).copy(),
         modes_6_26=vec.T[6:26])
    # C3: chain n=300 + perturbed conformations, e_anm / sd_enm
    n = 300
    base = orc.synthetic_chain(n, seed=0)
    res_name, chain_id, res_id = orc.synthetic_sequence(n, seed=0)
    out = {"base": base, "res_name": res_name, "chain_id": chain_id, "res_id": res_id}
    for c in (0, 1, 2, 4095):
        coord = orc.perturbed_conformation(base, c)
        atoms = make_atoms(coord, res_name, chain_id, res_id)
        for key in ("e_anm", "sd_enm"):
            ff = FF_BUILDERS[key](atoms)
            # NB: the ensemble is defined on the fp64 coordinates
            H, pairs = springcraft.compute_hessian(coord, ff)
            lam, vec = np.linalg.eigh(H)
            sub = np.arange(6, 26)
            sq = np.square(vec.T[sub]).reshape(20, n, 3).sum(-1)
            out[f"c{c}/{key}/n_pairs"] = np.array(len(pairs))
            out[f"c{c}/{key}/eigval"] = lam[:60]
            out[f"c{c}/{key}/modes_6_26"] = vec.T[6:26]
            out[f"c{c}/{key}/msf_6_26"] = (sq / lam[sub][:, None]).sum(0)
            if c == 0:
                out[f"c{c}/{key}/hessian"] = H.astype(np.float64)
    save("ref_c3_chain300.npz", **out)
    # C4 (scaled down): cloud, ParameterFree all pairs, lowest 100 modes
    n = 400
    coord = orc.synthetic_cloud(n, seed=0)
    H, _ = springcraft.compute_hessian(coord, springcraft.ParameterFreeForceField())
    lam, vec = np.linalg.eigh(H)
    sub = np.arange(6, 106)
    sq = np.square(vec.T[sub]).reshape(100, n, 3).sum(-1)
    save("ref_c4_cloud400.npz", coord=coord, eigval=lam[:130], modes_6_106=vec.T[6:106],
         msf_6_106=(sq / lam[sub][:, None]).sum(0), hessian_diag=np.diagonal(H).copy())
    ref_c5()


def ref_c5():
    """C5 (scaled down): DCC + linear response, full covariance and a mode subset ("from m modes")."""
    n = 400
    coord = orc.synthetic_chain(n, seed=3)
    anm = springcraft.ANM(coord, springcraft.InvariantForceField(13.0))
    gnm = springcraft.GNM(coord, springcraft.InvariantForceField(10.0))
    rng = np.random.default_rng(11)
    f = rng.normal(size=(n, 3))
    # the reference has no truncated linear response (nma.py:473 multiplies with the full pseudo-inverse); the
    # low-rank form V_S^T (L_S^-1 (V_S f)) is evaluated here from the reference's own eigen() output
    lam, vec = anm.eigen()
    sub = np.arange(6, 56)
    lr_sub = (vec[sub].T @ ((vec[sub] @ f.flatten()) / lam[sub])).reshape(n, 3)
    unit = np.zeros((n, 3))
    unit[42, 0] = 1.0                       # doc/index.rst:139-142
    save("ref_c5_chain400.npz", coord=coord, force=f,
         anm_dcc_sub=anm.dcc(mode_subset=np.arange(6, 56)),
         anm_dcc_sub_abs=anm.dcc(mode_subset=np.arange(6, 56), norm=False),
         anm_dcc_all=anm.dcc(),
         gnm_dcc_sub=gnm.dcc(mode_subset=np.arange(1, 51)),
         gnm_dcc_all_abs=gnm.dcc(norm=False),
         anm_lr=anm.linear_response(f),
         anm_lr_unit42=anm.linear_response(unit),
         anm_lr_sub=lr_sub,
         anm_msf=anm.mean_square_fluctuation(),
         gnm_msf=gnm.mean_square_fluctuation())


def ref_dense_table():
    """TabulatedForceField whose interaction_matrix was edited in place by the user (forcefield.py:429-434):
    the edited (n,n,k) float32 table is what force_constant() reads."""
    ca = load_ca("1l2y.pdb")
    out = {}
    for key in ("e_anm", "sd_enm"):
        ff = FF_BUILDERS[key](ca)
        M = ff.interaction_matrix
        rng = np.random.default_rng(5)
        for _ in range(12):
            i, j = rng.integers(0, len(ca), size=2)
            if i == j:
                continue
            fac = np.float32(rng.uniform(0.25, 3.0))
            M[i, j] *= fac
            M[j, i] = M[i, j]
        out[f"{key}/table"] = M.copy()
        out[f"{key}/hessian"], out[f"{key}/pairs"] = springcraft.compute_hessian(ca.coord, ff)
        out[f"{key}/kirchhoff"], _ = springcraft.compute_kirchhoff(ca.coord, ff)
    save("ref_dense_table.npz", **out)


FUZZ_KINDS = ["invariant", "hinsen", "hinsen_nocut", "pfree", "pfree_nocut", "e_anm", "e_anm_mean", "e_anm_mj",
              "e_anm_ke", "sd_enm", "d_enm", "s_enm_10", "s_enm_13"]
AA3 = ["ALA", "CYS", "ASP", "GLU", "PHE", "GLY", "HIS", "ILE", "LYS", "LEU",
       "MET", "ASN", "PRO", "GLN", "ARG", "SER", "THR", "VAL", "TRP", "TYR"]


def fuzz_force_field(kind, cutoff, atoms):
    if kind == "invariant":
        return springcraft.InvariantForceField(cutoff)
    if kind == "hinsen":
        return springcraft.HinsenForceField(cutoff)
    if kind == "hinsen_nocut":
        return springcraft.HinsenForceField()
    if kind == "pfree":
        return springcraft.ParameterFreeForceField(cutoff)
    if kind == "pfree_nocut":
        return springcraft.ParameterFreeForceField()
    if kind == "e_anm_mean":
        return springcraft.TabulatedForceField.e_anm(atoms, nonbonded_mean=True)
    return getattr(springcraft.TabulatedForceField, kind)(atoms)


def ref_fuzz():
    """Seeded random small cases (chains with breaks, several chain ids, gaps in the residue numbering, random
    sequences, every force-field kind, random patches and masses) through the unmodified reference."""
    rng = np.random.default_rng(20260)
    out = {"n_cases": np.array(26)}
    for i in range(26):
        n = int(rng.integers(10, 41))
        coord = orc.synthetic_chain(n, seed=100 + i, jitter=0.4).astype(np.float32)
        res_name = rng.choice(AA3, size=n)
        nchain = int(rng.integers(1, 4))
        cuts = np.sort(rng.choice(np.arange(2, n - 1), size=nchain - 1, replace=False)) if nchain > 1 else []
        chain_id = np.array(["A"] * n)
        for c, start in enumerate(cuts):
            chain_id[start:] = "ABC"[c + 1]
        res_id = np.arange(1, n + 1)
        for g in rng.choice(np.arange(1, n), size=int(rng.integers(0, 3)), replace=False):
            res_id[g:] += int(rng.integers(1, 4))          # numbering gaps: no bonded constant across them
        atoms = make_atoms(coord, res_name, chain_id, res_id)
        kind = FUZZ_KINDS[i % len(FUZZ_KINDS)]
        cutoff = float(np.round(rng.uniform(6.5, 14.0), 2))
        ff = fuzz_force_field(kind, cutoff, atoms)
        patched = bool(i % 2)
        shutdown = pair_off = pair_on = fcs = None
        if patched:
            shutdown = rng.choice(n, size=int(rng.integers(0, 3)), replace=False)
            pair_off = np.array([rng.choice(n, size=2, replace=False) for _ in range(2)])
            pair_on = np.array([rng.choice(n, size=2, replace=False) for _ in range(2)])
            fcs = np.round(rng.uniform(0.2, 9.0, size=2), 3)
            ff = springcraft.PatchedForceField(ff, contact_shutdown=shutdown if len(shutdown) else None,
                                               contact_pair_off=pair_off, contact_pair_on=pair_on, force_constants=fcs)
        masses = rng.uniform(60.0, 200.0, size=n) if i % 3 == 0 else None
        pre = f"case{i}/"
        out[pre + "coord"], out[pre + "res_name"], out[pre + "chain_id"], out[pre + "res_id"] = \
            coord, res_name, chain_id, res_id
        out[pre + "kind"], out[pre + "cutoff"], out[pre + "patched"] = np.array(kind), np.array(cutoff), np.array(patched)
        if patched:
            out[pre + "shutdown"], out[pre + "pair_off"], out[pre + "pair_on"], out[pre + "pair_on_fc"] = \
                shutdown, pair_off, pair_on, fcs
        if masses is not None:
            out[pre + "masses"] = masses
        H, pairs = springcraft.compute_hessian(atoms.coord, ff)
        K, _ = springcraft.compute_kirchhoff(atoms.coord, ff)
        out[pre + "pairs"], out[pre + "hessian"], out[pre + "kirchhoff"] = pairs, H, K
        anm = springcraft.ANM(atoms, ff, masses=masses)
        gnm = springcraft.GNM(atoms, ff, masses=masses)
        out[pre + "anm_matrix"], out[pre + "gnm_matrix"] = anm.hessian, gnm.kirchhoff
        out[pre + "anm_eigval"], out[pre + "gnm_eigval"] = anm.eigen()[0], gnm.eigen()[0]
        out[pre + "anm_msf"], out[pre + "gnm_msf"] = anm.mean_square_fluctuation(), gnm.mean_square_fluctuation()
    save("ref_fuzz.npz", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["structures", "ref_1l2y", "ref_two_chain", "thirdparty",
                             "ref_random500", "ref_7cal", "ref_synthetic", "ref_fuzz", "ref_dense_table"]
    for w in which:
        globals()[w]()
