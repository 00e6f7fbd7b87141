"""Pin the oracle: oracle/enm_oracle.py vs (i) the unmodified reference's
outputs (tests/golden/ref_*.npz) and (ii) the third-party golden vectors the
reference's own tests use (tests/golden/thirdparty_*.npz)."""
import numpy as np
import pytest

from oracle import enm_oracle as orc
from .conftest import golden

KINDS = {
    "invariant7": lambda s: orc.FFSpec("invariant", 7.0),
    "invariant13": lambda s: orc.FFSpec("invariant", 13.0),
    "hinsen": lambda s: orc.FFSpec("hinsen", None),
    "hinsen_cut12": lambda s: orc.FFSpec("hinsen", 12.0),
    "pfree": lambda s: orc.FFSpec("pfree", None),
    "pfree_cut10": lambda s: orc.FFSpec("pfree", 10.0),
    "e_anm": lambda s: orc.preset_spec("e_anm", *s),
    "e_anm_mean": lambda s: orc.preset_spec("e_anm", *s, nonbonded_mean=True),
    "e_anm_mj": lambda s: orc.preset_spec("e_anm_mj", *s),
    "e_anm_ke": lambda s: orc.preset_spec("e_anm_ke", *s),
    "sd_enm": lambda s: orc.preset_spec("sd_enm", *s),
    "d_enm": lambda s: orc.preset_spec("d_enm", *s),
    "s_enm_10": lambda s: orc.preset_spec("s_enm_10", *s),
    "s_enm_13": lambda s: orc.preset_spec("s_enm_13", *s),
}


def seq(st, name):
    return st[f"{name}_res_name"], st[f"{name}_chain_id"], st[f"{name}_res_id"]


@pytest.mark.parametrize("key", sorted(KINDS))
def test_1l2y_assembly_bit_exact(structures, key):
    ref = golden("ref_1l2y.npz")
    coord = structures["1l2y_coord"]
    spec = KINDS[key](seq(structures, "1l2y"))
    H, pairs = orc.compute_hessian(coord, spec)
    K, _ = orc.compute_kirchhoff(coord, spec)
    assert np.array_equal(pairs, ref[f"{key}/pairs"])
    assert np.array_equal(K, ref[f"{key}/kirchhoff"])
    assert np.array_equal(H, ref[f"{key}/hessian"])
    assert np.array_equal(orc.mass_weight(H, ref["masses"], 3), ref[f"{key}/mw_hessian"])
    assert np.array_equal(orc.mass_weight(K, ref["masses"], 1), ref[f"{key}/gnm_mw_kirchhoff"])


@pytest.mark.parametrize("key", ["invariant13", "hinsen", "e_anm", "sd_enm", "pfree"])
def test_1l2y_nma(structures, key):
    ref = golden("ref_1l2y.npz")
    H = ref[f"{key}/hessian"]
    lam, modes = orc.eigen(H)
    assert np.array_equal(lam, ref[f"{key}/anm_eigval"])
    assert np.allclose(orc.frequencies(lam, 6)[6:], ref[f"{key}/anm_freq"][6:], rtol=1e-14)
    assert np.allclose(orc.mean_square_fluctuation(lam, modes, 3), ref[f"{key}/anm_msf"], rtol=1e-12)
    assert np.allclose(orc.mean_square_fluctuation(lam, modes, 3, mode_subset=np.arange(6, 26)),
                       ref[f"{key}/anm_msf_sub"], rtol=1e-12)
    assert np.allclose(orc.mean_square_fluctuation(lam, modes, 3, tem=300, tem_factors=orc.K_B * orc.N_A),
                       ref[f"{key}/anm_msf_tem"], rtol=1e-12)
    assert np.allclose(orc.bfactor(lam, modes, 3), ref[f"{key}/anm_bfactor"], rtol=1e-12)
    cov = orc.covariance(H)
    assert np.allclose(cov, ref[f"{key}/anm_cov"], rtol=1e-9, atol=1e-12 * np.abs(cov).max())
    assert np.allclose(orc.dcc(lam, modes, 3, cov=cov), ref[f"{key}/anm_dcc"], atol=1e-10)
    assert np.allclose(orc.dcc(lam, modes, 3, cov=cov, norm=False), ref[f"{key}/anm_dcc_abs"],
                       rtol=1e-9, atol=1e-12 * np.abs(cov).max())
    assert np.allclose(orc.dcc(lam, modes, 3, mode_subset=np.arange(6, 36)), ref[f"{key}/anm_dcc_sub"], atol=1e-12)
    assert np.allclose(orc.dcc(lam, modes, 3, mode_subset=np.arange(6, 36), norm=False, tem=300),
                       ref[f"{key}/anm_dcc_sub_tem"], rtol=1e-10)
    assert np.allclose(orc.linear_response(cov, ref["force_unit"]), ref[f"{key}/anm_lr_unit"],
                       rtol=1e-8, atol=1e-12 * np.abs(cov).max())
    assert np.allclose(orc.linear_response(cov, ref["force_rand"]), ref[f"{key}/anm_lr_rand"],
                       rtol=1e-8, atol=1e-11 * np.abs(cov).max())
    p = orc.prs(cov)
    assert np.allclose(p, ref[f"{key}/anm_prs"], rtol=1e-8)
    eff, sens = orc.effector_sensor(p)
    assert np.allclose(eff, ref[f"{key}/anm_eff"], rtol=1e-8)
    assert np.allclose(sens, ref[f"{key}/anm_sens"], rtol=1e-8)
    # GNM
    K = ref[f"{key}/kirchhoff"]
    lam, modes = orc.eigen(K)
    assert np.array_equal(lam, ref[f"{key}/gnm_eigval"])
    assert np.allclose(orc.mean_square_fluctuation(lam, modes, 1), ref[f"{key}/gnm_msf"], rtol=1e-12)
    cov = orc.covariance(K)
    assert np.allclose(orc.dcc(lam, modes, 1, cov=cov), ref[f"{key}/gnm_dcc"], atol=1e-10)
    assert np.allclose(orc.dcc(lam, modes, 1, mode_subset=np.arange(1, 17)), ref[f"{key}/gnm_dcc_sub"], atol=1e-12)


def test_two_chain_patched(structures):
    ref = golden("ref_two_chain.npz")
    coord = ref["coord"]
    s = (ref["res_name"], ref["chain_id"], ref["res_id"])
    base = orc.FFSpec("invariant", 7.0)
    K, pairs = orc.compute_kirchhoff(coord, base)
    assert np.array_equal(pairs, ref["invariant7/pairs"])
    assert np.array_equal(K, ref["invariant7/kirchhoff"])
    patches = {
        "shutdown": dict(shutdown=ref["shutdown"]),
        "pair_off": dict(pair_off=ref["pair_off"]),
        "pair_on": dict(pair_on=ref["pair_on"], pair_on_fc=ref["pair_on_fc"]),
        "all": dict(shutdown=ref["shutdown"], pair_off=ref["pair_off"],
                    pair_on=ref["pair_on"], pair_on_fc=ref["pair_on_fc"]),
    }
    for tag, kw in patches.items():
        spec = orc.FFSpec("invariant", 7.0, patched=True, **kw)
        K, pairs = orc.compute_kirchhoff(coord, spec)
        assert np.array_equal(pairs, ref[f"patched_{tag}/pairs"]), tag
        assert np.array_equal(K, ref[f"patched_{tag}/kirchhoff"]), tag
    for key in ("e_anm", "sd_enm", "d_enm", "s_enm_13"):
        spec = KINDS[key](s)
        K, pairs = orc.compute_kirchhoff(coord, spec)
        assert np.array_equal(pairs, ref[f"{key}/pairs"])
        assert np.array_equal(K, ref[f"{key}/kirchhoff"]), key
        spec.patched = True
        spec.pair_off, spec.pair_on, spec.pair_on_fc = ref["pair_off"], ref["pair_on"], ref["pair_on_fc"]
        K, _ = orc.compute_kirchhoff(coord, spec)
        assert np.array_equal(K, ref[f"{key}_patched/kirchhoff"]), key
    coord = ref["shifted_coord"]
    for key in ("e_anm", "sd_enm", "hinsen", "invariant13"):
        spec = KINDS[key](s)
        H, pairs = orc.compute_hessian(coord, spec)
        assert np.array_equal(pairs, ref[f"shifted_{key}/pairs"])
        assert np.array_equal(H, ref[f"shifted_{key}/hessian"]), key
        spec.patched = True
        spec.shutdown, spec.pair_off = ref["shutdown"], ref["pair_off"]
        spec.pair_on, spec.pair_on_fc = ref["pair_on"], ref["pair_on_fc"]
        H, pairs = orc.compute_hessian(coord, spec)
        assert np.array_equal(pairs, ref[f"shifted_{key}_patched/pairs"])
        assert np.array_equal(H, ref[f"shifted_{key}_patched/hessian"]), key


@pytest.mark.parametrize("cutoff", [5, 10, 15])
def test_random500_vs_prody_and_reference(cutoff):
    tp = golden("thirdparty_random500.npz")
    ref = golden("ref_random500.npz")
    coord = tp["coord"]
    spec = orc.FFSpec("invariant", float(cutoff))
    K, pairs = orc.compute_kirchhoff(coord, spec)
    # contact set: identical to the reference's brute AND cell-list branches
    assert np.array_equal(pairs, ref[f"pairs_{cutoff}_0"])
    assert np.array_equal(pairs, ref[f"pairs_{cutoff}_1"])
    # Kirchhoff: bit-exact vs ProDy (test_interaction.py:11-40)
    assert np.array_equal(K, tp[f"prody_kirchhoff_{cutoff}"].astype(float))
    H, _ = orc.compute_hessian(coord, spec)
    b = H.reshape(500, 3, 500, 3).transpose(0, 2, 1, 3)
    assert np.array_equal(b[pairs[:, 0], pairs[:, 1]], ref[f"hessian_offdiag_{cutoff}"])
    assert np.array_equal(b[np.arange(500), np.arange(500)], ref[f"hessian_diag_{cutoff}"])
    if cutoff == 10:  # test_interaction.py:43-68 tolerances
        P = np.zeros((500, 500, 3, 3))
        P[tp["prody_hessian_10_bi"], tp["prody_hessian_10_bj"]] = tp["prody_hessian_10_blocks"]
        assert np.allclose(b, P, atol=1e-6, rtol=1e-3)


def test_thirdparty_1l2y(structures):
    """The reference's own third-party comparisons, applied to the oracle
    (test_gnm.py:23-152, test_anm.py:145-334, test_forcefield.py:360-422)."""
    tp = golden("thirdparty_1l2y.npz")
    coord = structures["1l2y_coord"]
    s = seq(structures, "1l2y")
    for cutoff in (4, 7, 13):
        K, _ = orc.compute_kirchhoff(coord, orc.FFSpec("invariant", float(cutoff)))
        assert np.array_equal(K, tp[f"prody_gnm_{cutoff}_ang_cutoff_kirchhoff_1l2y"])
    for cutoff in (4, 7):
        K, _ = orc.compute_kirchhoff(coord, orc.FFSpec("invariant", float(cutoff)))
        lam, modes = orc.eigen(K)
        assert np.allclose(lam[1:], tp[f"prody_gnm_{cutoff}_ang_cutoff_evals_1l2y"][1:])
        refv = tp[f"prody_gnm_{cutoff}_ang_cutoff_evecs_1l2y"].copy()
        refv *= np.sign(refv[:, 0])[:, None]
        m = modes * np.sign(modes[:, 0])[:, None]
        assert np.allclose(m[1:], refv[1:], atol=1e-6)
        assert np.allclose(orc.mean_square_fluctuation(lam, modes, 1),
                           tp[f"prody_gnm_{cutoff}_ang_cutoff_fluctuations_1l2y"])
        cov = orc.covariance(K)
        assert np.allclose(orc.dcc(lam, modes, 1, cov=cov), tp[f"prody_gnm_{cutoff}_ang_cutoff_dcc_norm_1l2y"])
        assert np.allclose(orc.dcc(lam, modes, 1, cov=cov, norm=False),
                           tp[f"prody_gnm_{cutoff}_ang_cutoff_dcc_absolute_1l2y"])
        assert np.allclose(orc.dcc(lam, modes, 1, mode_subset=np.arange(1, 17)),
                           tp[f"prody_gnm_{cutoff}_ang_cutoff_dcc_norm_subset_1l2y"])
    # ANM 13 A vs ProDy
    H, _ = orc.compute_hessian(coord, orc.FFSpec("invariant", 13.0))
    lam, modes = orc.eigen(H)
    assert np.allclose(lam[6:], tp["prody_anm_13_ang_cutoff_evals_1l2y"][6:])
    assert np.allclose(orc.mean_square_fluctuation(lam, modes, 3), tp["prody_anm_13_ang_cutoff_fluctuations_1l2y"])
    cov = orc.covariance(H)
    assert np.allclose(orc.dcc(lam, modes, 3, cov=cov), tp["prody_anm_13_ang_cutoff_dcc_norm_1l2y"])
    assert np.allclose(orc.dcc(lam, modes, 3, mode_subset=np.arange(6, 36)),
                       tp["prody_anm_13_ang_cutoff_dcc_norm_subset_1l2y"])
    p = orc.prs(cov)
    assert np.allclose(p, tp["prody_anm_13_ang_cutoff_prs_mat_1l2y"])
    eff, sens = orc.effector_sensor(p)
    assert np.allclose(eff, tp["prody_anm_13_ang_cutoff_prs_eff_1l2y"])
    assert np.allclose(sens, tp["prody_anm_13_ang_cutoff_prs_sens_1l2y"])
    # Hessians vs Bio3D / BioPhysConnectoR
    for key, name, atol in (("hinsen", "bio3d_anm_calpha_ff_hessian_1l2y", 1e-4),
                            ("sd_enm", "bio3d_anm_sdenm_ff_hessian_1l2y", 1e-8),
                            ("pfree", "bio3d_anm_pfanm_ff_hessian_1l2y", 1e-8),
                            ("e_anm", "biophysconnector_anm_eanm_hessian_1l2y", 1e-8),
                            ("e_anm_mj", "biophysconnector_anm_eanm_mj_hessian_1l2y", 1e-8),
                            # "Higher deviation for eANM_Ke-FF" (test_forcefield.py:387-388)
                            ("e_anm_ke", "biophysconnector_anm_eanm_ke_hessian_1l2y", 1e-4)):
        H, _ = orc.compute_hessian(coord, KINDS[key](s))
        assert np.allclose(H, tp[name], atol=atol), key
    # mass-weighted eigenvalues vs Bio3D (test_anm.py:87-142)
    masses = tp["bio3d_mass_1l2y"]
    for key, name in (("hinsen", "calpha"), ("sd_enm", "sdenm"), ("pfree", "pfanm")):
        H, _ = orc.compute_hessian(coord, KINDS[key](s), masses=masses)
        lam, modes = orc.eigen(H)
        assert np.allclose(lam[6:], tp[f"bio3d_anm_{name}_ff_evals_mw_1l2y"][6:], rtol=5e-3, atol=2e-3)
        assert np.allclose(orc.frequencies(lam, 6)[6:], tp[f"bio3d_anm_{name}_ff_frequencies_mw_1l2y"][6:],
                           rtol=5e-3, atol=2e-3)
        msf = orc.mean_square_fluctuation(lam, modes, 3, tem=300, tem_factors=orc.K_B * orc.N_A)
        assert np.allclose(msf / (1000 * masses), tp[f"bio3d_anm_{name}_ff_fluctuations_non_mw_1l2y"],
                           rtol=5e-3, atol=2e-3)
        sub = orc.mean_square_fluctuation(lam, modes, 3, mode_subset=np.arange(11, 33), tem=300,
                                          tem_factors=orc.K_B * orc.N_A)
        assert np.allclose(sub / (1000 * masses), tp[f"bio3d_anm_{name}_ff_fluctuations_subset_mw_1l2y"],
                           rtol=5e-3, atol=2e-3)
        cov = orc.covariance(H)
        assert np.allclose(orc.dcc(lam, modes, 3, cov=cov), tp[f"bio3d_anm_{name}_ff_dcc_mw_1l2y"],
                           rtol=5e-3, atol=2e-3)
        assert np.allclose(orc.dcc(lam, modes, 3, mode_subset=np.arange(6, 36)),
                           tp[f"bio3d_anm_{name}_ff_dcc_subset_mw_1l2y"], rtol=5e-3, atol=2e-3)
    # eANM MSF vs BioPhysConnectoR "bfacs" (test_anm.py:215-229,316)
    H, _ = orc.compute_hessian(coord, KINDS["e_anm"](s))
    lam, modes = orc.eigen(H)
    assert np.allclose(orc.mean_square_fluctuation(lam, modes, 3), tp["biophysconnector_anm_eanm_bfacs_1l2y"])
    assert np.allclose(lam[6:], tp["biophysconnector_anm_eanm_evals_1l2y"][6:])


def _fuzz_spec(ref, i):
    pre = f"case{i}/"
    kind, cutoff = str(ref[pre + "kind"]), float(ref[pre + "cutoff"])
    s = (ref[pre + "res_name"], ref[pre + "chain_id"], ref[pre + "res_id"])
    if kind == "invariant":
        spec = orc.FFSpec("invariant", cutoff)
    elif kind in ("hinsen", "pfree"):
        spec = orc.FFSpec(kind, cutoff)
    elif kind in ("hinsen_nocut", "pfree_nocut"):
        spec = orc.FFSpec(kind.split("_")[0], None)
    elif kind == "e_anm_mean":
        spec = orc.preset_spec("e_anm", *s, nonbonded_mean=True)
    else:
        spec = orc.preset_spec(kind, *s)
    if bool(ref[pre + "patched"]):
        spec.patched = True
        sd = ref[pre + "shutdown"]
        spec.shutdown = sd if len(sd) else None
        spec.pair_off, spec.pair_on, spec.pair_on_fc = ref[pre + "pair_off"], ref[pre + "pair_on"], ref[pre + "pair_on_fc"]
    return spec


@pytest.mark.parametrize("i", range(26))
def test_fuzz_cases_bit_exact(i):
    """26 seeded random cases (chain breaks, numbering gaps, random sequences, all force-field kinds, random
    patches and masses) generated by the unmodified reference (tests/golden/make_golden.py::ref_fuzz)."""
    ref = golden("ref_fuzz.npz")
    pre = f"case{i}/"
    spec = _fuzz_spec(ref, i)
    coord = ref[pre + "coord"]
    masses = ref[pre + "masses"] if pre + "masses" in ref else None
    H, pairs = orc.compute_hessian(coord, spec)
    K, _ = orc.compute_kirchhoff(coord, spec)
    assert np.array_equal(pairs, ref[pre + "pairs"])
    assert np.array_equal(K, ref[pre + "kirchhoff"])
    assert np.array_equal(H, ref[pre + "hessian"])
    Hm, _ = orc.compute_hessian(coord, spec, masses=masses)
    Km, _ = orc.compute_kirchhoff(coord, spec, masses=masses)
    assert np.array_equal(Hm, ref[pre + "anm_matrix"]) and np.array_equal(Km, ref[pre + "gnm_matrix"])
    lam, modes = orc.eigen(Hm)
    assert np.array_equal(lam, ref[pre + "anm_eigval"])
    assert np.allclose(orc.mean_square_fluctuation(lam, modes, 3), ref[pre + "anm_msf"], rtol=1e-11)
    lam, modes = orc.eigen(Km)
    assert np.array_equal(lam, ref[pre + "gnm_eigval"])
    assert np.allclose(orc.mean_square_fluctuation(lam, modes, 1), ref[pre + "gnm_msf"], rtol=1e-11)


def test_c5_chain400_products():
    """C5 scaled down (reference outputs, make_golden.py::ref_c5): DCC, MSF and the linear response from the full
    pseudo-inverse and from a mode subset."""
    ref = golden("ref_c5_chain400.npz")
    coord = ref["coord"]
    H, _ = orc.compute_hessian(coord, orc.FFSpec("invariant", 13.0))
    lam, modes = orc.eigen(H)
    cov = orc.covariance(H)
    scale = np.abs(ref["anm_lr"]).max()
    assert np.allclose(orc.linear_response(cov, ref["force"]), ref["anm_lr"], rtol=1e-9, atol=1e-11 * scale)
    unit = np.zeros((len(coord), 3))
    unit[42, 0] = 1.0
    assert np.allclose(orc.linear_response(cov, unit), ref["anm_lr_unit42"], rtol=1e-9,
                       atol=1e-11 * np.abs(ref["anm_lr_unit42"]).max())
    got = orc.linear_response_modes(lam, modes, ref["force"], np.arange(6, 56))
    assert np.allclose(got, ref["anm_lr_sub"], rtol=1e-9, atol=1e-11 * np.abs(ref["anm_lr_sub"]).max())
    assert np.allclose(orc.mean_square_fluctuation(lam, modes, 3), ref["anm_msf"], rtol=1e-10)
    assert np.allclose(orc.dcc(lam, modes, 3, mode_subset=np.arange(6, 56)), ref["anm_dcc_sub"], atol=1e-10)
    assert np.allclose(orc.dcc(lam, modes, 3, cov=cov), ref["anm_dcc_all"], atol=1e-10)


def test_dense_table_edited_in_place(structures):
    """forcefield.py:429-434: an interaction_matrix edited by the caller is the table force_constant() reads."""
    ref = golden("ref_dense_table.npz")
    coord = structures["1l2y_coord"].astype(np.float64)
    s = seq(structures, "1l2y")
    for key in ("e_anm", "sd_enm"):
        spec = orc.preset_spec(key, *s)
        spec.extra["dense_table"] = ref[f"{key}/table"]
        H, pairs = orc.compute_hessian(coord, spec)
        K, _ = orc.compute_kirchhoff(coord, spec)
        assert np.array_equal(pairs, ref[f"{key}/pairs"])
        assert np.array_equal(H, ref[f"{key}/hessian"])
        assert np.array_equal(K, ref[f"{key}/kirchhoff"])
