"""Checks of the partitioned (multi-GPU) paths that need NO oracle: the product on N ranks against the product on one
rank and against the committed golden fixtures.  Called collectively from every rank of an initialised NCCL process
group: by tests/test_multigpu.py (2 GPUs) and by `bench.py --gpus N` (N > 1), so that the driver's scaling run proves
the same paths (C3 shards, C4 row slabs + exchange, C5 row-partitioned contraction)."""
from os.path import dirname, join, realpath

import numpy as np

GOLDEN = join(dirname(realpath(__file__)), "golden")


def run_checks():
    """Returns a dict of what was verified; raises AssertionError on the first mismatch."""
    import torch
    import torch.distributed as dist
    import springcraft_b200 as sc
    from springcraft_b200 import _engine, parallel
    from springcraft_b200.dense_solver import DenseRowOperator

    rank, world = dist.get_rank(), dist.get_world_size()
    report = {"world": world}

    # ---- C4 (scaled down): dense all-pairs Hessian, row slabs + exchange of the block per operator application
    g = np.load(join(GOLDEN, "ref_c4_cloud400.npz"))
    lam4, modes4, it4 = sc.allpairs_lowest_modes(g["coord"], sc.ParameterFreeForceField(), 56)
    assert np.allclose(lam4[6:56], g["eigval"][6:56], rtol=1e-8, atol=0), "C4 eigenvalues vs reference golden"
    Q, _ = np.linalg.qr(modes4[6:56].T)
    Qa, _ = np.linalg.qr(g["modes_6_106"][:50].T)
    angle = float(np.linalg.norm(Q - Qa @ (Qa.T @ Q), 2))
    assert angle < 1e-6, "C4 subspace angle"
    report["c4_cloud400"] = {"eig_rel_max": float(np.max(np.abs(lam4[6:56] / g["eigval"][6:56] - 1.0))),
                             "subspace_sin": angle, "outer_iterations": int(it4)}
    # a borderline tolerance: rounding differences between the ranks must not split the collective decisions
    lam4b, _, _ = sc.allpairs_lowest_modes(g["coord"], sc.ParameterFreeForceField(), 56, tol=1e-6)
    assert np.allclose(lam4b[6:56], g["eigval"][6:56], rtol=1e-6), "C4 loose tolerance"
    # the FP64 filter (slab product with the fused peer-memory all-gather) and the same solve with the plain NCCL
    # all-gather instead of the fused epilogue
    lam4f, _, it4f = sc.allpairs_lowest_modes(g["coord"], sc.ParameterFreeForceField(), 56, filter="fp64")
    assert np.allclose(lam4f[6:56], g["eigval"][6:56], rtol=1e-8, atol=0), "C4 eigenvalues, FP64 filter"
    lam4n, _, it4n = sc.allpairs_lowest_modes(g["coord"], sc.ParameterFreeForceField(), 56, exchange="nccl",
                                              filter="fp64")
    assert it4n == it4f and np.allclose(lam4n, lam4f, rtol=1e-11, atol=1e-12), "peer vs nccl exchange"
    report["c4_cloud400"]["outer_iterations_fp64_filter"] = int(it4f)
    # one operator application, both exchanges, bit for bit
    gen = torch.Generator("cuda").manual_seed(7)
    Xb = torch.randn((3 * len(g["coord"]), 64), dtype=torch.float64, device="cuda", generator=gen)
    dist.broadcast(Xb, src=0)
    outs = []
    for ex in ("peer", "nccl"):
        op = DenseRowOperator(g["coord"], sc.ParameterFreeForceField(), 3, exchange=ex)
        outs.append(op.apply(Xb).clone())
        outs.append(op.apply(Xb, Xb, (0.5, 0.1, 0.25)).clone())
        op.close()
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[3]), "fused all-gather differs from NCCL all-gather"
    report["exchange"] = "peer == nccl bit for bit"

    # ---- C5: row-partitioned DCC against the single-rank contraction of the same operands
    c5 = np.load(join(GOLDEN, "ref_c5_chain400.npz"))
    anm = sc.ANM(c5["coord"], sc.InvariantForceField(13.0))
    lam, modes = anm.eigen(k=56)
    lam_d = torch.from_numpy(lam[6:]).cuda()
    modes_d = torch.from_numpy(np.ascontiguousarray(modes[6:])).cuda()
    full = parallel.dcc_row_partitioned(3, lam_d, modes_d, norm=True, gather=True).cpu().numpy()
    assert np.allclose(full, c5["anm_dcc_sub"], atol=1e-8), "gathered DCC vs reference golden"
    row0, row1, slab = parallel.dcc_row_partitioned(3, lam_d, modes_d, norm=False)
    one = _engine.modes_dcc(3, lam_d, modes_d, norm=False).cpu().numpy()
    assert np.allclose(slab.cpu().numpy(), one[row0:row1], rtol=1e-12, atol=1e-16), "row slab vs full contraction"
    report["c5_chain400"] = {"dcc_abs_max_err": float(np.max(np.abs(full - c5["anm_dcc_sub"])))}

    # ---- C3: ensemble shards, results gathered, against the whole ensemble on every rank
    from synthetic_inputs import perturbed_conformation
    B = 2 * world + 1                                  # ragged split
    confs = np.stack([perturbed_conformation(c5["coord"], c) for c in range(B)])
    a, b = parallel.shard_range(B, rank, world)
    res = sc.enm_ensemble(confs[a:b], sc.InvariantForceField(13.0), k=20)
    ev = parallel.gather_results(torch.from_numpy(res.eigenvalues).cuda(), B).cpu().numpy()
    ref = sc.enm_ensemble(confs, sc.InvariantForceField(13.0), k=20).eigenvalues
    assert np.allclose(ev, ref, rtol=1e-9), "sharded ensemble vs whole ensemble"
    report["c3_shards"] = {"structures": B, "eig_rel_max": float(np.max(np.abs(ev / ref - 1.0)))}
    return report
