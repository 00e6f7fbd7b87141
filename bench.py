#!/usr/bin/env python
"""Benchmark of record: ANM structures/s (Hessian + lowest 20 modes + MSF).

Workload (BASELINE.json configs[2], "C3"): an ensemble of 4,096 perturbed
conformations of a 300-residue synthetic CA chain, TabulatedForceField.e_anm
(13 A, residue-pair constants), the 20 lowest non-trivial modes and their MSF
per conformation.  One step = one pass of the whole hot path (contacts ->
assembly -> eigen -> MSF) over one batch of 4,096 conformations per GPU
(conformations shard across ranks with no communication: weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (see the keys below).  `--impl reference` times the CPU
restatement of the reference path (oracle/, identical NumPy/LAPACK calls) on the
host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from os.path import dirname, join, realpath

import numpy as np

ROOT = dirname(realpath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ANM structures/sec (Hessian+lowest 20 modes+MSF)"
N_RES = 300
N_CONF = 4096
K_MODES = 20
FF_NAME = "TabulatedForceField.e_anm"


def make_ensemble(rank, n_conf):
    """Conformations c = rank*n_conf .. : base chain + N(0, 0.5 A)."""
    from oracle import enm_oracle as orc  # input generator only (shared with the tests)
    base = orc.synthetic_chain(N_RES, seed=0)
    res_name, chain_id, res_id = orc.synthetic_sequence(N_RES, seed=0)
    rng = np.random.default_rng(1000 + rank)
    coords = base[None] + rng.normal(0.0, 0.5, size=(n_conf, N_RES, 3))
    return base, coords, (res_name, chain_id, res_id)


def peaks():
    try:
        with open(join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU reference path
def cpu_structure(args):
    """Reference algorithm for ONE conformation: dense compute_hessian -> full
    np.linalg.eigh -> modes 6..25 -> mean_square_fluctuation(mode_subset)
    (interaction.py:57-111, nma.py:61, nma.py:145-183)."""
    coord, seq = args
    from oracle import enm_oracle as orc
    spec = orc.preset_spec("e_anm", *seq)
    H, _ = orc.compute_hessian(coord, spec)
    lam, modes = orc.eigen(H)
    msf = orc.mean_square_fluctuation(lam, modes, 3, mode_subset=np.arange(6, 6 + K_MODES))
    return lam[6:6 + K_MODES], msf


def _cpu_worker_init():
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass


def time_cpu(coords, seq, n_sample, processes):
    """structures/s of the CPU reference path on `processes` host cores."""
    import multiprocessing as mp
    work = [(coords[i % len(coords)], seq) for i in range(n_sample)]
    if processes <= 1:
        t0 = time.perf_counter()
        for w in work:
            cpu_structure(w)
        return n_sample / (time.perf_counter() - t0)
    ctx = mp.get_context("fork")
    with ctx.Pool(processes, initializer=_cpu_worker_init) as pool:
        pool.map(cpu_structure, work[:processes])  # warm the workers (imports, table load)
        t0 = time.perf_counter()
        pool.map(cpu_structure, work, chunksize=1)
        return n_sample / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    _, coords, seq = make_ensemble(0, 64)
    per_step = max(2 * cores, 16)
    # calibrate so that warmup + steps stay within a few minutes
    t0 = time.perf_counter()
    cpu_structure((coords[0], seq))
    one = time.perf_counter() - t0
    budget = 150.0 / max(1, args.steps + args.warmup)
    per_step = int(max(cores, min(per_step, budget * cores / max(one, 1e-3))))
    vals = []
    for it in range(args.warmup + args.steps):
        v = time_cpu(coords, seq, per_step, cores)
        if it >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    sample = f"{per_step} conformations per step on {cores} worker processes (1 BLAS thread each)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "structures/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_step / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": value, "unit": "structures/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "structures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(n_gpus):
    return {
        "workload": f"C3 ensemble: {N_CONF} perturbed conformations x {N_RES}-residue CA chain per GPU, "
                    f"{FF_NAME} (13 A cutoff), lowest {K_MODES} non-trivial modes + MSF each",
        "conformations_per_gpu": N_CONF, "residues": N_RES, "force_field": FF_NAME, "modes": K_MODES,
        "tolerance": "residual <= 3e-9 * lambda_20 (eigenvalues ~1e-14, subspace angle ~1e-9, MSF < 1e-8 vs reference)",
        "parallelism": f"ensemble sharded by conformation over {n_gpus} GPU(s), no data-path collective",
        "l2_policy": "inputs larger than L2 (per-step working set ~7 GB >> 126 MB)",
    }


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import springcraft_b200 as sc
    from springcraft_b200 import _lib
    from springcraft_b200._engine import DeviceModel
    from springcraft_b200.ensemble import enm_ensemble_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    handle = _lib.require_device()

    base, coords, seq = make_ensemble(rank, N_CONF)
    atoms = sc.AtomArray(base, *seq)
    ff = sc.TabulatedForceField.e_anm(atoms)
    coords_pinned = torch.from_numpy(coords).pin_memory()
    eig_pinned = torch.empty((N_CONF, K_MODES), dtype=torch.float64).pin_memory()
    msf_pinned = torch.empty((N_CONF, N_RES), dtype=torch.float64).pin_memory()
    xyz = torch.from_numpy(np.ascontiguousarray(coords.transpose(0, 2, 1))).cuda()  # SoA, resident in HBM
    out = (torch.empty((N_CONF, K_MODES), dtype=torch.float64, device="cuda"),
           torch.empty((N_CONF, N_RES), dtype=torch.float64, device="cuda"),
           torch.empty(N_CONF, dtype=torch.int32, device="cuda"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return enm_ensemble_device(xyz, ff, k=K_MODES, out=out)

    def step_host():
        return sc.enm_ensemble(coords_pinned.numpy(), ff, k=K_MODES,
                               pinned_out=(eig_pinned.numpy(), msf_pinned.numpy(), None))

    def timed(fn, steps):
        barrier()
        l0 = handle.scb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            res = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = handle.scb_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, res

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches, res = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    eig_d, msf_d, iters_d, n_pairs, converged = res
    for _ in range(min(args.warmup, 2)):
        step_host()
    ms_e2e, _, res_h = timed(step_host, args.steps)

    # sanity: both arms agree, all structures converged
    assert converged and res_h.converged, "eigensolver did not converge for every conformation"
    assert np.allclose(res_h.eigenvalues, eig_d.cpu().numpy(), rtol=1e-9)
    total = N_CONF * world

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (BSR SpMM), timed live with CUDA events
    model = DeviceModel(coords[:N_CONF], ff, 3)
    b = 32
    X = torch.randn((N_CONF, 3 * N_RES, b), dtype=torch.float64, device="cuda")
    Y = torch.empty_like(X)
    for _ in range(3):
        model.spmm_paired(X, Y)
    reps = 20
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        model.spmm_paired(X, Y)
    e1.record()
    torch.cuda.synchronize()
    spmm_ms = e0.elapsed_time(e1) / reps
    P, n = model.P, N_RES
    # 72(P+n) + 4P + 8(n+1) [int64 rowptr] + 2*8*3n*b bytes per application, per batch
    alg_bytes = 72.0 * (P + N_CONF * n) + 4.0 * P + 8.0 * (N_CONF * n + 1) + 2 * 8.0 * 3 * n * b * N_CONF
    alg_flops = 18.0 * (P + N_CONF * n) * b
    pk, pk_kind = peaks()
    achieved = alg_bytes / (spmm_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "spmm_paired_kernel<3> (row-paired BSR 6x3 x 32-column block)", "achieved": achieved,
                "peak": pk["hbm_gbs"], "peak_kind": pk_kind, "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                # dram__bytes_read.sum + dram__bytes_write.sum of this launch, ncu --set full (profiles/r1h)
                "traffic": 7.415e9, "ms_per_launch": spmm_ms, "bytes_per_launch": alg_bytes,
                "fp64_gflops": alg_flops / (spmm_ms * 1e-3) / 1e9}

    # ---- CPU baseline beside it (bounded sample, rank 0 only)
    cores = os.cpu_count() or 1
    n_sample = max(cores * 2, 32)
    cpu_v = time_cpu(coords, seq, n_sample, cores)

    # ---- secondary: one 20,000-residue structure (north_star target: < 1 s on one B200)
    large = {}
    try:
        from oracle import enm_oracle as orc
        n_big = 20000
        big = orc.synthetic_chain(n_big, seed=0)
        for name, bff in (("InvariantForceField(13)", sc.InvariantForceField(13.0)),
                          ("TabulatedForceField.e_anm", sc.TabulatedForceField.e_anm(
                              sc.AtomArray(big, *orc.synthetic_sequence(n_big, seed=0))))):
            best = None
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                mb = DeviceModel(big, bff, 3)
                lam_b, X_b, res_b, it_b, _ = mb.eig_lowest(20, max_outer=400)
                msf_b = torch.empty((1, n_big), dtype=torch.float64, device="cuda")
                _lib.check(handle.scb_msf_cols(3, 1, n_big, 32, 0, 20, _lib.ptr(lam_b), _lib.ptr(X_b), 1.0,
                                               _lib.ptr(msf_b), _lib.stream_ptr()))
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            large[name] = {"residues": n_big, "ordered_pairs": int(mb.P), "seconds": best,
                           "outer_iterations": int(it_b[0]), "includes": "H2D of coordinates, contacts, assembly, "
                           "lowest 20 non-trivial modes, MSF (wall clock, best of 3)"}
    except Exception as exc:  # secondary information only
        large = {"error": repr(exc)}

    value = total * args.steps / (ms_dev * 1e-3)
    e2e_v = total * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "structures/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": e2e_v, "unit": "structures/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(coords_pinned.numel() * 8),
                "d2h_bytes_per_step": int((eig_pinned.numel() + msf_pinned.numel()) * 8)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_v, "unit": "structures/s", "cores": cores, "kind": "port",
                         "sample": f"{n_sample} conformations of the same ensemble, {cores} worker processes, "
                                   "dense Hessian + full np.linalg.eigh + MSF (reference algorithm)"},
        "single_structure_20k": large,
        "solver": {"outer_iterations_mean": float(iters_d.abs().double().mean().item()),
                   "outer_iterations_max": int(iters_d.abs().max().item()), "filter_degree": int(os.environ.get("SCB_DEGREE", 32)),
                   "block": 32, "ordered_pairs": int(n_pairs)},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
