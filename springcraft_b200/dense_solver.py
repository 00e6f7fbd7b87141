"""Lowest-k modes of DENSE interaction matrices (all-pairs force fields), row-partitioned over the
GPUs of one node.

Every rank owns a slab of rows of the N x N matrix, assembled locally from the replicated
coordinates (`scb_assemble_dense_allpairs`: no exchange at assembly, diagonal blocks are local row
sums).  The eigensolver is the same Chebyshev-filtered subspace iteration as the sparse path; per
operator application each rank computes its rows of Y = H X on the FP64 tensor cores
(Chebyshev recurrence fused) and the all-gather of the row slabs is fused into the SAME kernel
(`scb_dense_slab_apply_allgather`): every rank keeps its full output blocks in peer-mapped device
memory (`PeerBlockPool`, CUDA IPC) and the epilogue stores each finished tile into the block of every
rank over NVLink, followed by a one-element barrier all-reduce.  ``exchange="nccl"`` keeps the plain
`scb_dense_slab_apply` + NCCL all-gather for comparison.  The tall-skinny steps and the small Rayleigh-Ritz problem are
replicated on every rank (they are tiny next to the slab read).  This module only orchestrates C-ABI
kernels; the scalars of the filter live on the host (one device->host read per outer iteration).
"""

import ctypes as C
import math
import os

import numpy as np

from . import _engine, _lib
from .parallel import gather_results, row_slab, world

__all__ = ["DenseRowOperator", "PeerBlockPool", "eig_lowest_dense", "allpairs_lowest_modes"]


def _torch():
    import torch
    return torch


class _RawBlock:
    """CUDA array interface over a raw device pointer (lets torch view a peer-mapped buffer)."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerBlockPool:
    """``count`` full [N][b] fp64 blocks per rank in peer-mapped device memory.

    Every rank allocates its blocks (`scb_peer_alloc`), the 64-byte CUDA IPC handles are exchanged with
    one all-gather, and each rank maps the blocks of all others (`scb_peer_open`).  ``table(i)`` is the
    host array of `world` device pointers that `scb_dense_slab_apply_allgather` takes for block i.
    Collective: construct, use and close it from every rank in the same order."""

    def __init__(self, N, b, count=4, typestr="<f8"):
        """``typestr`` "<f8": [N][b] fp64 blocks (FP64 slab product); "<f4": [N][b] float32 (the transposed
        [2 b][ld] blocks of the TF32 filter are requested as N = 2 b rows of b = ld columns)."""
        torch = _torch()
        import torch.distributed as dist
        self.handle = _lib.require_device()
        self.rank, self.world = world()
        self.N, self.b, self.count = int(N), int(b), int(count)
        nbytes = self.N * self.b * int(typestr[2:])
        self._local, self._opened = [], []
        handles = torch.empty((self.count, 64), dtype=torch.uint8)
        for i in range(self.count):
            ptr = C.c_void_p()
            _lib.check(self.handle.scb_peer_alloc(nbytes, C.byref(ptr)))
            self._local.append(ptr.value)
            buf = (C.c_ubyte * 64)()
            _lib.check(self.handle.scb_peer_export(C.c_void_p(ptr.value), buf))
            handles[i] = torch.tensor(list(buf), dtype=torch.uint8)
        mine = handles.cuda()
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine)
        self._tables, self.blocks = [], []
        for i in range(self.count):
            ptrs = []
            for p in range(self.world):
                if p == self.rank:
                    ptrs.append(self._local[i])
                    continue
                raw = (C.c_ubyte * 64)(*gathered[p][i].cpu().tolist())
                out = C.c_void_p()
                _lib.check(self.handle.scb_peer_open(raw, C.byref(out)))
                self._opened.append(out.value)
                ptrs.append(out.value)
            self._tables.append((C.c_void_p * self.world)(*ptrs))
            self.blocks.append(torch.as_tensor(_RawBlock(self._local[i], (self.N, self.b), typestr), device="cuda"))
        self._token = torch.zeros(1, dtype=torch.float32, device="cuda")
        self._next = 0
        self._dist = dist
        self.barrier()

    def acquire(self, *busy):
        """Index of the next block (round robin) that is none of the tensors in ``busy``."""
        taken = {int(t.data_ptr()) for t in busy if t is not None}
        for _ in range(self.count):
            i = self._next
            self._next = (self._next + 1) % self.count
            if self._local[i] not in taken:
                return i
        raise RuntimeError("PeerBlockPool exhausted")

    def table(self, i):
        return self._tables[i]

    def barrier(self):
        """Stream-ordered rendezvous: returns (on the stream) once every rank has finished the kernels it
        enqueued before its own call, i.e. all remote stores into the local blocks are complete."""
        self._dist.all_reduce(self._token)

    def close(self):
        if not self._local:
            return
        torch = _torch()
        self.blocks = []
        torch.cuda.synchronize()
        self.barrier()
        torch.cuda.synchronize()
        for ptr in self._opened:
            _lib.check(self.handle.scb_peer_close(C.c_void_p(ptr)))
        self.barrier()
        torch.cuda.synchronize()
        for ptr in self._local:
            _lib.check(self.handle.scb_peer_free(C.c_void_p(ptr)))
        self._local, self._opened, self._tables = [], [], []


class DenseRowOperator:
    """Rows [row0, row1) (node units) of the dense interaction matrix of ONE structure."""

    def __init__(self, coord, force_field, D=3, masses=None, exchange=None):
        torch = _torch()
        self.handle = _lib.require_device()
        exchange = exchange or os.environ.get("SCB_DENSE_EXCHANGE", "peer")
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        self.exchange = exchange
        self._pools = {}
        self._workspaces = {}
        coord = np.asarray(coord, dtype=np.float64)
        if coord.ndim != 2 or coord.shape[1] != 3:
            raise ValueError(f"Expected coordinates with shape (n,3), got {coord.shape}")
        if force_field.cutoff_distance is not None:
            raise ValueError("the dense row-partitioned path is for force fields without a cutoff")
        if force_field.contact_shutdown is not None or force_field.contact_pair_off is not None \
                or force_field.contact_pair_on is not None:
            raise NotImplementedError("contact patches are not supported on the dense all-pairs path")
        built = force_field._descriptor(len(coord))
        if built is None:
            raise NotImplementedError("user-defined ForceField subclasses are not supported on the dense path")
        self.desc, self._keep = built
        self.D, self.n = int(D), len(coord)
        self.N = self.D * self.n
        self.rank, self.world = world()
        self.row0, self.row1 = row_slab(self.n, self.rank, self.world)
        self.xyz = torch.from_numpy(np.ascontiguousarray(coord.T)).cuda()[None].contiguous()  # SoA [1][3][n]
        self.masses = None if masses is None else _lib.to_device(np.asarray(masses, dtype=np.float64), torch.float64)
        rows = (self.row1 - self.row0) * self.D
        self.slab = torch.empty((rows, self.N), dtype=torch.float64, device="cuda")
        _lib.check(self.handle.scb_assemble_dense_allpairs(
            self.D, _lib.ptr(self.xyz), self.n, C.byref(self.desc), _lib.ptr(self.masses), self.row0, self.row1,
            _lib.ptr(self.slab), _lib.stream_ptr()))

    # -- operator ---------------------------------------------------------------
    def apply(self, X, W=None, coeffs=None):
        """Y = H X (coeffs None) or alpha (H X - c X) - beta W; returns the full [N][b] block."""
        torch = _torch()
        b = int(X.shape[1])
        alpha, cshift, beta = coeffs if coeffs is not None else (1.0, 0.0, 0.0)
        ws = self._workspaces.get(b)
        if ws is None:   # split-K scratch (zeroed once: the tile counters return to 0 after every launch)
            nbytes = self.handle.scb_dense_slab_workspace_bytes(self.N, self.row0 * self.D, self.row1 * self.D, b)
            ws = self._workspaces[b] = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        if self.world > 1 and self.exchange == "peer":
            pool = self._pools.get(b)
            if pool is None:
                pool = self._pools[b] = PeerBlockPool(self.N, b)
            i = pool.acquire(X, W)
            _lib.check(self.handle.scb_dense_slab_apply_allgather(
                self.N, self.row0 * self.D, self.row1 * self.D, _lib.ptr(self.slab), _lib.ptr(X), _lib.ptr(W),
                pool.table(i), self.world, b, int(coeffs is not None), float(alpha), float(cshift), float(beta),
                _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
            pool.barrier()
            return pool.blocks[i]
        local = torch.empty(((self.row1 - self.row0) * self.D, b), dtype=torch.float64, device="cuda")
        _lib.check(self.handle.scb_dense_slab_apply(
            self.N, self.row0 * self.D, self.row1 * self.D, _lib.ptr(self.slab), _lib.ptr(X), _lib.ptr(W),
            _lib.ptr(local), b, int(coeffs is not None), float(alpha), float(cshift), float(beta),
            _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        if self.world == 1:
            return local
        if self.n % self.world == 0:
            # equal slabs: one NCCL all-gather straight into the full block (rows are rank-major contiguous)
            import torch.distributed as dist
            full = torch.empty((self.N, b), dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(full, local)
            return full
        # ragged split: all-gather of padded slabs (node granularity keeps the slabs aligned)
        full = gather_results(local.view(self.row1 - self.row0, self.D * b), self.n, dim=0)
        return full.view(self.N, b)

    # -- residual-form filter on the TF32 tensor cores -------------------------------
    def slab32(self, split=True):
        """Single-precision copy of the row slab (built on first use): with ``split`` the TF32-representable part
        and the remainder (two slabs, 8 N^2 / G bytes), else one plain FP32 slab."""
        torch = _torch()
        key = "_slab32_split" if split else "_slab32_plain"
        if getattr(self, key, None) is None:
            ld = int(self.handle.scb_tf32_ld(self.N))
            hi = torch.empty((self.slab.shape[0], ld), dtype=torch.float32, device="cuda")
            lo = torch.empty_like(hi) if split else None
            _lib.check(self.handle.scb_dense_slab_to_f32(self.N, self.slab.shape[0], _lib.ptr(self.slab),
                                                         _lib.ptr(hi), _lib.ptr(lo), _lib.stream_ptr()))
            setattr(self, key, (hi, lo))
        return getattr(self, key)

    def filter_residual_tf32(self, A, HX, theta, rn2, lo, ub, degree, split=True):
        """A <- A + |r| z with z = q(H) r / p(theta) from `degree` Chebyshev steps on [lo, ub] (see dense_tf32.cu);
        A are Ritz vectors, HX = H A, rn2 their squared residual norms.  One GPU, 128-column blocks.
        ``split``: 3-term TF32 product (FP32-class accuracy, two slab streams per step) instead of a single one."""
        torch = _torch()
        h, st = self.handle, _lib.stream_ptr
        b = int(A.shape[1])
        if b != 128:
            raise NotImplementedError("the TF32 filter needs 128-column blocks")
        ld = int(h.scb_tf32_ld(self.N))
        rows = 2 * b if split else b
        cache = getattr(self, "_tf32_buffers", None)
        if cache is None or cache["rows"] != rows:
            if self.world > 1:
                # the z blocks live in peer-mapped memory: every rank stores its rows into the blocks of all ranks
                pool = self._pools[("tf32", rows)] = PeerBlockPool(rows, ld, count=2, typestr="<f4")
                z = pool.blocks
            else:
                pool = None
                z = [torch.empty((rows, ld), dtype=torch.float32, device="cuda") for _ in range(2)]
            cache = self._tf32_buffers = {
                "rows": rows, "pool": pool, "z": z,
                "rhat": torch.empty((b, ld), dtype=torch.float32, device="cuda"),
                "cA": torch.empty((64, b), dtype=torch.float32, device="cuda"),
                "cB": torch.empty((64, b), dtype=torch.float32, device="cuda")}
        degree = int(min(max(degree, 2), 64))
        pool = cache["pool"]
        iprev, icur = 0, 1
        z = cache["z"]
        hi, lo_slab = self.slab32(split)
        if pool is not None:
            pool.barrier()       # nobody still reads the blocks of the previous filter
        # replicated: every rank prepares the complete blocks from its copies of A, HX, theta, rn2
        _lib.check(h.scb_resform_prepare(self.N, b, degree, _lib.ptr(A), _lib.ptr(HX), _lib.ptr(theta), _lib.ptr(rn2),
                                         float(lo), float(ub), _lib.ptr(cache["rhat"]), _lib.ptr(z[icur]),
                                         _lib.ptr(z[iprev]), _lib.ptr(cache["cA"]), _lib.ptr(cache["cB"]), int(split),
                                         st()))
        if pool is not None:
            pool.barrier()       # every rank has initialised its blocks before remote rows arrive
        cshift = 0.5 * (ub + lo)
        r0, r1 = self.row0 * self.D, self.row1 * self.D
        for k in range(1, degree):
            # z_{k+1} overwrites z_{k-1}; rows [r0, r1) of the result go into the block of every rank
            if pool is None:
                _lib.check(h.scb_dense_slab_tf32_apply(self.N, r0, r1, _lib.ptr(hi), _lib.ptr(lo_slab), b,
                                                       _lib.ptr(z[icur]), _lib.ptr(z[iprev]), _lib.ptr(cache["rhat"]),
                                                       _lib.ptr(z[iprev]), _lib.ptr(cache["cA"][k]),
                                                       _lib.ptr(cache["cB"][k]), cshift, 1, st()))
            else:
                _lib.check(h.scb_dense_slab_tf32_apply_allgather(
                    self.N, r0, r1, _lib.ptr(hi), _lib.ptr(lo_slab), b, _lib.ptr(z[icur]), _lib.ptr(z[iprev]),
                    _lib.ptr(cache["rhat"]), pool.table(iprev), self.world, _lib.ptr(cache["cA"][k]),
                    _lib.ptr(cache["cB"][k]), cshift, 1, st()))
                pool.barrier()
            iprev, icur = icur, iprev
        _lib.check(h.scb_resform_finish(self.N, b, _lib.ptr(rn2), _lib.ptr(z[icur]), _lib.ptr(A), int(split), st()))
        return A

    def close(self):
        """Release the peer-mapped blocks (collective; tensors returned by `apply` die with them)."""
        self._tf32_buffers = None
        for pool in self._pools.values():
            pool.close()
        self._pools = {}

    def spectrum_bound(self):
        """Gershgorin upper bound of the spectrum (max over ranks)."""
        torch = _torch()
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        _lib.check(self.handle.scb_dense_gershgorin(self.N, self.slab.shape[0], _lib.ptr(self.slab), _lib.ptr(out),
                                                    _lib.stream_ptr()))
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(out, op=dist.ReduceOp.MAX)
        return float(out.item())

    def rigid_basis(self):
        torch = _torch()
        nz = 6 if self.D == 3 else 1
        Z = torch.empty((1, self.N, nz), dtype=torch.float64, device="cuda")
        _lib.check(self.handle.scb_rigid_basis(self.D, _lib.ptr(self.xyz), 1, self.n, _lib.ptr(self.masses),
                                               _lib.ptr(Z), _lib.stream_ptr()))
        return Z[0]


def lanczos_bound(op, start, steps=12, factor=1.05):
    """Estimate of the largest eigenvalue from `steps` steps of column-wise Lanczos on the (random) block `start`
    (every column an independent run), times a safety factor: the Gershgorin bound of an all-pairs Hessian is
    ~3x too large, which costs ~1.7x more filter steps.  Collective when the operator is partitioned."""
    torch = _torch()
    h, st = op.handle, _lib.stream_ptr
    N = int(start.shape[0])
    b = min(int(start.shape[1]), 64)     # 64 independent Lanczos runs are plenty; the slab product costs ~ b
    f64 = dict(dtype=torch.float64, device="cuda")
    V = start[:, :b].contiguous()
    Vp = torch.zeros_like(V)
    alpha = torch.zeros((steps, b), **f64)
    beta2 = torch.zeros((steps, b), **f64)
    nrm = torch.empty(b, **f64)
    _lib.check(h.scb_coldot(1, N, b, _lib.ptr(V), _lib.ptr(V), _lib.ptr(nrm), st()))
    _lib.check(h.scb_lanczos_axpy(1, N, b, 2, _lib.ptr(V), None, None, None, None, _lib.ptr(nrm), st()))
    for j in range(steps):
        W = op.apply(V)
        if W.data_ptr() == V.data_ptr():
            raise RuntimeError("operator returned its input block")
        W = W if op.world == 1 else W.clone()      # peer-pool blocks rotate: keep this one
        _lib.check(h.scb_coldot(1, N, b, _lib.ptr(V), _lib.ptr(W), _lib.ptr(alpha[j]), st()))
        _lib.check(h.scb_lanczos_axpy(1, N, b, 0, _lib.ptr(V), _lib.ptr(Vp), _lib.ptr(W), _lib.ptr(alpha[j]),
                                      _lib.ptr(beta2[j - 1]) if j else None, None, st()))
        _lib.check(h.scb_coldot(1, N, b, _lib.ptr(W), _lib.ptr(W), _lib.ptr(beta2[j]), st()))
        if j + 1 < steps:
            _lib.check(h.scb_lanczos_axpy(1, N, b, 1, _lib.ptr(V), _lib.ptr(Vp), _lib.ptr(W), None, None,
                                          _lib.ptr(beta2[j]), st()))
    out = torch.empty(1, **f64)
    _lib.check(h.scb_lanczos_bound(1, b, steps, _lib.ptr(alpha), _lib.ptr(beta2), float(factor), _lib.ptr(out), st()))
    if op.world > 1:
        import torch.distributed as dist
        dist.broadcast(out, src=0)                 # one value for every rank
    return float(out.item())


def eig_lowest_dense(op, k, Z=None, b=None, tol=3e-9, degree=24, max_outer=300, seed=0x5CB200, lanczos_steps=12,
                     filter=None):
    """The k lowest modes of the operator deflated by Z ([N][nz], orthonormal): returns
    (theta[b], X[N][b], resid[b], outer_iterations).  Columns 0..k-1 are converged to
    ``||H x - theta x|| <= tol * theta_k``.

    ``filter``: "fp64" = Chebyshev filter of the block itself with the FP64 slab kernel; "tf32" = residual-form
    filter (the correction of every Ritz pair) on the TF32 tensor cores as a 3-term split product (FP32-class
    accuracy), "tf32x1" = the same with a single TF32 product (enough for small systems); FP64 everywhere else --
    128-column blocks (default "tf32"; ``SCB_DENSE_FILTER`` overrides).  On several GPUs every rank filters its
    row slab and stores the result into the peer-mapped blocks of all ranks (fused all-gather)."""
    torch = _torch()
    h = op.handle
    st = _lib.stream_ptr
    N = op.N
    filter = filter or os.environ.get("SCB_DENSE_FILTER") or "tf32"
    if filter not in ("fp64", "tf32", "tf32x1"):
        raise ValueError("filter must be 'fp64', 'tf32' or 'tf32x1'")
    if b is None:
        b = 128 if (filter != "fp64" or k + 8 > 64) else 64
    if k > b or b not in (64, 128):
        raise NotImplementedError(f"k={k} needs a block wider than 128 columns")
    if filter != "fp64" and b != 128:
        filter = "fp64"
    nz = 0 if Z is None else int(Z.shape[1])
    if N < b + nz:
        raise NotImplementedError("system smaller than the solver block: use the full-spectrum solver")
    f64 = dict(dtype=torch.float64, device="cuda")
    A = torch.empty((N, b), **f64)
    scratch = torch.zeros(8 * b, **f64)
    G = torch.empty((b, b), **f64)
    Cm = torch.empty((b, b), **f64)
    rn2 = torch.empty(b, **f64)
    _lib.check(h.scb_rand_block(N * b, seed, _lib.ptr(A), st()))
    ub_safe = op.spectrum_bound() * (1.0 + 1e-10)
    ub = min(ub_safe, lanczos_bound(op, A, steps=lanczos_steps)) if lanczos_steps else ub_safe
    theta = None
    lo = a0 = 0.0
    lo_seen = None
    cur = A

    def orthonormalise(src, dst, also=None):
        _lib.check(h.scb_gram(1, N, b, _lib.ptr(src), _lib.ptr(src), _lib.ptr(G), st()))
        _lib.check(h.scb_chol_orth(1, b, _lib.ptr(G), _lib.ptr(Cm), st()))
        _lib.check(h.scb_rotate(1, N, b, _lib.ptr(Cm), _lib.ptr(src), _lib.ptr(dst), _lib.ptr(also), _lib.ptr(also), st()))

    for outer in range(max_outer + 1):
        if outer > 0 and filter != "fp64":
            # A <- A + correction (nearly orthonormal)
            op.filter_residual_tf32(A, HX, theta, rn2, lo, ub, degree, split=(filter == "tf32"))
            cur = A
        elif outer > 0:
            half, c = 0.5 * (ub - lo), 0.5 * (ub + lo)
            sigma1 = half / (a0 - c)
            sigma = sigma1
            prev = A
            curb = op.apply(prev, None, (sigma1 / half, c, 0.0))
            for _ in range(1, degree):
                sigma2 = 1.0 / (2.0 / sigma1 - sigma)
                nxt = op.apply(curb, prev, (2.0 * sigma2 / half, c, sigma * sigma2))
                prev, curb, sigma = curb, nxt, sigma2
            cur = curb
        if nz:
            _lib.check(h.scb_deflate(1, N, b, nz, _lib.ptr(Z), _lib.ptr(cur), _lib.ptr(scratch), st()))
        if not (outer > 0 and filter != "fp64"):
            orthonormalise(cur, A)      # the residual-form update keeps A well conditioned: CholQR below suffices
        HX = op.apply(A)
        orthonormalise(A, A, also=HX)          # second pass (CholQR2); H (A C) = (H A) C
        _lib.check(h.scb_gram(1, N, b, _lib.ptr(A), _lib.ptr(HX), _lib.ptr(G), st()))
        lam, modes = _engine.eig_full_dense(G.clone())
        theta = lam[0].contiguous()
        _lib.check(h.scb_transpose_small(b, _lib.ptr(modes[0]), _lib.ptr(Cm), st()))
        _lib.check(h.scb_rotate(1, N, b, _lib.ptr(Cm), _lib.ptr(A), _lib.ptr(A), _lib.ptr(HX), _lib.ptr(HX), st()))
        _lib.check(h.scb_residual_norms(1, N, b, _lib.ptr(A), _lib.ptr(HX), _lib.ptr(theta), _lib.ptr(rn2), st()))
        if op.world > 1:
            # The replicated tall-skinny steps may differ by rounding between ranks; every decision that steers
            # the collectives (convergence, filter bounds) is taken on rank 0's values so that all ranks agree.
            import torch.distributed as dist
            both = torch.cat([theta, rn2])
            dist.broadcast(both, src=0)
            theta, rn2 = both[:b].contiguous(), both[b:].contiguous()
        th = theta.cpu().numpy()
        res = np.sqrt(np.maximum(rn2.cpu().numpy(), 0.0))
        a0 = float(th[0])
        if os.environ.get("SCB_DENSE_TRACE"):
            print(f"[dense] outer {outer} filter {filter} theta0 {th[0]:.4g} theta_k {th[k - 1]:.4g} theta_b {th[b - 1]:.4g} "
                  f"ub {ub:.4g} res_max {res[:k].max():.3e}", flush=True)
        if outer > 0 and ub < ub_safe and th[b - 1] > 0.5 * ub:
            # a filtered block must sit far below the bound: the Lanczos estimate was too small, repair it
            ub = min(ub_safe, 1.15 * max(ub, float(th[b - 1])))
        # Lower edge of the damped interval = largest Ritz value of the block.  Once the block has been filtered the
        # legitimate value only decreases; a jump upwards is a component from the top of the spectrum that leaked in
        # (bound estimate slightly too small, single-precision noise) and must not drag the interval up with it.
        top = float(th[b - 1])
        lo = top if outer < 1 or lo_seen is None else min(top, lo_seen)
        if outer >= 1:
            lo_seen = lo
        lo = min(lo, 0.98 * ub)
        if not lo > a0:
            lo = a0 + 0.5 * (ub - a0)
        if res[:k].max() <= tol * max(abs(th[k - 1]), 1e-6 * ub):
            return theta, A, torch.from_numpy(res).cuda(), outer
    raise RuntimeError(_lib.lib().scb_status_string(_lib.SCB_ERR_NOT_CONVERGED).decode())


def allpairs_lowest_modes(coord, force_field, k, kind="anm", masses=None, tol=3e-9, exchange=None, filter=None):
    """``eigen(k=...)`` for all-pairs force fields on the dense row-partitioned path: the k lowest
    modes INCLUDING the trivial ones (analytic rigid-body basis, eigenvalue 0), rows = modes.
    Call it from every rank of the process group; the result is replicated.  ``filter``: see eig_lowest_dense."""
    torch = _torch()
    D = 3 if kind == "anm" else 1
    ntriv = 6 if D == 3 else 1
    op = DenseRowOperator(coord, force_field, D, masses, exchange=exchange)
    Z = op.rigid_basis()
    kk = max(k - ntriv, 1)
    try:
        theta, X, _, iters = eig_lowest_dense(op, kk, Z=Z, tol=tol, filter=filter)
        lam = torch.cat([torch.zeros(ntriv, dtype=torch.float64, device="cuda"), theta[:kk]])
        modes = torch.cat([Z.T.contiguous(), X[:, :kk].T.contiguous()])
        return lam[:k].cpu().numpy(), modes[:k].cpu().numpy(), iters
    finally:
        op.close()


del math
