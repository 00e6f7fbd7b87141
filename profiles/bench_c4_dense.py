"""C4 at full size: dense all-pairs Hessian (ParameterFree), lowest-k modes, row-partitioned over the ranks.
    python profiles/bench_c4_dense.py <n> <k>      (or under torch.distributed.run for N GPUs)"""
import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import springcraft_b200 as sc
from oracle import enm_oracle as orc
from springcraft_b200.dense_solver import DenseRowOperator, eig_lowest_dense
import os
if int(os.environ.get('WORLD_SIZE','1'))>1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
n = int(sys.argv[1]); k = int(sys.argv[2])
rng = np.random.default_rng(0)
side = (n/0.008)**(1/3)
# quick cloud with min distance via jittered grid (fast generator for big n)
g = int(np.ceil(n**(1/3))); pts = np.stack(np.meshgrid(*[np.arange(g)]*3, indexing='ij'), -1).reshape(-1,3)[:n]*(side/g)
coord = pts + rng.uniform(-0.25, 0.25, pts.shape)*(side/g)*0.5
torch.cuda.synchronize(); t0=time.perf_counter()
op = DenseRowOperator(coord, sc.ParameterFreeForceField(), 3)
torch.cuda.synchronize(); t1=time.perf_counter()
b = 64 if k+8<=64 else 128
X = torch.randn((3*n, b), dtype=torch.float64, device='cuda')
for _ in range(2): op.apply(X)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5): op.apply(X)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
if int(os.environ.get("RANK","0"))==0: print(f"world={os.environ.get('WORLD_SIZE','1')} exchange={op.exchange} n={n} N={3*n} slab {op.slab.numel()*8/1e9:.1f} GB assemble {t1-t0:.3f}s; apply b={b}: {dt*1e3:.2f} ms = {2*(3*n)**2*b/dt/1e12:.1f} TFLOP/s, {op.slab.numel()*8/dt/1e9:.0f} GB/s")
Z = op.rigid_basis()
torch.cuda.synchronize(); t=time.perf_counter()
theta, A, res, iters = eig_lowest_dense(op, k, Z=Z)
torch.cuda.synchronize(); elapsed = time.perf_counter()-t
ub = op.spectrum_bound()   # collective: every rank calls it
exchange = op.exchange
op.close()                 # collective: releases the peer-mapped blocks
if int(os.environ.get("WORLD_SIZE","1"))>1: dist.destroy_process_group()
if int(os.environ.get("RANK","0"))==0: print(f"lowest {k} modes: {elapsed:.2f} s, {iters} outer iterations, ub={ub:.1f}, theta[0]={theta[0].item():.4g}, theta[k-1]={theta[k-1].item():.4g}, maxres={res[:k].max().item():.2e}")
# ---- the same solve with the FP64 filter, and one TF32 filter step timed alone
if int(os.environ.get("WORLD_SIZE","1")) == 1 and os.environ.get("C4_COMPARE", "1") == "1":
    op = DenseRowOperator(coord, sc.ParameterFreeForceField(), 3)
    from springcraft_b200 import _lib
    h = _lib.require_device()
    ld = int(h.scb_tf32_ld(op.N)); s32, s32lo = op.slab32(True)
    Zc = torch.randn((256, ld), dtype=torch.float32, device="cuda"); Zp = torch.randn_like(Zc); Rh = torch.randn((128, ld), dtype=torch.float32, device="cuda")
    cA = torch.rand(128, device="cuda") * 0.1; cB = torch.rand(128, device="cuda")
    def step():
        _lib.check(h.scb_dense_slab_tf32_apply(op.N, 0, op.N, _lib.ptr(s32), _lib.ptr(s32lo), 128, _lib.ptr(Zc), _lib.ptr(Zp), _lib.ptr(Rh), _lib.ptr(Zp), _lib.ptr(cA), _lib.ptr(cB), 0.7, 1, _lib.stream_ptr()))
    for _ in range(3): step()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)/10
    print(f"3xTF32 filter step (tcgen05): {ms:.3f} ms = {3*2*op.N**2*128/ms/1e9:.0f} TFLOP/s issued, slab stream {2*s32.numel()*4/ms/1e6:.0f} GB/s")
    Z = op.rigid_basis()
    torch.cuda.synchronize(); t=time.perf_counter()
    theta, A, res, iters = eig_lowest_dense(op, k, Z=Z, filter="fp64")
    torch.cuda.synchronize()
    print(f"FP64 filter: lowest {k} modes: {time.perf_counter()-t:.2f} s, {iters} outer iterations")
    op.close()
