// Two-sided cyclic Jacobi in shared memory (shared by the Rayleigh-Ritz kernel
// and the block-Jacobi pivot kernel).
#pragma once
#include "common.cuh"

namespace scb {

__device__ __forceinline__ double block_sum_256(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    return t;
}

// Two-sided cyclic Jacobi (round-robin ordering) on the symmetric matrix A
// (order M, leading dimension LD) in shared memory, accumulating V (M rows; V
// must be initialised by the caller).  Called by all 256 threads of the CTA.
// One round = M/2 disjoint rotations: the parameters are computed by M/2
// threads, then A <- J^T A J is applied in ONE pass over the 2x2 sub-blocks
// {p_k,q_k} x {p_l,q_l} (row and column rotation fused, 4 loads + 4 stores per
// sub-block) together with V <- V J: two barriers per round.
template <int M, int LD>
__device__ void jacobi_eigen_smem(double* A, double* V, double* cs, double* sn, int* pp, int* qq, double* red,
                                  int max_sweeps = 30, bool cross_only = false) {
    constexpr int half = M / 2;   // M is even
    const int tid = threadIdx.x;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int q = tid; q < M * M; q += 256) {
            const int i = q / M, j = q % M;
            const double v = A[i * LD + j];
            if (i == j) dg += v * v; else off += v * v;
        }
        off = block_sum_256(off, red);
        dg = block_sum_256(dg, red);
        if (off <= 1e-30 * dg || off == 0.0) break;
        // cross_only: only the pairs (i, M/2 + j) between the two halves are rotated (M/2 rounds); the block
        // Jacobi driver uses it for pivots whose diagonal blocks were treated earlier in the same sweep
        const int rounds = cross_only ? half : M - 1;
        for (int step = 0; step < rounds; ++step) {
            if (tid < half) {
                int p, q;
                if (cross_only) { p = tid; q = half + ((tid + step) & (half - 1)); }
                else if (tid == 0) { p = M - 1; q = step % (M - 1); }
                else { p = (step + tid) % (M - 1); q = (step - tid + (M - 1)) % (M - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                const double apq = A[p * LD + q];
                double c = 1.0, s = 0.0;
                if (apq != 0.0) {
                    // t = sign(tau) / (|tau| + sqrt(1 + tau^2)), tau = d / h, written with one sqrt and one division
                    const double d = A[q * LD + q] - A[p * LD + p], h = 2.0 * apq;
                    const double r = sqrt(fma(d, d, h * h));
                    double t = (r > 0.0) ? fabs(h) / (fabs(d) + r) : 1.0;
                    if ((d < 0.0) != (h < 0.0)) t = -t;
                    c = rsqrt(fma(t, t, 1.0));
                    s = t * c;
                }
                cs[tid] = c; sn[tid] = s; pp[tid] = p; qq[tid] = q;
            }
            __syncthreads();
            // A <- J^T A J on the 2x2 sub-blocks (k: row pair, l: column pair); loads of all sub-blocks of a
            // thread are issued before the first store so that the shared-memory latency overlaps
            {
                constexpr int NA = (half * half) / 256;
                static_assert((half * half) % 256 == 0, "sub-block count must be a multiple of the CTA size");
                double a[NA][4], ck[NA], sk[NA], cl[NA], sl[NA];
                int ip[NA], iq[NA], ir[NA], it[NA];
#pragma unroll
                for (int u = 0; u < NA; ++u) {
                    const int w = u * 256 + tid;
                    const int k = w / half, l = w % half;
                    ck[u] = cs[k]; sk[u] = sn[k]; cl[u] = cs[l]; sl[u] = sn[l];
                    ip[u] = pp[k] * LD; iq[u] = qq[k] * LD; ir[u] = pp[l]; it[u] = qq[l];
                }
#pragma unroll
                for (int u = 0; u < NA; ++u) {
                    a[u][0] = A[ip[u] + ir[u]]; a[u][1] = A[ip[u] + it[u]];
                    a[u][2] = A[iq[u] + ir[u]]; a[u][3] = A[iq[u] + it[u]];
                }
#pragma unroll
                for (int u = 0; u < NA; ++u) {
                    const double bpr = ck[u] * a[u][0] - sk[u] * a[u][2], bqr = sk[u] * a[u][0] + ck[u] * a[u][2];
                    const double bpt = ck[u] * a[u][1] - sk[u] * a[u][3], bqt = sk[u] * a[u][1] + ck[u] * a[u][3];
                    A[ip[u] + ir[u]] = cl[u] * bpr - sl[u] * bpt;
                    A[ip[u] + it[u]] = sl[u] * bpr + cl[u] * bpt;
                    A[iq[u] + ir[u]] = cl[u] * bqr - sl[u] * bqt;
                    A[iq[u] + it[u]] = sl[u] * bqr + cl[u] * bqt;
                }
            }
            // V <- V J
            {
                constexpr int NV = (half * M) / 256;
                double v[NV][2], c[NV], sg[NV];
                int ir[NV], it[NV];
#pragma unroll
                for (int u = 0; u < NV; ++u) {
                    const int w = u * 256 + tid;
                    const int i = w / half, l = w % half;
                    c[u] = cs[l]; sg[u] = sn[l];
                    ir[u] = i * LD + pp[l]; it[u] = i * LD + qq[l];
                }
#pragma unroll
                for (int u = 0; u < NV; ++u) { v[u][0] = V[ir[u]]; v[u][1] = V[it[u]]; }
#pragma unroll
                for (int u = 0; u < NV; ++u) {
                    V[ir[u]] = c[u] * v[u][0] - sg[u] * v[u][1];
                    V[it[u]] = sg[u] * v[u][0] + c[u] * v[u][1];
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
}

}  // namespace scb
