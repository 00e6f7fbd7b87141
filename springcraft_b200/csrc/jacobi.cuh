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

template <int NT>
__device__ __forceinline__ double block_sum_nt(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += red[w];
    return t;
}

// Two-sided cyclic Jacobi (round-robin ordering) on the symmetric matrix A
// (order M, leading dimension LD) in shared memory, accumulating V (M rows; V
// must be initialised by the caller).  Called by all NT threads of the CTA.
// cs/sn/pp/qq hold TWO parameter sets of M/2 entries each, red NT/32 doubles.
//
// One round = M/2 disjoint rotations.  A <- J^T A J is applied in ONE pass over
// the 2x2 sub-blocks {p_k,q_k} x {p_l,q_l} (row and column rotation fused: 4
// loads + 4 stores per sub-block).  V <- V J of round r does not feed the
// rotation parameters, so it runs in the same phase as the parameter
// computation of round r+1 (double-buffered parameters): two barriers per
// round, with the long sqrt/divide chain of the parameters hidden.
// cross_only: only the pairs (i, M/2 + j) between the two halves are rotated
// (M/2 rounds); the block-Jacobi driver uses it for pivots whose diagonal
// blocks were treated earlier in the same sweep.
template <int M, int LD, int NT>
__device__ void jacobi_eigen_smem(double* A, double* V, double* cs, double* sn, int* pp, int* qq, double* red,
                                  int max_sweeps = 30, bool cross_only = false) {
    constexpr int half = M / 2;   // M is even
    static_assert((half * half) % NT == 0 && (half * M) % NT == 0, "work must divide over the CTA");
    const int tid = threadIdx.x;
    const int rounds = cross_only ? half : M - 1;

    auto parameters = [&](int step, int buf) {
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
        cs[buf * half + tid] = c; sn[buf * half + tid] = s; pp[buf * half + tid] = p; qq[buf * half + tid] = q;
    };

    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int q = tid; q < M * M; q += NT) {
            const int i = q / M, j = q % M;
            const double v = A[i * LD + j];
            if (i == j) dg += v * v; else off += v * v;
        }
        off = block_sum_nt<NT>(off, red);
        dg = block_sum_nt<NT>(dg, red);
        if (off <= 1e-30 * dg || off == 0.0) break;
        if (tid < half) parameters(0, 0);
        __syncthreads();
        for (int step = 0; step < rounds; ++step) {
            const int o = (step & 1) * half;
            // A <- J^T A J on the 2x2 sub-blocks (k: row pair, l: column pair); loads of all sub-blocks of a
            // thread are issued before the first store so that the shared-memory latency overlaps
            {
                constexpr int NA = (half * half) / NT;
                double a[NA][4], ck[NA], sk[NA], cl[NA], sl[NA];
                int ip[NA], iq[NA], ir[NA], it[NA];
#pragma unroll
                for (int u = 0; u < NA; ++u) {
                    const int w = u * NT + tid;
                    const int k = o + w / half, l = o + w % half;
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
            __syncthreads();
            // parameters of the next round (reads the updated A) || V <- V J of this round
            if (tid < half && step + 1 < rounds) parameters(step + 1, (step + 1) & 1);
            {
                constexpr int NV = (half * M) / NT;
                double v[NV][2], c[NV], sg[NV];
                int ir[NV], it[NV];
#pragma unroll
                for (int u = 0; u < NV; ++u) {
                    const int w = u * NT + tid;
                    const int i = w / half, l = o + w % half;
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
