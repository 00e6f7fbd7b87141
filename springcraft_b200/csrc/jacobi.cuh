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
// (leading dimension LD, order m) in shared memory, accumulating V (V must be
// initialised by the caller).  Called by all 256 threads of the CTA.
template <int LD>
__device__ void jacobi_eigen_smem(double* A, double* V, int m, double* cs, double* sn, int* pp, int* qq,
                                  double* red, int vrows, int max_sweeps = 30) {
    const int half = m / 2;       // m is even
    const int tid = threadIdx.x;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int q = tid; q < m * m; q += 256) {
            const int i = q / m, j = q % m;
            const double v = A[i * LD + j];
            if (i == j) dg += v * v; else off += v * v;
        }
        off = block_sum_256(off, red);
        dg = block_sum_256(dg, red);
        if (off <= 1e-30 * dg || off == 0.0) break;
        for (int step = 0; step < m - 1; ++step) {
            if (tid < half) {
                int p, q;
                if (tid == 0) { p = m - 1; q = step % (m - 1); }
                else { p = (step + tid) % (m - 1); q = (step - tid + (m - 1)) % (m - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                const double apq = A[p * LD + q];
                double c = 1.0, s = 0.0;
                if (apq != 0.0) {
                    const double app = A[p * LD + p], aqq = A[q * LD + q];
                    const double tau = (aqq - app) / (2.0 * apq);
                    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c = 1.0 / sqrt(1.0 + t * t);
                    s = t * c;
                }
                cs[tid] = c; sn[tid] = s; pp[tid] = p; qq[tid] = q;
            }
            __syncthreads();
            // rows: A <- J^T A
            for (int w = tid; w < half * m; w += 256) {
                const int k = w / m, j = w % m;
                const double c = cs[k], s = sn[k];
                if (s != 0.0) {
                    const int p = pp[k], q = qq[k];
                    const double ap = A[p * LD + j], aq = A[q * LD + j];
                    A[p * LD + j] = c * ap - s * aq;
                    A[q * LD + j] = s * ap + c * aq;
                }
            }
            __syncthreads();
            // columns: A <- A J ; V <- V J
            for (int w = tid; w < half * (m + vrows); w += 256) {
                const int k = w / (m + vrows), i = w % (m + vrows);
                const double c = cs[k], s = sn[k];
                if (s != 0.0) {
                    const int p = pp[k], q = qq[k];
                    double* M = (i < m) ? (A + i * LD) : (V + (i - m) * LD);
                    const double ap = M[p], aq = M[q];
                    M[p] = c * ap - s * aq;
                    M[q] = s * ap + c * aq;
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
}


}  // namespace scb
