// Tall-skinny FP64 steps of the subspace iteration on the FP64 tensor cores (mma.sync.m8n8k4.f64), for the
// 32-column blocks of the ensemble solver.  One CTA per structure, so every reduction is CTA-local and runs in a
// fixed order (no atomics: results are reproducible run to run).
//   gram2:        S = X^T X and T = X^T (H X) in ONE pass over X and HX (upper 8x8 blocks only, mirrored)
//   rotate_resid: X <- X C, HX <- HX C and the squared residual norms |HX c_j - theta_j X c_j|^2 of the rotated
//                 pairs from the accumulator fragments (no extra pass over the blocks)
// Fragment layout of m8n8k4 (PTX ISA): A[lane/4][lane%4], B[lane%4][lane/4], C/D[lane/4][2*(lane%4) + {0,1}].
#include "subspace.cuh"

namespace scb {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int kTsWarps = 8;

// upper-triangular 8x8 block pairs of a 32x32 matrix: (0,0)(0,1)(0,2)(0,3)(1,1)(1,2)(1,3)(2,2)(2,3)(3,3)
__global__ void __launch_bounds__(kTsWarps * 32)
gram2_dmma_kernel(int N, const double* __restrict__ X, const double* __restrict__ HX, double* __restrict__ S,
                  double* __restrict__ T, const int32_t* __restrict__ done) {
    __shared__ double buf[20][64];
    const int64_t s = blockIdx.x;
    if (done && done[s]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lr = lane & 3, lc = lane >> 2;
    const double* Xs = X + s * (int64_t)N * 32;
    const double* Hs = HX + s * (int64_t)N * 32;
    double accS[10][2], accT[10][2];
#pragma unroll
    for (int q = 0; q < 10; ++q) { accS[q][0] = accS[q][1] = 0.0; accT[q][0] = accT[q][1] = 0.0; }
    const int nslab = (N + 3) >> 2;
    for (int k = warp; k < nslab; k += kTsWarps) {
        const int row = 4 * k + lr;
        const bool ok = row < N;
        double xf[4], hf[4];
#pragma unroll
        for (int bl = 0; bl < 4; ++bl) {
            xf[bl] = ok ? Xs[(int64_t)row * 32 + 8 * bl + lc] : 0.0;
            hf[bl] = ok ? Hs[(int64_t)row * 32 + 8 * bl + lc] : 0.0;
        }
        int q = 0;
#pragma unroll
        for (int bi = 0; bi < 4; ++bi)
#pragma unroll
            for (int bj = bi; bj < 4; ++bj) {
                dmma884(accS[q][0], accS[q][1], xf[bi], xf[bj]);
                dmma884(accT[q][0], accT[q][1], xf[bi], hf[bj]);
                ++q;
            }
    }
    // fixed-order sum over the warps
    for (int w = 0; w < kTsWarps; ++w) {
        if (warp == w) {
#pragma unroll
            for (int q = 0; q < 10; ++q) {
                double* bs = &buf[q][2 * lane];
                double* bt = &buf[10 + q][2 * lane];
                if (w == 0) { bs[0] = accS[q][0]; bs[1] = accS[q][1]; bt[0] = accT[q][0]; bt[1] = accT[q][1]; }
                else { bs[0] += accS[q][0]; bs[1] += accS[q][1]; bt[0] += accT[q][0]; bt[1] += accT[q][1]; }
            }
        }
        __syncthreads();
    }
    // buf[q][2*lane + e] = block q, element (i = lane/4, j = 2*(lane%4) + e)  ->  full symmetric 32x32 matrices
    double* Ss = S + s * 1024;
    double* Ts = T + s * 1024;
    for (int idx = threadIdx.x; idx < 20 * 64; idx += kTsWarps * 32) {
        const int qq = idx / 64, el = idx % 64;
        const int q = qq % 10;
        int bi = 0, rem = q;
        while (rem >= 4 - bi) { rem -= 4 - bi; ++bi; }
        const int bj = bi + rem;
        const int l = el >> 1, e = el & 1;
        const int i = 8 * bi + (l >> 2), j = 8 * bj + 2 * (l & 3) + e;
        const double v = buf[qq][el];
        double* M = qq < 10 ? Ss : Ts;
        M[i * 32 + j] = v;
        if (bi != bj) M[j * 32 + i] = v;
    }
}

__global__ void __launch_bounds__(kTsWarps * 32)
rotate_resid_dmma_kernel(int N, const double* __restrict__ C, double* X, double* HX, const double* __restrict__ theta,
                         double* __restrict__ rn2, const int32_t* __restrict__ done) {
    __shared__ double red[kTsWarps][32];
    const int64_t s = blockIdx.x;
    if (done && done[s]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lr = lane & 3, lc = lane >> 2;
    const double* Cs = C + s * 1024;
    double* Xs = X + s * (int64_t)N * 32;
    double* Hs = HX + s * (int64_t)N * 32;
    double cf[8][4];          // B fragments of C: C[4 kk + lane%4][8 jb + lane/4]
#pragma unroll
    for (int kk = 0; kk < 8; ++kk)
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) cf[kk][jb] = Cs[(4 * kk + lr) * 32 + 8 * jb + lc];
    double th[4][2], acc[4][2];
#pragma unroll
    for (int jb = 0; jb < 4; ++jb)
#pragma unroll
        for (int e = 0; e < 2; ++e) { th[jb][e] = theta[s * 32 + 8 * jb + 2 * lr + e]; acc[jb][e] = 0.0; }
    const int ntile = (N + 7) >> 3;
    for (int t = warp; t < ntile; t += kTsWarps) {
        const int row = 8 * t + lc;
        const bool ok = row < N;
        double af[8], hf[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            af[kk] = ok ? Xs[(int64_t)row * 32 + 4 * kk + lr] : 0.0;
            hf[kk] = ok ? Hs[(int64_t)row * 32 + 4 * kk + lr] : 0.0;
        }
        __syncwarp();   // every lane of the warp has read its part of the 8 rows before they are overwritten
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) {
            double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                dmma884(d0, d1, af[kk], cf[kk][jb]);
                dmma884(e0, e1, hf[kk], cf[kk][jb]);
            }
            if (ok) {
                *reinterpret_cast<double2*>(Xs + (int64_t)row * 32 + 8 * jb + 2 * lr) = make_double2(d0, d1);
                *reinterpret_cast<double2*>(Hs + (int64_t)row * 32 + 8 * jb + 2 * lr) = make_double2(e0, e1);
            }
            const double r0 = e0 - th[jb][0] * d0, r1 = e1 - th[jb][1] * d1;
            acc[jb][0] = fma(r0, r0, acc[jb][0]);
            acc[jb][1] = fma(r1, r1, acc[jb][1]);
        }
    }
    // columns 8 jb + 2 (lane%4) + e: sum over the 8 row positions of the warp (lanes with equal lane%4), then warps
#pragma unroll
    for (int jb = 0; jb < 4; ++jb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double v = acc[jb][e];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (lc == 0) red[warp][8 * jb + 2 * lr + e] = v;
        }
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = 0.0;
        for (int w = 0; w < kTsWarps; ++w) t += red[w][threadIdx.x];
        rn2[s * 32 + threadIdx.x] = t;
    }
}

int gram2_dmma(int B, int64_t N, const double* X, const double* HX, double* S, double* T, const int32_t* done,
               cudaStream_t st) {
    gram2_dmma_kernel<<<B, kTsWarps * 32, 0, st>>>((int)N, X, HX, S, T, done);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

int rotate_resid_dmma(int B, int64_t N, const double* C, double* X, double* HX, const double* theta, double* rn2,
                      const int32_t* done, cudaStream_t st) {
    rotate_resid_dmma_kernel<<<B, kTsWarps * 32, 0, st>>>((int)N, C, X, HX, theta, rn2, done);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

}  // namespace scb
