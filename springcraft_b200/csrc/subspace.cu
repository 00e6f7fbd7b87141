// K3b: tall-skinny and small dense building blocks of the lowest-k eigensolver:
// Gram matrices, Cholesky-orthonormalisation / Rayleigh-Ritz (two-sided cyclic
// Jacobi in shared memory), basis rotation, deflation of the analytic null
// space, residual norms and the per-structure solver state.  "Only the small
// Rayleigh-Ritz solve is done densely" (BASELINE.json north_star).
#include <stdlib.h>

#include "subspace.cuh"
#include "jacobi.cuh"

namespace scb {

// ---------------------------------------------------------------------------
// Reductions over the rows of a tall block are split over several CTAs per structure.  Every CTA writes its partial
// result to its own slot and a second kernel adds the slots in index order: fixed summation order, no atomicAdd on
// doubles, results reproducible run to run.
// out[s][q] = sum_c part[s][c][q]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sum_chunks_kernel(int64_t per_struct, int chunks, const double* __restrict__ part, double* __restrict__ out,
                  const int32_t* __restrict__ done) {
    const int64_t s = blockIdx.y;
    if (done && done[s]) return;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= per_struct) return;
    const double* p = part + s * chunks * per_struct + q;
    double t = 0.0;
    for (int c = 0; c < chunks; ++c) t += p[(int64_t)c * per_struct];
    out[s * per_struct + q] = t;
}

static int sum_chunks(int B, int64_t per_struct, int chunks, const double* part, double* out, const int32_t* done,
                      cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(per_struct, 256), (unsigned)B);
    sum_chunks_kernel<<<grid, 256, 0, st>>>(per_struct, chunks, part, out, done);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

// ---------------------------------------------------------------------------
// G[s] = A[s]^T B[s]   (A, B: [N][BW] row-major; G: [BW][BW]); partial Gram matrix of the CTA's rows -> its slot
// ---------------------------------------------------------------------------
constexpr int kGramRows = 512;  // rows of the tall matrices reduced by one CTA

template <int BW>
__global__ void __launch_bounds__(256)
gram_kernel(int64_t N, const double* __restrict__ A, const double* __restrict__ Bm, double* __restrict__ G,
            const int32_t* __restrict__ done) {
    constexpr int TT = BW / 16;
    constexpr int R = 16;
    __shared__ double sA[R][BW], sB[R][BW];
    const int s = blockIdx.y;
    if (done && done[s]) return;
    const double* As = A + (int64_t)s * N * BW;
    const double* Bs = Bm + (int64_t)s * N * BW;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t r0 = (int64_t)blockIdx.x * kGramRows;
    const int64_t r1 = min(N, r0 + kGramRows);
    double acc[TT][TT];
#pragma unroll
    for (int i = 0; i < TT; ++i)
#pragma unroll
        for (int j = 0; j < TT; ++j) acc[i][j] = 0.0;
    for (int64_t rb = r0; rb < r1; rb += R) {
        const int rows = (int)min((int64_t)R, r1 - rb);
        __syncthreads();
        for (int q = threadIdx.x; q < R * BW; q += 256) {
            const int r = q / BW, c = q % BW;
            const bool in = r < rows;
            sA[r][c] = in ? As[(rb + r) * BW + c] : 0.0;
            sB[r][c] = in ? Bs[(rb + r) * BW + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double a[TT], b[TT];
#pragma unroll
            for (int i = 0; i < TT; ++i) a[i] = sA[r][ty * TT + i];
#pragma unroll
            for (int j = 0; j < TT; ++j) b[j] = sB[r][tx * TT + j];
#pragma unroll
            for (int i = 0; i < TT; ++i)
#pragma unroll
                for (int j = 0; j < TT; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
    }
    double* Gs = G + ((int64_t)s * gridDim.x + blockIdx.x) * BW * BW;
#pragma unroll
    for (int i = 0; i < TT; ++i)
#pragma unroll
        for (int j = 0; j < TT; ++j) Gs[(ty * TT + i) * BW + tx * TT + j] = acc[i][j];
}

int gram(int B, int64_t N, int b, const double* A, const double* Bm, double* G, const int32_t* done,
         cudaStream_t st) {
    if (b != 32 && b != 64 && b != 128) return SCB_ERR_UNSUPPORTED;
    const int chunks = (int)ceil_div(N, kGramRows);
    dim3 grid((unsigned)chunks, (unsigned)B);
    double* part = G;          // a single chunk writes the result directly
    if (chunks > 1) SCB_TRY(pool_alloc((void**)&part, sizeof(double) * (size_t)B * chunks * b * b, st));
    if (b == 32) gram_kernel<32><<<grid, 256, 0, st>>>(N, A, Bm, part, done);
    else if (b == 64) gram_kernel<64><<<grid, 256, 0, st>>>(N, A, Bm, part, done);
    else gram_kernel<128><<<grid, 256, 0, st>>>(N, A, Bm, part, done);
    SCB_LAUNCH_CHECK();
    if (chunks > 1) {
        const int status = sum_chunks(B, (int64_t)b * b, chunks, part, G, done, st);
        pool_free(part, st);
        return status;
    }
    return SCB_OK;
}

// ---------------------------------------------------------------------------
// small dense: Cholesky of S, optional Rayleigh-Ritz of (T, S) by Jacobi
// ---------------------------------------------------------------------------
// mode 0: C = L^-T               (orthonormalise: X <- X C)
// mode 1: Rayleigh-Ritz of the pencil (T, S): theta ascending, C = L^-T V
template <int BW>
__global__ void __launch_bounds__(256)
rr_kernel(const double* __restrict__ Sg, const double* __restrict__ Tg, double* __restrict__ theta,
          double* __restrict__ Cout, const int32_t* __restrict__ done, int mode, int nact) {
    constexpr int LD = BW + 1;
    extern __shared__ double sm[];
    double* S = sm;
    double* T = S + BW * LD;
    double* V = T + BW * LD;
    __shared__ double cs[BW], sn[BW], red[8], ev[BW];   // two rotation-parameter sets of BW/2
    __shared__ int pp[BW], qq[BW], rank[BW];
    const int s = blockIdx.x;
    if (done && done[s]) return;
    const int tid = threadIdx.x;
    const double* Ss = Sg + (int64_t)s * BW * BW;
    for (int q = tid; q < BW * BW; q += 256) {
        const int i = q / BW, j = q % BW;
        S[i * LD + j] = 0.5 * (Ss[i * BW + j] + Ss[j * BW + i]);
    }
    if (mode == 1) {
        const double* Ts = Tg + (int64_t)s * BW * BW;
        for (int q = tid; q < BW * BW; q += 256) {
            const int i = q / BW, j = q % BW;
            T[i * LD + j] = 0.5 * (Ts[i * BW + j] + Ts[j * BW + i]);
        }
    }
    __syncthreads();
    // trace-based floor for the pivots (rank-deficient blocks stay finite)
    double tr = 0.0;
    for (int q = tid; q < BW; q += 256) tr += S[q * LD + q];
    tr = block_sum_256(tr, red);
    const double floor_piv = 1e-28 * tr + 1e-300;
    // right-looking Cholesky, lower triangle, in place
    for (int k = 0; k < BW; ++k) {
        if (tid == 0) S[k * LD + k] = sqrt(fmax(S[k * LD + k], floor_piv));
        __syncthreads();
        const double d = S[k * LD + k];
        for (int i = k + 1 + tid; i < BW; i += 256) S[i * LD + k] /= d;
        __syncthreads();
        const int rem = BW - 1 - k;
        for (int q = tid; q < rem * rem; q += 256) {
            const int i = k + 1 + q / rem, j = k + 1 + q % rem;
            if (j <= i) S[i * LD + j] -= S[i * LD + k] * S[j * LD + k];
        }
        __syncthreads();
    }
    // V <- L^-1 (lower triangular), column j by thread j
    for (int q = tid; q < BW * BW; q += 256) V[(q / BW) * LD + q % BW] = 0.0;
    __syncthreads();
    if (tid < BW) {
        const int j = tid;
        for (int i = j; i < BW; ++i) {
            double acc = (i == j) ? 1.0 : 0.0;
            for (int k = j; k < i; ++k) acc -= S[i * LD + k] * V[k * LD + j];
            V[i * LD + j] = acc / S[i * LD + i];
        }
    }
    __syncthreads();
    double* Cs = Cout + (int64_t)s * BW * BW;
    if (mode == 0) {
        for (int q = tid; q < BW * BW; q += 256) {
            const int p = q / BW, c = q % BW;
            Cs[q] = (p <= c) ? V[c * LD + p] : 0.0;
        }
        return;
    }
    // S <- Linv * T ; T <- S * Linv^T
    for (int q = tid; q < BW * BW; q += 256) {
        const int i = q / BW, j = q % BW;
        double acc = 0.0;
        for (int k = 0; k <= i; ++k) acc = fma(V[i * LD + k], T[k * LD + j], acc);
        S[i * LD + j] = acc;
    }
    __syncthreads();
    for (int q = tid; q < BW * BW; q += 256) {
        const int i = q / BW, j = q % BW;
        double acc = 0.0;
        for (int k = 0; k <= j; ++k) acc = fma(S[i * LD + k], V[j * LD + k], acc);
        T[i * LD + j] = acc;
    }
    __syncthreads();
    // enforce exact symmetry, keep Linv in S, V <- I
    for (int q = tid; q < BW * BW; q += 256) {
        const int i = q / BW, j = q % BW;
        S[i * LD + j] = V[i * LD + j];
        if (j < i) { const double a = 0.5 * (T[i * LD + j] + T[j * LD + i]); T[i * LD + j] = a; T[j * LD + i] = a; }
    }
    __syncthreads();
    for (int q = tid; q < BW * BW; q += 256) V[(q / BW) * LD + q % BW] = (q / BW == q % BW) ? 1.0 : 0.0;
    __syncthreads();
    jacobi_eigen_smem<BW, LD, 256>(T, V, cs, sn, pp, qq, red);
    // ascending order (ties broken by index)
    if (tid < BW) ev[tid] = T[tid * LD + tid];
    __syncthreads();
    if (tid < BW) {
        // only the leading nact entries are real; padding keeps its position
        int r = tid;
        if (tid < nact) {
            r = 0;
            for (int k = 0; k < nact; ++k) r += (ev[k] < ev[tid]) || (ev[k] == ev[tid] && k < tid);
        }
        rank[tid] = r;
        theta[(int64_t)s * BW + r] = ev[tid];
    }
    __syncthreads();
    // C[:, rank[q]] = Linv^T V[:, q]
    for (int w = tid; w < BW * BW; w += 256) {
        const int p = w / BW, q = w % BW;
        double acc = 0.0;
        for (int k = p; k < BW; ++k) acc = fma(S[k * LD + p], V[k * LD + q], acc);
        Cs[p * BW + rank[q]] = acc;
    }
}

int small_rr(int B, int b, const double* S, const double* T, double* theta, double* C, const int32_t* done,
             int mode, cudaStream_t st, int nact) {
    if (nact <= 0) nact = b;
    if (b == 32) {
        const size_t smem = sizeof(double) * 3 * 32 * 33;
        rr_kernel<32><<<B, 256, smem, st>>>(S, T, theta, C, done, mode, nact);
    } else if (b == 64) {
        const size_t smem = sizeof(double) * 3 * 64 * 65;
        // per-device function attribute: set on every launch (cheap), no process-wide "configured" flag
        SCB_CUDA(cudaFuncSetAttribute(rr_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rr_kernel<64><<<B, 256, smem, st>>>(S, T, theta, C, done, mode, nact);
    } else {
        return SCB_ERR_UNSUPPORTED;
    }
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

// ---------------------------------------------------------------------------
// Xout = Xin * C  (and optionally Yout = Yin * C); C: [b][b]; in place allowed
// ---------------------------------------------------------------------------
template <int BW>
__global__ void __launch_bounds__(256)
rotate_kernel(int64_t N, const double* __restrict__ C, const double* Xin, double* Xout, const double* Yin,
              double* Yout, const int32_t* __restrict__ done) {
    constexpr int CC = BW / 32;
    extern __shared__ double sm[];
    double* sC = sm;                 // [BW][BW]
    double* sx = sm + BW * BW;       // [8][BW]
    const int s = blockIdx.y;
    if (done && done[s]) return;
    const double* Cs = C + (int64_t)s * BW * BW;
    for (int q = threadIdx.x; q < BW * BW; q += 256) sC[q] = Cs[q];
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const int64_t r0 = (int64_t)blockIdx.x * kGramRows;
    const int64_t r1 = min(N, r0 + kGramRows);
    for (int pass = 0; pass < 2; ++pass) {
        const double* in = pass ? Yin : Xin;
        double* out = pass ? Yout : Xout;
        if (!in) continue;
        in += (int64_t)s * N * BW;
        out += (int64_t)s * N * BW;
        for (int64_t r = r0 + warp; r < r1; r += 8) {
            __syncwarp();
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) sx[warp * BW + lane + 32 * cc] = in[r * BW + lane + 32 * cc];
            __syncwarp();
            double acc[CC];
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) acc[cc] = 0.0;
#pragma unroll 8
            for (int p = 0; p < BW; ++p) {
                const double x = sx[warp * BW + p];
#pragma unroll
                for (int cc = 0; cc < CC; ++cc) acc[cc] = fma(x, sC[p * BW + lane + 32 * cc], acc[cc]);
            }
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) out[r * BW + lane + 32 * cc] = acc[cc];
        }
    }
}

int rotate(int B, int64_t N, int b, const double* C, const double* Xin, double* Xout, const double* Yin,
           double* Yout, const int32_t* done, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(N, kGramRows), (unsigned)B);
    const size_t smem = sizeof(double) * ((size_t)b * b + 8 * b);
    if (b == 32) {
        rotate_kernel<32><<<grid, 256, smem, st>>>(N, C, Xin, Xout, Yin, Yout, done);
    } else if (b == 64) {
        rotate_kernel<64><<<grid, 256, smem, st>>>(N, C, Xin, Xout, Yin, Yout, done);
    } else if (b == 128) {
        // per-device function attribute: set on every launch (cheap), no process-wide "configured" flag
        SCB_CUDA(cudaFuncSetAttribute(rotate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rotate_kernel<128><<<grid, 256, smem, st>>>(N, C, Xin, Xout, Yin, Yout, done);
    } else {
        return SCB_ERR_UNSUPPORTED;
    }
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

// ---------------------------------------------------------------------------
// deflation: X <- X - Z (Z^T X),  Z: [N][nz] orthonormal, nz <= 8
// ---------------------------------------------------------------------------
constexpr int kMaxNz = 8;

template <int CC>
__global__ void __launch_bounds__(256)
ztx_kernel(int64_t N, int nz, const double* __restrict__ Z, const double* __restrict__ X, double* __restrict__ P,
           const int32_t* __restrict__ done) {
    constexpr int BW = 32 * CC;
    const int s = blockIdx.y;
    if (done && done[s]) return;
    const double* Zs = Z + (int64_t)s * N * nz;
    const double* Xs = X + (int64_t)s * N * BW;
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const int64_t r0 = (int64_t)blockIdx.x * kGramRows;
    const int64_t r1 = min(N, r0 + kGramRows);
    double acc[kMaxNz][CC];
#pragma unroll
    for (int z = 0; z < kMaxNz; ++z)
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) acc[z][cc] = 0.0;
    for (int64_t r = r0 + warp; r < r1; r += 8) {
        double x[CC];
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) x[cc] = Xs[r * BW + lane + 32 * cc];
#pragma unroll
        for (int z = 0; z < kMaxNz; ++z)
            if (z < nz) {
                const double zv = Zs[r * nz + z];
#pragma unroll
                for (int cc = 0; cc < CC; ++cc) acc[z][cc] = fma(zv, x[cc], acc[z][cc]);
            }
    }
    // warps of the CTA in index order -> the CTA's slot P[s][chunk][z][column]
    __shared__ double red[kMaxNz][BW];
    for (int w = 0; w < 8; ++w) {
        if (warp == w) {
#pragma unroll
            for (int z = 0; z < kMaxNz; ++z)
#pragma unroll
                for (int cc = 0; cc < CC; ++cc) {
                    double* r = &red[z][lane + 32 * cc];
                    *r = (w == 0 ? 0.0 : *r) + acc[z][cc];
                }
        }
        __syncthreads();
    }
    double* Ps = P + ((int64_t)s * gridDim.x + blockIdx.x) * kMaxNz * BW;
    for (int q = threadIdx.x; q < kMaxNz * BW; q += 256) Ps[q] = red[q / BW][q % BW];
}

template <int CC>
__global__ void __launch_bounds__(256)
subz_kernel(int64_t N, int nz, const double* __restrict__ Z, const double* __restrict__ P, double* __restrict__ X,
            const int32_t* __restrict__ done) {
    constexpr int BW = 32 * CC;
    const int s = blockIdx.y;
    if (done && done[s]) return;
    const double* Zs = Z + (int64_t)s * N * nz;
    double* Xs = X + (int64_t)s * N * BW;
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    double p[kMaxNz][CC];
#pragma unroll
    for (int z = 0; z < kMaxNz; ++z)
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) p[z][cc] = (z < nz) ? P[((int64_t)s * kMaxNz + z) * BW + lane + 32 * cc] : 0.0;
    const int64_t r0 = (int64_t)blockIdx.x * kGramRows;
    const int64_t r1 = min(N, r0 + kGramRows);
    for (int64_t r = r0 + warp; r < r1; r += 8) {
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
            double v = Xs[r * BW + lane + 32 * cc];
#pragma unroll
            for (int z = 0; z < kMaxNz; ++z)
                if (z < nz) v = fma(-Zs[r * nz + z], p[z][cc], v);
            Xs[r * BW + lane + 32 * cc] = v;
        }
    }
}

int deflate(int B, int64_t N, int b, int nz, const double* Z, double* X, double* P, const int32_t* done,
            cudaStream_t st) {
    if (nz <= 0 || !Z) return SCB_OK;
    if (nz > kMaxNz) return SCB_ERR_UNSUPPORTED;
    if (b != 32 && b != 64 && b != 128) return SCB_ERR_UNSUPPORTED;
    const int chunks = (int)ceil_div(N, kGramRows);
    dim3 grid((unsigned)chunks, (unsigned)B);
    double* part = P;
    if (chunks > 1) SCB_TRY(pool_alloc((void**)&part, sizeof(double) * (size_t)B * chunks * kMaxNz * b, st));
    if (b == 32) ztx_kernel<1><<<grid, 256, 0, st>>>(N, nz, Z, X, part, done);
    else if (b == 64) ztx_kernel<2><<<grid, 256, 0, st>>>(N, nz, Z, X, part, done);
    else ztx_kernel<4><<<grid, 256, 0, st>>>(N, nz, Z, X, part, done);
    SCB_LAUNCH_CHECK();
    if (chunks > 1) {
        const int status = sum_chunks(B, (int64_t)kMaxNz * b, chunks, part, P, done, st);
        pool_free(part, st);
        if (status != SCB_OK) return status;
    }
    if (b == 32) subz_kernel<1><<<grid, 256, 0, st>>>(N, nz, Z, P, X, done);
    else if (b == 64) subz_kernel<2><<<grid, 256, 0, st>>>(N, nz, Z, P, X, done);
    else subz_kernel<4><<<grid, 256, 0, st>>>(N, nz, Z, P, X, done);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

// ---------------------------------------------------------------------------
// squared residual norms  rn2[s][q] = || HX[:,q] - theta_q X[:,q] ||^2
// ---------------------------------------------------------------------------
template <int CC>
__global__ void __launch_bounds__(256)
resid_kernel(int64_t N, const double* __restrict__ X, const double* __restrict__ HX, const double* __restrict__ theta,
             double* __restrict__ rn2, const int32_t* __restrict__ done) {
    constexpr int BW = 32 * CC;
    __shared__ double red[8][BW];
    const int s = blockIdx.y;
    if (done && done[s]) return;
    const double* Xs = X + (int64_t)s * N * BW;
    const double* Hs = HX + (int64_t)s * N * BW;
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    double th[CC], acc[CC];
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) { th[cc] = theta[(int64_t)s * BW + lane + 32 * cc]; acc[cc] = 0.0; }
    const int64_t r0 = (int64_t)blockIdx.x * kGramRows;
    const int64_t r1 = min(N, r0 + kGramRows);
    for (int64_t r = r0 + warp; r < r1; r += 8)
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
            const double d = Hs[r * BW + lane + 32 * cc] - th[cc] * Xs[r * BW + lane + 32 * cc];
            acc[cc] = fma(d, d, acc[cc]);
        }
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) red[warp][lane + 32 * cc] = acc[cc];
    __syncthreads();
    for (int c = threadIdx.x; c < BW; c += 256) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w][c];
        rn2[((int64_t)s * gridDim.x + blockIdx.x) * BW + c] = t;
    }
}

int residual_norms(int B, int64_t N, int b, const double* X, const double* HX, const double* theta, double* rn2,
                   const int32_t* done, cudaStream_t st) {
    // rn2 of converged structures survives: their CTAs and their final sums are skipped
    if (b != 32 && b != 64 && b != 128) return SCB_ERR_UNSUPPORTED;
    const int chunks = (int)ceil_div(N, kGramRows);
    dim3 grid((unsigned)chunks, (unsigned)B);
    double* part = rn2;
    if (chunks > 1) SCB_TRY(pool_alloc((void**)&part, sizeof(double) * (size_t)B * chunks * b, st));
    if (b == 32) resid_kernel<1><<<grid, 256, 0, st>>>(N, X, HX, theta, part, done);
    else if (b == 64) resid_kernel<2><<<grid, 256, 0, st>>>(N, X, HX, theta, part, done);
    else resid_kernel<4><<<grid, 256, 0, st>>>(N, X, HX, theta, part, done);
    SCB_LAUNCH_CHECK();
    if (chunks > 1) {
        const int status = sum_chunks(B, b, chunks, part, rn2, done, st);
        pool_free(part, st);
        return status;
    }
    return SCB_OK;
}

// ---------------------------------------------------------------------------
// solver state
// ---------------------------------------------------------------------------
__global__ void state_init_kernel(int B, const double* __restrict__ gersh, EigState* st, int32_t* done,
                                  int32_t* n_active, int32_t* skip32, int32_t* skip64, int allow32, int degree) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) *n_active = B;
    if (s >= B) return;
    EigState e;
    e.ub = gersh[s] * (1.0 + 1e-10) + 1e-300;
    e.ub_safe = e.ub;
    e.prev_res = 0.0;
    e.lo = 0.0;
    e.a0 = 0.0;
    e.iters = 0;
    e.converged = 0;
    e.degree_next = degree;
    e.degree_used = degree;
    st[s] = e;
    done[s] = 0;
    // filter precision of the next outer iteration: FP32 first (when allowed), FP64 later
    skip32[s] = allow32 ? 0 : 1;
    skip64[s] = allow32 ? 1 : 0;
}

__global__ void zero_active_rn2_kernel(int B, int b, double* rn2, const int32_t* done) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B * b) return;
    if (!done[q / b]) rn2[q] = 0.0;
}

// after a Rayleigh-Ritz step: convergence test, new filter bounds
__global__ void state_update_kernel(int B, int b, int k, double tol, const double* __restrict__ theta,
                                    const double* __restrict__ rn2, EigState* st, int32_t* done,
                                    int32_t* n_active, double* __restrict__ resid, int32_t* skip32,
                                    int32_t* skip64, int allow32, double switch_tol, int degree) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B || done[s]) return;
    const double* th = theta + (int64_t)s * b;
    EigState e = st[s];
    e.iters += 1;
    double worst = 0.0;
    // The scale must not collapse when the deflated operator still has zero modes among the wanted ones
    // (atoms isolated by contact_shutdown, disconnected chains): theta_k ~ 0 there, and a purely relative test
    // could never pass.  1e-6 * ub is far above the attainable residual (~1e-14 * ub) and far below any
    // non-trivial eigenvalue of a connected network.
    const double scale = fmax(fabs(th[k - 1]), 1e-6 * e.ub);
    bool finite = true;   // a NaN Ritz pair must never pass the convergence test (fmax drops NaNs)
    for (int q = 0; q < b; ++q) {
        const double r = sqrt(rn2[(int64_t)s * b + q]);
        resid[(int64_t)s * b + q] = r;
        if (q < k) {
            worst = fmax(worst, r);
            finite = finite && (r == r) && (th[q] == th[q]);
        }
    }
    e.a0 = th[0];
    // repair a too small (estimated) upper bound: after a filter pass the largest Ritz value of the block
    // must sit far below ub; if it does not, the estimate was violated -> fall back towards the safe bound
    if (e.iters > 1 && e.ub < e.ub_safe && th[b - 1] > 0.5 * e.ub) e.ub = fmin(e.ub_safe, 1.15 * fmax(e.ub, th[b - 1]));
    double lo = th[b - 1];
    // keep the damped interval [lo, ub] non-degenerate
    lo = fmin(lo, 0.98 * e.ub);
    if (!(lo > e.a0)) lo = e.a0 + 0.5 * (e.ub - e.a0);
    e.lo = lo;
    bool finished = false;
    if (finite && worst <= tol * scale) {
        e.converged = 1;
        done[s] = 1;
        finished = true;
        atomicSub(n_active, 1);
    }
    // The FP32 filter stalls at a residual of ~5e-8 * ub (2e-5 on the C3 eANM matrices, 2e-4 on the stiff
    // sdENM ones): hand over to FP64 at switch_tol * ub, i.e. with a ~50x margin, and also as soon as an FP32
    // iteration fails to halve the residual.
    const bool stagnating = e.prev_res > 0.0 && worst > 0.5 * e.prev_res;
    const bool use32 = allow32 && !finished && worst > switch_tol * e.ub && !stagnating;
    // Last iterations: a full-degree filter overshoots the tolerance.  From the reduction per degree observed
    // in the iteration just finished, pick the smallest degree that is predicted to reach 0.3 * tol.
    int next_degree = degree;
    if (!use32 && !finished && e.prev_res > 0.0 && worst < 0.5 * e.prev_res && e.iters > 2) {
        const double log_rho = log(worst / e.prev_res) / (double)max(e.degree_used, 1);   // < 0
        const double need = log(0.3 * tol * scale / worst);                                 // < 0
        const int d = (int)ceil(need / log_rho) + 1;
        next_degree = min(degree, max(8, d));
    }
    // Conditioning guard: a degree-m filter amplifies the lowest Ritz direction by T_m(x0), x0 = (c - a0) / half,
    // against the direction at `lo`.  When the block spans a large part of the spectrum (small systems) that
    // ratio exceeds 1/eps and the filtered block loses rank; keep it below ~1e9 (no effect on protein-sized
    // systems, where x0 - 1 is a few percent and the cap is > 50).
    if (!finished && e.iters >= 2) {
        const double half = 0.5 * (e.ub - lo), c = 0.5 * (e.ub + lo);
        const double x0 = (c - e.a0) / half;
        if (x0 > 1.0) {
            const double growth = log(x0 + sqrt(x0 * x0 - 1.0));   // acosh(x0); T_m(x0) ~ exp(m * growth) / 2
            const int cap = (int)fmin(1e6, floor(log(2e9) / growth));
            next_degree = min(next_degree, max(4, cap));
        }
    }
    e.degree_used = next_degree;   // the coming iteration runs next_degree steps
    e.degree_next = next_degree;
    e.prev_res = worst;
    skip32[s] = (finished || !use32) ? 1 : 0;
    skip64[s] = (finished || use32) ? 1 : 0;
    st[s] = e;
}

// Chebyshev coefficients of every filter step: coef[s][d][0..2] = alpha, c, beta
__global__ void cheb_coef_kernel(int B, int degree, const EigState* __restrict__ st, const int32_t* __restrict__ done,
                                 double* __restrict__ coef) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B || done[s]) return;
    const EigState e = st[s];
    const double half = 0.5 * (e.ub - e.lo), c = 0.5 * (e.ub + e.lo);
    const double sigma1 = half / (e.a0 - c);
    double sigma = sigma1;
    double* cf = coef + (int64_t)s * degree * 3;
    // a structure that needs only deg_s < degree steps idles through the first (degree - deg_s) launches:
    // alpha == 0 tells the SpMM kernel to copy its rows forward (Y = X) so that buffer parity stays common
    const int deg_s = min(max(e.degree_next, 2), degree);
    const int start = degree - deg_s;
    for (int d = 0; d < start; ++d) { cf[3 * d] = 0.0; cf[3 * d + 1] = 0.0; cf[3 * d + 2] = 0.0; }
    cf[3 * start] = sigma1 / half; cf[3 * start + 1] = c; cf[3 * start + 2] = 0.0;
    for (int d = start + 1; d < degree; ++d) {
        const double sigma2 = 1.0 / (2.0 / sigma1 - sigma);
        cf[3 * d] = 2.0 * sigma2 / half;
        cf[3 * d + 1] = c;
        cf[3 * d + 2] = sigma * sigma2;
        sigma = sigma2;
    }
}

// deterministic pseudo-random start block
__global__ void rand_init_kernel(int64_t total, uint64_t seed, double* __restrict__ X) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < total) X[q] = uniform_pm1(seed, (uint64_t)q);
}

__global__ void gather_results_kernel(int B, int b, const double* __restrict__ theta, const EigState* __restrict__ st,
                                      double* __restrict__ eigval, int32_t* __restrict__ iters) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < B * b) eigval[q] = theta[q];
    if (q < B) iters[q] = st[q].converged ? st[q].iters : -st[q].iters;
}

int state_init(int B, const double* gersh, EigState* st, int32_t* done, int32_t* n_active, int32_t* skip32,
               int32_t* skip64, int allow32, int degree, cudaStream_t s) {
    state_init_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, s>>>(B, gersh, st, done, n_active, skip32, skip64, allow32,
                                                                 degree);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}
int zero_active_rn2(int B, int b, double* rn2, const int32_t* done, cudaStream_t s) {
    zero_active_rn2_kernel<<<(unsigned)ceil_div((int64_t)B * b, 256), 256, 0, s>>>(B, b, rn2, done);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}
int state_update(int B, int b, int k, double tol, const double* theta, const double* rn2, EigState* st,
                 int32_t* done, int32_t* n_active, double* resid, int32_t* skip32, int32_t* skip64, int allow32,
                 double switch_tol, int degree, cudaStream_t s) {
    state_update_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, s>>>(B, b, k, tol, theta, rn2, st, done, n_active, resid,
                                                                  skip32, skip64, allow32, switch_tol, degree);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}
int cheb_coef(int B, int degree, const EigState* st, const int32_t* done, double* coef, cudaStream_t s) {
    cheb_coef_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, s>>>(B, degree, st, done, coef);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}
int rand_init(int64_t total, uint64_t seed, double* X, cudaStream_t s) {
    rand_init_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(total, seed, X);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}
int gather_results(int B, int b, const double* theta, const EigState* st, double* eigval, int32_t* iters,
                   cudaStream_t s) {
    gather_results_kernel<<<(unsigned)ceil_div((int64_t)B * b, 256), 256, 0, s>>>(B, b, theta, st, eigval, iters);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

// ---------------------------------------------------------------------------
// Spectrum upper bound: k steps of column-wise Lanczos on the block (every one of
// the b columns is an independent Lanczos run started from a random vector).
// The largest Ritz value over all columns under-estimates lambda_max by a few
// percent after ~10 steps; the solver uses min(Gershgorin, 1.05 * estimate) and
// repairs the bound if a filtered block ever shows that it was too small.
// ---------------------------------------------------------------------------
// out[s][c] += sum_r A[r][c] * B[r][c]
template <int CC>
__global__ void __launch_bounds__(256)
coldot_kernel(int64_t N, const double* __restrict__ A, const double* __restrict__ Bm, double* __restrict__ out) {
    constexpr int BW = 32 * CC;
    __shared__ double red[8][BW];
    const int s = blockIdx.y;
    const double* As = A + (int64_t)s * N * BW;
    const double* Bs = Bm + (int64_t)s * N * BW;
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    double acc[CC];
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) acc[cc] = 0.0;
    const int64_t r0 = (int64_t)blockIdx.x * kGramRows;
    const int64_t r1 = min(N, r0 + kGramRows);
    for (int64_t r = r0 + warp; r < r1; r += 8)
#pragma unroll
        for (int cc = 0; cc < CC; ++cc)
            acc[cc] = fma(As[r * BW + lane + 32 * cc], Bs[r * BW + lane + 32 * cc], acc[cc]);
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) red[warp][lane + 32 * cc] = acc[cc];
    __syncthreads();
    for (int c = threadIdx.x; c < BW; c += 256) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w][c];
        out[((int64_t)s * gridDim.x + blockIdx.x) * BW + c] = t;
    }
}

// mode 0: W <- W - alpha[c] V - beta_prev[c] Vprev          (alpha = dot(V,W) just computed)
// mode 1: Vprev <- V ; V <- W / sqrt(nrm2[c])               (nrm2 = ||W||^2 just computed)
// mode 2: V <- V / sqrt(nrm2[c])                            (normalise the start block)
template <int CC>
__global__ void __launch_bounds__(256)
lanczos_axpy_kernel(int64_t N, int mode, double* __restrict__ V, double* __restrict__ Vprev, double* __restrict__ Wm,
                    const double* __restrict__ alpha, const double* __restrict__ beta_prev,
                    const double* __restrict__ nrm2) {
    constexpr int BW = 32 * CC;
    const int s = blockIdx.y;
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const int64_t off = (int64_t)s * N * BW;
    const int64_t r0 = (int64_t)blockIdx.x * kGramRows;
    const int64_t r1 = min(N, r0 + kGramRows);
    double a[CC], bp[CC], inv[CC];
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) {
        const int64_t q = (int64_t)s * BW + lane + 32 * cc;
        a[cc] = alpha ? alpha[q] : 0.0;
        bp[cc] = beta_prev ? sqrt(fmax(beta_prev[q], 0.0)) : 0.0;
        inv[cc] = nrm2 ? rsqrt(fmax(nrm2[q], 1e-300)) : 0.0;
    }
    for (int64_t r = r0 + warp; r < r1; r += 8)
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
            const int64_t idx = off + r * BW + lane + 32 * cc;
            if (mode == 0) {
                Wm[idx] = Wm[idx] - a[cc] * V[idx] - bp[cc] * Vprev[idx];
            } else if (mode == 1) {
                Vprev[idx] = V[idx];
                V[idx] = Wm[idx] * inv[cc];
            } else {
                V[idx] = V[idx] * inv[cc];
            }
        }
}

// largest eigenvalue of each column's k x k Lanczos tridiagonal (bisection on the
// Sturm count), max over columns -> tightened upper bound in the solver state
__global__ void lanczos_bound_kernel(int B, int b, int k, const double* __restrict__ alpha,
                                     const double* __restrict__ beta2, EigState* st, double ub_factor,
                                     double* __restrict__ plain_out) {
    // alpha[j][s][c], beta2[j][s][c] = beta_j^2 (coupling between steps j and j+1)
    const int s = blockIdx.x;
    const int c = threadIdx.x;
    __shared__ double best[128];
    double th = 0.0;
    if (c < b) {
        double lo = 0.0, hi = 0.0;
        for (int j = 0; j < k; ++j) {  // Gershgorin interval of T
            const double a = alpha[((int64_t)j * B + s) * b + c];
            const double bl = j > 0 ? sqrt(fmax(beta2[((int64_t)(j - 1) * B + s) * b + c], 0.0)) : 0.0;
            const double br = j < k - 1 ? sqrt(fmax(beta2[((int64_t)j * B + s) * b + c], 0.0)) : 0.0;
            hi = (j == 0) ? a + bl + br : fmax(hi, a + bl + br);
            lo = (j == 0) ? a - bl - br : fmin(lo, a - bl - br);
        }
        // bisection: number of eigenvalues < x from the signs of the LDL^T pivots
        for (int it = 0; it < 60; ++it) {
            const double x = 0.5 * (lo + hi);
            int below = 0;
            double d = 1.0;
            for (int j = 0; j < k; ++j) {
                const double a = alpha[((int64_t)j * B + s) * b + c];
                const double b2 = j > 0 ? fmax(beta2[((int64_t)(j - 1) * B + s) * b + c], 0.0) : 0.0;
                d = (a - x) - (j > 0 ? b2 / d : 0.0);
                if (d == 0.0) d = -1e-300;
                below += d < 0.0;
            }
            if (below >= k) hi = x; else lo = x;  // all k eigenvalues below x -> x is above theta_max
        }
        th = hi;
    }
    best[c] = th;
    __syncthreads();
    if (c == 0) {
        double m = 0.0;
        for (int q = 0; q < b; ++q) m = fmax(m, best[q]);
        if (plain_out) { plain_out[s] = ub_factor * m; return; }
        EigState e = st[s];
        const double est = ub_factor * m;
        if (est > 0.0 && est < e.ub) e.ub = est;
        st[s] = e;
    }
}

int coldot(int B, int64_t N, int b, const double* A, const double* Bm, double* out, cudaStream_t st) {
    if (b != 32 && b != 64 && b != 128) return SCB_ERR_UNSUPPORTED;
    const int chunks = (int)ceil_div(N, kGramRows);
    dim3 grid((unsigned)chunks, (unsigned)B);
    double* part = out;
    if (chunks > 1) SCB_TRY(pool_alloc((void**)&part, sizeof(double) * (size_t)B * chunks * b, st));
    if (b == 32) coldot_kernel<1><<<grid, 256, 0, st>>>(N, A, Bm, part);
    else if (b == 64) coldot_kernel<2><<<grid, 256, 0, st>>>(N, A, Bm, part);
    else coldot_kernel<4><<<grid, 256, 0, st>>>(N, A, Bm, part);
    SCB_LAUNCH_CHECK();
    if (chunks > 1) {
        const int status = sum_chunks(B, b, chunks, part, out, nullptr, st);
        pool_free(part, st);
        return status;
    }
    return SCB_OK;
}

int lanczos_axpy(int B, int64_t N, int b, int mode, double* V, double* Vprev, double* W, const double* alpha,
                 const double* beta_prev, const double* nrm2, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(N, kGramRows), (unsigned)B);
    if (b == 32) lanczos_axpy_kernel<1><<<grid, 256, 0, st>>>(N, mode, V, Vprev, W, alpha, beta_prev, nrm2);
    else if (b == 64) lanczos_axpy_kernel<2><<<grid, 256, 0, st>>>(N, mode, V, Vprev, W, alpha, beta_prev, nrm2);
    else if (b == 128) lanczos_axpy_kernel<4><<<grid, 256, 0, st>>>(N, mode, V, Vprev, W, alpha, beta_prev, nrm2);
    else return SCB_ERR_UNSUPPORTED;
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

int lanczos_bound(int B, int b, int k, const double* alpha, const double* beta2, EigState* state, cudaStream_t st) {
    double ub_factor = 1.03;
    if (const char* env = getenv("SCB_UBFACTOR")) ub_factor = atof(env) > 1.0 ? atof(env) : ub_factor;
    lanczos_bound_kernel<<<B, 128, 0, st>>>(B, b, k, alpha, beta2, state, ub_factor, nullptr);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

int lanczos_bound_plain(int B, int b, int k, const double* alpha, const double* beta2, double factor, double* out,
                        cudaStream_t st) {
    if (b > 128) return SCB_ERR_UNSUPPORTED;
    lanczos_bound_kernel<<<B, 128, 0, st>>>(B, b, k, alpha, beta2, nullptr, factor, out);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}


// ---------------------------------------------------------------------------
// analytic null space: translations + rotations (D=3) or the constant vector
// (D=1), scaled by sqrt(m) for mass-weighted operators, orthonormalised.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void rigid_row(int a, const double r[3], double w, double out[6]) {
    // row (3i+a) of the raw basis [T_x T_y T_z R_x R_y R_z]; R_b = e_b x r
    out[0] = (a == 0) ? w : 0.0;
    out[1] = (a == 1) ? w : 0.0;
    out[2] = (a == 2) ? w : 0.0;
    // e_x x r = (0,-rz,ry); e_y x r = (rz,0,-rx); e_z x r = (-ry,rx,0)
    const double rx[3] = {0.0, -r[2], r[1]};
    const double ry[3] = {r[2], 0.0, -r[0]};
    const double rz[3] = {-r[1], r[0], 0.0};
    out[3] = w * rx[a];
    out[4] = w * ry[a];
    out[5] = w * rz[a];
}

__global__ void __launch_bounds__(256)
rigid_basis_kernel(int D, const double* __restrict__ xyz, int n, const double* __restrict__ masses,
                   double* __restrict__ Z) {
    __shared__ double red[8];
    __shared__ double G[36], Li[36], cen[3];
    const int s = blockIdx.x;
    const double* X = xyz + (size_t)s * 3 * n;
    const int tid = threadIdx.x;
    if (D == 1) {
        double acc = 0.0;
        for (int i = tid; i < n; i += 256) acc += masses ? masses[i] : 1.0;
        acc = block_sum_256(acc, red);
        const double inv = 1.0 / sqrt(acc);
        for (int i = tid; i < n; i += 256) Z[(size_t)s * n + i] = (masses ? sqrt(masses[i]) : 1.0) * inv;
        return;
    }
    for (int a = 0; a < 3; ++a) {
        double acc = 0.0;
        for (int i = tid; i < n; i += 256) acc += X[(size_t)a * n + i];
        acc = block_sum_256(acc, red);
        if (tid == 0) cen[a] = acc / n;
    }
    __syncthreads();
    // Gram matrix of the raw basis
    double g[21];
    for (int q = 0; q < 21; ++q) g[q] = 0.0;
    for (int i = tid; i < n; i += 256) {
        const double r[3] = {X[i] - cen[0], X[n + i] - cen[1], X[2 * (size_t)n + i] - cen[2]};
        const double w = masses ? sqrt(masses[i]) : 1.0;
        for (int a = 0; a < 3; ++a) {
            double row[6];
            rigid_row(a, r, w, row);
            int q = 0;
            for (int p = 0; p < 6; ++p)
                for (int c = 0; c <= p; ++c) g[q++] += row[p] * row[c];
        }
    }
    for (int q = 0; q < 21; ++q) {
        const double v = block_sum_256(g[q], red);
        if (tid == 0) {
            // unpack lower-triangular index
            int p = 0, acc = 0;
            while (acc + p + 1 <= q) { acc += p + 1; ++p; }
            const int c = q - acc;
            G[p * 6 + c] = v;
            G[c * 6 + p] = v;
        }
    }
    __syncthreads();
    if (tid == 0) {
        // Cholesky G = L L^T, then Li = L^-1
        double L[36];
        for (int q = 0; q < 36; ++q) L[q] = 0.0;
        double tr = 0.0;
        for (int p = 0; p < 6; ++p) tr += G[p * 6 + p];
        for (int j = 0; j < 6; ++j) {
            double d = G[j * 6 + j];
            for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
            d = sqrt(fmax(d, 1e-24 * tr + 1e-300));
            L[j * 6 + j] = d;
            for (int i = j + 1; i < 6; ++i) {
                double v = G[i * 6 + j];
                for (int k = 0; k < j; ++k) v -= L[i * 6 + k] * L[j * 6 + k];
                L[i * 6 + j] = v / d;
            }
        }
        for (int j = 0; j < 6; ++j)
            for (int i = 0; i < 6; ++i) {
                if (i < j) { Li[i * 6 + j] = 0.0; continue; }
                double v = (i == j) ? 1.0 : 0.0;
                for (int k = j; k < i; ++k) v -= L[i * 6 + k] * Li[k * 6 + j];
                Li[i * 6 + j] = v / L[i * 6 + i];
            }
    }
    __syncthreads();
    // Z = raw * L^-T  ->  Z[:, c] = sum_p raw[:, p] * Li[c][p]
    double* Zs = Z + (size_t)s * 3 * n * 6;
    for (int i = tid; i < n; i += 256) {
        const double r[3] = {X[i] - cen[0], X[n + i] - cen[1], X[2 * (size_t)n + i] - cen[2]};
        const double w = masses ? sqrt(masses[i]) : 1.0;
        for (int a = 0; a < 3; ++a) {
            double row[6];
            rigid_row(a, r, w, row);
            for (int c = 0; c < 6; ++c) {
                double v = 0.0;
                for (int p = 0; p <= c; ++p) v += row[p] * Li[c * 6 + p];
                Zs[((size_t)3 * i + a) * 6 + c] = v;
            }
        }
    }
}

}  // namespace scb

extern "C" int scb_rigid_basis(int D, const double* xyz, int B, int n, const double* masses, double* Z,
                               void* stream) {
    if (!xyz || !Z || B < 1 || n < 1 || (D != 1 && D != 3)) return SCB_ERR_INVALID;
    scb::rigid_basis_kernel<<<B, 256, 0, scb::as_stream(stream)>>>(D, xyz, n, masses, Z);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}
