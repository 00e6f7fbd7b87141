// K3a: BSR (DxD blocks) sparse matrix times block vector, fused with the
// Chebyshev three-term recurrence:
//     Y = alpha * (H X - cshift * X) - beta * W
// (alpha, cshift, beta) are per-structure device scalars so that one launch
// advances the filter of every structure of an ensemble; coef == nullptr gives
// the plain product Y = H X.  Replaces the `mech_matrix @ v` products buried in
// LAPACK behind np.linalg.eigh (nma.py:61).
//
// Mapping: one warp per block row; lane l owns columns l, l+32, ... of the
// block vector (row-major X[N][b], so every X-row load is a coalesced 256-byte
// segment per 32 columns).  The row's blocks and column indices are staged
// through shared memory in chunks of 32 contacts with coalesced loads and read
// back as warp-wide broadcasts (padded to 10 doubles per 3x3 block so the
// broadcasts are 16-byte aligned LDS.128).
#include "common.cuh"

namespace scb {

template <int D, int C>
__global__ void __launch_bounds__(256)
spmm_kernel(int n, int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
            const double* __restrict__ offdiag, const double* __restrict__ diag,
            const double* __restrict__ X, const double* __restrict__ W, double* __restrict__ Y,
            const double* __restrict__ coef, int coef_stride, const int32_t* __restrict__ done) {
    constexpr int DD = D * D;
    constexpr int BS = (D == 3) ? 10 : 1;  // padded block stride in shared memory
    constexpr int b = 32 * C;
    __shared__ __align__(16) double sblk[8][32 * BS];
    __shared__ int32_t scol[8][32];

    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const int64_t row = (int64_t)blockIdx.x * 8 + warp;
    if (row >= nrows) return;
    const int64_t s = row / n;
    if (done && done[s]) return;
    const int i = (int)(row % n);
    const int64_t N = (int64_t)D * n;
    const double* Xs = X + s * N * b;

    double acc[D][C], xi[D][C];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int cc = 0; cc < C; ++cc) {
            xi[a][cc] = Xs[((int64_t)D * i + a) * b + lane + 32 * cc];
            acc[a][cc] = 0.0;
        }
    {
        double dg[DD];
#pragma unroll
        for (int q = 0; q < DD; ++q) dg[q] = diag[row * DD + q];
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int c = 0; c < D; ++c)
#pragma unroll
                for (int cc = 0; cc < C; ++cc) acc[a][cc] = fma(dg[a * D + c], xi[c][cc], acc[a][cc]);
    }

    const int64_t pb = rowptr[row], pe = rowptr[row + 1];
    for (int64_t p0 = pb; p0 < pe; p0 += 32) {
        const int cnt = (int)min((int64_t)32, pe - p0);
        __syncwarp();
        if ((int)lane < cnt) scol[warp][lane] = col[p0 + lane];
        if (D == 3) {
            const double* src = offdiag + p0 * DD;
            for (int q = lane; q < cnt * DD; q += 32) sblk[warp][(q / DD) * BS + (q % DD)] = src[q];
        } else {
            if ((int)lane < cnt) sblk[warp][lane] = offdiag[p0 + lane];
        }
        __syncwarp();
#pragma unroll 2
        for (int t = 0; t < cnt; ++t) {
            const int j = scol[warp][t];
            double xj[D][C];
#pragma unroll
            for (int c = 0; c < D; ++c)
#pragma unroll
                for (int cc = 0; cc < C; ++cc) xj[c][cc] = Xs[((int64_t)D * j + c) * b + lane + 32 * cc];
            if (D == 3) {
                const double2* bp = reinterpret_cast<const double2*>(&sblk[warp][t * BS]);
                const double2 b01 = bp[0], b23 = bp[1], b45 = bp[2], b67 = bp[3];
                const double b8 = sblk[warp][t * BS + 8];
                const double h[9] = {b01.x, b01.y, b23.x, b23.y, b45.x, b45.y, b67.x, b67.y, b8};
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int cc = 0; cc < C; ++cc) acc[a][cc] = fma(h[a * 3 + c], xj[c][cc], acc[a][cc]);
            } else {
                const double h = sblk[warp][t];
#pragma unroll
                for (int cc = 0; cc < C; ++cc) acc[0][cc] = fma(h, xj[0][cc], acc[0][cc]);
            }
        }
    }

    double alpha = 1.0, cshift = 0.0, beta = 0.0;
    if (coef) {
        const double* cf = coef + s * coef_stride;
        alpha = cf[0];
        cshift = cf[1];
        beta = cf[2];
    }
    double* Ys = Y + s * N * b;
    const double* Ws = W ? W + s * N * b : nullptr;
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int cc = 0; cc < C; ++cc) {
            const int64_t idx = ((int64_t)D * i + a) * b + lane + 32 * cc;
            double v = acc[a][cc];
            if (coef) {
                v = alpha * (v - cshift * xi[a][cc]);
                if (Ws && beta != 0.0) v -= beta * Ws[idx];
            }
            Ys[idx] = v;
        }
}


// Ensemble variant: one 1024-thread CTA per (structure, chunk of rows).  All 32
// warps of an SM work on the SAME structure, so the X rows they gather (the
// whole block vector of a 300-residue structure is 230 KB) are served by L1
// instead of L2: the row-per-warp kernel above interleaves 8 structures per SM
// and is bound by L2->SM bandwidth on small structures.
constexpr int kStructWarps = 32;
constexpr int kStructChunk = 16;  // contacts staged per warp and iteration

template <int D, int C>
__global__ void __launch_bounds__(1024, 1)
spmm_struct_kernel(int n, int rows_per_cta, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                   const double* __restrict__ offdiag, const double* __restrict__ diag,
                   const double* __restrict__ X, const double* __restrict__ W, double* __restrict__ Y,
                   const double* __restrict__ coef, int coef_stride, const int32_t* __restrict__ done) {
    constexpr int DD = D * D;
    constexpr int BS = (D == 3) ? 10 : 1;
    constexpr int b = 32 * C;
    __shared__ __align__(16) double sblk[kStructWarps][kStructChunk * BS];
    __shared__ int32_t scol[kStructWarps][kStructChunk];
    const int64_t s = blockIdx.y;
    if (done && done[s]) return;
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const int64_t N = (int64_t)D * n;
    const double* Xs = X + s * N * b;
    double* Ys = Y + s * N * b;
    const double* Ws = W ? W + s * N * b : nullptr;
    double alpha = 1.0, cshift = 0.0, beta = 0.0;
    if (coef) {
        const double* cf = coef + s * coef_stride;
        alpha = cf[0]; cshift = cf[1]; beta = cf[2];
    }
    const int i0 = blockIdx.x * rows_per_cta;
    const int i1 = min(n, i0 + rows_per_cta);
    for (int i = i0 + warp; i < i1; i += kStructWarps) {
        const int64_t row = s * n + i;
        double acc[D][C], xi[D][C];
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int cc = 0; cc < C; ++cc) {
                xi[a][cc] = Xs[((int64_t)D * i + a) * b + lane + 32 * cc];
                acc[a][cc] = 0.0;
            }
        {
            double dg[DD];
#pragma unroll
            for (int q = 0; q < DD; ++q) dg[q] = diag[row * DD + q];
#pragma unroll
            for (int a = 0; a < D; ++a)
#pragma unroll
                for (int c = 0; c < D; ++c)
#pragma unroll
                    for (int cc = 0; cc < C; ++cc) acc[a][cc] = fma(dg[a * D + c], xi[c][cc], acc[a][cc]);
        }
        const int64_t pb = rowptr[row], pe = rowptr[row + 1];
        for (int64_t p0 = pb; p0 < pe; p0 += kStructChunk) {
            const int cnt = (int)min((int64_t)kStructChunk, pe - p0);
            __syncwarp();
            if ((int)lane < cnt) scol[warp][lane] = col[p0 + lane];
            if (D == 3) {
                const double* src = offdiag + p0 * DD;
                for (int q = lane; q < cnt * DD; q += 32) sblk[warp][(q / DD) * BS + (q % DD)] = src[q];
            } else {
                if ((int)lane < cnt) sblk[warp][lane] = offdiag[p0 + lane];
            }
            __syncwarp();
#pragma unroll 4
            for (int t = 0; t < cnt; ++t) {
                const int j = scol[warp][t];
                double xj[D][C];
#pragma unroll
                for (int c = 0; c < D; ++c)
#pragma unroll
                    for (int cc = 0; cc < C; ++cc) xj[c][cc] = Xs[((int64_t)D * j + c) * b + lane + 32 * cc];
                if (D == 3) {
                    const double2* bp = reinterpret_cast<const double2*>(&sblk[warp][t * BS]);
                    const double2 b01 = bp[0], b23 = bp[1], b45 = bp[2], b67 = bp[3];
                    const double b8 = sblk[warp][t * BS + 8];
                    const double h[9] = {b01.x, b01.y, b23.x, b23.y, b45.x, b45.y, b67.x, b67.y, b8};
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int cc = 0; cc < C; ++cc) acc[a][cc] = fma(h[a * 3 + c], xj[c][cc], acc[a][cc]);
                } else {
                    const double h = sblk[warp][t];
#pragma unroll
                    for (int cc = 0; cc < C; ++cc) acc[0][cc] = fma(h, xj[0][cc], acc[0][cc]);
                }
            }
        }
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int cc = 0; cc < C; ++cc) {
                const int64_t idx = ((int64_t)D * i + a) * b + lane + 32 * cc;
                double v = acc[a][cc];
                if (coef) {
                    v = alpha * (v - cshift * xi[a][cc]);
                    if (Ws && beta != 0.0) v -= beta * Ws[idx];
                }
                Ys[idx] = v;
            }
    }
}

template <int D, int C>
static int spmm_struct_launch(int B, int n, const int64_t* rowptr, const int32_t* col, const double* offdiag,
                              const double* diag, const double* X, const double* W, double* Y, const double* coef,
                              int coef_stride, const int32_t* done, cudaStream_t st) {
    // per-device function attribute: set on every launch (cheap), no process-wide "configured" flag --   give L1 as much of the unified array as possible
    cudaFuncSetAttribute(spmm_struct_kernel<D, C>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    // whole structures per CTA up to 512 rows; larger ones in chunks of 256 rows
    const int rows_per_cta = n <= 512 ? n : 256;
    dim3 grid((unsigned)ceil_div(n, rows_per_cta), (unsigned)B);
    spmm_struct_kernel<D, C><<<grid, 1024, 0, st>>>(n, rows_per_cta, rowptr, col, offdiag, diag, X, W, Y, coef,
                                                   coef_stride, done);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

template <int D>
static int spmm_dispatch(int C, int n, int64_t nrows, const int64_t* rowptr, const int32_t* col,
                         const double* offdiag, const double* diag, const double* X, const double* W, double* Y,
                         const double* coef, int coef_stride, const int32_t* done, cudaStream_t st) {
    const unsigned grid = (unsigned)ceil_div(nrows, 8);
#define SCB_SPMM_CASE(CV)                                                                          \
    case CV:                                                                                       \
        spmm_kernel<D, CV><<<grid, 256, 0, st>>>(n, nrows, rowptr, col, offdiag, diag, X, W, Y, coef, \
                                                 coef_stride, done);                               \
        break;
    switch (C) {
        SCB_SPMM_CASE(1)
        SCB_SPMM_CASE(2)
        SCB_SPMM_CASE(4)
        default:
            return SCB_ERR_UNSUPPORTED;
    }
#undef SCB_SPMM_CASE
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

// internal entry used by the eigensolver
int spmm_cheb(int D, int B, int n, const int64_t* rowptr, const int32_t* col, const double* offdiag,
              const double* diag, const double* X, const double* W, double* Y, int b, const double* coef,
              int coef_stride, const int32_t* done, cudaStream_t st) {
    if (b % 32 != 0) return SCB_ERR_UNSUPPORTED;
    const int64_t nrows = (int64_t)B * n;
    // ensembles of small structures: keep each SM on one structure (L1-resident X)
    if (B >= 64 && n >= 64 && n <= 4096 && (b == 32 || b == 64)) {
        if (D == 3 && b == 32) return spmm_struct_launch<3, 1>(B, n, rowptr, col, offdiag, diag, X, W, Y, coef, coef_stride, done, st);
        if (D == 3 && b == 64) return spmm_struct_launch<3, 2>(B, n, rowptr, col, offdiag, diag, X, W, Y, coef, coef_stride, done, st);
        if (D == 1 && b == 32) return spmm_struct_launch<1, 1>(B, n, rowptr, col, offdiag, diag, X, W, Y, coef, coef_stride, done, st);
        if (D == 1 && b == 64) return spmm_struct_launch<1, 2>(B, n, rowptr, col, offdiag, diag, X, W, Y, coef, coef_stride, done, st);
    }
    if (D == 1) return spmm_dispatch<1>(b / 32, n, nrows, rowptr, col, offdiag, diag, X, W, Y, coef, coef_stride, done, st);
    if (D == 3) return spmm_dispatch<3>(b / 32, n, nrows, rowptr, col, offdiag, diag, X, W, Y, coef, coef_stride, done, st);
    return SCB_ERR_INVALID;
}

}  // namespace scb

extern "C" int scb_spmm(int D, int B, int n, const int64_t* rowptr, const int32_t* col, const double* offdiag,
                        const double* diag, const double* X, double* Y, int b, void* stream) {
    if (!rowptr || !col || !offdiag || !diag || !X || !Y || B < 1 || n < 1) return SCB_ERR_INVALID;
    return scb::spmm_cheb(D, B, n, rowptr, col, offdiag, diag, X, nullptr, Y, b, nullptr, 0, nullptr,
                          scb::as_stream(stream));
}
