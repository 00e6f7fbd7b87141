// K2: fused force constant + Kirchhoff (D=1) / Hessian 3x3-block (D=3) assembly.
// Replaces ForceField.force_constant (forcefield.py) + compute_kirchhoff /
// compute_hessian (interaction.py:14-111) + mass weighting (anm.py:89-113).
//
// One warp per CSR row.  Lanes evaluate the force constant and the off-diagonal
// block of one contact each (the row's contacts in chunks of 32) and write the
// BSR values; the diagonal block is then accumulated by D*D lanes walking the
// row's blocks in ascending column order -- the same order np.sum(axis=0) uses
// in the reference (interaction.py:52,104) -- so the result is bit-identical.
// Arithmetic that the parity contract pins uses explicitly rounded intrinsics
// (no FMA contraction).
#include "common.cuh"

namespace scb {

struct FFView {
    scb_ff_desc d;
};

__device__ __forceinline__ int bin_of(const scb_ff_desc& ff, double sq) {
    // np.searchsorted(edges**2, sq) with side="left": number of edges^2 < sq
    // (forcefield.py:523)
    int lo = 0, hi = ff.nbins;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (ff.edges_sq[mid] < sq) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// force constant of the ordered pair (i, j); e = global pair index
__device__ __forceinline__ double base_force_constant(const scb_ff_desc& ff, int n, int i, int j,
                                                      double sq, int64_t e, int32_t* status_flag) {
    switch (ff.kind) {
        case SCB_FF_INVARIANT:  // forcefield.py:283-284
            return 1.0;
        case SCB_FF_HINSEN: {   // forcefield.py:318-326
            const double r = fmax(sqrt(sq), 2.9);
            if (r < 4.0) return __dadd_rn(__dmul_rn(r, 860.0), -2390.0);
            return __dmul_rn(pow(r, -6.0), 1280000.0);
        }
        case SCB_FF_PFREE:      // forcefield.py:361-362
            return 1.0 / sq;
        case SCB_FF_TABULATED:
        case SCB_FF_TABULATED_DENSE: {  // forcefield.py:497-533
            if (i == j) return 0.0;     // forcefield.py:512-513
            int b = 0;
            if (ff.nbins > 1) {
                b = bin_of(ff, sq);
                if (b >= ff.nbins) {
                    if (status_flag) atomicExch(status_flag, (int)SCB_ERR_ABOVE_CUTOFF);
                    return 0.0;
                }
            }
            if (ff.kind == SCB_FF_TABULATED_DENSE)
                return (double)ff.dense_table[((size_t)i * n + j) * ff.nbins + b];
            const int lo = min(i, j), hi = max(i, j);
            const int tlo = ff.res_type[lo], thi = ff.res_type[hi];
            if (hi == lo + 1 && ff.bonded_next[lo])  // forcefield.py:504-509
                return (double)ff.bonded[(tlo * 20 + thi) * ff.nbins + b];
            const int ti = ff.res_type[i], tj = ff.res_type[j];
            const float* tab = (ff.chain[i] == ff.chain[j]) ? ff.intra : ff.inter;
            return (double)tab[(ti * 20 + tj) * ff.nbins + b];
        }
        case SCB_FF_EXTERNAL:
            return ff.external_fc[e];
    }
    return 0.0;
}

__device__ __forceinline__ double force_constant(const scb_ff_desc& ff, int n, int i, int j, double sq,
                                                 int64_t e, int32_t* status_flag) {
    if (!ff.patched) return base_force_constant(ff, n, i, j, sq, e, status_flag);
    // PatchedForceField.force_constant (forcefield.py:183-226)
    double fc = 0.0;
    if (ff.cutoff_sq < 0.0 || sq <= ff.cutoff_sq) fc = base_force_constant(ff, n, i, j, sq, e, status_flag);
    for (int q = 0; q < ff.n_pair_on; ++q) {
        const int a = ff.pair_on[2 * q], b = ff.pair_on[2 * q + 1];
        if ((a == i && b == j) || (a == j && b == i)) {
            const double v = ff.pair_on_fc[q];
            if (v != -1.0) fc = v;  // sentinel, forcefield.py:214-223
        }
    }
    return fc;
}

template <int D>
__global__ void __launch_bounds__(256)
assemble_rows_kernel(const double* __restrict__ xyz, int n, int64_t nrows, FFView ffv,
                     const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                     double* __restrict__ offdiag, double* __restrict__ diag, int32_t* status_flag) {
    constexpr int DD = D * D;
    const scb_ff_desc& ff = ffv.d;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const unsigned lane = lane_id();
    const int s = (int)(row / n), i = (int)(row % n);
    const double* X = xyz + (size_t)s * 3 * n;
    const double xi = X[i], yi = X[n + i], zi = X[2 * n + i];
    const int64_t b = rowptr[row], e = rowptr[row + 1];
    for (int64_t p = b + lane; p < e; p += 32) {
        const int j = col[p];
        // disp = x_j - x_i ; sq = (d0^2 + d1^2) + d2^2  (interaction.py:183-184)
        const double d[3] = {X[j] - xi, X[n + j] - yi, X[2 * n + j] - zi};
        const double sq = __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])),
                                    __dmul_rn(d[2], d[2]));
        const double fc = force_constant(ff, n, i, j, sq, p, status_flag);
        double* out = offdiag + p * DD;
        if (D == 1) {
            out[0] = -fc;  // interaction.py:50
        } else {
            const double t = (-fc) / sq;  // interaction.py:96-101: ((-fc/sq) * d_a) * d_b
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double ta = __dmul_rn(t, d[a]);
#pragma unroll
                for (int c = 0; c < 3; ++c) out[a * 3 + c] = __dmul_rn(ta, d[c]);
            }
        }
    }
    __syncwarp();
    // diagonal block: -(sum over the row's blocks, ascending column)
    if (lane < DD) {
        double acc = 0.0;
        for (int64_t p = b; p < e; ++p) acc = __dadd_rn(acc, offdiag[p * DD + lane]);
        diag[row * DD + lane] = -acc;
    }
}

// H *= outer(w, w), w = 1/sqrt(m) repeated D times; the product w_i*w_j is
// rounded first (anm.py:89-96, 112-113)
template <int D>
__global__ void __launch_bounds__(256)
mass_scale_kernel(int n, int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                  const double* __restrict__ masses, double* __restrict__ offdiag, double* __restrict__ diag) {
    constexpr int DD = D * D;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const unsigned lane = lane_id();
    const int i = (int)(row % n);
    const double wi = 1.0 / sqrt(masses[i]);
    const int64_t b = rowptr[row], e = rowptr[row + 1];
    for (int64_t p = b + lane; p < e; p += 32) {
        const double w = __dmul_rn(wi, 1.0 / sqrt(masses[col[p]]));
#pragma unroll
        for (int q = 0; q < DD; ++q) offdiag[p * DD + q] = __dmul_rn(offdiag[p * DD + q], w);
    }
    if (lane < DD) diag[row * DD + lane] = __dmul_rn(diag[row * DD + lane], __dmul_rn(wi, wi));
}

// upper bound of the spectrum: max row sum of |entries| (Gershgorin)
template <int D>
__global__ void __launch_bounds__(256)
gershgorin_kernel(int n, int64_t nrows, const int64_t* __restrict__ rowptr, const double* __restrict__ offdiag,
                  const double* __restrict__ diag, double* __restrict__ gersh) {
    constexpr int DD = D * D;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const unsigned lane = lane_id();
    const int64_t b = rowptr[row], e = rowptr[row + 1];
    double acc[D];
#pragma unroll
    for (int a = 0; a < D; ++a) acc[a] = 0.0;
    for (int64_t p = b + lane; p < e; p += 32)
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int c = 0; c < D; ++c) acc[a] += fabs(offdiag[p * DD + a * D + c]);
    double best = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        double v = warp_sum(acc[a]);
#pragma unroll
        for (int c = 0; c < D; ++c) v += fabs(diag[row * DD + a * D + c]);
        best = fmax(best, v);
    }
    if (lane == 0) atomic_max_nonneg(&gersh[row / n], best);
}

template <int D>
__global__ void __launch_bounds__(256)
densify_kernel(int n, int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
               const double* __restrict__ offdiag, const double* __restrict__ diag, double* __restrict__ dense) {
    constexpr int DD = D * D;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const unsigned lane = lane_id();
    const int64_t s = row / n;
    const int i = (int)(row % n);
    const int64_t N = (int64_t)D * n;
    double* M = dense + s * N * N;
    const int64_t b = rowptr[row], e = rowptr[row + 1];
    for (int64_t p = b + lane; p < e; p += 32) {
        const int j = col[p];
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int c = 0; c < D; ++c) M[((int64_t)D * i + a) * N + (int64_t)D * j + c] = offdiag[p * DD + a * D + c];
    }
    if (lane < DD) {
        const int a = lane / D, c = lane % D;
        M[((int64_t)D * i + a) * N + (int64_t)D * i + c] = diag[row * DD + lane];
    }
}

// all-pairs force fields: dense row slab, one warp per node row i in [row0,row1).
// Off-diagonal blocks are written straight into the dense slab; the diagonal
// block is the negated in-order sum (thread-sequential over j for bit-exactness
// would serialise 20k terms, so lanes own strided j and the partial sums are
// combined in a fixed lane order: deterministic, within 1e-12 of the reference).
template <int D>
__global__ void __launch_bounds__(256)
dense_allpairs_kernel(const double* __restrict__ xyz, int n, FFView ffv, const double* __restrict__ masses,
                      int row0, int row1, double* __restrict__ dense) {
    constexpr int DD = D * D;
    const scb_ff_desc& ff = ffv.d;
    const int i = row0 + blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= row1) return;
    const unsigned lane = lane_id();
    const int64_t N = (int64_t)D * n;
    double* M = dense + (int64_t)(i - row0) * D * N;
    const double xi = xyz[i], yi = xyz[n + i], zi = xyz[2 * (size_t)n + i];
    const double wi = masses ? 1.0 / sqrt(masses[i]) : 1.0;
    double acc[DD];
#pragma unroll
    for (int q = 0; q < DD; ++q) acc[q] = 0.0;
    for (int j = lane; j < n; j += 32) {
        if (j == i) continue;
        const double d[3] = {xyz[j] - xi, xyz[n + j] - yi, xyz[2 * (size_t)n + j] - zi};
        const double sq = __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])),
                                    __dmul_rn(d[2], d[2]));
        const double fc = force_constant(ff, n, i, j, sq, 0, nullptr);
        const double w = masses ? __dmul_rn(wi, 1.0 / sqrt(masses[j])) : 1.0;
        if (D == 1) {
            const double v = -fc;
            acc[0] += v;
            M[j] = masses ? __dmul_rn(v, w) : v;
        } else {
            const double t = (-fc) / sq;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double ta = __dmul_rn(t, d[a]);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double v = __dmul_rn(ta, d[c]);
                    acc[a * 3 + c] += v;
                    M[(int64_t)a * N + 3 * (int64_t)j + c] = masses ? __dmul_rn(v, w) : v;
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < DD; ++q) {
        double v = acc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        acc[q] = v;
    }
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const double v = -acc[a * D + c];
                M[(int64_t)a * N + (int64_t)D * i + c] = masses ? __dmul_rn(v, __dmul_rn(wi, wi)) : v;
            }
    }
}

// ForceField.force_constant(atom_i, atom_j, sq_distance) for explicit triples
__global__ void __launch_bounds__(256)
force_constant_kernel(FFView ffv, int n, const int32_t* __restrict__ ai, const int32_t* __restrict__ aj,
                      const double* __restrict__ sq, int64_t P, double* __restrict__ out, int32_t* status_flag) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    out[p] = force_constant(ffv.d, n, ai[p], aj[p], sq[p], p, status_flag);
}

// disp[P][3] = x_j - x_i (optional), sq[P] = (d0^2 + d1^2) + d2^2   (interaction.py:182-184)
__global__ void __launch_bounds__(256)
pair_geometry_kernel(const double* __restrict__ xyz, int n, int64_t nrows, const int64_t* __restrict__ rowptr,
                     const int32_t* __restrict__ col, double* __restrict__ disp, double* __restrict__ sq) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int s = (int)(row / n), i = (int)(row % n);
    const double* X = xyz + (size_t)s * 3 * n;
    const double xi = X[i], yi = X[n + i], zi = X[2 * n + i];
    for (int64_t p = rowptr[row] + lane_id(); p < rowptr[row + 1]; p += 32) {
        const int j = col[p];
        const double d0 = X[j] - xi, d1 = X[n + j] - yi, d2 = X[2 * n + j] - zi;
        if (disp) { disp[3 * p] = d0; disp[3 * p + 1] = d1; disp[3 * p + 2] = d2; }
        sq[p] = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
    }
}

template <int D>
static int assemble_impl(const double* xyz, int B, int n, const scb_ff_desc* ff, const int64_t* rowptr,
                         const int32_t* col, const double* masses, double* offdiag, double* diag,
                         double* gersh, int32_t* status_flag, cudaStream_t st) {
    const int64_t nrows = (int64_t)B * n;
    const unsigned grid = (unsigned)ceil_div(nrows, 8);
    FFView ffv{*ff};
    assemble_rows_kernel<D><<<grid, 256, 0, st>>>(xyz, n, nrows, ffv, rowptr, col, offdiag, diag, status_flag);
    SCB_LAUNCH_CHECK();
    if (masses) {
        mass_scale_kernel<D><<<grid, 256, 0, st>>>(n, nrows, rowptr, col, masses, offdiag, diag);
        SCB_LAUNCH_CHECK();
    }
    if (gersh) {
        SCB_CUDA(cudaMemsetAsync(gersh, 0, sizeof(double) * B, st));
        gershgorin_kernel<D><<<grid, 256, 0, st>>>(n, nrows, rowptr, offdiag, diag, gersh);
        SCB_LAUNCH_CHECK();
    }
    return SCB_OK;
}

}  // namespace scb

using namespace scb;

extern "C" int scb_assemble(int D, const double* xyz, int B, int n, const scb_ff_desc* ff, const int64_t* rowptr,
                            const int32_t* col, const double* masses, double* offdiag, double* diag,
                            double* gersh, int32_t* status_flag, void* stream) {
    if (!xyz || !ff || !rowptr || !col || !offdiag || !diag || B < 1 || n < 1) return SCB_ERR_INVALID;
    if (D == 1) return assemble_impl<1>(xyz, B, n, ff, rowptr, col, masses, offdiag, diag, gersh, status_flag, as_stream(stream));
    if (D == 3) return assemble_impl<3>(xyz, B, n, ff, rowptr, col, masses, offdiag, diag, gersh, status_flag, as_stream(stream));
    return SCB_ERR_INVALID;
}

extern "C" int scb_densify(int D, int B, int n, const int64_t* rowptr, const int32_t* col, const double* offdiag,
                           const double* diag, double* dense, void* stream) {
    if (!rowptr || !col || !offdiag || !diag || !dense || B < 1 || n < 1) return SCB_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    const int64_t nrows = (int64_t)B * n;
    const int64_t N = (int64_t)D * n;
    SCB_CUDA(cudaMemsetAsync(dense, 0, sizeof(double) * (size_t)B * N * N, st));
    const unsigned grid = (unsigned)ceil_div(nrows, 8);
    if (D == 1) densify_kernel<1><<<grid, 256, 0, st>>>(n, nrows, rowptr, col, offdiag, diag, dense);
    else if (D == 3) densify_kernel<3><<<grid, 256, 0, st>>>(n, nrows, rowptr, col, offdiag, diag, dense);
    else return SCB_ERR_INVALID;
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_assemble_dense_allpairs(int D, const double* xyz, int n, const scb_ff_desc* ff,
                                           const double* masses, int row0, int row1, double* dense, void* stream) {
    if (!xyz || !ff || !dense || n < 1 || row0 < 0 || row1 > n || row0 >= row1) return SCB_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    FFView ffv{*ff};
    const unsigned grid = (unsigned)ceil_div(row1 - row0, 8);
    if (D == 1) dense_allpairs_kernel<1><<<grid, 256, 0, st>>>(xyz, n, ffv, masses, row0, row1, dense);
    else if (D == 3) dense_allpairs_kernel<3><<<grid, 256, 0, st>>>(xyz, n, ffv, masses, row0, row1, dense);
    else return SCB_ERR_INVALID;
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_force_constant(const scb_ff_desc* ff, int n, const int32_t* atom_i, const int32_t* atom_j,
                                  const double* sq, int64_t P, double* out, int32_t* status_flag, void* stream) {
    if (!ff || !atom_i || !atom_j || !sq || !out || P < 0) return SCB_ERR_INVALID;
    if (ff->kind == SCB_FF_EXTERNAL) return SCB_ERR_INVALID;
    if (P == 0) return SCB_OK;
    FFView ffv{*ff};
    force_constant_kernel<<<(unsigned)ceil_div(P, 256), 256, 0, as_stream(stream)>>>(ffv, n, atom_i, atom_j, sq, P,
                                                                                    out, status_flag);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_pair_geometry(const double* xyz, int B, int n, const int64_t* rowptr, const int32_t* col,
                                 double* disp, double* sq, void* stream) {
    if (!xyz || !rowptr || !col || !sq || B < 1 || n < 1) return SCB_ERR_INVALID;
    const int64_t nrows = (int64_t)B * n;
    pair_geometry_kernel<<<(unsigned)ceil_div(nrows, 8), 256, 0, as_stream(stream)>>>(xyz, n, nrows, rowptr, col,
                                                                                     disp, sq);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}
