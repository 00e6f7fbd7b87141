// Whole hot path for an ensemble: contacts -> assembly -> lowest-k modes -> MSF.
// The host-buffer variant is the drop-in for the reference-side loop
//   for c in conformations: ANM(c, ff).eigen(); mean_square_fluctuation(...)
// (anm.py:62-148 + nma.py:29-184), with H2D/D2H inside.
#include <stdlib.h>

#include <mutex>

#include "subspace.cuh"

namespace scb {

__global__ void aos_to_soa_kernel(const double* __restrict__ aos, int64_t total_atoms, int n, double* __restrict__ soa) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total_atoms) return;
    const int64_t s = q / n;
    const int i = (int)(q % n);
#pragma unroll
    for (int a = 0; a < 3; ++a) soa[(s * 3 + a) * n + i] = aos[q * 3 + a];
}

// small systems: modes_out[s][q][r] = modes_full[s][k0+q][r], eigval_out[s][q] = lam_full[s][k0+q]
__global__ void slice_full_kernel(int B, int N, int k0, int k, const double* __restrict__ lam_full,
                                  const double* __restrict__ modes_full, double* __restrict__ eigval,
                                  double* __restrict__ modes) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)B * k * N;
    if (q >= total) return;
    const int64_t s = q / ((int64_t)k * N);
    const int64_t rem = q % ((int64_t)k * N);
    const int m = (int)(rem / N), r = (int)(rem % N);
    modes[q] = modes_full[(s * N + k0 + m) * N + r];
    if (r == 0) eigval[s * k + m] = lam_full[s * N + k0 + m];
}

// eigval_out[s][q] = theta[s][k0+q]
__global__ void slice_eigval_kernel(int B, int b, int k0, int k, const double* __restrict__ theta,
                                    double* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B * k) return;
    out[q] = theta[(int64_t)(q / k) * b + k0 + q % k];
}

// Scratch of the whole-path entry points comes from a PRIVATE memory pool per device (created on first use,
// guarded by a mutex): its release threshold keeps the blocks cached between calls without touching the default
// pool that the caller's allocator (PyTorch) uses.  scb_trim_pool() hands the cached blocks back to the driver.
static std::mutex g_pool_mutex;
static cudaMemPool_t g_pools[64] = {};

static cudaMemPool_t device_pool() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        g_pools[dev] = pool;
    }
    return g_pools[dev];
}

int pool_alloc(void** p, size_t bytes, cudaStream_t st) {
    cudaMemPool_t pool = device_pool();
    cudaError_t e = pool ? cudaMallocFromPoolAsync(p, bytes, pool, st) : cudaMallocAsync(p, bytes, st);
    if (e != cudaSuccess) { set_last_cuda_error(e, __FILE__, __LINE__); return SCB_ERR_CUDA; }
    return SCB_OK;
}
void pool_free(void* p, cudaStream_t st) { if (p) cudaFreeAsync(p, st); }

struct Scratch {
    cudaStream_t st;
    cudaMemPool_t pool;
    void* ptrs[24];
    int count = 0;
    explicit Scratch(cudaStream_t s) : st(s), pool(device_pool()) {}
    template <typename T>
    int alloc(T** p, size_t elems) {
        void* v = nullptr;
        cudaError_t e = pool ? cudaMallocFromPoolAsync(&v, elems * sizeof(T) + 256, pool, st)
                             : cudaMallocAsync(&v, elems * sizeof(T) + 256, st);
        if (e != cudaSuccess) { set_last_cuda_error(e, __FILE__, __LINE__); return SCB_ERR_CUDA; }
        ptrs[count++] = v;
        *p = static_cast<T*>(v);
        // SCB_POISON=1 (tests): fill fresh scratch with 0xFF bytes (NaN doubles, -1 ints) so that a read of
        // memory the path forgot to initialise shows up as a wrong result instead of depending on pool history
        static const bool poison = getenv("SCB_POISON") && atoi(getenv("SCB_POISON")) != 0;
        if (poison) cudaMemsetAsync(v, 0xFF, elems * sizeof(T) + 256, st);
        return SCB_OK;
    }
    ~Scratch() { for (int i = 0; i < count; ++i) cudaFreeAsync(ptrs[i], st); }
};

}  // namespace scb

using namespace scb;

extern "C" int scb_trim_pool(void) {
    cudaMemPool_t pool = device_pool();
    if (!pool) return SCB_OK;
    SCB_CUDA(cudaDeviceSynchronize());
    SCB_CUDA(cudaMemPoolTrimTo(pool, 0));
    return SCB_OK;
}

extern "C" int scb_coords_to_soa(const double* coord_aos, int B, int n, double* xyz_soa, void* stream) {
    if (!coord_aos || !xyz_soa || B < 1 || n < 1) return SCB_ERR_INVALID;
    const int64_t total = (int64_t)B * n;
    aos_to_soa_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(coord_aos, total, n, xyz_soa);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_enm_ensemble(int D, const double* xyz, int B, int n, const scb_ff_desc* ff,
                                const scb_patch* patch, const double* masses, int k, double tol, double* eigval,
                                double* msf, double* modes, int32_t* iters, int64_t* n_pairs_out, void* stream) {
    if (!xyz || !ff || !eigval || !msf || B < 1 || n < 1 || k < 1 || (D != 1 && D != 3)) return SCB_ERR_INVALID;
    if (ff->cutoff_sq < 0.0) return SCB_ERR_UNSUPPORTED;  // all-pairs force fields: dense slab path
    if (B > 65535) return SCB_ERR_UNSUPPORTED;   // structures are indexed by blockIdx.y: callers chunk larger batches
    cudaStream_t st = as_stream(stream);
    const int nz = (D == 3) ? 6 : 1;
    const int b = (k + 8 <= 32) ? 32 : 64;
    if (k + 4 > b) return SCB_ERR_UNSUPPORTED;
    // Small systems take the dense full-spectrum path: when the solver block spans a large part of the space
    // (N <= 8 b) the Chebyshev filter separates the block's own modes by more than 1/eps and subspace iteration
    // loses rank, while a 256 x 256 Jacobi solve costs next to nothing.
    const bool small = (int64_t)D * n <= 8 * (int64_t)b;
    if (small && nz + k > D * n) return SCB_ERR_INVALID;
    const int64_t nrows = (int64_t)B * n;
    const int64_t N = (int64_t)D * n;
    Scratch sc(st);
    int32_t* rowcount; int64_t* rowptr; void* scan_tmp; int32_t* flag;
    SCB_TRY(sc.alloc(&rowcount, (size_t)nrows));
    SCB_TRY(sc.alloc(&rowptr, (size_t)nrows + 1));
    SCB_TRY(sc.alloc((char**)&scan_tmp, scb_scan_scratch_bytes(nrows)));
    SCB_TRY(sc.alloc(&flag, 1));
    SCB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), st));
    SCB_TRY(scb_contacts_count(xyz, B, n, ff->cutoff_sq, patch, 0, rowcount, st));
    SCB_TRY(scb_contacts_scan(rowcount, nrows, rowptr, scan_tmp, st));
    int64_t P = 0;
    SCB_CUDA(cudaMemcpyAsync(&P, rowptr + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaStreamSynchronize(st));
    if (n_pairs_out) *n_pairs_out = P;
    int32_t* col; double *offdiag, *diag, *gersh, *Z, *theta, *X, *resid; int32_t* it; void* ws;
    SCB_TRY(sc.alloc(&col, (size_t)(P > 0 ? P : 1)));
    SCB_TRY(sc.alloc(&offdiag, (size_t)(P > 0 ? P : 1) * D * D));
    SCB_TRY(sc.alloc(&diag, (size_t)nrows * D * D));
    SCB_TRY(sc.alloc(&gersh, (size_t)B));
    SCB_TRY(sc.alloc(&Z, (size_t)B * N * nz));
    SCB_TRY(sc.alloc(&theta, (size_t)B * b));
    SCB_TRY(sc.alloc(&resid, (size_t)B * b));
    SCB_TRY(sc.alloc(&X, (size_t)B * N * b));
    SCB_TRY(sc.alloc(&it, (size_t)B));
    const size_t ws_bytes = scb_eig_lowest_workspace_bytes(D, B, n, b, nz, P);
    SCB_TRY(sc.alloc((char**)&ws, ws_bytes));
    SCB_TRY(scb_contacts_fill(xyz, B, n, ff->cutoff_sq, patch, 0, rowptr, col, st));
    SCB_TRY(scb_assemble(D, xyz, B, n, ff, rowptr, col, masses, offdiag, diag, gersh, flag, st));
    if (small) {
        // tiny structures (e.g. 20-residue Trp-cage): densify and run the shared-memory / block Jacobi solver
        const int64_t NN = N * N;
        double *dense, *lam_full, *modes_full, *modes_k; void* fws;
        SCB_TRY(sc.alloc(&dense, (size_t)B * NN));
        SCB_TRY(sc.alloc(&lam_full, (size_t)B * N));
        SCB_TRY(sc.alloc(&modes_full, (size_t)B * NN));
        SCB_TRY(sc.alloc(&modes_k, (size_t)B * k * N));
        const size_t fbytes = scb_eig_full_workspace_bytes(B, (int)N);
        SCB_TRY(sc.alloc((char**)&fws, fbytes));
        SCB_TRY(scb_densify(D, B, n, rowptr, col, offdiag, diag, dense, st));
        SCB_TRY(scb_eig_full(B, (int)N, dense, lam_full, modes_full, fws, fbytes, st));
        const int64_t total = (int64_t)B * k * N;
        slice_full_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(B, (int)N, nz, k, lam_full, modes_full,
                                                                         eigval, modes_k);
        SCB_LAUNCH_CHECK();
        SCB_TRY(scb_msf(D, B, n, k, eigval, modes_k, 1.0, msf, st));
        if (modes) SCB_CUDA(cudaMemcpyAsync(modes, modes_k, sizeof(double) * (size_t)total, cudaMemcpyDeviceToDevice, st));
        if (iters) SCB_CUDA(cudaMemsetAsync(iters, 0, sizeof(int32_t) * B, st));
        int32_t hflag = 0;
        SCB_CUDA(cudaMemcpyAsync(&hflag, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaStreamSynchronize(st));
        return hflag != 0 ? hflag : SCB_OK;
    }
    SCB_TRY(scb_rigid_basis(D, xyz, B, n, masses, Z, st));
    // filter degree: measured on the C3 batch with the structure-resident FP32 filter (step time): 24: 256 ms,
    // 28: 250, 32: 247, 36: 247, 40: 240, 48: 275
    int degree = 40;
    if (const char* env = getenv("SCB_DEGREE")) degree = atoi(env) >= 2 ? atoi(env) : degree;
    int status = scb_eig_lowest(D, B, n, P, rowptr, col, offdiag, diag, gersh, Z, nz, k, b, tol, 200, degree,
                                0x5cb200ull, theta, X, resid, it, ws, ws_bytes, st);
    if (status != SCB_OK && status != SCB_ERR_NOT_CONVERGED) return status;
    slice_eigval_kernel<<<(unsigned)ceil_div((int64_t)B * k, 256), 256, 0, st>>>(B, b, 0, k, theta, eigval);
    SCB_LAUNCH_CHECK();
    SCB_TRY(scb_msf_cols(D, B, n, b, 0, k, theta, X, 1.0, msf, st));
    if (modes) SCB_TRY(scb_export_modes(B, (int)N, b, 0, k, X, modes, st));
    if (iters) SCB_CUDA(cudaMemcpyAsync(iters, it, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
    int32_t hflag = 0;
    SCB_CUDA(cudaMemcpyAsync(&hflag, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaStreamSynchronize(st));
    if (hflag != 0) return hflag;
    return status;
}

extern "C" int scb_enm_ensemble_host(int D, const double* coord_host, int B, int n, const scb_ff_desc* ff,
                                     const scb_patch* patch, const double* masses_dev, int k, double tol,
                                     double* eigval_host, double* msf_host, double* modes_host,
                                     int32_t* iters_host, int64_t* n_pairs_out, void* stream) {
    if (!coord_host || !eigval_host || !msf_host) return SCB_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    const int64_t N = (int64_t)D * n;
    Scratch sc(st);
    double *aos, *soa, *eigval, *msf, *modes = nullptr;
    int32_t* iters = nullptr;
    if (iters_host) SCB_TRY(sc.alloc(&iters, (size_t)B));
    SCB_TRY(sc.alloc(&aos, (size_t)B * n * 3));
    SCB_TRY(sc.alloc(&soa, (size_t)B * n * 3));
    SCB_TRY(sc.alloc(&eigval, (size_t)B * k));
    SCB_TRY(sc.alloc(&msf, (size_t)B * n));
    if (modes_host) SCB_TRY(sc.alloc(&modes, (size_t)B * k * N));
    SCB_CUDA(cudaMemcpyAsync(aos, coord_host, sizeof(double) * (size_t)B * n * 3, cudaMemcpyHostToDevice, st));
    SCB_TRY(scb_coords_to_soa(aos, B, n, soa, st));
    int status = scb_enm_ensemble(D, soa, B, n, ff, patch, masses_dev, k, tol, eigval, msf, modes, iters,
                                  n_pairs_out, st);
    if (status != SCB_OK && status != SCB_ERR_NOT_CONVERGED) return status;
    SCB_CUDA(cudaMemcpyAsync(eigval_host, eigval, sizeof(double) * (size_t)B * k, cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaMemcpyAsync(msf_host, msf, sizeof(double) * (size_t)B * n, cudaMemcpyDeviceToHost, st));
    if (modes_host)
        SCB_CUDA(cudaMemcpyAsync(modes_host, modes, sizeof(double) * (size_t)B * k * N, cudaMemcpyDeviceToHost, st));
    if (iters_host)
        SCB_CUDA(cudaMemcpyAsync(iters_host, iters, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaStreamSynchronize(st));
    return status;
}
