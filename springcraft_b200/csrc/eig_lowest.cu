// K3: lowest-k eigensolver -- Chebyshev-filtered subspace iteration on the BSR
// SpMM with a dense Rayleigh-Ritz step, batched over the structures of an
// ensemble.  Replaces np.linalg.eigh (nma.py:61) for the "lowest modes" use.
//
// Per outer iteration and structure (all launches cover the whole batch;
// converged structures are skipped through the device-side `done` flags):
//   1. filter   X <- p_d(H) X          d fused SpMM+recurrence launches
//   2. deflate  X <- X - Z Z^T X       analytic null space (rigid-body modes)
//   3. orth     S = X^T X = L L^T,  X <- X L^-T
//   4. HX = H X
//   5. RR       T = X^T HX, S2 = X^T X -> Jacobi -> theta, C ; X <- X C, HX <- HX C
//   6. residuals ||HX - theta X||, convergence flags, new filter bounds
// The host polls one int32 (number of active structures) per outer iteration.
#include <stdlib.h>

#include "resident.cuh"

namespace scb {


constexpr int kLanczosSteps = 8;      // column-wise Lanczos steps of the spectrum-bound estimator (6: +4 % iterations, 10/12: same iterations, 1-3 % slower)
constexpr int kLanczosStepsMax = 16;  // workspace is sized for this many (SCB_LANCZOS)

struct EigWork {
    double *A, *Bf, *Cf, *HX;    // block vectors [B][N][b]; A is the canonical basis (caller's X)
    double *S, *T, *Cm, *theta, *rn2, *P, *coef;
    double *lz_alpha, *lz_beta2;  // Lanczos coefficients [steps][B][b]
    EigState* state;
    int32_t *done, *n_active;
    int32_t* pcount;          // row-paired operator (spmm_paired.cu)
    char* pent;
    char* pent32;             // single-precision copy of the records (spmm_paired_f32.cu)
    float *Xa32, *Xb32, *Xc32;
    int32_t *skip32, *skip64; // per-structure precision of the next filter
    ResLayout res;            // structure-resident operator (resident.cuh); res.cols == 0: not eligible
};

static size_t carve(Arena& ar, EigWork* w, int D, int B, int n, int b, int nz, int degree_cap, double* X,
                    int64_t P) {
    const size_t N = (size_t)D * n;
    const size_t vec = (size_t)B * N * b;
    w->A = X;
    w->Bf = ar.take<double>(vec);
    w->Cf = ar.take<double>(vec);
    w->HX = ar.take<double>(vec);
    w->S = ar.take<double>((size_t)B * b * b);
    w->T = ar.take<double>((size_t)B * b * b);
    w->Cm = ar.take<double>((size_t)B * b * b);
    w->theta = ar.take<double>((size_t)B * b);
    w->rn2 = ar.take<double>((size_t)B * b);
    w->P = ar.take<double>((size_t)B * 8 * b);
    w->coef = ar.take<double>((size_t)B * degree_cap * 3);
    w->lz_alpha = ar.take<double>((size_t)kLanczosStepsMax * B * b);
    w->lz_beta2 = ar.take<double>((size_t)kLanczosStepsMax * B * b);
    w->state = ar.take<EigState>(B);
    w->done = ar.take<int32_t>(B);
    w->n_active = ar.take<int32_t>(1);
    const size_t cap = paired_capacity(B, n, P);
    w->pcount = ar.take<int32_t>((size_t)B * ((n + 1) / 2));
    w->pent = ar.take<char>(cap * paired_entry_bytes(D));
    w->pent32 = ar.take<char>(cap * paired_entry32_bytes(D));
    w->Xa32 = ar.take<float>(vec);
    w->Xb32 = ar.take<float>(vec);
    w->Xc32 = ar.take<float>(vec);
    w->skip32 = ar.take<int32_t>(B);
    w->skip64 = ar.take<int32_t>(B);
    (void)nz;
    // structure-resident path (ANM ensembles of small structures)
    ResLayout& L = w->res;
    L.cols = (D == 3) ? resident_cols(n, b) : 0;
    if (L.cols > 0) {
        L.rpw = 128 / L.cols;
        resident_shape(n, L.cols, &L.G, &L.nsplit, &L.rec_mul, &L.rec_pad);
        L.rec = ar.take<ResRec>(resident_capacity(B, n, P, L.cols));
        L.gstart = ar.take<int32_t>((size_t)B * (L.G + 1));
        L.order = ar.take<uint16_t>((size_t)B * L.G * L.rpw);
        L.diag32 = ar.take<float>((size_t)B * n * 12);
        L.flag = ar.take<int32_t>(1);
        L.est = ar.take<double>(B);
        L.app_counter = ar.take<unsigned long long>(1);
    }
    return ar.off;
}

constexpr int kDegreeCap = 64;

}  // namespace scb

using namespace scb;

extern "C" size_t scb_eig_lowest_workspace_bytes(int D, int B, int n, int b, int nz, int64_t P) {
    Arena ar(nullptr, 0);
    EigWork w;
    return carve(ar, &w, D, B, n, b, nz, kDegreeCap, nullptr, P) + 256;
}

extern "C" int scb_eig_lowest(int D, int B, int n, int64_t P, const int64_t* rowptr, const int32_t* col,
                              const double* offdiag, const double* diag, const double* gersh, const double* Z,
                              int nz, int k, int b, double tol, int max_outer, int degree, uint64_t seed,
                              double* eigval, double* X, double* resid, int32_t* iters, void* workspace,
                              size_t workspace_bytes, void* stream) {
    if (!rowptr || !col || !offdiag || !diag || !gersh || !eigval || !X || !resid || !iters || !workspace)
        return SCB_ERR_INVALID;
    if ((D != 1 && D != 3) || B < 1 || n < 1 || k < 1 || k > b) return SCB_ERR_INVALID;
    if (b != 32 && b != 64) return SCB_ERR_UNSUPPORTED;
    if (degree < 2 || degree > kDegreeCap || max_outer < 1) return SCB_ERR_INVALID;
    const int64_t N = (int64_t)D * n;
    if (N < b + nz) return SCB_ERR_UNSUPPORTED;  // tiny systems: use scb_eig_full
    cudaStream_t st = as_stream(stream);
    Arena ar(workspace, workspace_bytes);
    EigWork w;
    carve(ar, &w, D, B, n, b, nz, kDegreeCap, X, P);
    if (!ar.ok()) return SCB_ERR_WORKSPACE;
    const int32_t* done = w.done;
    // row-paired copy of the operator for the register-blocked SpMM
    SCB_TRY(build_paired(D, B, n, rowptr, col, offdiag, diag, w.pcount, w.pent, st));
    auto apply = [&](const double* Xin, const double* Win, double* Yout, const double* cf, int stride) {
        return spmm_paired(D, B, n, rowptr, w.pcount, w.pent, Xin, Win, Yout, b, cf, stride, done, st);
    };

    // early outer iterations filter in FP32 (see spmm_paired_f32.cu); SCB_FP32=0 disables it
    int allow32 = 1;
    if (const char* env = getenv("SCB_FP32")) allow32 = atoi(env) != 0;
    double switch_tol = 3e-6;  // x spectrum bound: ~50x above the FP32 stagnation level (~5e-8 * ub)
    if (const char* env = getenv("SCB_SWITCH_TOL")) switch_tol = atof(env) > 0.0 ? atof(env) : switch_tol;
    int lz_steps = kLanczosSteps;
    if (const char* env = getenv("SCB_LANCZOS")) lz_steps = (atoi(env) >= 2 && atoi(env) <= kLanczosStepsMax) ? atoi(env) : lz_steps;
    // structure-resident residual-form path (resident.cuh): ANM blocks of the form -v v^T, structure fits in smem
    bool resident = w.res.cols > 0;
    if (const char* env = getenv("SCB_RESIDENT")) resident = resident && atoi(env) != 0;
    if (resident) {
        SCB_TRY(resident_build(B, n, rowptr, col, offdiag, diag, w.res, st));
        int32_t bad = 0;
        SCB_CUDA(cudaMemcpyAsync(&bad, w.res.flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaStreamSynchronize(st));
        if (bad) resident = false;   // a block with positive t (negative force constant): general 3x3 kernels
    }
    if (resident) allow32 = 0;
    if (allow32) SCB_TRY(build_paired32(D, paired_capacity(B, n, P), w.pent, w.pent32, st));
    SCB_TRY(state_init(B, gersh, w.state, w.done, w.n_active, w.skip32, w.skip64, allow32, degree, st));
    SCB_TRY(rand_init((int64_t)B * N * b, seed, w.A, st));
    SCB_CUDA(cudaMemsetAsync(w.rn2, 0, sizeof(double) * (size_t)B * b, st));

    // ---- spectrum bound: column-wise Lanczos on a copy of the random block (V=Bf, Vprev=Cf, W=HX)
    if (resident && N > 4 * lz_steps) {
        double ub_factor = 1.03;
        if (const char* env = getenv("SCB_UBFACTOR")) ub_factor = atof(env) > 1.0 ? atof(env) : ub_factor;
        SCB_TRY(resident_lanczos(B, n, b, rowptr, w.res, lz_steps, seed ^ 0x9e3779b97f4a7c15ull, ub_factor, w.state, st));
    } else if (N > 4 * lz_steps) {
        const size_t vec_bytes = sizeof(double) * (size_t)B * N * b;
        SCB_CUDA(cudaMemcpyAsync(w.Bf, w.A, vec_bytes, cudaMemcpyDeviceToDevice, st));
        SCB_CUDA(cudaMemsetAsync(w.Cf, 0, vec_bytes, st));
        SCB_TRY(coldot(B, N, b, w.Bf, w.Bf, w.rn2, st));
        SCB_TRY(lanczos_axpy(B, N, b, 2, w.Bf, nullptr, nullptr, nullptr, nullptr, w.rn2, st));
        for (int j = 0; j < lz_steps; ++j) {
            double* aj = w.lz_alpha + (size_t)j * B * b;
            double* bj = w.lz_beta2 + (size_t)j * B * b;
            const double* bprev = j > 0 ? w.lz_beta2 + (size_t)(j - 1) * B * b : nullptr;
            SCB_TRY(apply(w.Bf, nullptr, w.HX, nullptr, 0));
            SCB_TRY(coldot(B, N, b, w.Bf, w.HX, aj, st));
            SCB_TRY(lanczos_axpy(B, N, b, 0, w.Bf, w.Cf, w.HX, aj, bprev, nullptr, st));
            SCB_TRY(coldot(B, N, b, w.HX, w.HX, bj, st));
            if (j + 1 < lz_steps) SCB_TRY(lanczos_axpy(B, N, b, 1, w.Bf, w.Cf, w.HX, nullptr, nullptr, bj, st));
        }
        SCB_TRY(lanczos_bound(B, b, lz_steps, w.lz_alpha, w.lz_beta2, w.state, st));
        SCB_CUDA(cudaMemsetAsync(w.rn2, 0, sizeof(double) * (size_t)B * b, st));
    }

    // one pinned word per host thread for the convergence poll (allocated once, reused by every call)
    static thread_local int32_t* h_active = nullptr;
    if (!h_active) SCB_CUDA(cudaMallocHost(&h_active, sizeof(int32_t)));
    *h_active = B;
    int status = SCB_OK;
    double* cur = w.A;  // block to orthonormalise at the top of the Rayleigh-Ritz stage
    if (resident) {
        // Per outer iteration: HX (FP64) -> Rayleigh-Ritz of the pencil (X^T H X, X^T X) -> residuals ->
        // convergence / filter bounds -> ONE launch of the structure-resident FP32 filter, which returns the
        // deflated block X + |r| z.  The corrected block is nearly orthonormal (z ~ the error of X), so no
        // separate orthonormalisation pass is needed: the Cholesky factor of X^T X inside rr_kernel does it.
        if ((status = deflate(B, N, b, nz, Z, w.A, w.P, nullptr, st)) != SCB_OK) return status;
        const bool prof = profile_enabled();
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        double prof_ms = 0.0;
        long long prof_launches = 0;
        SCB_CUDA(cudaMemsetAsync(w.res.app_counter, 0, sizeof(unsigned long long), st));
        if (prof) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); }
        for (int outer = 0; outer <= max_outer; ++outer) {
            if ((status = apply(w.A, nullptr, w.HX, nullptr, 0)) != SCB_OK) break;
            if (b == 32) {
                if ((status = gram2_dmma(B, N, w.A, w.HX, w.S, w.T, done, st)) != SCB_OK) break;
                if ((status = small_rr(B, b, w.S, w.T, w.theta, w.Cm, done, 1, st)) != SCB_OK) break;
                if ((status = rotate_resid_dmma(B, N, w.Cm, w.A, w.HX, w.theta, w.rn2, done, st)) != SCB_OK) break;
            } else {
                if ((status = gram(B, N, b, w.A, w.A, w.S, done, st)) != SCB_OK) break;
                if ((status = gram(B, N, b, w.A, w.HX, w.T, done, st)) != SCB_OK) break;
                if ((status = small_rr(B, b, w.S, w.T, w.theta, w.Cm, done, 1, st)) != SCB_OK) break;
                if ((status = rotate(B, N, b, w.Cm, w.A, w.A, w.HX, w.HX, done, st)) != SCB_OK) break;
                if ((status = zero_active_rn2(B, b, w.rn2, done, st)) != SCB_OK) break;
                if ((status = residual_norms(B, N, b, w.A, w.HX, w.theta, w.rn2, done, st)) != SCB_OK) break;
            }
            if ((status = state_update(B, b, k, tol, w.theta, w.rn2, w.state, w.done, w.n_active, resid, w.skip32,
                                       w.skip64, 0, switch_tol, degree, st)) != SCB_OK)
                break;
            if (cudaMemcpyAsync(h_active, w.n_active, sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaStreamSynchronize(st) != cudaSuccess) {
                set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__);
                status = SCB_ERR_CUDA;
                break;
            }
            if (*h_active == 0 || outer == max_outer) break;
            if (prof) cudaEventRecord(ev0, st);
            if ((status = resident_filter(B, n, b, rowptr, w.res, w.A, w.HX, w.theta, w.rn2, w.state, done, Z, nz,
                                          w.A, k, tol, st)) != SCB_OK)
                break;
            if (prof) {   // scb_profile: device time of the filter launches (the stream is synchronised once per
                          // outer iteration anyway; the extra synchronisation only exists while profiling)
                cudaEventRecord(ev1, st);
                cudaEventSynchronize(ev1);
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev0, ev1);
                prof_ms += ms;
                ++prof_launches;
            }
        }
        if (prof) {
            unsigned long long apps = 0;
            cudaMemcpyAsync(&apps, w.res.app_counter, sizeof(apps), cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            profile_add(prof_ms, prof_launches, apps);
            cudaEventDestroy(ev0);
            cudaEventDestroy(ev1);
        }
        if (status != SCB_OK) return status;
        SCB_TRY(gather_results(B, b, w.theta, w.state, eigval, iters, st));
        return *h_active == 0 ? SCB_OK : SCB_ERR_NOT_CONVERGED;
    }
    for (int outer = 0; outer <= max_outer; ++outer) {
        if (outer > 0) {
            // ---- 1. Chebyshev filter of degree `degree` applied to the basis in A
            if ((status = cheb_coef(B, degree, w.state, done, w.coef, st)) != SCB_OK) break;
            double* prev = w.A;
            double* curb = w.Bf;
            double* next = w.Cf;
            float *p32 = w.Xa32, *c32 = w.Xb32, *n32 = w.Xc32;
            const int64_t per_struct = N * b;
            if (allow32) {
                // structures in the FP32 phase: basis -> float, filter on the float buffers
                if ((status = block_to_f32(B, per_struct, w.A, p32, w.skip32, st)) != SCB_OK) break;
                if ((status = spmm_paired_f32(D, B, n, rowptr, w.pcount, w.pent32, p32, nullptr, c32, b, w.coef,
                                              degree * 3, w.skip32, st)) != SCB_OK) break;
            }
            if ((status = spmm_paired(D, B, n, rowptr, w.pcount, w.pent, prev, nullptr, curb, b, w.coef, degree * 3,
                                      w.skip64, st)) != SCB_OK) break;
            for (int d = 1; d < degree; ++d) {
                if (allow32) {
                    if ((status = spmm_paired_f32(D, B, n, rowptr, w.pcount, w.pent32, c32, p32, n32, b,
                                                  w.coef + 3 * d, degree * 3, w.skip32, st)) != SCB_OK) break;
                    float* t32 = p32; p32 = c32; c32 = n32; n32 = t32;
                }
                if ((status = spmm_paired(D, B, n, rowptr, w.pcount, w.pent, curb, prev, next, b, w.coef + 3 * d,
                                          degree * 3, w.skip64, st)) != SCB_OK) break;
                double* t = prev; prev = curb; curb = next; next = t;
            }
            if (status != SCB_OK) break;
            if (allow32 && (status = block_to_f64(B, per_struct, c32, curb, w.skip32, st)) != SCB_OK) break;
            cur = curb;
        }
        // ---- 2. deflate the analytic null space
        if ((status = deflate(B, N, b, nz, Z, cur, w.P, done, st)) != SCB_OK) break;
        // ---- 3. orthonormalise: A <- cur * L^-T
        if ((status = gram(B, N, b, cur, cur, w.S, done, st)) != SCB_OK) break;
        if ((status = small_rr(B, b, w.S, nullptr, w.theta, w.Cm, done, 0, st)) != SCB_OK) break;
        if ((status = rotate(B, N, b, w.Cm, cur, w.A, nullptr, nullptr, done, st)) != SCB_OK) break;
        // ---- 4. HX = H A
        if ((status = apply(w.A, nullptr, w.HX, nullptr, 0)) != SCB_OK) break;
        // ---- 5. Rayleigh-Ritz on the (nearly orthonormal) basis
        if ((status = gram(B, N, b, w.A, w.A, w.S, done, st)) != SCB_OK) break;
        if ((status = gram(B, N, b, w.A, w.HX, w.T, done, st)) != SCB_OK) break;
        if ((status = small_rr(B, b, w.S, w.T, w.theta, w.Cm, done, 1, st)) != SCB_OK) break;
        if ((status = rotate(B, N, b, w.Cm, w.A, w.A, w.HX, w.HX, done, st)) != SCB_OK) break;
        // ---- 6. residuals + state
        if ((status = zero_active_rn2(B, b, w.rn2, done, st)) != SCB_OK) break;
        if ((status = residual_norms(B, N, b, w.A, w.HX, w.theta, w.rn2, done, st)) != SCB_OK) break;
        if ((status = state_update(B, b, k, tol, w.theta, w.rn2, w.state, w.done, w.n_active, resid, w.skip32,
                                   w.skip64, allow32, switch_tol, degree, st)) != SCB_OK)
            break;
        if (cudaMemcpyAsync(h_active, w.n_active, sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__);
            status = SCB_ERR_CUDA;
            break;
        }
        if (*h_active == 0) break;
    }
    const int active = *h_active;
    if (status != SCB_OK) return status;
    SCB_TRY(gather_results(B, b, w.theta, w.state, eigval, iters, st));
    return active == 0 ? SCB_OK : SCB_ERR_NOT_CONVERGED;
}
