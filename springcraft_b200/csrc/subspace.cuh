// Internal interfaces between the eigensolver translation units.
#pragma once
#include "common.cuh"

namespace scb {

struct EigState {
    double ub;       // upper bound of the spectrum
    double lo;       // lower edge of the damped interval (largest Ritz value of the block)
    double a0;       // lowest Ritz value (scaling point of the filter)
    double ub_safe;  // guaranteed bound (Gershgorin); ub may be a tighter Lanczos estimate
    double prev_res; // worst wanted residual of the previous outer iteration (FP32 stagnation guard)
    int32_t iters;
    int32_t converged;
    int32_t degree_next;  // filter degree of the next outer iteration (<= the launch degree; last iterations shrink)
    int32_t degree_used;
};

int spmm_cheb(int D, int B, int n, const int64_t* rowptr, const int32_t* col, const double* offdiag,
              const double* diag, const double* X, const double* W, double* Y, int b, const double* coef,
              int coef_stride, const int32_t* done, cudaStream_t st);
size_t paired_capacity(int B, int n, int64_t P);
size_t paired_entry_bytes(int D);
int build_paired(int D, int B, int n, const int64_t* rowptr, const int32_t* col, const double* offdiag,
                 const double* diag, int32_t* pcount, void* pent, cudaStream_t st);
int spmm_paired(int D, int B, int n, const int64_t* rowptr, const int32_t* pcount, const void* pent,
                const double* X, const double* W, double* Y, int b, const double* coef, int coef_stride,
                const int32_t* done, cudaStream_t st);
size_t paired_entry32_bytes(int D);
int build_paired32(int D, size_t cap, const void* pent, void* pent32, cudaStream_t st);
int block_to_f32(int B, int64_t per_struct, const double* in, float* out, const int32_t* skip, cudaStream_t st);
int block_to_f64(int B, int64_t per_struct, const float* in, double* out, const int32_t* skip, cudaStream_t st);
int spmm_paired_f32(int D, int B, int n, const int64_t* rowptr, const int32_t* pcount, const void* pent32,
                    const float* X, const float* W, float* Y, int b, const double* coef, int coef_stride,
                    const int32_t* skip, cudaStream_t st);
int gram(int B, int64_t N, int b, const double* A, const double* Bm, double* G, const int32_t* done, cudaStream_t st);
int small_rr(int B, int b, const double* S, const double* T, double* theta, double* C, const int32_t* done,
             int mode, cudaStream_t st, int nact = 0);
int rotate(int B, int64_t N, int b, const double* C, const double* Xin, double* Xout, const double* Yin,
           double* Yout, const int32_t* done, cudaStream_t st);
int deflate(int B, int64_t N, int b, int nz, const double* Z, double* X, double* P, const int32_t* done,
            cudaStream_t st);
int residual_norms(int B, int64_t N, int b, const double* X, const double* HX, const double* theta, double* rn2,
                   const int32_t* done, cudaStream_t st);
int state_init(int B, const double* gersh, EigState* st, int32_t* done, int32_t* n_active, int32_t* skip32,
               int32_t* skip64, int allow32, int degree, cudaStream_t s);
int zero_active_rn2(int B, int b, double* rn2, const int32_t* done, cudaStream_t s);
int state_update(int B, int b, int k, double tol, const double* theta, const double* rn2, EigState* st,
                 int32_t* done, int32_t* n_active, double* resid, int32_t* skip32, int32_t* skip64, int allow32,
                 double switch_tol, int degree, cudaStream_t s);
int cheb_coef(int B, int degree, const EigState* st, const int32_t* done, double* coef, cudaStream_t s);
int rand_init(int64_t total, uint64_t seed, double* X, cudaStream_t s);
int coldot(int B, int64_t N, int b, const double* A, const double* Bm, double* out, cudaStream_t st);
int lanczos_axpy(int B, int64_t N, int b, int mode, double* V, double* Vprev, double* W, const double* alpha,
                 const double* beta_prev, const double* nrm2, cudaStream_t st);
int lanczos_bound_plain(int B, int b, int k, const double* alpha, const double* beta2, double factor, double* out,
                        cudaStream_t st);
int lanczos_bound(int B, int b, int k, const double* alpha, const double* beta2, EigState* state, cudaStream_t st);
// FP64 tensor-core versions for 32-column blocks (tallskinny_dmma.cu): S = X^T X and T = X^T HX in one pass;
// X <- X C, HX <- HX C fused with the squared residual norms of the rotated pairs
int gram2_dmma(int B, int64_t N, const double* X, const double* HX, double* S, double* T, const int32_t* done,
               cudaStream_t st);
int rotate_resid_dmma(int B, int64_t N, const double* C, double* X, double* HX, const double* theta, double* rn2,
                      const int32_t* done, cudaStream_t st);
int gather_results(int B, int b, const double* theta, const EigState* st, double* eigval, int32_t* iters,
                   cudaStream_t s);

}  // namespace scb
