// K3c: full symmetric eigendecomposition of a dense matrix (nma.py:61,
// np.linalg.eigh): all eigenvalues ascending + eigenvectors as rows.
//
//  * N <= 64  : one CTA per matrix, two-sided cyclic Jacobi entirely in shared
//               memory (the Rayleigh-Ritz kernel with S = I).
//  * N > 64   : tridiagonalisation + divide and conquer + back-transformation -- see eig_full_tridiag.cu
//               (N <= 9,200; beyond that, or without cooperative launch: block Jacobi, eig_full_block.cu).
#include <stdlib.h>

#include "subspace.cuh"

namespace scb {

int eig_full_block(int B, int N, double* A, double* eigval, double* modes, void* workspace, size_t workspace_bytes,
                   cudaStream_t st);
size_t eig_full_block_workspace_bytes(int B, int N);
// eig_full_tridiag.cu: Householder tridiagonalisation + divide and conquer + back-transformation
bool eig_full_tridiag_supported(int N);
int eig_full_tridiag(int B, int N, double* A, double* eigval, double* modes, void* workspace, size_t workspace_bytes,
                     cudaStream_t st);
size_t eig_full_tridiag_workspace_bytes(int B, int N);

// solver: SCB_EIG_AUTO / SCB_EIG_JACOBI / SCB_EIG_TRIDIAG.  SCB_EIG_FULL=jacobi forces the block-Jacobi solver in
// automatic mode (A/B measurements).  The tridiagonal solver is a cooperative launch that needs every SM: a caller
// whose other streams keep SMs busy for an unknown time can ask for the block-Jacobi solver (ordinary launches).
static bool use_tridiag(int N, int solver) {
    if (solver == SCB_EIG_JACOBI || !eig_full_tridiag_supported(N)) return false;
    if (solver == SCB_EIG_TRIDIAG) return true;
    const char* env = getenv("SCB_EIG_FULL");
    if (env && env[0] == 'j') return false;
    return true;
}

// T[s] = lower triangle of A[s] mirrored, padded to PxP with huge decoupled diagonal; S[s] = I
__global__ void pad_symmetric_kernel(int N, int P, const double* __restrict__ A, double* __restrict__ T,
                                     double* __restrict__ S) {
    const int s = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P * P) return;
    const int i = q / P, j = q % P;
    double v = 0.0;
    if (i < N && j < N) v = (j <= i) ? A[((int64_t)s * N + i) * N + j] : A[((int64_t)s * N + j) * N + i];
    // padding rows/columns are zero: decoupled (a_pq == 0 is never rotated) and kept in place by the sort
    T[(int64_t)s * P * P + q] = v;
    S[(int64_t)s * P * P + q] = (i == j) ? 1.0 : 0.0;
}

// modes[s][k][r] = C[s][r][k]
__global__ void unpad_modes_kernel(int N, int P, const double* __restrict__ C, const double* __restrict__ theta,
                                   double* __restrict__ modes, double* __restrict__ eigval) {
    const int s = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= N * N) return;
    const int k = q / N, r = q % N;
    modes[(int64_t)s * N * N + q] = C[(int64_t)s * P * P + (int64_t)r * P + k];
    if (r == 0) eigval[(int64_t)s * N + k] = theta[(int64_t)s * P + k];
}

}  // namespace scb

using namespace scb;

extern "C" size_t scb_eig_full_workspace_bytes_ex(int solver, int B, int N) {
    if (N <= 64) {
        const size_t P = N <= 32 ? 32 : 64;
        return 3 * (((size_t)B * P * P * sizeof(double) + 255) & ~size_t(255)) +
               (((size_t)B * P * sizeof(double) + 255) & ~size_t(255)) + 256;
    }
    if (use_tridiag(N, solver)) return eig_full_tridiag_workspace_bytes(B, N);
    return eig_full_block_workspace_bytes(B, N);
}

extern "C" size_t scb_eig_full_workspace_bytes(int B, int N) { return scb_eig_full_workspace_bytes_ex(SCB_EIG_AUTO, B, N); }

extern "C" int scb_eig_full_ex(int solver, int B, int N, double* A, double* eigval, double* modes, void* workspace,
                               size_t workspace_bytes, void* stream) {
    if (!A || !eigval || !modes || !workspace || B < 1 || N < 1 || solver < SCB_EIG_AUTO || solver > SCB_EIG_TRIDIAG)
        return SCB_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    if (N > 64 && use_tridiag(N, solver)) return eig_full_tridiag(B, N, A, eigval, modes, workspace, workspace_bytes, st);
    if (N > 64) return eig_full_block(B, N, A, eigval, modes, workspace, workspace_bytes, st);
    const int P = N <= 32 ? 32 : 64;
    Arena ar(workspace, workspace_bytes);
    double* T = ar.take<double>((size_t)B * P * P);
    double* S = ar.take<double>((size_t)B * P * P);
    double* C = ar.take<double>((size_t)B * P * P);
    double* theta = ar.take<double>((size_t)B * P);
    if (!ar.ok()) return SCB_ERR_WORKSPACE;
    dim3 g1((unsigned)ceil_div(P * P, 256), (unsigned)B);
    pad_symmetric_kernel<<<g1, 256, 0, st>>>(N, P, A, T, S);
    SCB_LAUNCH_CHECK();
    SCB_TRY(small_rr(B, P, S, T, theta, C, nullptr, 1, st, N));
    dim3 g2((unsigned)ceil_div(N * N, 256), (unsigned)B);
    unpad_modes_kernel<<<g2, 256, 0, st>>>(N, P, C, theta, modes, eigval);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_eig_full(int B, int N, double* A, double* eigval, double* modes, void* workspace,
                            size_t workspace_bytes, void* stream) {
    return scb_eig_full_ex(SCB_EIG_AUTO, B, N, A, eigval, modes, workspace, workspace_bytes, stream);
}
