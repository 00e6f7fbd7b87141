/*
 * scb200.h -- C ABI of libscb200.so, the B200 (sm_100a) implementation of the
 * springcraft elastic-network hot path.
 *
 * The reference (biotite-dev/springcraft 0.3.0) has no FFI layer: the seam is
 * its Python API.  Every entry point below replaces the
 * Python/NumPy code cited next to it; springcraft_b200/ binds them with ctypes
 * (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns an int status: 0 = OK, <0 = scb_status error.
 *     No exception crosses the ABI.
 *   - unless the name ends in _host, every pointer is a DEVICE pointer owned by
 *     the caller and `stream` is a cudaStream_t passed as void*.  Functions are
 *     asynchronous with respect to the host unless documented otherwise.
 *   - re-entrant per stream; the only process-wide state is the launch counter, the last-error text and the
 *     private scratch pool of the whole-path entry points (scb_trim_pool).
 *   - B = number of structures in the batch (ensemble), n = nodes per structure,
 *     D = 1 (GNM / Kirchhoff) or 3 (ANM / Hessian), N = D*n,
 *     P = number of ORDERED contact pairs of the whole batch.
 *   - coordinates are SoA: xyz[B][3][n] fp64.
 *   - the neighbour list is one CSR over the B*n rows of the batch:
 *     rowptr[B*n+1] int64 (global offsets), col[P] int32 (node index inside the
 *     structure), columns ascending inside a row  ==> walking it reproduces the
 *     reference's np.where(adj) pair order (interaction.py:177-178).
 *   - matrices are BSR with DxD blocks: offdiag[P][D*D] (row-major block),
 *     diag[B*n][D*D].
 *   - block vectors are row-major X[B][N][b] (b contiguous).
 */
#ifndef SCB200_H
#define SCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCB_VERSION 100

typedef enum scb_status {
    SCB_OK = 0,
    SCB_ERR_INVALID = -1,       /* bad argument                      -> ValueError   */
    SCB_ERR_CUDA = -2,          /* CUDA runtime error                -> RuntimeError */
    SCB_ERR_ABOVE_CUTOFF = -3,  /* Tabulated bin above last edge     -> ValueError (forcefield.py:526-533) */
    SCB_ERR_WORKSPACE = -4,     /* workspace too small               -> RuntimeError */
    SCB_ERR_NOT_CONVERGED = -5, /* eigensolver hit max iterations    -> RuntimeError */
    SCB_ERR_UNSUPPORTED = -6    /* size / mode not implemented       -> NotImplementedError */
} scb_status;

/* force-field kinds: forcefield.py:264-289, 292-330, 333-366, 369-545 */
enum {
    SCB_FF_INVARIANT = 0,
    SCB_FF_HINSEN = 1,
    SCB_FF_PFREE = 2,
    SCB_FF_TABULATED = 3,        /* 20x20xk residue-pair tables + per-atom attributes */
    SCB_FF_TABULATED_DENSE = 4,  /* explicit (n,n,k) float32 interaction_matrix (forcefield.py:429-434) */
    SCB_FF_EXTERNAL = 5          /* caller-supplied fc[P] (user ForceField subclass, doc/advanced.rst) */
};

/* POD force-field descriptor.  All pointers are device pointers (or NULL). */
typedef struct scb_ff_desc {
    int32_t kind;
    int32_t nbins;             /* tabulated: number of distance bins k (>=1)            */
    int32_t patched;           /* PatchedForceField semantics (forcefield.py:183-226)   */
    int32_t n_pair_on;         /* number of switched-on pairs                           */
    double cutoff_sq;          /* cutoff**2, or < 0 for "no cutoff" (all pairs)         */
    const float *bonded;       /* [20][20][k]                                           */
    const float *intra;        /* [20][20][k]                                           */
    const float *inter;        /* [20][20][k]                                           */
    const double *edges_sq;    /* [k] squared bin edges (NULL when k==1)                */
    const uint8_t *res_type;   /* [n] residue index in ACDEFGHIKLMNPQRSTVWY order       */
    const int32_t *chain;      /* [n] chain label id                                    */
    const uint8_t *bonded_next;/* [n] atom a and a+1 are peptide bonded                 */
    const float *dense_table;  /* [n][n][k] for SCB_FF_TABULATED_DENSE                  */
    const double *external_fc; /* [P] for SCB_FF_EXTERNAL                               */
    const int32_t *pair_on;    /* [n_pair_on][2]                                        */
    const double *pair_on_fc;  /* [n_pair_on]                                           */
} scb_ff_desc;

/* contact patches applied to the adjacency (interaction.py:193-213) */
typedef struct scb_patch {
    int32_t n_pair_off;
    int32_t n_pair_on;
    const uint8_t *dead;       /* [n] 1 = contact_shutdown atom, or NULL                */
    const int32_t *pair_off;   /* [n_pair_off][2] or NULL                               */
    const int32_t *pair_on;    /* [n_pair_on][2]  or NULL                               */
} scb_patch;

int scb_version(void);
const char *scb_status_string(int status);
/* last CUDA error text seen by this thread (for SCB_ERR_CUDA) */
const char *scb_last_cuda_error(void);
/* number of kernels this library has launched in this process (bench accounting) */
uint64_t scb_launch_count(void);
/* Instrumentation for bench.py (single host thread): while enabled, scb_eig_lowest brackets every launch of the
 * structure-resident filter kernel with CUDA events on its stream.  Each call returns and resets the accumulated
 * device time (ms), the number of filter launches and the number of (structure x operator application) units
 * they processed; enable = 1 / 0 switches the instrumentation, a negative value leaves it unchanged. */
int scb_profile(int enable, double *filter_ms, int64_t *filter_launches, int64_t *filter_applications);

/* ---------------------------------------------------------------------------
 * K1  contact search.   Replaces interaction.py:149-178 (+ biotite CellList,
 * interaction.py:155-159) and _patch_adjacency_matrix (interaction.py:193-213).
 * Criterion: fp64 ((dx*dx)+(dy*dy))+(dz*dz) <= cutoff_sq, no FMA contraction.
 * Two-phase count -> scan -> fill.
 * ------------------------------------------------------------------------- */
/* rowcount[B*n] int32.  use_cell_list != 0 selects the cell-list kernels
 * (B must be 1); both variants produce identical results. */
int scb_contacts_count(const double *xyz, int B, int n, double cutoff_sq,
                       const scb_patch *patch, int use_cell_list,
                       int32_t *rowcount, void *stream);
/* exclusive scan of rowcount -> rowptr[B*n+1] (int64); scratch >= scb_scan_scratch_bytes */
size_t scb_scan_scratch_bytes(int64_t nrows);
int scb_contacts_scan(const int32_t *rowcount, int64_t nrows, int64_t *rowptr,
                      void *scratch, void *stream);
int scb_contacts_fill(const double *xyz, int B, int n, double cutoff_sq,
                      const scb_patch *patch, int use_cell_list,
                      const int64_t *rowptr, int32_t *col, void *stream);
/* pairs[P][2] int64 (interaction.py:177-178); struct-local node indices */
int scb_pairs_materialize(const int64_t *rowptr, const int32_t *col, int B, int n,
                          int64_t *pairs, void *stream);

/* ---------------------------------------------------------------------------
 * K2  fused force constant + Kirchhoff / Hessian block assembly.
 * Replaces ForceField.force_constant (forcefield.py:183-226,283-284,318-326,
 * 361-362,515-533) + compute_kirchhoff / compute_hessian (interaction.py:47-52,
 * 93-109) + mass weighting (anm.py:89-96,112-113; gnm.py:85-89,105-106).
 * Off-diagonal block = ((-fc/sq)*d_a)*d_b; diagonal = -(sum over the row's
 * blocks in ascending column order)  => bit-identical to the reference.
 * masses: [n] (shared by the batch) or NULL.  gersh[B]: optional upper bound
 * of the spectrum (Gershgorin), or NULL.  status_flag: device int32, set to
 * SCB_ERR_ABOVE_CUTOFF when a tabulated lookup falls above the last edge.
 * ------------------------------------------------------------------------- */
int scb_assemble(int D, const double *xyz, int B, int n, const scb_ff_desc *ff,
                 const int64_t *rowptr, const int32_t *col, const double *masses,
                 double *offdiag, double *diag, double *gersh, int32_t *status_flag,
                 void *stream);
/* ForceField.force_constant(atom_i, atom_j, sq_distance) for explicit triples
 * (forcefield.py:67-95): out[P] fp64 (tabulated values are float32 widened). */
int scb_force_constant(const scb_ff_desc *ff, int n, const int32_t *atom_i,
                       const int32_t *atom_j, const double *sq, int64_t P, double *out,
                       int32_t *status_flag, void *stream);
/* disp[P][3] = x_j - x_i (may be NULL) and sq[P] for the CSR pairs
 * (interaction.py:182-184); feeds user-defined ForceField callbacks. */
int scb_pair_geometry(const double *xyz, int B, int n, const int64_t *rowptr,
                      const int32_t *col, double *disp, double *sq, void *stream);
/* dense [B][N][N] row-major (the reference's return layout, interaction.py:106-109) */
int scb_densify(int D, int B, int n, const int64_t *rowptr, const int32_t *col,
                const double *offdiag, const double *diag, double *dense, void *stream);
/* all-pairs force fields (cutoff None): dense row slab rows [row0,row1) of the
 * (N x N) matrix of ONE structure, written to dense[(row1-row0)*D][N]; diagonal
 * blocks included (local row sums).  SURVEY 8e config C4. */
int scb_assemble_dense_allpairs(int D, const double *xyz, int n, const scb_ff_desc *ff,
                                const double *masses, int row0, int row1,
                                double *dense, void *stream);

/* ---------------------------------------------------------------------------
 * K3  eigensolvers.  Replace nma.eigen / np.linalg.eigh (nma.py:29-63).
 * ------------------------------------------------------------------------- */
/* y = H x on block vectors:  Y[B][N][b] = BSR * X[B][N][b]  (b = 32, 64 or 128) */
int scb_spmm(int D, int B, int n, const int64_t *rowptr, const int32_t *col,
             const double *offdiag, const double *diag, const double *X, double *Y,
             int b, void *stream);
/* Row-paired operator used by the eigensolver (two consecutive block rows merged
 * into (2D x D) blocks, diagonal folded in; spmm_paired.cu).  `paired` is one
 * device buffer of scb_paired_bytes(); the two offsets (may be NULL) locate the
 * per-pair counts and the merged-contact records {blocks, column} inside it. */
size_t scb_paired_bytes(int D, int B, int n, int64_t P, size_t *count_off, size_t *entry_off);
int scb_paired_build(int D, int B, int n, int64_t P, const int64_t *rowptr, const int32_t *col,
                     const double *offdiag, const double *diag, void *paired, void *stream);
int scb_spmm_paired(int D, int B, int n, int64_t P, const int64_t *rowptr, const void *paired,
                    const double *X, double *Y, int b, void *stream);
/* orthonormal basis of the analytic null space (rigid-body modes, mass weighted
 * when masses != NULL): Z[B][N][nz], nz = 6 (D=3) or 1 (D=1) */
int scb_rigid_basis(int D, const double *xyz, int B, int n, const double *masses,
                    double *Z, void *stream);
/* lowest-k eigensolver: Chebyshev-filtered subspace iteration on the BSR SpMM
 * with a dense Rayleigh-Ritz (Jacobi) step.  Computes the b lowest modes of the
 * operator deflated by Z (pass nz=0/Z=NULL for no deflation); the first `k`
 * are converged to ||H x - theta x|| <= tol * theta_k.
 *   eigval[B][b], X[B][N][b] (column q = mode q), resid[B][b], iters[B].
 * Host-synchronous (polls a device convergence counter once per outer iteration).
 */
size_t scb_eig_lowest_workspace_bytes(int D, int B, int n, int b, int nz, int64_t P);
int scb_eig_lowest(int D, int B, int n, int64_t P, const int64_t *rowptr, const int32_t *col,
                   const double *offdiag, const double *diag, const double *gersh,
                   const double *Z, int nz, int k, int b, double tol, int max_outer,
                   int degree, uint64_t seed, double *eigval, double *X, double *resid,
                   int32_t *iters, void *workspace, size_t workspace_bytes, void *stream);
/* ---- dense row-slab operator + solver building blocks (all-pairs force fields).
 * The Python host code drives the same Chebyshev-filtered subspace iteration with these, putting a
 * torch.distributed all-gather of the row slabs of Y between scb_dense_slab_apply calls. ---- */
/* Y[rows][b] = H[rows,:] X                      (fused == 0)
 *            = alpha (H[rows,:] X - c X[rows]) - beta W[rows]   (fused != 0);  rows = [row0,row1) of
 * the N x N matrix, slab = those rows (row-major, N columns); X, W are full [N][b]; b % 64 == 0.
 * workspace (scb_dense_slab_workspace_bytes, zero-filled ONCE by the caller, reusable across calls on one
 * stream; may be NULL): lets the launcher split K over several CTAs per output tile when the slab has too
 * few tiles to fill the GPU (deterministic reduction: partial tiles are added in split order). */
size_t scb_dense_slab_workspace_bytes(int64_t N, int64_t row0, int64_t row1, int b);
int scb_dense_slab_apply(int64_t N, int64_t row0, int64_t row1, const double *slab,
                         const double *X, const double *W, double *Y, int b, int fused,
                         double alpha, double cshift, double beta, void *workspace,
                         size_t workspace_bytes, void *stream);
/* Same product with the all-gather of the row slabs fused into the epilogue: Y_all is a HOST array of
 * `world` device pointers, Y_all[p] = the full [N][b] output block of rank p mapped into this process
 * (scb_peer_open; the entry of this rank is its own scb_peer_alloc buffer).  Rows [row0,row1) are stored
 * into every block over NVLink peer memory; the caller runs a barrier collective before any rank reads
 * its block (replaces ncclAllGather of SURVEY 8e C4).  world <= 16; X, W must not alias an output. */
int scb_dense_slab_apply_allgather(int64_t N, int64_t row0, int64_t row1, const double *slab,
                                   const double *X, const double *W, double *const *Y_all, int world,
                                   int b, int fused, double alpha, double cshift, double beta,
                                   void *workspace, size_t workspace_bytes, void *stream);
/* Peer-mapped device buffers for the call above (one process per GPU): allocate (zero-filled), export a
 * 64-byte CUDA IPC handle, open the handle of another rank (enables peer access), close, free. */
int scb_peer_alloc(size_t bytes, void **dptr);
int scb_peer_free(void *dptr);
int scb_peer_export(void *dptr, unsigned char *handle64);
int scb_peer_open(const unsigned char *handle64, void **dptr);
int scb_peer_close(void *dptr);
/* out[0] = max row sum of |entries| over the slab rows (Gershgorin bound of the local rows) */
int scb_dense_gershgorin(int64_t N, int64_t rows, const double *slab, double *out, void *stream);
/* G[B][b][b] = A^T Bm  for block vectors [B][N][b]  (b = 32, 64, 128) */
int scb_gram(int B, int64_t N, int b, const double *A, const double *Bm, double *G, void *stream);
/* S = L L^T  ->  C = L^-T  (X C is orthonormal);  b <= 160 */
int scb_chol_orth(int B, int b, const double *S, double *C, void *stream);
/* Xout = Xin C (and Yout = Yin C when Yin != NULL); in place allowed */
int scb_rotate(int B, int64_t N, int b, const double *C, const double *Xin, double *Xout,
               const double *Yin, double *Yout, void *stream);
/* X <- X - Z (Z^T X);  scratch >= B*8*b doubles */
int scb_deflate(int B, int64_t N, int b, int nz, const double *Z, double *X, double *scratch,
                void *stream);
/* rn2[B][b] = squared column norms of HX - X diag(theta) */
int scb_residual_norms(int B, int64_t N, int b, const double *X, const double *HX,
                       const double *theta, double *rn2, void *stream);
/* out = in^T for a b x b matrix (mode rows -> rotation columns) */
int scb_transpose_small(int b, const double *in, double *out, void *stream);
/* ---- residual-form Chebyshev filter of the dense row-slab operator on the TF32 tensor cores (dense_tf32.cu).
 * For a Ritz pair (theta, x) with r = Hx - theta x the filtered vector is x + z, z = q(H) r / p(theta); z is
 * proportional to the residual, so it may be computed with a single-precision (TF32) copy of the slab while H X,
 * Rayleigh-Ritz and the residuals stay FP64.  Blocks of the filter are FP32 and TRANSPOSED: zT[b][ld],
 * ld = scb_tf32_ld(N) (N rounded up to a multiple of 4), b = 128.
 *   scb_dense_slab_to_f32:     slab32[rows][ld] <- slab[rows][N]
 *   scb_resform_prepare:       rhatT = ((HX - X theta') / |r|)^T with theta' = min(theta, lo), z1T (first filter step),
 *                              z0T = 0 and the per-column coefficient tables cA, cB [deg][b] of steps 1..deg-1
 *   scb_dense_slab_tf32_apply: rows [row0,row1) of  outT = cA (H zcur - cshift zcur + rhat) - cB zprev  (fused != 0,
 *                              cA/cB = the step's row of the tables, cshift = (ub+lo)/2; outT may alias zprevT) or
 *                              outT = H zcur (fused == 0); tcgen05.mma kind::tf32, TMA operands, TMEM accumulator
 *   scb_resform_finish:        X[N][b] += |r| z^T */
int64_t scb_tf32_ld(int64_t N);
int scb_dense_slab_to_f32(int64_t N, int64_t rows, const double *slab, float *slab32, float *slab32_lo,
                          void *stream);
int scb_resform_prepare(int64_t N, int b, int deg, const double *X, const double *HX, const double *theta,
                        const double *rn2, double lo, double ub, float *rhatT, float *z1T, float *z0T,
                        float *cA, float *cB, int split, void *stream);
int scb_dense_slab_tf32_apply(int64_t N, int64_t row0, int64_t row1, const float *slab32, const float *slab32_lo,
                              int b, const float *zcurT, const float *zprevT, const float *rhatT, float *outT,
                              const float *cA, const float *cB, double cshift, int fused, void *stream);
/* The same step with the all-gather of the row slabs fused into the epilogue: out_all is a HOST array of `world`
 * (<= 8) device pointers, out_all[p] = the output block of rank p mapped into this process (scb_peer_open; rank's own
 * buffer first or anywhere); rows [row0,row1) are stored into every block over NVLink.  The caller runs a barrier
 * collective before any rank reads its block.  zprevT is this rank's copy of the block that is being overwritten. */
int scb_dense_slab_tf32_apply_allgather(int64_t N, int64_t row0, int64_t row1, const float *slab32,
                                        const float *slab32_lo, int b, const float *zcurT, const float *zprevT,
                                        const float *rhatT, float *const *out_all, int world, const float *cA,
                                        const float *cB, double cshift, int fused, void *stream);
int scb_resform_finish(int64_t N, int b, const double *rn2, const float *zT, double *X, int split, void *stream);
/* Column-wise Lanczos (every column of a block is an independent Lanczos run; spectrum bound of an operator the
 * caller applies, e.g. the dense row-slab operator).  b = 32, 64 or 128.
 *   scb_coldot:        out[B][b] = column-wise dot products of A and Bm ([B][N][b])
 *   scb_lanczos_axpy:  mode 0: W <- W - alpha V - sqrt(beta_prev) Vprev   (beta_prev may be NULL)
 *                      mode 1: Vprev <- V, V <- W / sqrt(nrm2)       mode 2: V <- V / sqrt(nrm2)
 *   scb_lanczos_bound: out[B] = factor * max over the columns of the largest eigenvalue of the steps x steps
 *                      tridiagonal (alpha[steps][B][b], beta2[steps][B][b] = squared couplings) */
int scb_coldot(int B, int64_t N, int b, const double *A, const double *Bm, double *out, void *stream);
int scb_lanczos_axpy(int B, int64_t N, int b, int mode, double *V, double *Vprev, double *W,
                     const double *alpha, const double *beta_prev, const double *nrm2, void *stream);
int scb_lanczos_bound(int B, int b, int steps, const double *alpha, const double *beta2, double factor,
                      double *out, void *stream);
/* deterministic pseudo-random block in (-1, 1) */
int scb_rand_block(int64_t total, uint64_t seed, double *X, void *stream);

/* full symmetric eigendecomposition of dense A[B][N][N] (lower triangle is
 * referenced, like LAPACK dsyevd behind np.linalg.eigh): eigval[B][N] ascending,
 * modes[B][N][N] with ROW k = mode k (nma.py:63).  A is destroyed.
 * N <= 64: Jacobi in shared memory; larger N (up to 9,200): Householder
 * tridiagonalisation (one persistent cooperative kernel) + divide and conquer +
 * compact-WY back-transformation on the FP64 tensor cores; the host reads back one
 * status word at the end (the call returns after the work has completed).
 * Matrices of a batch share every launch. */
size_t scb_eig_full_workspace_bytes(int B, int N);
int scb_eig_full(int B, int N, double *A, double *eigval, double *modes,
                 void *workspace, size_t workspace_bytes, void *stream);
/* the same with an explicit solver: the tridiagonal solver is a cooperative launch
 * that needs every SM of the device; a caller whose other streams hold SMs for an
 * unknown time can ask for SCB_EIG_JACOBI (ordinary launches only) */
#define SCB_EIG_AUTO 0
#define SCB_EIG_JACOBI 1
#define SCB_EIG_TRIDIAG 2
size_t scb_eig_full_workspace_bytes_ex(int solver, int B, int N);
int scb_eig_full_ex(int solver, int B, int N, double *A, double *eigval, double *modes,
                    void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------
 * K4  fluctuation / covariance products (nma.py:108-359, 422-473).
 * modes are ROW-major per mode: modes[B][m][N]; lam[B][m].  `scale` carries
 * tem*tem_factors (and 8*pi^2/3 for B-factors, nma.py:228).
 * ------------------------------------------------------------------------- */
/* msf[B][n] = scale * sum_k (sum_a modes[k][D*i+a]^2) / lam[k]   (nma.py:145-183) */
int scb_msf(int D, int B, int n, int m, const double *lam, const double *modes,
            double scale, double *msf, void *stream);
/* same, reading the eigensolver's column layout X[B][N][b], modes k0..k0+m-1 */
int scb_msf_cols(int D, int B, int n, int b, int k0, int m, const double *eigval,
                 const double *X, double scale, double *msf, void *stream);
/* dcc rows [row0,row1) of sum_k U_k U_k^T / lam_k (FP64 tensor-core DMMA):
 * out[(row1-row0)][n]; norm != 0 divides by sqrt(d_ii d_jj) (nma.py:338-357) */
int scb_dcc(int D, int n, int m, const double *lam, const double *modes, int norm,
            double scale, int row0, int row1, double *out, void *workspace,
            size_t workspace_bytes, void *stream);
size_t scb_dcc_workspace_bytes(int D, int n, int m);
/* covariance rows [row0,row1) = sum_k u_k u_k^T / lam_k   out[(row1-row0)][N]
 * (the pinv of anm.py:132-136 restricted to the given modes) */
int scb_covariance(int N, int m, const double *lam, const double *modes, int row0,
                   int row1, double *out, void *workspace, size_t workspace_bytes,
                   void *stream);
/* linear response dr[N] = sum_k u_k (u_k . f) / lam_k   (nma.py:473) */
int scb_linear_response(int N, int m, const double *lam, const double *modes,
                        const double *force, double *out, void *workspace,
                        size_t workspace_bytes, void *stream);
/* Products of a covariance matrix ASSIGNED by the caller (anm.py:138-148: `enm.covariance = C`), which the reference
 * then uses as is: dcc (nma.py:324-357: 3x3 traces for ANM, optional normalisation by sqrt(d_ii d_jj) BEFORE the
 * temperature scale; diag_scratch[n] needed when norm != 0) and linear response y = C f (nma.py:473). */
int scb_dcc_from_covariance(int D, int n, const double *cov, int norm, double scale, double *out,
                            double *diag_scratch, void *stream);
int scb_symv(int N, const double *cov, const double *f, double *y, void *stream);
/* perturbation response scanning matrix (nma.py:511-531): out[n][n] from the covariance cov[3n][3n];
 * norm != 0 divides row i by out[i][i] */
int scb_prs(int n, const double *cov, int norm, double *out, void *stream);
/* normal-mode trajectory (nma.py:402-419): out[frames][n][3] for one mode[3n]; triangle != 0 selects
 * the "triangle" movement, else "sine" */
int scb_normal_mode(int n, int frames, const double *mode, double amplitude, int triangle, double *out,
                    void *stream);
/* modes[B][m][N] <- X[B][N][b] columns k0..k0+m-1 (transpose to the reference's
 * "eigenvectors as rows" layout) */
int scb_export_modes(int B, int N, int b, int k0, int m, const double *X, double *modes,
                     void *stream);

/* ---------------------------------------------------------------------------
 * Whole-path entry point with HOST buffers (H2D and D2H inside): contacts ->
 * assembly -> lowest-k modes -> MSF for an ensemble of B conformations of one
 * n-node structure.  This is what a reference-side plugin would call in place
 * of   for c in conformations: ANM(c, ff).eigen()/mean_square_fluctuation(...)
 *   coord_host[B][n][3] fp64 (the reference's (n,3) layout per structure)
 *   ff: descriptor whose pointers are DEVICE pointers (tables are uploaded once
 *       by the caller and shared by the batch)
 *   eigval_host[B][k], msf_host[B][n], modes_host[B][k][N] or NULL
 *   k non-trivial modes (indices ntriv..ntriv+k-1 of the reference's ordering)
 * Returns 0 or a negative scb_status; n_pairs_out receives P (may be NULL).
 * B <= 65,535 per call (structures are indexed by blockIdx.y; SCB_ERR_UNSUPPORTED beyond): callers
 * split larger ensembles, as springcraft_b200.enm_ensemble does.  Systems with D*n <= 8*block take the
 * dense full-spectrum solver instead of subspace iteration.
 * ------------------------------------------------------------------------- */
/* same path with DEVICE buffers (xyz SoA [B][3][n]; outputs on the device) */
int scb_enm_ensemble(int D, const double *xyz, int B, int n, const scb_ff_desc *ff,
                     const scb_patch *patch, const double *masses_dev, int k, double tol,
                     double *eigval /*[B][k]*/, double *msf /*[B][n]*/,
                     double *modes /*[B][k][N] or NULL*/, int32_t *iters /*[B] or NULL*/,
                     int64_t *n_pairs_out, void *stream);
/* AoS (n,3) -> SoA [3][n] per structure (struc.coord layout -> kernel layout) */
int scb_coords_to_soa(const double *coord_aos, int B, int n, double *xyz_soa, void *stream);
/* iters_host[B] (may be NULL): outer iterations per structure, NEGATIVE for a structure that did not converge
 * (the call then returns SCB_ERR_NOT_CONVERGED and its rows hold the last iterate). */
int scb_enm_ensemble_host(int D, const double *coord_host, int B, int n,
                          const scb_ff_desc *ff, const scb_patch *patch,
                          const double *masses_dev, int k, double tol,
                          double *eigval_host, double *msf_host, double *modes_host,
                          int32_t *iters_host, int64_t *n_pairs_out, void *stream);
/* The whole-path entry points take their scratch from a library-private memory pool per device that keeps freed
 * blocks cached between calls; this hands the cached blocks back to the driver (host-synchronous). */
int scb_trim_pool(void);

#ifdef __cplusplus
}
#endif
#endif /* SCB200_H */
