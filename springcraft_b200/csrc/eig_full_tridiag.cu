// K3c (64 < N <= 9,200): full symmetric eigendecomposition by reduction to tridiagonal
// form, divide and conquer on the tridiagonal matrix, and back-transformation.
// Replaces LAPACK dsyevd behind np.linalg.eigh (nma.py:61) and np.linalg.pinv
// (anm.py:135) when every mode is requested; the two-sided block Jacobi of
// eig_full_block.cu (145 N^3 flop) stays as the fallback for larger orders.
//
//  1. sytrd_kernel -- Householder tridiagonalisation, ONE persistent
//     cooperative launch.  The rows of the (full, symmetric) matrix are dealt
//     cyclically to the G CTAs of a group; per column j a CTA applies the
//     rank-2 update of reflector j-1 to its rows and multiplies the updated
//     rows with reflector j in the same pass (one read + one write of the
//     trailing matrix per column), writes its slice of p = tau A v and of the
//     next column, and meets the group at ONE flag barrier; every CTA then
//     forms w, the next column and the next reflector redundantly from the
//     exchanged 2 x N doubles.  Rows live in shared memory as far as they fit
//     (re-packed every 128 columns as the trailing matrix shrinks), the rest is
//     served from L2: the reduction never streams the matrix from HBM.  Groups
//     of CTAs work on different matrices of a batch concurrently.
//  2. divide and conquer (stedc_core.cuh): leaves of order <= 32 by Jacobi in
//     shared memory, then per level  prepare (rank sort + deflation scan) ->
//     rotate -> secular roots (warp per root) -> z-hat -> eigenvector matrix
//     of the rank-one problem -> Z_new = U^T Z as FP64 tensor-core GEMMs, two
//     half-width products per merge (columns of U grouped by child); all merges
//     of a level and all matrices of a batch share each launch.
//  3. back-transformation x^T <- x^T (I - V_b T_b^T V_b^T) by compact-WY blocks of 128
//     reflectors, last block first: T_b V_b^T for all blocks in one batched GEMM, then
//     per block a split-K product X (T_b V_b^T)^T, a small reduction and a rank-128 update.
// No host read-back while the solver runs (one status word at the end): every data-dependent size (survivors of a
// deflation, rotations) stays on the device.
#include <stdlib.h>

#include "jacobi.cuh"
#include "stedc_core.cuh"
#include "subspace.cuh"

namespace scb {

// ------------------------------------------------------------------------------------------------------------
// general FP64 tensor-core GEMM: C[M][N] = alpha * op(A) op(B) + beta * C, row-major C
//   AK: A is K-major  A[m*lda + k]   else M-major  A[k*lda + m]
//   BK: B is K-major  B[n*ldb + k]   else N-major  B[k*ldb + n], optionally with gathered rows  B[bidx[k]*ldb + n]
// Leading dimensions and the offsets of the operand origins must be even (16-byte cp.async chunks).
// ------------------------------------------------------------------------------------------------------------
struct GemmTask {
    const double* A;
    const double* B;
    double* C;
    const int32_t* bidx;
    int M, N, K;
    int lda, ldb, ldc;
    int kmodB;   // > 0: the K index of a K-major B wraps modulo kmodB (B repeated along K: split-K partial sums)
    double alpha, beta;
};

namespace {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// 8-byte variant for operands whose origin is only 8-byte aligned (odd column offset)
__device__ __forceinline__ void cp_async8z(void* smem, const void* gmem, int bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(bytes));
}

// copies `bytes` (0, 8 or 16) and zero-fills the rest of the 16-byte chunk
__device__ __forceinline__ void cp_async16z(void* smem, const void* gmem, int bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(bytes));
}

constexpr int kTM = 128, kTN = 128, kTK = 16;
constexpr int kLDK = 20;    // K-major tile: [128][20]
constexpr int kLDM = 136;   // M/N-major tile: [16][136]
constexpr int kTileDoubles = kTM * kLDK;   // 2560 >= 16 * 136
constexpr size_t kGemmSmem = sizeof(double) * 4 * kTileDoubles;

template <bool AK, bool BK>
__global__ void __launch_bounds__(256)
dgemm_dmma_kernel(GemmTask single, const GemmTask* __restrict__ tasks, int64_t strideA, int64_t strideB,
                  int64_t strideC) {
    GemmTask t;
    if (tasks) t = tasks[blockIdx.z];
    else {
        t = single;
        t.A += strideA * blockIdx.z; t.B += strideB * blockIdx.z; t.C += strideC * blockIdx.z;
    }
    const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
    if (m0 >= t.M || n0 >= t.N) return;
    extern __shared__ __align__(16) double gsm[];
    double* sA = gsm;                      // [2][kTileDoubles]
    double* sB = gsm + 2 * kTileDoubles;   // [2][kTileDoubles]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;   // 4 x 2 warps, warp tile 32 x 64
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const bool a_odd = AK && ((reinterpret_cast<uintptr_t>(t.A) & 15) != 0);
    auto load_stage = [&](int stage, int kk) {
        double* a_s = sA + stage * kTileDoubles;
        double* b_s = sB + stage * kTileDoubles;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = tid + q * 256;
            if (AK) {
                const int r = c >> 3, kc = (c & 7) * 2;
                const int kg = kk + kc;
                int bytes = 0;
                if (m0 + r < t.M && kg < t.K) bytes = (kg + 1 < t.K) ? 16 : 8;
                const double* src = bytes ? t.A + (int64_t)(m0 + r) * t.lda + kg : t.A;
                if (a_odd) {   // origin at an odd column: two 8-byte copies
                    cp_async8z(&a_s[r * kLDK + kc], src, bytes ? 8 : 0);
                    cp_async8z(&a_s[r * kLDK + kc + 1], bytes == 16 ? src + 1 : t.A, bytes == 16 ? 8 : 0);
                } else {
                    cp_async16z(&a_s[r * kLDK + kc], src, bytes);
                }
            } else {
                const int kr = c >> 6, mc = (c & 63) * 2;
                int bytes = 0;
                if (kk + kr < t.K && m0 + mc < t.M) bytes = (m0 + mc + 1 < t.M) ? 16 : 8;
                cp_async16z(&a_s[kr * kLDM + mc], bytes ? (const void*)(t.A + (int64_t)(kk + kr) * t.lda + m0 + mc) : (const void*)t.A, bytes);
            }
            if (BK) {
                const int r = c >> 3, kc = (c & 7) * 2;
                const int kg = kk + kc;
                int bytes = 0;
                if (n0 + r < t.N && kg < t.K) bytes = (kg + 1 < t.K) ? 16 : 8;
                const int kb = t.kmodB > 0 ? kg % t.kmodB : kg;
                cp_async16z(&b_s[r * kLDK + kc], bytes ? (const void*)(t.B + (int64_t)(n0 + r) * t.ldb + kb) : (const void*)t.B, bytes);
            } else {
                const int kr = c >> 6, nc = (c & 63) * 2;
                int bytes = 0;
                if (kk + kr < t.K && n0 + nc < t.N) bytes = (n0 + nc + 1 < t.N) ? 16 : 8;
                const double* src = t.B;
                if (bytes) {
                    const int64_t row = t.bidx ? (int64_t)t.bidx[kk + kr] : (int64_t)(kk + kr);
                    src = t.B + row * t.ldb + n0 + nc;
                }
                cp_async16z(&b_s[kr * kLDM + nc], src, bytes);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };

    const int nk = (t.K + kTK - 1) / kTK;
    if (nk > 0) load_stage(0, 0);
    for (int it = 0; it < nk; ++it) {
        const int stage = it & 1;
        if (it + 1 < nk) {
            load_stage(stage ^ 1, (it + 1) * kTK);
            asm volatile("cp.async.wait_group 1;\n" ::);
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
        const double* a_s = sA + stage * kTileDoubles;
        const double* b_s = sB + stage * kTileDoubles;
#pragma unroll
        for (int k4 = 0; k4 < kTK; k4 += 4) {
            double af[4], bf[8];
            const int kq = k4 + (lane & 3);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = wm * 32 + i * 8 + (lane >> 2);
                af[i] = AK ? a_s[r * kLDK + kq] : a_s[kq * kLDM + r];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int cidx = wn * 64 + j * 8 + (lane >> 2);
                bf[j] = BK ? b_s[cidx * kLDK + kq] : b_s[kq * kLDM + cidx];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = m0 + wm * 32 + i * 8 + (lane >> 2);
        if (r >= t.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = n0 + wn * 64 + j * 8 + 2 * (lane & 3);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (c + e >= t.N) continue;
                double* dst = t.C + (int64_t)r * t.ldc + c + e;
                double v = t.alpha * acc[i][j][e];
                if (t.beta != 0.0) v += t.beta * *dst;
                *dst = v;
            }
        }
    }
}

template <bool AK, bool BK>
int launch_gemm(const GemmTask& single, const GemmTask* tasks, int maxM, int maxN, int batch, int64_t sA, int64_t sB,
                int64_t sC, cudaStream_t st) {
    if (maxM < 1 || maxN < 1 || batch < 1) return SCB_OK;
    SCB_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<AK, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    dim3 grid((unsigned)ceil_div(maxN, kTN), (unsigned)ceil_div(maxM, kTM), (unsigned)batch);
    dgemm_dmma_kernel<AK, BK><<<grid, 256, kGemmSmem, st>>>(single, tasks, sA, sB, sC);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

// ------------------------------------------------------------------------------------------------------------
// 1. tridiagonalisation
// ------------------------------------------------------------------------------------------------------------
constexpr int kTrdEpoch = 128;   // columns between two re-packings of the shared-memory row cache

struct TrdParams {
    int N, LD, B, G, ngroups;
    int cache_doubles;   // capacity of the row cache in doubles
    double* A;           // [B][N][LD] full symmetric matrix (destroyed)
    double* Vt;          // [B][N][LD] row j = reflector j (zero up to j, 1 at j+1); zeroed by the caller
    double* tau;         // [B][LD]
    double* d;           // [B][LD]
    double* e;           // [B][LD]
    double* xch;         // [ngroups][kXchCopies][2][2][LD]  (p, next column) of even / odd steps
    unsigned* flags;     // [ngroups][148][kInboxPad] barrier inboxes, zeroed by the caller
    unsigned* abort_flag;   // set by a barrier that gave up waiting (zeroed by the caller)
    long long* dbg;      // optional [grid][4] cycle counters: phase 1, pass, barrier, re-pack (SCB_TRD_DEBUG)
};

constexpr int kInboxPad = 160;   // row pitch of the barrier inboxes (>= 148 CTAs, whole 128-byte lines)
constexpr int kXchCopies = 4;    // replicas of the exchange vectors: 1/4 of the CTAs read each (no hot L2 lines)

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned* p, unsigned v) {
    asm volatile("st.volatile.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// All CTAs of a group.  Push model: thread t of CTA `me` writes the epoch into the inbox of CTA t and then polls
// slot t of its OWN inbox, so every CTA spins on lines nobody else reads (one shared flag line polled by 148 SMs
// saturates its L2 slice and slows every other access of the step).
__device__ __forceinline__ void group_arrive(unsigned* inbox, int me, int G, unsigned epoch) {
    __syncthreads();
    if ((int)threadIdx.x < G) {
        __threadfence();   // the CTA's writes (ordered before this thread by the barrier) become visible first
        st_volatile_u32(inbox + (size_t)threadIdx.x * kInboxPad + me, epoch);
    }
}
__device__ __forceinline__ void group_wait(unsigned* inbox, int me, int G, unsigned epoch, unsigned* abort_flag) {
    if ((int)threadIdx.x < G) {
        const unsigned* mine = inbox + (size_t)me * kInboxPad + threadIdx.x;
        // acquire: the data loads after the CTA barrier below see the writes.  A member that never arrives (it can
        // only be a fault elsewhere) must not hang the device: after ~10 s of spinning the wait is abandoned, every
        // later barrier of the launch falls through and the host reports SCB_ERR_CUDA.
        unsigned spins = 0;
        while (ld_acquire_u32(mine) < epoch) {
            if ((++spins & 0xFFu) == 0) {
                if (ld_volatile_u32(abort_flag) != 0) break;
                if (spins > (1u << 24)) { st_volatile_u32(abort_flag, 1u); break; }
            }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void group_barrier(unsigned* inbox, int me, int G, unsigned epoch, unsigned* abort_flag) {
    group_arrive(inbox, me, G, epoch);
    group_wait(inbox, me, G, epoch, abort_flag);
}

// (Measured dead end: letting the exchanged values carry their own arrival flag -- every slot armed with a NaN
// payload and polled by its consumers, no barrier -- is SLOWER than the barrier below: 148 x 1024 threads spinning
// on shared lines delay the producers' stores; C2 65 ms instead of 51 ms.)
// Also measured and dropped: four rows per pass task sharing the vector loads (C2 38.9 against 38.5 ms, 64 x N=900
// 65.9 against 61.5 ms), fence.acq_rel + relaxed polling instead of __threadfence() + acquire polling (C2 41.9
// against 40.4 ms), two co-resident 256-thread CTAs per SM (no gain), and one 16-CTA thread-block cluster
// per matrix with the exchange pushed through DSMEM and the hardware cluster barrier (64 x N=900: 98 ms against
// 89 ms; 256 x N=300: 69 ms against 35 ms -- the per-column cost is the CTA's own chain of reductions).
constexpr int kXchBufs = 2;   // even / odd steps

// sum over the CTA with ONE barrier: `red` must not be reused before another barrier (callers alternate two arrays)
template <int NT>
__device__ __forceinline__ double block_sum_nt1(double v, double* red) {
    v = warp_sum(v);
    if (lane_id() == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    return warp_sum(lane_id() < NT / 32 ? red[lane_id()] : 0.0);
}

constexpr int kTrdMaxRows = 128;  // rows of one CTA (make_plan keeps ceil(N / G) below this)
constexpr int kTrdMaxSeg = 8;     // column segments a row pair is split into when the CTA has few rows

template <int kTrdThreads>
__global__ void __launch_bounds__(kTrdThreads, 1) sytrd_kernel(TrdParams P) {
    extern __shared__ __align__(16) double tsm[];
    __shared__ double redA[32], redB[32], bc[4];
    __shared__ double partial[kTrdMaxSeg][kTrdMaxRows];
    __shared__ double pcol[kTrdMaxRows];   // column j+1 of the CTA's rows (for p = tau scal (A a - beta A[:, j+1]))
    const int N = P.N, LD = P.LD, G = P.G;
    const int grp = blockIdx.x / G, c = blockIdx.x % G;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double* vprev = tsm;
    double* wprev = tsm + LD;
    double* vcur = tsm + 2 * LD;
    double* cache = tsm + 3 * LD;
    unsigned* flags = P.flags + (size_t)grp * kNumSM * kInboxPad;
    const size_t xcopy = (size_t)kXchBufs * 2 * LD;                 // one replica: [step % 2][p | column][LD]
    double* xch_all = P.xch + (size_t)grp * kXchCopies * xcopy;
    const int ncopies = G > 80 ? kXchCopies : (G > 40 ? 2 : 1);      // ~37 readers per replica
    const double* xch = xch_all + (size_t)(c % ncopies) * xcopy;     // the replica this CTA reads
    unsigned epoch = 0;
#ifdef SCB_TRD_TIMING   // per-phase cycle counters (they cost registers: development builds only)
    long long t_ph1 = 0, t_pass = 0, t_bar = 0, t_pack = 0, t0 = clock64(), t1;
#define TRD_LAP(acc) do { t1 = clock64(); acc += t1 - t0; t0 = t1; } while (0)
#else
#define TRD_LAP(acc) do { } while (0)
#endif
    constexpr int UNR = (kTrdThreads <= 512) ? 4 : 2;     // chunks in flight per lane (register budget)
    const int nown = (c < N) ? (N - c + G - 1) / G : 0;   // rows c, c+G, ...

    for (int s = grp; s < P.B; s += P.ngroups) {
        double* A = P.A + (size_t)s * N * LD;
        double* Vt = P.Vt + (size_t)s * N * LD;
        double* tau = P.tau + (size_t)s * LD;
        double* dd = P.d + (size_t)s * LD;
        double* ee = P.e + (size_t)s * LD;
        int cache_col0 = 0, cache_len = LD, cache_rows = 0;
        double tau_prev = 0.0;
        // tracked incrementally (no division by G in the column loop): owner = j % G, rmin_j = first owned row > j
        int owner = 0, rmin_j = (c == 0) ? 1 : 0, seg_np = -1, seg_nc = -1, seg_n = 1;
        for (int i = tid; i < 3 * LD; i += kTrdThreads) tsm[i] = 0.0;   // the padding [N, LD) of the vectors stays zero
        __syncthreads();

        for (int j = 0; j < N; ++j) {
            // ---- re-pack the row cache: rows are only needed from column j+1 on
            if ((j % kTrdEpoch) == 0 && j + 1 < N) {
                for (int q = warp; q < cache_rows; q += kTrdThreads / 32) {   // write back
                    const int r = nown - 1 - q;
                    double* dst = A + (size_t)(c + G * r) * LD;
                    const double* src = cache + (size_t)q * cache_len - cache_col0;
                    for (int cc = cache_col0 + lane; cc < LD; cc += 32) dst[cc] = src[cc];
                }
                __syncthreads();
                cache_col0 = (j + 1) & ~1;
                cache_len = LD - cache_col0;
                const int rmin = (j + 1 > c) ? (j + 1 - c + G - 1) / G : 0;   // first owned row with index > j
                int alive = nown - rmin;
                if (alive < 0) alive = 0;
                cache_rows = P.cache_doubles / cache_len;
                if (cache_rows > alive) cache_rows = alive;
                for (int q = warp; q < cache_rows; q += kTrdThreads / 32) {
                    const int r = nown - 1 - q;
                    const double* src = A + (size_t)(c + G * r) * LD;
                    double* dst = cache + (size_t)q * cache_len - cache_col0;
                    for (int cc = cache_col0 + lane; cc < LD; cc += 32) dst[cc] = src[cc];
                }
                __syncthreads();
            }
            TRD_LAP(t_pack);
            // ---- phase 1 (every CTA, redundantly): w_{j-1}, d_j, column j, reflector j.  Thread t owns the
            //      components j + t, j + t + 1024, ... through all steps, so only the reductions need barriers
            double dj = 0.0;
            {
                double acc = 0.0;
                if (j == 0) {
                    for (int i = tid; i < N; i += kTrdThreads) vcur[i] = A[(size_t)i * LD];   // column 0
                } else {
                    const double* pbuf = xch + (size_t)((j - 1) % kXchBufs) * 2 * LD;
                    const double* cbuf = pbuf + LD;
                    // all L2 loads of a thread are issued before the first use (one round trip, not one per element)
                    constexpr int EPT = 4;
                    for (int base = j + tid; base < N; base += kTrdThreads * EPT) {
                        double pr[EPT], cr[EPT];
#pragma unroll
                        for (int u = 0; u < EPT; ++u) {
                            const int i = base + u * kTrdThreads;
                            if (i < N) { pr[u] = __ldcg(pbuf + i); cr[u] = __ldcg(cbuf + i); }
                        }
#pragma unroll
                        for (int u = 0; u < EPT; ++u) {
                            const int i = base + u * kTrdThreads;
                            if (i < N) {
                                wprev[i] = pr[u];
                                vcur[i] = cr[u];
                                acc = fma(pr[u], vprev[i], acc);
                            }
                        }
                    }
                    if (tid == 0) bc[0] = wprev[j];   // p_j
                }
                const double dot = block_sum_nt1<kTrdThreads>(acc, redA);
                const double kappa = 0.5 * tau_prev * dot;
                const double wj = (j > 0) ? bc[0] - kappa : 0.0;   // vprev[j] == 1
                double sig = 0.0;
                if (j > 0) {
                    for (int i = j + tid; i < N; i += kTrdThreads) {
                        const double v = vprev[i];
                        const double w = wprev[i] - kappa * v;
                        wprev[i] = w;
                        const double a = vcur[i] - (v * wj + w);   // column j of A_j (for i == j: d_j)
                        if (i == j) { bc[1] = a; vcur[i] = 0.0; }
                        else {
                            vcur[i] = a;
                            if (i == j + 1) bc[2] = a; else sig = fma(a, a, sig);
                        }
                    }
                } else {
                    for (int i = tid; i < N; i += kTrdThreads) {
                        const double a = vcur[i];
                        if (i == 0) { bc[1] = a; vcur[i] = 0.0; }
                        else if (i == 1) bc[2] = a;
                        else sig = fma(a, a, sig);
                    }
                }
                const double sigma = block_sum_nt1<kTrdThreads>(sig, redB);
                dj = bc[1];
                // The pass below runs on the RAW column a (vcur): v = scal (a - beta e_{j+1}) gives
                //   A v = scal (A a - beta A[:, j+1]),
                // so beta, tau and the scaling (a square root and two divisions) leave the critical path, and v is
                // normalised in the shadow of the group barrier.
                const double x0 = bc[2];
                if (j == N - 1) {
                    if (owner == c && tid == 0) dd[j] = dj;
                    break;
                }
                TRD_LAP(t_ph1);
                // ---- pass j: rows i > j of this CTA: apply update j-1, multiply with reflector j.
                //      Task = (pair of rows, column segment); a lane handles two adjacent columns of both rows.
                double* pout = xch_all + (size_t)(j % kXchBufs) * 2 * LD;   // replica 0; replica k at + k * xcopy
                double* cout = pout + LD;
                const int rmin = rmin_j;
                const int na = nown - rmin;                   // rows alive
                const int npairs = (na + 1) >> 1;
                const int cb = (j + 1) & ~1;                  // first column pair (if it starts at column j: vcur[j] == 0)
                const int nchunk = (LD - cb + 63) >> 6;       // chunks of 64 columns
                // column segments per row pair: minimise (rounds of the warps) x (segment length); a segment keeps
                // at least one full set of chunks in flight (11 pairs on 16 warps: 4 segments = 3 rounds of 1/4)
                if (npairs != seg_np || nchunk != seg_nc) {   // changes every ~32 columns only
                    constexpr int nw = kTrdThreads / 32;
                    int best = 1 << 30;
                    seg_np = npairs; seg_nc = nchunk; seg_n = 1;
#pragma unroll   // divisions by compile-time constants
                    for (int sgc = 1; sgc <= kTrdMaxSeg; ++sgc) {
                        if (npairs <= 0) break;
                        const int len = (nchunk + sgc - 1) / sgc;
                        if (sgc > 1 && len < UNR) break;
                        // + 4: set-up and the two warp reductions of a task cost about eight loop iterations
                        const int cost = ((npairs * sgc + nw - 1) / nw) * ((len + UNR - 1) / UNR + 8);
                        if (cost < best) { best = cost; seg_n = sgc; }
                    }
                }
                const int nseg = seg_n;
                const int cps = (nchunk + nseg - 1) / nseg;
                for (int task = warp; task < npairs * nseg; task += kTrdThreads / 32) {
                    const int pr = task / nseg, sg = task - pr * nseg;
                    const int r0 = rmin + 2 * pr;
                    const bool two = (r0 + 1 < nown);
                    const int r1 = two ? r0 + 1 : r0;
                    const int i0 = c + G * r0, i1 = c + G * r1;
                    const int q0 = nown - 1 - r0, q1 = nown - 1 - r1;
                    double* row0 = (q0 < cache_rows) ? cache + (size_t)q0 * cache_len - cache_col0 : A + (size_t)i0 * LD;
                    double* row1 = (q1 < cache_rows) ? cache + (size_t)q1 * cache_len - cache_col0 : A + (size_t)i1 * LD;
                    const double v0 = vprev[i0], w0 = wprev[i0], v1 = vprev[i1], w1 = wprev[i1];
                    double acc0 = 0.0, acc1 = 0.0;
                    const int ch_end = min(nchunk, (sg + 1) * cps);
                    for (int ch = sg * cps; ch < ch_end; ch += UNR) {
                        double2 a0[UNR], a1[UNR];
#pragma unroll
                        for (int u = 0; u < UNR; ++u) {
                            const int cc = cb + ((ch + u) << 6) + 2 * lane;
                            if (ch + u < ch_end && cc < LD) {
                                a0[u] = *reinterpret_cast<const double2*>(row0 + cc);
                                a1[u] = *reinterpret_cast<const double2*>(row1 + cc);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < UNR; ++u) {
                            const int cc = cb + ((ch + u) << 6) + 2 * lane;
                            if (ch + u < ch_end && cc < LD) {
                                const double2 vp = *reinterpret_cast<const double2*>(vprev + cc);
                                const double2 wp = *reinterpret_cast<const double2*>(wprev + cc);
                                const double2 vc = *reinterpret_cast<const double2*>(vcur + cc);
                                double2 x = a0[u];
                                x.x -= (v0 * wp.x + w0 * vp.x);
                                x.y -= (v0 * wp.y + w0 * vp.y);
                                *reinterpret_cast<double2*>(row0 + cc) = x;
                                acc0 = fma(x.x, vc.x, acc0);
                                acc0 = fma(x.y, vc.y, acc0);
                                if (cc == cb && sg == 0) {                                 // column j+1
                                    const double cv = (cb == j) ? x.y : x.x;
                                    pcol[2 * pr] = cv;
                                    for (int k = 0; k < ncopies; ++k) __stcg(cout + k * xcopy + i0, cv);
                                }
                                if (two) {
                                    double2 y = a1[u];
                                    y.x -= (v1 * wp.x + w1 * vp.x);
                                    y.y -= (v1 * wp.y + w1 * vp.y);
                                    *reinterpret_cast<double2*>(row1 + cc) = y;
                                    acc1 = fma(y.x, vc.x, acc1);
                                    acc1 = fma(y.y, vc.y, acc1);
                                    if (cc == cb && sg == 0) {
                                        const double cv = (cb == j) ? y.y : y.x;
                                        pcol[2 * pr + 1] = cv;
                                        for (int k = 0; k < ncopies; ++k) __stcg(cout + k * xcopy + i1, cv);
                                    }
                                }
                            }
                        }
                    }
                    acc0 = warp_sum(acc0);
                    acc1 = warp_sum(acc1);
                    if (lane == 0) {
                        partial[sg][2 * pr] = acc0;
                        if (two) partial[sg][2 * pr + 1] = acc1;
                    }
                }
                double tau_cur = 0.0, beta = x0, scal = 0.0;
                if (sigma > 0.0) {
                    const double nrm = sqrt(fma(x0, x0, sigma));
                    beta = -copysign(nrm, x0);
                    tau_cur = (beta - x0) / beta;
                    scal = 1.0 / (x0 - beta);
                }
                __syncthreads();
                if (tid < na) {
                    double acc = 0.0;
                    for (int sg = 0; sg < nseg; ++sg) acc += partial[sg][tid];
                    const int i = c + G * (rmin + tid);
                    const double pv = tau_cur * scal * (acc - beta * pcol[tid]);
                    for (int k = 0; k < ncopies; ++k) __stcg(pout + k * xcopy + i, pv);
                }
                TRD_LAP(t_pass);
                ++epoch;
                group_arrive(flags, c, G, epoch);
                // in the shadow of the barrier: v_j = scal (a - beta e_{j+1}) (v_j[j+1] = 1), the reflector's row of V^T
                {
                    const bool mine = (owner == c);
                    double* vrow = Vt + (size_t)j * LD;
                    for (int i = j + 1 + tid; i < N; i += kTrdThreads) {
                        const double v = (i == j + 1) ? 1.0 : vcur[i] * scal;
                        vcur[i] = v;
                        if (mine) vrow[i] = v;
                    }
                    if (mine && tid == 0) { dd[j] = dj; ee[j] = beta; tau[j] = tau_cur; }
                }
                group_wait(flags, c, G, epoch, P.abort_flag);
                TRD_LAP(t_bar);
                double* tmp = vprev; vprev = vcur; vcur = tmp;
                tau_prev = tau_cur;
                if (++owner == G) owner = 0;
                if (c + G * rmin_j <= j + 1) ++rmin_j;   // row c + G rmin_j is dead from column j+1 on
            }
        }
        ++epoch;
        group_barrier(flags, c, G, epoch, P.abort_flag);   // the exchange buffers are reused by the next matrix of this group
    }
#ifdef SCB_TRD_TIMING
    if (P.dbg && tid == 0) {
        P.dbg[blockIdx.x * 4 + 0] = t_ph1; P.dbg[blockIdx.x * 4 + 1] = t_pass;
        P.dbg[blockIdx.x * 4 + 2] = t_bar; P.dbg[blockIdx.x * 4 + 3] = t_pack;
    }
#endif
#undef TRD_LAP
}

// Ap[s][i][j] = symmetric extension of the lower triangle of A[s] (N x N), padded to LD columns with zeros
__global__ void trd_init_kernel(int N, int LD, const double* __restrict__ A, double* __restrict__ Ap) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (int64_t)N * LD) return;
    const int i = (int)(q / LD), j = (int)(q % LD);
    const double* As = A + (int64_t)blockIdx.y * N * N;
    double v = 0.0;
    if (j < N) v = (j <= i) ? As[(int64_t)i * N + j] : As[(int64_t)j * N + i];
    Ap[(int64_t)blockIdx.y * N * LD + q] = v;
}

// ------------------------------------------------------------------------------------------------------------
// 2. divide and conquer on the tridiagonal matrix
// ------------------------------------------------------------------------------------------------------------
constexpr int kLeaf = 32;

// node t of `level` (root = level 0) covers [lo, hi); all split points are even (16-byte aligned operand origins)
__host__ __device__ inline void node_range(int N, int level, int t, int* lo, int* hi) {
    int a = 0, b = N;
    for (int bit = level - 1; bit >= 0; --bit) {
        const int mid = a + (((b - a) / 2) & ~1);
        if ((t >> bit) & 1) a = mid; else b = mid;
    }
    *lo = a;
    *hi = b;
}

struct DcWork {           // per matrix (index s) unless noted; vectors have LD entries
    int N, LD, L;
    double *d, *e;        // tridiagonal (d is torn in place, e replaced by |e|)
    double* sgn;          // row signs that make all off-diagonals non-negative
    double *D0, *D1;      // eigenvalues of the current / next level, per sub-problem range
    double *Z0, *Z1;      // [N][LD] eigenvectors as rows, block diagonal over the sub-problems
    double* U;            // [N][LD] differences d_i - lambda_j, then the eigenvectors of the rank-one problems
    double *dsc, *dl, *w, *zh;
    int32_t *nd, *dfl;    // survivors / deflated entries (local indices) per range
    int32_t *pos, *ndg;   // column of survivor t in the grouped eigenvector matrix; rows gathered in grouped order
    stedc::Rotation* rot;
    int32_t *kc, *nr;     // [B][nodes of the level] survivors and rotations of each merge
    GemmTask* tasks;      // [B * nodes]
    int64_t vstride, mstride;   // strides between matrices of the batch for vectors / matrices
};

// |e|, signs, tearing of every boundary of every level
__global__ void __launch_bounds__(1024) dc_setup_kernel(DcWork W) {
    const int s = blockIdx.x, N = W.N;
    double* d = W.d + s * W.vstride;
    double* e = W.e + s * W.vstride;
    double* sgn = W.sgn + s * W.vstride;
    // sgn[i] = (-1)^(number of negative off-diagonals before i)
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        int neg = 0;
        for (int q = 0; q < i; ++q) neg += (e[q] < 0.0) ? 1 : 0;
        sgn[i] = (neg & 1) ? -1.0 : 1.0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i + 1 < N; i += blockDim.x) e[i] = fabs(e[i]);
    __syncthreads();
    for (int level = 1; level <= W.L; ++level)
        for (int t = 1 + 2 * threadIdx.x; t < (1 << level); t += 2 * blockDim.x) {
            int lo, hi;
            node_range(N, level, t, &lo, &hi);
            const double b = e[lo - 1];
            d[lo - 1] -= b;
            d[lo] -= b;
        }
}

__global__ void __launch_bounds__(256) dc_leaf_kernel(DcWork W) {
    constexpr int LDS = kLeaf + 1;
    constexpr int NT = 256;
    extern __shared__ __align__(16) double lsm[];
    double* S = lsm;
    double* V = lsm + kLeaf * LDS;
    __shared__ double cs[kLeaf], sn[kLeaf], red[NT / 32];
    __shared__ int pp[kLeaf], qq[kLeaf];
    const int s = blockIdx.y, N = W.N, LD = W.LD;
    int lo, hi;
    node_range(N, W.L, blockIdx.x, &lo, &hi);
    const int n = hi - lo;
    const double* d = W.d + s * W.vstride;
    const double* e = W.e + s * W.vstride;
    for (int q = threadIdx.x; q < kLeaf * kLeaf; q += NT) {
        const int r = q / kLeaf, cidx = q % kLeaf;
        double v = 0.0;
        if (r < n && cidx < n) {
            if (r == cidx) v = d[lo + r];
            else if (r == cidx + 1) v = e[lo + cidx];
            else if (cidx == r + 1) v = e[lo + r];
        }
        S[r * LDS + cidx] = v;
        V[r * LDS + cidx] = (r == cidx) ? 1.0 : 0.0;
    }
    __syncthreads();
    jacobi_eigen_smem<kLeaf, LDS, NT>(S, V, cs, sn, pp, qq, red, 60, false);
    double* D = W.D0 + s * W.vstride;
    double* Z = W.Z0 + s * W.mstride;
    for (int q = threadIdx.x; q < n * n; q += NT) {
        const int m = q / n, i = q % n;
        Z[(int64_t)(lo + m) * LD + lo + i] = V[i * LDS + m];
    }
    for (int m = threadIdx.x; m < n; m += NT) D[lo + m] = S[m * LDS + m];
}

// one CTA per merge: z, rank sort, deflation scan, GEMM task
__global__ void __launch_bounds__(1024)
dc_prepare_kernel(DcWork W, int level, const double* __restrict__ Dcur, const double* __restrict__ Zcur,
                  double* __restrict__ Znew) {
    extern __shared__ __align__(16) double psm[];
    const int s = blockIdx.y, t = blockIdx.x, N = W.N, LD = W.LD;
    const int nodes = 1 << level;
    int lo, hi, l1, mid;
    node_range(N, level, t, &lo, &hi);
    node_range(N, level + 1, 2 * t, &l1, &mid);
    const int n = hi - lo;
    double* dd = psm;                      // [n]
    double* z = psm + n;                   // [n]
    int* order = reinterpret_cast<int*>(psm + 2 * n);   // [n]
    unsigned char* mixed = reinterpret_cast<unsigned char*>(order + n);   // [n]
    const int n1 = mid - lo;
    const double* D = Dcur + s * W.vstride;
    const double* Z = Zcur + s * W.mstride;
    const double rho = 2.0 * W.e[s * W.vstride + mid - 1];
    for (int l = threadIdx.x; l < n; l += blockDim.x) {
        dd[l] = D[lo + l];
        z[l] = Z[(int64_t)(lo + l) * LD + (lo + l < mid ? mid - 1 : mid)] * 0.70710678118654752440;
        mixed[l] = 0;
    }
    __syncthreads();
    for (int l = threadIdx.x; l < n; l += blockDim.x) {
        const double v = dd[l];
        int r = 0;
        for (int q = 0; q < n; ++q) {
            const double u = dd[q];
            r += (u < v) || (u == v && q < l);
        }
        order[r] = l;
    }
    __syncthreads();
    int32_t* nd = W.nd + s * W.vstride + lo;
    int32_t* dfl = W.dfl + s * W.vstride + lo;
    stedc::Rotation* rot = W.rot + s * W.vstride + lo;
    __shared__ int sh_k, sh_cnt[3];
    int32_t* pos = W.pos + s * W.vstride + lo;
    int32_t* ndg = W.ndg + s * W.vstride + lo;
    if (threadIdx.x == 0) {
        int nrot = 0;
        const int k = stedc::deflation_scan(n, order, dd, z, rho, nd, dfl, rot, &nrot, n1, mixed);
        sh_k = k;
        W.kc[s * nodes + t] = k;
        W.nr[s * nodes + t] = nrot;
    }
    __syncthreads();
    // The eigenvector rows of the two children live in disjoint column ranges (except the few rows mixed by a
    // rotation across the children), so the columns of U are grouped [first child | mixed | second child] and the
    // merge is two products of half the width and about half the depth.  Category of survivor q: 0 / 1 / 2.
    {
        const int k = sh_k;
        int* cat = order;   // the sort order is dead: reuse it for the categories
        for (int q = threadIdx.x; q < k; q += blockDim.x) {
            const int l = nd[q];
            cat[q] = mixed[l] ? 1 : (l < n1 ? 0 : 2);
        }
        __syncthreads();
        if (threadIdx.x < 3) {   // category sizes
            int cnt = 0;
            for (int q = 0; q < k; ++q) cnt += (cat[q] == (int)threadIdx.x);
            sh_cnt[threadIdx.x] = cnt;
        }
        __syncthreads();
        const int off[3] = {0, sh_cnt[0], sh_cnt[0] + sh_cnt[1]};
        for (int q = threadIdx.x; q < k; q += blockDim.x) {
            const int cq = cat[q];
            int r = 0;
            for (int u = 0; u < q; ++u) r += (cat[u] == cq);
            const int pq = off[cq] + r;
            pos[q] = pq;
            ndg[pq] = nd[q];
        }
    }
    if (threadIdx.x == 0) {
        const int k = sh_k, k1 = sh_cnt[0], km = sh_cnt[1];
        GemmTask g;
        g.lda = LD; g.ldb = LD; g.ldc = LD;
        g.kmodB = 0;
        g.alpha = 1.0; g.beta = 0.0;
        g.M = k;
        const double* Ub = W.U + s * W.mstride + (int64_t)lo * LD + lo;
        const double* Zb = Z + (int64_t)lo * LD + lo;
        double* Cb = Znew + s * W.mstride + (int64_t)lo * LD + lo;
        g.A = Ub; g.B = Zb; g.C = Cb; g.bidx = ndg; g.N = n1; g.K = k1 + km;
        W.tasks[(s * nodes + t) * 2] = g;
        g.A = Ub + k1; g.B = Zb + n1; g.C = Cb + n1; g.bidx = ndg + k1; g.N = n - n1; g.K = k - k1;
        W.tasks[(s * nodes + t) * 2 + 1] = g;
    }
    __syncthreads();
    const int k = sh_k;
    double* dl = W.dl + s * W.vstride + lo;
    double* w = W.w + s * W.vstride + lo;
    double* dsc = W.dsc + s * W.vstride + lo;
    for (int l = threadIdx.x; l < n; l += blockDim.x) {
        dsc[l] = dd[l];
        if (l < k) {
            const int src = nd[l];
            dl[l] = dd[src];
            w[l] = z[src];
        }
    }
}

// rotations of the deflation scan, applied to the eigenvector rows: one thread per component
__global__ void __launch_bounds__(256) dc_rotate_kernel(DcWork W, int level, double* __restrict__ Zcur) {
    const int s = blockIdx.z, t = blockIdx.y, LD = W.LD;
    const int nodes = 1 << level;
    const int nrot = W.nr[s * nodes + t];
    if (nrot == 0) return;
    int lo, hi;
    node_range(W.N, level, t, &lo, &hi);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi - lo) return;
    const stedc::Rotation* rot = W.rot + s * W.vstride + lo;
    double* Z = Zcur + s * W.mstride + (int64_t)lo * LD + lo + i;
    for (int r = 0; r < nrot; ++r) {
        const stedc::Rotation R = rot[r];
        const double a = Z[(int64_t)R.p * LD], b = Z[(int64_t)R.q * LD];
        Z[(int64_t)R.p * LD] = R.c * a + R.s * b;
        Z[(int64_t)R.q * LD] = R.c * b - R.s * a;
    }
}

struct WarpLanes {
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31u); }
    __device__ __forceinline__ int lanes() const { return 32; }
    __device__ __forceinline__ double sum(double v) const { return warp_sum(v); }
};

// one warp per root of the secular equation
__global__ void __launch_bounds__(256) dc_secular_kernel(DcWork W, int level, double* __restrict__ Dnew) {
    const int s = blockIdx.z, t = blockIdx.y, LD = W.LD;
    const int nodes = 1 << level;
    const int k = W.kc[s * nodes + t];
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= k) return;
    int lo, hi, l1, mid;
    node_range(W.N, level, t, &lo, &hi);
    node_range(W.N, level + 1, 2 * t, &l1, &mid);
    const double rho = 2.0 * W.e[s * W.vstride + mid - 1];
    const double* dl = W.dl + s * W.vstride + lo;
    const double* w = W.w + s * W.vstride + lo;
    double* delta = W.U + s * W.mstride + (int64_t)(lo + j) * LD + lo;
    WarpLanes cx;
    const double lam = stedc::secular_root(cx, k, j, dl, w, rho, delta, W.pos + s * W.vstride + lo);
    if ((threadIdx.x & 31) == 0) Dnew[s * W.vstride + lo + j] = lam;
}

// z-hat_i = sign(w_i) sqrt( prod_j (dl_i - lam_j) / prod_{j != i} (dl_i - dl_j) ): 32 components x 32 slices of j
__global__ void __launch_bounds__(1024) dc_zhat_kernel(DcWork W, int level) {
    __shared__ double part[32][33];
    const int s = blockIdx.z, t = blockIdx.y, LD = W.LD;
    const int nodes = 1 << level;
    const int k = W.kc[s * nodes + t];
    if ((int)blockIdx.x * 32 >= k) return;
    int lo, hi;
    node_range(W.N, level, t, &lo, &hi);
    const int ix = threadIdx.x & 31, jy = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + ix;
    const double* dl = W.dl + s * W.vstride + lo;
    const double* U = W.U + s * W.mstride + (int64_t)lo * LD + lo;
    double p = 1.0;
    if (i < k) {
        const double di = dl[i];
        const int pi = W.pos[s * W.vstride + lo + i];   // column of entry i in the grouped matrix
        for (int j = jy; j < k; j += 32) {
            const double num = U[(int64_t)j * LD + pi];
            p *= (j == i) ? num : num / (di - dl[j]);
        }
    }
    part[jy][ix] = p;
    __syncthreads();
    if (jy == 0 && i < k) {
        double q = 1.0;
        for (int y = 0; y < 32; ++y) q *= part[y][ix];
        W.zh[s * W.vstride + lo + i] = copysign(sqrt(fabs(q)), W.w[s * W.vstride + lo + i]);
    }
}

// U[j][i] = zh_i / (dl_i - lam_j), normalised: one warp per root
__global__ void __launch_bounds__(256) dc_vectors_kernel(DcWork W, int level) {
    const int s = blockIdx.z, t = blockIdx.y, LD = W.LD;
    const int nodes = 1 << level;
    const int k = W.kc[s * nodes + t];
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= k) return;
    int lo, hi;
    node_range(W.N, level, t, &lo, &hi);
    const int lane = threadIdx.x & 31;
    const double* zh = W.zh + s * W.vstride + lo;
    const int32_t* pos = W.pos + s * W.vstride + lo;
    double* row = W.U + s * W.mstride + (int64_t)(lo + j) * LD + lo;
    double acc = 0.0;
    for (int i = lane; i < k; i += 32) {
        const int pi = pos[i];
        const double u = zh[i] / row[pi];
        row[pi] = u;
        acc += u * u;
    }
    acc = warp_sum(acc);
    const double sc = 1.0 / sqrt(acc);
    for (int i = lane; i < k; i += 32) row[i] *= sc;
}

// deflated eigenpairs are copied behind the k new ones
__global__ void __launch_bounds__(128)
dc_copy_deflated_kernel(DcWork W, int level, const double* __restrict__ Zcur, double* __restrict__ Znew,
                        double* __restrict__ Dnew) {
    const int s = blockIdx.z, t = blockIdx.y, LD = W.LD;
    const int nodes = 1 << level;
    const int k = W.kc[s * nodes + t];
    int lo, hi;
    node_range(W.N, level, t, &lo, &hi);
    const int n = hi - lo;
    const int q = blockIdx.x;
    if (q >= n - k) return;
    const int src = W.dfl[s * W.vstride + lo + q];
    const double* from = Zcur + s * W.mstride + (int64_t)(lo + src) * LD + lo;
    double* to = Znew + s * W.mstride + (int64_t)(lo + k + q) * LD + lo;
    for (int i = threadIdx.x; i < n; i += blockDim.x) to[i] = from[i];
    if (threadIdx.x == 0) Dnew[s * W.vstride + lo + k + q] = W.dsc[s * W.vstride + lo + src];
}

// ascending rank of every eigenvalue (ties by index); eigval[rank] = D
__global__ void __launch_bounds__(256)
dc_rank_kernel(int N, int64_t vstride, const double* __restrict__ D, int32_t* __restrict__ rank,
               double* __restrict__ eigval) {
    const int s = blockIdx.y;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= N) return;
    const double* Ds = D + s * vstride;
    const double v = Ds[m];
    int r = 0;
    for (int q = 0; q < N; ++q) {
        const double u = Ds[q];
        r += (u < v) || (u == v && q < m);
    }
    rank[s * vstride + m] = r;
    eigval[(int64_t)s * N + r] = v;
}

// X[rank[m]][i] = sgn[i] * Z[m][i]
__global__ void __launch_bounds__(256)
dc_permute_kernel(int N, int LD, int64_t vstride, int64_t mstride, const double* __restrict__ Z,
                  const int32_t* __restrict__ rank, const double* __restrict__ sgn, double* __restrict__ X) {
    const int s = blockIdx.z, m = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= LD) return;
    const int r = rank[s * vstride + m];
    X[s * mstride + (int64_t)r * LD + i] = (i < N) ? sgn[s * vstride + i] * Z[s * mstride + (int64_t)m * LD + i] : 0.0;
}

// ------------------------------------------------------------------------------------------------------------
// 3. back-transformation
// ------------------------------------------------------------------------------------------------------------
constexpr int kWY = 128;   // reflectors per compact-WY block

// T of the block  H_j0 ... H_j0+nb-1 = I - V T V^T  (forward, columnwise): T[a][a] = tau_a,
// T[0:a, a] = -tau_a T[0:a, 0:a] (V^T v_a)[0:a].  One CTA per block; Gm = V^T V of the block.
__global__ void __launch_bounds__(256)
wy_tfactor_kernel(int N, int64_t vstride, int nblk, int S, const double* __restrict__ tau, double* __restrict__ Gm,
                  double* __restrict__ T) {
    extern __shared__ __align__(16) double wsm[];   // T [kWY][kWY+1]
    constexpr int LDT = kWY + 1;
    __shared__ double col[kWY];
    const int s = blockIdx.y, b = blockIdx.x;
    const int j0 = b * kWY;
    const int nb = min(kWY, (N - 1) - j0);   // reflectors j0 .. j0+nb-1 (there are N-1, the last has tau = 0)
    double* G = Gm + ((int64_t)s * nblk + b) * S * kWY * kWY;   // S split-K partials, summed into the first
    for (int q = threadIdx.x; q < kWY * kWY; q += blockDim.x) {
        double acc = G[q];
        for (int sp = 1; sp < S; ++sp) acc += G[(int64_t)sp * kWY * kWY + q];
        G[q] = acc;
    }
    double* Tg = T + ((int64_t)s * nblk + b) * kWY * kWY;
    const double* tv = tau + s * vstride + j0;
    for (int q = threadIdx.x; q < kWY * LDT; q += blockDim.x) wsm[q] = 0.0;
    __syncthreads();
    // G is symmetric: column a is read as ROW a (contiguous), staged in shared memory one step ahead
    __shared__ double grow[2][kWY];
    if (threadIdx.x < kWY) grow[0][threadIdx.x] = G[threadIdx.x];
    __syncthreads();
    for (int a = 0; a < nb; ++a) {
        const double ta = tv[a];
        const double* ga = grow[a & 1];
        double nxt = 0.0;
        if (a + 1 < nb && threadIdx.x < kWY) nxt = G[(int64_t)(a + 1) * kWY + threadIdx.x];
        // col[r] = -ta * sum_{c=r}^{a-1} T[r][c] G[c][a]   (T upper triangular)
        for (int r = threadIdx.x; r < a; r += blockDim.x) {
            double acc = 0.0;
            for (int cidx = r; cidx < a; ++cidx) acc = fma(wsm[r * LDT + cidx], ga[cidx], acc);
            col[r] = -ta * acc;
        }
        __syncthreads();
        for (int r = threadIdx.x; r < a; r += blockDim.x) wsm[r * LDT + a] = col[r];
        if (threadIdx.x == 0) wsm[a * LDT + a] = ta;
        if (threadIdx.x < kWY) grow[(a + 1) & 1][threadIdx.x] = nxt;
        __syncthreads();
    }
    for (int q = threadIdx.x; q < kWY * kWY; q += blockDim.x) Tg[q] = wsm[(q / kWY) * LDT + q % kWY];
}

// GEMM tasks of the Gram matrices Gm[b] = V_b V_b^T of all blocks and matrices
__global__ void wy_gram_tasks_kernel(int N, int LD, int nblk, int live, int S, int64_t mstride,
                                     const double* __restrict__ Vt, double* __restrict__ Gm, GemmTask* __restrict__ tasks) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nblk * live * S) return;
    const int sb = q / S, sp = q % S;
    const int s = sb / nblk, b = sb % nblk;
    const int j0 = b * kWY;
    const int nb = min(kWY, N - 1 - j0);
    const int K = N - j0;
    const int Kc = (((K + S - 1) / S) + 15) & ~15;
    const int k0 = sp * Kc;
    int Ks = K - k0;
    if (Ks > Kc) Ks = Kc;
    if (Ks < 0) Ks = 0;
    GemmTask g;
    g.A = Vt + s * mstride + (int64_t)j0 * LD + j0 + (Ks > 0 ? k0 : 0);   // the reflectors vanish below component j0+1
    g.B = g.A;
    g.C = Gm + (int64_t)q * kWY * kWY;
    g.bidx = nullptr;
    g.M = nb; g.N = nb; g.K = Ks;
    g.lda = LD; g.ldb = LD; g.ldc = kWY;
    g.kmodB = 0;
    g.alpha = 1.0; g.beta = 0.0;
    tasks[q] = g;
}

// GEMM tasks of  VT_b = T_b V_b^T  (rows a of block b: sum_c T[a][c] Vt[j0+c][:]) for every block and matrix
__global__ void wy_vt_tasks_kernel(int N, int LD, int nblk, int live, int64_t mstride, const double* __restrict__ T,
                                   const double* __restrict__ Vt, double* __restrict__ VTt, GemmTask* __restrict__ tasks) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nblk * live) return;
    const int s = q / nblk, b = q % nblk;
    const int j0 = b * kWY;
    const int nb = min(kWY, N - 1 - j0);
    GemmTask g;
    g.A = T + (int64_t)q * kWY * kWY;
    g.B = Vt + s * mstride + (int64_t)j0 * LD + j0;
    g.C = VTt + s * mstride + (int64_t)j0 * LD + j0;
    g.bidx = nullptr;
    g.M = nb; g.N = N - j0; g.K = nb;
    g.lda = kWY; g.ldb = LD; g.ldc = LD;
    g.kmodB = 0;
    g.alpha = 1.0; g.beta = 0.0;
    tasks[q] = g;
}

// Wb[m][a] = sum over the K splits of Wa[m][split][a]
__global__ void __launch_bounds__(256)
wy_reduce_kernel(int N, int S, const double* __restrict__ Wa, double* __restrict__ Wb) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (int64_t)N * kWY) return;
    const int m = (int)(q / kWY), a = (int)(q % kWY);
    const double* src = Wa + (int64_t)blockIdx.y * N * S * kWY + (int64_t)m * S * kWY + a;
    double acc = 0.0;
    for (int sp = 0; sp < S; ++sp) acc += src[sp * kWY];
    Wb[(int64_t)blockIdx.y * N * kWY + q] = acc;
}

// GEMM tasks of  Wa = X V_b  for every block, matrix and K split: Wa[s][m][split * kWY + a]
__global__ void wy_w_tasks_kernel(int N, int LD, int nblk, int live, int S, int64_t mstride, const double* __restrict__ X,
                                  const double* __restrict__ Vt, double* __restrict__ Wa, double* __restrict__ Wb,
                                  GemmTask* __restrict__ tasks) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nblk * live * S) return;
    const int b = q / (live * S), s = (q / S) % live, sp = q % S;
    const int j0 = b * kWY;
    const int nb = min(kWY, N - 1 - j0);
    const int K = N - j0;
    const int Kc = (((K + S - 1) / S) + 15) & ~15;
    const int k0 = sp * Kc;
    int Ks = K - k0;
    if (Ks > Kc) Ks = Kc;
    if (Ks < 0) Ks = 0;
    GemmTask g;
    g.A = X + s * mstride + j0 + (Ks > 0 ? k0 : 0);
    g.B = Vt + s * mstride + (int64_t)j0 * LD + j0 + (Ks > 0 ? k0 : 0);
    g.C = S > 1 ? Wa + (int64_t)s * N * S * kWY + sp * kWY : Wb + (int64_t)s * N * kWY;   // one split: no reduction
    g.bidx = nullptr;
    g.M = N; g.N = nb; g.K = Ks;
    g.lda = LD; g.ldb = LD; g.ldc = S * kWY;
    g.kmodB = 0;
    g.alpha = 1.0; g.beta = 0.0;
    tasks[q] = g;
}

// modes[s][m][i] = X[s][m][i]
__global__ void __launch_bounds__(256)
export_rows_kernel(int N, int LD, int64_t mstride, const double* __restrict__ X, double* __restrict__ modes) {
    const int s = blockIdx.z, m = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    modes[((int64_t)s * N + m) * N + i] = X[s * mstride + (int64_t)m * LD + i];
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct TrdPlan {
    int LD, L, G, ngroups, group, nblk, cache_doubles;
    int wsplit;   // K splits of the skinny product X V_b of the back-transformation
    size_t smem;
};

int dc_levels(int N) {
    int L = 0, n = N;
    while (n > kLeaf) {
        n = n - ((n / 2) & ~1);   // the larger child
        ++L;
    }
    return L;
}

// SMs the cooperative kernel can count on (B200: 148); never more than the inbox pitch was sized for
int trd_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n < 1) {
        cudaGetLastError();
        return kNumSM;   // no device visible (size queries on a build host)
    }
    return n < kNumSM ? n : kNumSM;
}

constexpr size_t kTrdSmemBudget = 216 * 1024;   // + 8.5 KB static (partial sums, reduction scratch) <= 227 KB

TrdPlan make_plan(int B, int N) {
    TrdPlan p;
    const int sms = trd_sm_count();
    p.LD = (N + 3) & ~3;
    p.L = dc_levels(N);
    p.nblk = (int)ceil_div(N - 1, kWY);
    // matrices that share the launches of the batched stages: as many as fit ~3 GB of workspace, at most 64
    const size_t per = sizeof(double) * 4 * (size_t)N * p.LD;
    size_t g = ((size_t)3 << 30) / per;
    if (g < 1) g = 1;
    if (g > 64) g = 64;
    p.group = (int)(g < (size_t)B ? g : (size_t)B);
    // CTA groups of the tridiagonalisation (matrices reduced concurrently).  The per-column cost is dominated by
    // latencies that do not shrink with the work, so more, smaller groups win as long as the matrices in flight
    // stay L2-resident (rows that do not fit the shared-memory cache are served from L2): N = 900, 64 matrices:
    // 4 groups 118 ms, 8 groups 83 ms, 12 groups 79 ms, 16 groups 67..230 ms (L2 thrashing): 12 are chosen.  A single matrix
    // always gets every SM.
    p.ngroups = 1;
    const size_t vec = sizeof(double) * 3 * p.LD;
    {
        const size_t footprint = sizeof(double) * (size_t)N * p.LD;
        // the rows cached in shared memory (all SMs together ~27 MB) do not occupy L2
        int ng = (int)((((size_t)56 << 20) + (size_t)sms * (kTrdSmemBudget - vec)) / footprint);
        if (ng > p.group) ng = p.group;
        while (ng > 1 && (sms / ng < 4 || ceil_div(N, sms / ng) > kTrdMaxRows)) --ng;
        if (ng >= 1) p.ngroups = ng;
    }
    if (const char* env = getenv("SCB_TRD_GROUPS")) {   // A/B switch: CTA groups (matrices reduced concurrently)
        const int ng = atoi(env);
        if (ng >= 1 && ng <= 37 && ceil_div(N, sms / ng) <= kTrdMaxRows) p.ngroups = ng < p.group ? ng : p.group;
    }
    p.G = sms / p.ngroups;
    if (p.G > N) p.G = N;
    // X V_b has only N/128 x 1 tiles: split K until the launch fills the GPU
    p.wsplit = 1;
    while (p.wsplit < 6 && ceil_div(N, kTM) * p.group * p.wsplit < sms && N / (p.wsplit + 1) >= 256) ++p.wsplit;
    p.smem = kTrdSmemBudget;
    p.cache_doubles = (int)((p.smem - vec) / sizeof(double));
    return p;
}

struct TrdWork {
    double *A, *Vt, *Z0, *Z1;          // [group][N][LD]; A doubles as U and X
    double *tau, *d, *e, *sgn, *D0, *D1, *dsc, *dl, *w, *zh;   // [group][LD]
    int32_t *nd, *dfl, *rank, *pos, *ndg;   // [group][LD]
    stedc::Rotation* rot;              // [group][LD]
    int32_t *kc, *nr;                  // [group][2^L]
    GemmTask* tasks;                   // [group][2^L][2]
    GemmTask* wytasks;                 // [group][nblk]
    GemmTask* wtasks;                  // [nblk][group][wsplit]
    double *Gm, *T;                    // [group][nblk][wsplit][kWY][kWY] partial Gram matrices, [group][nblk][kWY][kWY]
    double *Wa, *Wb;                   // [group][N][wsplit * kWY] partial products, [group][N][kWY]
    double* xch;                       // [ngroups][kXchCopies][2][2][LD]
    unsigned* flags;                   // [ngroups][148][kInboxPad]
};

void trd_carve(Arena& ar, TrdWork* w, int N, const TrdPlan& p) {
    const size_t g = (size_t)p.group, m = (size_t)N * p.LD, v = (size_t)p.LD;
    const size_t nodes = (size_t)1 << p.L;
    w->A = ar.take<double>(g * m);
    w->Vt = ar.take<double>(g * m);
    w->Z0 = ar.take<double>(g * m);
    w->Z1 = ar.take<double>(g * m);
    w->tau = ar.take<double>(g * v);
    w->d = ar.take<double>(g * v);
    w->e = ar.take<double>(g * v);
    w->sgn = ar.take<double>(g * v);
    w->D0 = ar.take<double>(g * v);
    w->D1 = ar.take<double>(g * v);
    w->dsc = ar.take<double>(g * v);
    w->dl = ar.take<double>(g * v);
    w->w = ar.take<double>(g * v);
    w->zh = ar.take<double>(g * v);
    w->nd = ar.take<int32_t>(g * v);
    w->dfl = ar.take<int32_t>(g * v);
    w->rank = ar.take<int32_t>(g * v);
    w->pos = ar.take<int32_t>(g * v);
    w->ndg = ar.take<int32_t>(g * v);
    w->rot = ar.take<stedc::Rotation>(g * v);
    w->kc = ar.take<int32_t>(g * nodes);
    w->nr = ar.take<int32_t>(g * nodes);
    w->tasks = ar.take<GemmTask>(g * nodes * 2);
    w->wytasks = ar.take<GemmTask>(g * p.nblk * p.wsplit);
    w->wtasks = ar.take<GemmTask>(g * p.nblk * p.wsplit);
    w->Gm = ar.take<double>(g * p.nblk * kWY * kWY * p.wsplit);
    w->T = ar.take<double>(g * p.nblk * kWY * kWY);
    w->Wa = ar.take<double>(g * (size_t)N * kWY * p.wsplit);
    w->Wb = ar.take<double>(g * (size_t)N * kWY);
    w->xch = ar.take<double>((size_t)p.ngroups * kXchCopies * kXchBufs * 2 * v);
    w->flags = ar.take<unsigned>((size_t)p.ngroups * kNumSM * kInboxPad + 64);   // + the abort word
}

}  // namespace

bool eig_full_tridiag_supported(int N) {
    // three vectors of the reduction must fit in shared memory next to at least one cached row, and the device must
    // be able to keep one CTA per SM co-resident (cooperative launch); otherwise the block-Jacobi solver is used
    const size_t LD = (size_t)((N + 3) & ~3);
    // (N <= 9,200: the three vectors alone fill the shared memory, rows are then cached only once they have shrunk)
    // (block Jacobi is 3-5x slower already at N = 72...256: 3.7 / 6.4 ms against 0.8 / 2.2 ms)
    if (N <= 64 || sizeof(double) * 3 * LD + 2048 > kTrdSmemBudget) return false;
    int dev = 0, coop = 1, smem = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess) coop = 1;
        if (cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess &&
            (size_t)smem < kTrdSmemBudget + 9 * 1024)
            coop = 0;
    }
    cudaGetLastError();
    return coop != 0;
}

size_t eig_full_tridiag_workspace_bytes(int B, int N) {
    Arena ar(nullptr, 0);
    TrdWork w;
    trd_carve(ar, &w, N, make_plan(B, N));
    return ar.off + 256;
}

int eig_full_tridiag(int B, int N, double* A, double* eigval, double* modes, void* workspace, size_t workspace_bytes,
                     cudaStream_t st) {
    const TrdPlan p = make_plan(B, N);
    Arena ar(workspace, workspace_bytes);
    TrdWork w;
    trd_carve(ar, &w, N, p);
    if (!ar.ok()) return SCB_ERR_WORKSPACE;
    const int LD = p.LD, L = p.L;
    const int64_t mstride = (int64_t)N * LD, vstride = LD;
    // 384 threads per CTA (168 registers, no spills): 512 is 1-3 % slower (spills at 128 registers), 1,024 (barrier
    // skew) and 256 (too few warps for the row pass) clearly slower
    constexpr int trd_threads = 384;
    const void* trd_fn = (const void*)sytrd_kernel<384>;
    SCB_CUDA(cudaFuncSetAttribute(trd_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    SCB_CUDA(cudaFuncSetAttribute(dc_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 21 * N + 64));
    const size_t lsmem = sizeof(double) * 2 * kLeaf * (kLeaf + 1);
    SCB_CUDA(cudaFuncSetAttribute(dc_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem));
    const size_t tsmem = sizeof(double) * kWY * (kWY + 1);
    SCB_CUDA(cudaFuncSetAttribute(wy_tfactor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));

    for (int s0 = 0; s0 < B; s0 += p.group) {
        const int live = (B - s0 < p.group) ? B - s0 : p.group;
        // ---- 1. tridiagonalisation
        trd_init_kernel<<<dim3((unsigned)ceil_div((int64_t)N * LD, 256), (unsigned)live), 256, 0, st>>>(
            N, LD, A + (int64_t)s0 * N * N, w.A);
        SCB_LAUNCH_CHECK();
        SCB_CUDA(cudaMemsetAsync(w.Vt, 0, sizeof(double) * (size_t)live * mstride, st));
        SCB_CUDA(cudaMemsetAsync(w.Z0, 0, sizeof(double) * (size_t)live * mstride, st));
        SCB_CUDA(cudaMemsetAsync(w.Z1, 0, sizeof(double) * (size_t)live * mstride, st));
        SCB_CUDA(cudaMemsetAsync(w.flags, 0, sizeof(unsigned) * ((size_t)p.ngroups * kNumSM * kInboxPad + 64), st));
        TrdParams tp;
        tp.N = N; tp.LD = LD; tp.B = live; tp.G = p.G;
        tp.ngroups = p.ngroups < live ? p.ngroups : live;
        tp.cache_doubles = p.cache_doubles;
        tp.A = w.A; tp.Vt = w.Vt; tp.tau = w.tau; tp.d = w.d; tp.e = w.e; tp.xch = w.xch; tp.flags = w.flags;
        tp.abort_flag = w.flags + (size_t)p.ngroups * kNumSM * kInboxPad;
        tp.dbg = nullptr;
        long long* dbg = nullptr;
        if (getenv("SCB_TRD_DEBUG")) { cudaMalloc(&dbg, sizeof(long long) * 4 * kNumSM); tp.dbg = dbg; }
        void* kargs[] = {&tp};
        SCB_CUDA(cudaLaunchCooperativeKernel(trd_fn, dim3((unsigned)(tp.ngroups * tp.G)), dim3((unsigned)trd_threads), kargs,
                                             p.smem, st));
        count_launches(1);
        if (dbg) {
            long long h[4 * kNumSM];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
            const int nc = tp.ngroups * tp.G;
            for (int q = 0; q < nc; q += (nc > 8 ? nc / 8 : 1))
                fprintf(stderr, "sytrd cta %3d: phase1 %lld pass %lld barrier %lld repack %lld cycles\n", q, h[4 * q], h[4 * q + 1],
                        h[4 * q + 2], h[4 * q + 3]);
            cudaFree(dbg);
        }

        // ---- 2. divide and conquer
        DcWork W;
        W.N = N; W.LD = LD; W.L = L;
        W.d = w.d; W.e = w.e; W.sgn = w.sgn; W.D0 = w.D0; W.D1 = w.D1; W.Z0 = w.Z0; W.Z1 = w.Z1; W.U = w.A;
        W.dsc = w.dsc; W.dl = w.dl; W.w = w.w; W.zh = w.zh; W.nd = w.nd; W.dfl = w.dfl; W.pos = w.pos; W.ndg = w.ndg; W.rot = w.rot;
        W.kc = w.kc; W.nr = w.nr; W.tasks = w.tasks; W.vstride = vstride; W.mstride = mstride;
        dc_setup_kernel<<<live, 1024, 0, st>>>(W);
        SCB_LAUNCH_CHECK();
        dc_leaf_kernel<<<dim3(1u << L, (unsigned)live), 256, lsmem, st>>>(W);
        SCB_LAUNCH_CHECK();
        double *Dcur = w.D0, *Dnew = w.D1, *Zcur = w.Z0, *Znew = w.Z1;
        for (int level = L - 1; level >= 0; --level) {
            const unsigned nodes = 1u << level;
            int lo, hi;
            node_range(N, level, (int)nodes - 1, &lo, &hi);   // the last node of a level is the largest
            const int nmax = hi - lo;
            dc_prepare_kernel<<<dim3(nodes, (unsigned)live), 1024, 21 * (size_t)nmax + 64, st>>>(W, level, Dcur, Zcur, Znew);
            SCB_LAUNCH_CHECK();
            dc_rotate_kernel<<<dim3((unsigned)ceil_div(nmax, 256), nodes, (unsigned)live), 256, 0, st>>>(W, level, Zcur);
            SCB_LAUNCH_CHECK();
            dc_secular_kernel<<<dim3((unsigned)ceil_div(nmax, 8), nodes, (unsigned)live), 256, 0, st>>>(W, level, Dnew);
            SCB_LAUNCH_CHECK();
            dc_zhat_kernel<<<dim3((unsigned)ceil_div(nmax, 32), nodes, (unsigned)live), 1024, 0, st>>>(W, level);
            SCB_LAUNCH_CHECK();
            dc_vectors_kernel<<<dim3((unsigned)ceil_div(nmax, 8), nodes, (unsigned)live), 256, 0, st>>>(W, level);
            SCB_LAUNCH_CHECK();
            // tasks are stored [matrix][node]: contiguous over the live matrices only if nodes is the stride
            // two tasks per merge (left / right half of the columns); the right child is the larger one
            SCB_TRY((launch_gemm<true, false>(GemmTask{}, w.tasks, nmax, nmax - ((nmax / 2) & ~1), (int)nodes * live * 2, 0, 0, 0,
                                             st)));
            dc_copy_deflated_kernel<<<dim3((unsigned)nmax, nodes, (unsigned)live), 128, 0, st>>>(W, level, Zcur, Znew, Dnew);
            SCB_LAUNCH_CHECK();
            double* td = Dcur; Dcur = Dnew; Dnew = td;
            double* tz = Zcur; Zcur = Znew; Znew = tz;
        }
        dc_rank_kernel<<<dim3((unsigned)ceil_div(N, 256), (unsigned)live), 256, 0, st>>>(
            N, vstride, Dcur, w.rank, eigval + (int64_t)s0 * N);
        SCB_LAUNCH_CHECK();
        double* X = w.A;   // U is dead after the last merge
        dc_permute_kernel<<<dim3((unsigned)ceil_div(LD, 256), (unsigned)N, (unsigned)live), 256, 0, st>>>(
            N, LD, vstride, mstride, Zcur, w.rank, w.sgn, X);
        SCB_LAUNCH_CHECK();

        // ---- 3. back-transformation  x^T <- x^T Q_b^T = x^T (I - V_b T_b^T V_b^T) per block, last block first
        const int nblk = p.nblk;
        const int S = p.wsplit;
        double* VTt = Znew;   // both eigenvector buffers of the divide and conquer are dead: rows of  T_b V_b^T
        {
            // Gram matrices of all blocks (split along K) in one launch, T factors, then  VT_b = T_b V_b^T  in one launch
            wy_gram_tasks_kernel<<<(unsigned)ceil_div(nblk * live * S, 128), 128, 0, st>>>(N, LD, nblk, live, S, mstride, w.Vt,
                                                                                           w.Gm, w.wytasks);
            SCB_LAUNCH_CHECK();
            SCB_TRY((launch_gemm<true, true>(GemmTask{}, w.wytasks, kWY, kWY, nblk * live * S, 0, 0, 0, st)));
            wy_tfactor_kernel<<<dim3((unsigned)nblk, (unsigned)live), 256, tsmem, st>>>(N, vstride, nblk, S, w.tau, w.Gm, w.T);
            SCB_LAUNCH_CHECK();
            wy_vt_tasks_kernel<<<(unsigned)ceil_div(nblk * live, 128), 128, 0, st>>>(N, LD, nblk, live, mstride, w.T, w.Vt, VTt,
                                                                                     w.wytasks);
            SCB_LAUNCH_CHECK();
            SCB_TRY((launch_gemm<true, false>(GemmTask{}, w.wytasks, kWY, N, nblk * live, 0, 0, 0, st)));
        }
        wy_w_tasks_kernel<<<(unsigned)ceil_div(nblk * live * S, 128), 128, 0, st>>>(N, LD, nblk, live, S, mstride, X, VTt, w.Wa,
                                                                                   w.Wb, w.wtasks);
        SCB_LAUNCH_CHECK();
        for (int b = nblk - 1; b >= 0; --b) {
            const int j0 = b * kWY;
            const int nb = (N - 1 - j0 < kWY) ? N - 1 - j0 : kWY;
            // Wa[m][split][a] = sum over the split's components i of X[m][i] (T_b V_b^T)[a][i]
            SCB_TRY((launch_gemm<true, true>(GemmTask{}, w.wtasks + (size_t)b * live * S, N, nb, live * S, 0, 0, 0, st)));
            if (S > 1) {
                wy_reduce_kernel<<<dim3((unsigned)ceil_div((int64_t)N * kWY, 256), (unsigned)live), 256, 0, st>>>(N, S, w.Wa, w.Wb);
                SCB_LAUNCH_CHECK();
            }
            // X[m][i] -= sum_a Wb[m][a] Vt[j0+a][i]
            GemmTask g;
            g.bidx = nullptr; g.kmodB = 0;
            g.A = w.Wb; g.B = w.Vt + (int64_t)j0 * LD + j0; g.C = X + j0;
            g.M = N; g.N = N - j0; g.K = nb; g.lda = kWY; g.ldb = LD; g.ldc = LD;
            g.alpha = -1.0; g.beta = 1.0;
            SCB_TRY((launch_gemm<true, false>(g, nullptr, N, N - j0, live, (int64_t)N * kWY, mstride, mstride, st)));
        }
        export_rows_kernel<<<dim3((unsigned)ceil_div(N, 256), (unsigned)N, (unsigned)live), 256, 0, st>>>(
            N, LD, mstride, X, modes + (int64_t)s0 * N * N);
        SCB_LAUNCH_CHECK();
        // the only host synchronisation of the solver: did a barrier of the cooperative kernel give up?
        unsigned aborted = 0;
        SCB_CUDA(cudaMemcpyAsync(&aborted, w.flags + (size_t)p.ngroups * kNumSM * kInboxPad, sizeof(unsigned),
                                 cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaStreamSynchronize(st));
        if (aborted) {
            set_last_cuda_error(cudaErrorLaunchTimeout, __FILE__, __LINE__);
            return SCB_ERR_CUDA;
        }
    }
    return SCB_OK;
}

}  // namespace scb
