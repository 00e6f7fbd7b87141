// K3 (dense operator): row-slab product of a dense interaction matrix with a block vector, fused
// with the Chebyshev recurrence.  This is the operator of the lowest-k solver for all-pairs force
// fields (cutoff None: ParameterFree / Hinsen), where the Hessian is dense and
// each GPU owns a slab of rows:   Y[rows] = alpha * (H[rows, :] X - c X[rows]) - beta W[rows].
// Single GPU: Y is the local slab [rows][b].  Several GPUs: the all-gather of the row slabs is FUSED
// into the epilogue (scb_dense_slab_apply_allgather): every rank holds a full [N][b] output block in
// peer-mapped memory (cudaIpc), and each output tile is stored straight into the full block of EVERY
// rank over NVLink as soon as its accumulators are final, so the exchange overlaps the remaining tiles
// of the product; a tiny barrier collective afterwards replaces the NCCL all-gather.
//
// DMMA (mma.sync.m8n8k4.f64) tiles of 64 rows x 64 columns, K (= N) in steps of 16, cp.async double
// buffering.  The slab is streamed from HBM exactly once per application (8 N^2 / G bytes per GPU);
// at b >= 64 columns the kernel is FP64-tensor bound (intensity b/4 flop/B).
//
// Also exported here: the building blocks of the solver loop that the Python host code drives for
// this path (Gram matrices, Cholesky orthonormalisation, rotation, deflation, residual norms).
#include <stdlib.h>
#include <string.h>

#include "subspace.cuh"

namespace scb {

constexpr int kDsBM = 64, kDsBN = 64, kDsBK = 16;
constexpr int kDsLDA = 20;   // 16 + 4  : conflict-free A fragments
constexpr int kDsLDB = 68;   // 64 + 4  : conflict-free B fragments

__device__ __forceinline__ void ds_cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(bytes));
}

__device__ __forceinline__ void ds_cp_async8(void* smem, const void* gmem, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(bytes));
}

__device__ __forceinline__ void ds_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int kMaxPeers = 16;
struct PeerBlocks {
    double* full[kMaxPeers];   // full [N][b] output block of every rank (peer-mapped), entry `rank` is local
    int world;
};

// slab: [rows][N] row-major (rows = row1 - row0 matrix rows), X/W: [N][b].
// PEERS == false: Y is the local slab [rows][b].  PEERS == true: rows [row0, row0+rows) of every rank's full block.
template <bool PEERS>
__global__ void __launch_bounds__(128)
dense_slab_apply_kernel(int64_t N, int64_t row0, int64_t rows, int b, const double* __restrict__ slab,
                        const double* __restrict__ X, const double* __restrict__ W, double* __restrict__ Y,
                        double alpha, double cshift, double beta, int fused, PeerBlocks peers, int splits,
                        double* __restrict__ partial, unsigned int* __restrict__ counters) {
    __shared__ __align__(16) double sA[2][kDsBM * kDsLDA];
    __shared__ __align__(16) double sB[2][kDsBK * kDsLDB];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;  // 2 x 2 warps, warp tile 32 x 32
    const int64_t m0 = (int64_t)blockIdx.x * kDsBM;
    const int n0 = blockIdx.y * kDsBN;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const bool even_n = (N & 1) == 0;
    auto load_stage = [&](int stage, int64_t kk) {
        // A: 64 rows x 16 doubles = 512 16-byte chunks (4 per thread)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = tid + q * 128;
            const int r = c >> 3, kc = (c & 7) * 2;
            if (even_n) {
                const bool ok = (m0 + r) < rows && (kk + kc) < N;
                ds_cp_async16(&sA[stage][r * kDsLDA + kc], ok ? (const void*)(slab + (m0 + r) * N + kk + kc) : (const void*)slab, ok);
            } else {   // odd N: rows of the slab are only 8-byte aligned
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const bool ok = (m0 + r) < rows && (kk + kc + e) < N;
                    ds_cp_async8(&sA[stage][r * kDsLDA + kc + e], ok ? (const void*)(slab + (m0 + r) * N + kk + kc + e) : (const void*)slab, ok);
                }
            }
        }
        // B: 16 rows (k) x 64 doubles = 512 chunks (4 per thread)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = tid + q * 128;
            const int r = c >> 5, nc = (c & 31) * 2;
            const bool ok = (kk + r) < N;
            ds_cp_async16(&sB[stage][r * kDsLDB + nc], ok ? (const void*)(X + (kk + r) * b + n0 + nc) : (const void*)X, ok);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };

    // split-K: blockIdx.z owns the k-steps [it0, it1) (splits == 1: all of them)
    const int64_t nk_all = ceil_div(N, (int64_t)kDsBK);
    const int64_t it0 = nk_all * blockIdx.z / splits, nk = nk_all * (blockIdx.z + 1) / splits;
    load_stage((int)(it0 & 1), it0 * kDsBK);
    for (int64_t it = it0; it < nk; ++it) {
        const int stage = (int)(it & 1);
        if (it + 1 < nk) {
            load_stage(stage ^ 1, (it + 1) * kDsBK);
            asm volatile("cp.async.wait_group 1;\n" ::);
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
        const double* a_s = &sA[stage][(wm * 32) * kDsLDA];
        const double* b_s = &sB[stage][wn * 32];
#pragma unroll
        for (int k4 = 0; k4 < kDsBK; k4 += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = a_s[(i * 8 + (lane >> 2)) * kDsLDA + k4 + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = b_s[(k4 + (lane & 3)) * kDsLDB + j * 8 + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) ds_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
    auto emit = [&](int64_t r, int c, double v0, double v1) {   // Chebyshev epilogue + store of two columns
        if (fused) {
            const int64_t g = (row0 + r) * b + c;  // global row of X / W
            v0 = alpha * (v0 - cshift * X[g]);
            v1 = alpha * (v1 - cshift * X[g + 1]);
            if (W && beta != 0.0) { v0 -= beta * W[g]; v1 -= beta * W[g + 1]; }
        }
        if (PEERS) {
            const int64_t g = (row0 + r) * b + c;
#pragma unroll 1
            for (int p = 0; p < peers.world; ++p)
                *reinterpret_cast<double2*>(&peers.full[p][g]) = make_double2(v0, v1);
        } else {
            *reinterpret_cast<double2*>(&Y[r * b + c]) = make_double2(v0, v1);
        }
    };
    if (splits == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t r = m0 + wm * 32 + i * 8 + (lane >> 2);
            if (r >= rows) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                emit(r, n0 + wn * 32 + j * 8 + 2 * (lane & 3), acc[i][j][0], acc[i][j][1]);
        }
    } else {
        // deterministic split-K: every split parks its 64x64 partial tile; the CTA that arrives last adds the
        // partials in split order (the same order whichever CTA is last) and runs the epilogue
        __shared__ int is_last;
        const int64_t tile = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
        const int64_t ntiles = (int64_t)gridDim.x * gridDim.y;
        double* mine = partial + ((int64_t)blockIdx.z * ntiles + tile) * (kDsBM * kDsBN);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int lr = wm * 32 + i * 8 + (lane >> 2), lc = wn * 32 + j * 8 + 2 * (lane & 3);
                *reinterpret_cast<double2*>(&mine[lr * kDsBN + lc]) = make_double2(acc[i][j][0], acc[i][j][1]);
            }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned int seen = atomicAdd(&counters[tile], 1u);
            is_last = (seen == (unsigned int)splits - 1);
            if (is_last) counters[tile] = 0;   // ready for the next launch
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            const double* base = partial + tile * (kDsBM * kDsBN);
            for (int q = tid; q < kDsBM * kDsBN / 2; q += 128) {
                const int lr = q / (kDsBN / 2), lc = 2 * (q % (kDsBN / 2));
                double v0 = 0.0, v1 = 0.0;
                for (int sp = 0; sp < splits; ++sp) {
                    const double2 t = *reinterpret_cast<const double2*>(&base[(int64_t)sp * ntiles * (kDsBM * kDsBN) + lr * kDsBN + lc]);
                    v0 += t.x;
                    v1 += t.y;
                }
                const int64_t r = m0 + lr;
                if (r < rows) emit(r, n0 + lc, v0, v1);
            }
        }
    }
    if (PEERS) __threadfence_system();   // remote stores performed before the kernel retires
}

// max row sum of |entries| of a dense slab (Gershgorin bound of the local rows)
__global__ void __launch_bounds__(256)
dense_gershgorin_kernel(int64_t N, int64_t rows, const double* __restrict__ slab, double* __restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    double acc = 0.0;
    for (int64_t c = lane_id(); c < N; c += 32) acc += fabs(slab[r * N + c]);
    acc = warp_sum(acc);
    if (lane_id() == 0) atomic_max_nonneg(out, acc);
}

// S (b x b, SPD) -> C = L^-T with S = L L^T, so that X C has orthonormal columns.  One CTA, the
// factor lives in shared memory (b <= 160), L^-1 is built column by column straight into C.
__global__ void __launch_bounds__(256)
chol_orth_kernel(int b, const double* __restrict__ Sg, double* __restrict__ Cout) {
    extern __shared__ double sm[];
    const int LD = b + 1;
    double* S = sm;
    __shared__ double red[8];
    const int s = blockIdx.x;
    const int tid = threadIdx.x;
    const double* Ss = Sg + (int64_t)s * b * b;
    double* Cs = Cout + (int64_t)s * b * b;
    for (int q = tid; q < b * b; q += 256) {
        const int i = q / b, j = q % b;
        S[i * LD + j] = 0.5 * (Ss[i * b + j] + Ss[j * b + i]);
    }
    __syncthreads();
    double tr = 0.0;
    for (int q = tid; q < b; q += 256) tr += S[q * LD + q];
    tr = warp_sum(tr);
    if (lane_id() == 0) red[tid >> 5] = tr;
    __syncthreads();
    tr = 0.0;
    for (int w = 0; w < 8; ++w) tr += red[w];
    const double floor_piv = 1e-28 * tr + 1e-300;
    for (int k = 0; k < b; ++k) {
        if (tid == 0) S[k * LD + k] = sqrt(fmax(S[k * LD + k], floor_piv));
        __syncthreads();
        const double d = S[k * LD + k];
        for (int i = k + 1 + tid; i < b; i += 256) S[i * LD + k] /= d;
        __syncthreads();
        const int rem = b - 1 - k;
        for (int q = tid; q < rem * rem; q += 256) {
            const int i = k + 1 + q / rem, j = k + 1 + q % rem;
            if (j <= i) S[i * LD + j] -= S[i * LD + k] * S[j * LD + k];
        }
        __syncthreads();
    }
    // column j of L^-1 by forward substitution; C[p][c] = (L^-T)[p][c] = Linv[c][p]
    for (int q = tid; q < b * b; q += 256) Cs[q] = 0.0;
    __syncthreads();
    for (int j = tid; j < b; j += 256) {
        for (int i = j; i < b; ++i) {
            double acc = (i == j) ? 1.0 : 0.0;
            for (int k = j; k < i; ++k) acc -= S[i * LD + k] * Cs[j * b + k];  // Linv[k][j] stored at C[j][k]
            Cs[j * b + i] = acc / S[i * LD + i];
        }
    }
}

// C[p][q] = modes[q][p]  (eigenvector rows -> rotation matrix columns)
__global__ void transpose_small_kernel(int b, const double* __restrict__ in, double* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= b * b) return;
    out[(q % b) * b + q / b] = in[q];
}

}  // namespace scb

using namespace scb;

// Split-K factor: a slab with few 64x64 output tiles (many GPUs, or a small matrix) would leave SMs with
// uneven tile counts; cut K until there are >= 12 work units per SM, i.e. 3 waves of the 4 resident CTAs
// (measured: 7,500 x 60,000 slab 4.51 -> 3.63 ms, 15,000 rows 9.06 -> 7.2 ms; the reduction is deterministic).
static int slab_splits(int64_t N, int64_t rows, int b) {
    const int64_t tiles = ceil_div(rows, (int64_t)kDsBM) * (b / kDsBN);
    int splits = 1;
    if (const char* env = getenv("SCB_SLAB_SPLITS")) return atoi(env) > 0 ? atoi(env) : 1;
    while (tiles * splits < 12 * kNumSM && splits < 8 && N / (2 * splits) >= 64 * kDsBK) splits *= 2;
    return splits;
}

static size_t slab_partial_bytes(int64_t rows, int b, int splits) {
    const int64_t tiles = ceil_div(rows, (int64_t)kDsBM) * (b / kDsBN);
    return (size_t)256 + sizeof(double) * (size_t)splits * tiles * kDsBM * kDsBN + 256 + sizeof(unsigned int) * tiles;
}

extern "C" size_t scb_dense_slab_workspace_bytes(int64_t N, int64_t row0, int64_t row1, int b) {
    if (N < 1 || row1 <= row0 || b % 64 != 0) return 0;
    return slab_partial_bytes(row1 - row0, b, 8);   // enough for any split factor the launcher may pick
}

template <bool PEERS>
static int slab_launch(int64_t N, int64_t row0, int64_t rows, int b, const double* slab, const double* X,
                       const double* W, double* Y, double alpha, double cshift, double beta, int fused,
                       const PeerBlocks& peers, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    int splits = workspace ? slab_splits(N, rows, b) : 1;
    while (splits > 1 && slab_partial_bytes(rows, b, splits) > workspace_bytes) splits /= 2;
    const int64_t tiles = ceil_div(rows, (int64_t)kDsBM) * (b / kDsBN);
    double* partial = nullptr;
    unsigned int* counters = nullptr;
    if (splits > 1) {
        Arena ar(workspace, workspace_bytes);
        partial = ar.take<double>((size_t)splits * tiles * kDsBM * kDsBN);
        counters = ar.take<unsigned int>((size_t)tiles);
        if (!ar.ok()) return SCB_ERR_WORKSPACE;
    }
    dim3 grid((unsigned)ceil_div(rows, kDsBM), (unsigned)(b / kDsBN), (unsigned)splits);
    dense_slab_apply_kernel<PEERS><<<grid, 128, 0, st>>>(N, row0, rows, b, slab, X, W, Y, alpha, cshift, beta, fused,
                                                         peers, splits, partial, counters);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_dense_slab_apply(int64_t N, int64_t row0, int64_t row1, const double* slab, const double* X,
                                    const double* W, double* Y, int b, int fused, double alpha, double cshift,
                                    double beta, void* workspace, size_t workspace_bytes, void* stream) {
    if (!slab || !X || !Y || N < 1 || row0 < 0 || row1 > N || row0 >= row1 || b % 64 != 0) return SCB_ERR_INVALID;
    return slab_launch<false>(N, row0, row1 - row0, b, slab, X, W, Y, alpha, cshift, beta, fused, PeerBlocks{},
                              workspace, workspace_bytes, as_stream(stream));
}

extern "C" int scb_dense_slab_apply_allgather(int64_t N, int64_t row0, int64_t row1, const double* slab,
                                              const double* X, const double* W, double* const* Y_all, int world,
                                              int b, int fused, double alpha, double cshift, double beta,
                                              void* workspace, size_t workspace_bytes, void* stream) {
    if (!slab || !X || !Y_all || N < 1 || row0 < 0 || row1 > N || row0 >= row1 || b % 64 != 0 || world < 1 ||
        world > kMaxPeers)
        return SCB_ERR_INVALID;
    PeerBlocks peers{};
    peers.world = world;
    for (int p = 0; p < world; ++p) {
        if (!Y_all[p] || Y_all[p] == X || Y_all[p] == W) return SCB_ERR_INVALID;
        peers.full[p] = Y_all[p];
    }
    return slab_launch<true>(N, row0, row1 - row0, b, slab, X, W, nullptr, alpha, cshift, beta, fused, peers, workspace,
                             workspace_bytes, as_stream(stream));
}

// ---- peer-mapped device buffers (CUDA IPC): one process per GPU, every rank maps the buffers of the others
extern "C" int scb_peer_alloc(size_t bytes, void** dptr) {
    if (!dptr || bytes == 0) return SCB_ERR_INVALID;
    SCB_CUDA(cudaMalloc(dptr, bytes));
    SCB_CUDA(cudaMemset(*dptr, 0, bytes));
    return SCB_OK;
}

extern "C" int scb_peer_free(void* dptr) {
    if (!dptr) return SCB_ERR_INVALID;
    SCB_CUDA(cudaFree(dptr));
    return SCB_OK;
}

extern "C" int scb_peer_export(void* dptr, unsigned char* handle64) {
    if (!dptr || !handle64) return SCB_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    SCB_CUDA(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle64, &h, sizeof(h));
    return SCB_OK;
}

extern "C" int scb_peer_open(const unsigned char* handle64, void** dptr) {
    if (!dptr || !handle64) return SCB_ERR_INVALID;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    SCB_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SCB_OK;
}

extern "C" int scb_peer_close(void* dptr) {
    if (!dptr) return SCB_ERR_INVALID;
    SCB_CUDA(cudaIpcCloseMemHandle(dptr));
    return SCB_OK;
}

extern "C" int scb_dense_gershgorin(int64_t N, int64_t rows, const double* slab, double* out, void* stream) {
    if (!slab || !out || N < 1 || rows < 1) return SCB_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    SCB_CUDA(cudaMemsetAsync(out, 0, sizeof(double), st));
    dense_gershgorin_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, st>>>(N, rows, slab, out);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_gram(int B, int64_t N, int b, const double* A, const double* Bm, double* G, void* stream) {
    if (!A || !Bm || !G || B < 1 || N < 1) return SCB_ERR_INVALID;
    return gram(B, N, b, A, Bm, G, nullptr, as_stream(stream));
}

extern "C" int scb_chol_orth(int B, int b, const double* S, double* C, void* stream) {
    if (!S || !C || B < 1 || b < 1 || b > 160) return SCB_ERR_INVALID;
    const size_t smem = sizeof(double) * (size_t)b * (b + 1);
    // per-device function attribute: set on every launch (cheap), no process-wide "configured" flag
    SCB_CUDA(cudaFuncSetAttribute(chol_orth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 161 * 8));
    chol_orth_kernel<<<B, 256, smem, as_stream(stream)>>>(b, S, C);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_rotate(int B, int64_t N, int b, const double* C, const double* Xin, double* Xout,
                          const double* Yin, double* Yout, void* stream) {
    if (!C || !Xin || !Xout || B < 1 || N < 1) return SCB_ERR_INVALID;
    return rotate(B, N, b, C, Xin, Xout, Yin, Yout, nullptr, as_stream(stream));
}

extern "C" int scb_deflate(int B, int64_t N, int b, int nz, const double* Z, double* X, double* scratch, void* stream) {
    if (!X || !scratch || B < 1 || N < 1) return SCB_ERR_INVALID;
    return deflate(B, N, b, nz, Z, X, scratch, nullptr, as_stream(stream));
}

extern "C" int scb_residual_norms(int B, int64_t N, int b, const double* X, const double* HX, const double* theta,
                                  double* rn2, void* stream) {
    if (!X || !HX || !theta || !rn2 || B < 1 || N < 1) return SCB_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    SCB_CUDA(cudaMemsetAsync(rn2, 0, sizeof(double) * (size_t)B * b, st));
    return residual_norms(B, N, b, X, HX, theta, rn2, nullptr, st);
}

extern "C" int scb_transpose_small(int b, const double* in, double* out, void* stream) {
    if (!in || !out || b < 1) return SCB_ERR_INVALID;
    transpose_small_kernel<<<(unsigned)ceil_div((int64_t)b * b, 256), 256, 0, as_stream(stream)>>>(b, in, out);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_rand_block(int64_t total, uint64_t seed, double* X, void* stream) {
    if (!X || total < 1) return SCB_ERR_INVALID;
    return rand_init(total, seed, X, as_stream(stream));
}

// ---- column-wise Lanczos building blocks (spectrum bound of operators applied by the caller, e.g. the dense
// row-slab operator): the same kernels the batched sparse solver uses internally
extern "C" int scb_coldot(int B, int64_t N, int b, const double* A, const double* Bm, double* out, void* stream) {
    if (!A || !Bm || !out || B < 1 || N < 1) return SCB_ERR_INVALID;
    return scb::coldot(B, N, b, A, Bm, out, scb::as_stream(stream));
}

extern "C" int scb_lanczos_axpy(int B, int64_t N, int b, int mode, double* V, double* Vprev, double* W,
                                const double* alpha, const double* beta_prev, const double* nrm2, void* stream) {
    if (!V || B < 1 || N < 1 || mode < 0 || mode > 2) return SCB_ERR_INVALID;
    if (mode == 0 && (!Vprev || !W || !alpha)) return SCB_ERR_INVALID;
    if (mode == 1 && (!Vprev || !W || !nrm2)) return SCB_ERR_INVALID;
    if (mode == 2 && !nrm2) return SCB_ERR_INVALID;
    return scb::lanczos_axpy(B, N, b, mode, V, Vprev, W, alpha, beta_prev, nrm2, scb::as_stream(stream));
}

extern "C" int scb_lanczos_bound(int B, int b, int steps, const double* alpha, const double* beta2, double factor,
                                 double* out, void* stream) {
    if (!alpha || !beta2 || !out || B < 1 || steps < 1 || steps > 64) return SCB_ERR_INVALID;
    return scb::lanczos_bound_plain(B, b, steps, alpha, beta2, factor, out, scb::as_stream(stream));
}
