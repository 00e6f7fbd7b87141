// K3c (large): two-sided BLOCK Jacobi for dense symmetric matrices, full
// spectrum.  Replaces LAPACK dsyevd behind np.linalg.eigh (nma.py:61) and
// np.linalg.pinv (anm.py:135) when every mode is requested.
//
// The matrix is cut into column blocks of width 32.  A round-robin tournament
// pairs the blocks; for each pair (I,J) the 64x64 pivot sub-matrix is fully
// diagonalised in shared memory by cyclic Jacobi (one CTA per pair), giving an
// orthogonal R.  The update A <- R^T A R, V <- V R is applied by ONE launch of
// 64x64x64 tile products on the FP64 tensor cores (DMMA m8n8k4): the pairs of a
// step are disjoint, so tile (k,l) of A (rows of pair k, columns of pair l) is
// A_kl <- R_k^T A_kl R_l independently of all other tiles; only k <= l is
// computed and mirrored (A stays exactly symmetric), V tiles need one product.
// Step 0 of a sweep rotates every index pair inside its pivots, the other
// steps only the pairs ACROSS the two blocks, i.e. each index pair is rotated
// once per sweep (cyclic-by-blocks ordering).
// Sweeps repeat until the off-diagonal Frobenius norm is ~1e-14 of the total.
// A batch is solved in groups of up to 32 matrices that share every launch
// (grid.y = matrix): the pivots of one N=900 matrix occupy only 15 SMs, so
// batching is what fills the GPU for ensembles of small structures.
#include <stdlib.h>

#include "subspace.cuh"
#include "jacobi.cuh"

namespace scb {

constexpr int kBW = 32;         // block width
constexpr int kPW = 2 * kBW;    // pivot order
constexpr int kTLD = 68;        // padded leading dimension of the smem tiles (conflict-free DMMA fragments)

__device__ __forceinline__ double bj_block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    return t;
}

__device__ __forceinline__ void block_pair(int nb, int step, int k, int* I, int* J) {
    // circle method: player nb-1 is fixed, the others rotate
    int p, q;
    if (k == 0) { p = nb - 1; q = step % (nb - 1); }
    else { p = (step + k) % (nb - 1); q = (step - k + (nb - 1)) % (nb - 1); }
    *I = min(p, q);
    *J = max(p, q);
}

// Ap (Np x Np) <- symmetric extension of the lower triangle of A (N x N), zero padded; V <- I
__global__ void bj_init_kernel(int N, int Np, const double* __restrict__ A, double* __restrict__ Ap,
                               double* __restrict__ V) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (int64_t)Np * Np) return;
    A += (int64_t)blockIdx.y * N * N;
    Ap += (int64_t)blockIdx.y * Np * Np;
    V += (int64_t)blockIdx.y * Np * Np;
    const int i = (int)(q / Np), j = (int)(q % Np);
    double v = 0.0;
    if (i < N && j < N) v = (j <= i) ? A[(int64_t)i * N + j] : A[(int64_t)j * N + i];
    Ap[q] = v;
    V[q] = (i == j) ? 1.0 : 0.0;
}

// one CTA per block pair: diagonalise the 64x64 pivot, write R and R^T [pair][2][64][64]
constexpr int kPivotThreads = 512;

__global__ void __launch_bounds__(kPivotThreads)
bj_pivot_kernel(int Np, int nb, int step, const double* __restrict__ Ap, double* __restrict__ R,
                int32_t* __restrict__ active, int inner_sweeps, int cross_only, const int32_t* __restrict__ mdone) {
    constexpr int LD = kPW + 1;
    constexpr int NT = kPivotThreads;
    if (mdone[blockIdx.y]) return;   // this matrix of the group has converged
    Ap += (int64_t)blockIdx.y * Np * Np;
    R += (int64_t)blockIdx.y * (nb / 2) * 2 * kPW * kPW;
    active += (int64_t)blockIdx.y * (nb / 2);
    extern __shared__ double sm[];
    double* S = sm;
    double* V = S + kPW * LD;
    __shared__ double cs[kPW], sn[kPW], red[NT / 32];   // two rotation-parameter sets of kPW/2
    __shared__ int pp[kPW], qq[kPW];
    int I, J;
    block_pair(nb, step, blockIdx.x, &I, &J);
    const int tid = threadIdx.x;
    for (int q = tid; q < kPW * kPW; q += NT) {
        const int r = q / kPW, c = q % kPW;
        const int gr = (r < kBW ? I * kBW + r : J * kBW + r - kBW);
        const int gc = (c < kBW ? I * kBW + c : J * kBW + c - kBW);
        S[r * LD + c] = Ap[(int64_t)gr * Np + gc];
        V[r * LD + c] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    // the diagonal tile is symmetric up to rounding only: average the two triangles
    for (int q = tid; q < kPW * kPW; q += NT) {
        const int r = q / kPW, c = q % kPW;
        if (r < c) {
            const double a = 0.5 * (S[r * LD + c] + S[c * LD + r]);
            S[r * LD + c] = a;
            S[c * LD + r] = a;
        }
    }
    __syncthreads();
    // skip pivots whose coupling block is already negligible
    double off = 0.0, dg = 0.0;
    for (int q = tid; q < kPW * kPW; q += NT) {
        const int r = q / kPW, c = q % kPW;
        const double v = S[r * LD + c];
        if (r == c) dg += v * v; else off += v * v;
    }
    off = block_sum_nt<NT>(off, red);
    dg = block_sum_nt<NT>(dg, red);
    const bool work = off > 1e-32 * dg && off > 0.0;
    // inexact pivot diagonalisation: ONE cyclic sweep per visit is enough for the outer iteration to converge
    // (quadratically at the end) and costs 5x less than a full inner solve
    if (work) jacobi_eigen_smem<kPW, LD, NT>(S, V, cs, sn, pp, qq, red, inner_sweeps, cross_only != 0);
    double* Rp = R + (int64_t)blockIdx.x * 2 * kPW * kPW;
    for (int q = tid; q < kPW * kPW; q += NT) {
        Rp[q] = V[(q / kPW) * LD + q % kPW];
        Rp[kPW * kPW + q] = V[(q % kPW) * LD + q / kPW];
    }
    if (tid == 0) active[blockIdx.x] = work ? 1 : 0;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// out(64x64) = L(64x64, row-major in sL) * Rm(64x64, row-major in sR); warp w owns output rows 8w..8w+7
__device__ __forceinline__ void tile_product(const double* sL, const double* sR, double acc[8][2]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
#pragma unroll 4
    for (int k4 = 0; k4 < kPW; k4 += 4) {
        const double a = sL[(warp * 8 + (lane >> 2)) * kTLD + k4 + (lane & 3)];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double b = sR[(k4 + (lane & 3)) * kTLD + j * 8 + (lane >> 2)];
            dmma884(acc[j][0], acc[j][1], a, b);
        }
    }
}

__device__ __forceinline__ void bj_cp_async16(void* smem, const void* gmem) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}

__device__ __forceinline__ int pair_index(int I, int J, int r) { return r < kBW ? I * kBW + r : J * kBW + r - kBW; }

// One launch per step.  CTAs [0, nA): tiles k <= l of A, A_kl <- R_k^T A_kl R_l, mirrored into A_lk.
// CTAs [nA, nA + (Np/64) * npairs): V[64 rows][columns of pair l] <- (same) R_l.
__global__ void __launch_bounds__(256)
bj_update_kernel(int Np, int nb, int step, double* __restrict__ Ap, double* __restrict__ V,
                 const double* __restrict__ R, const int32_t* __restrict__ active, const int32_t* __restrict__ mdone) {
    constexpr int TLD = kPW + 1;   // transposition buffer: odd leading dimension
    if (mdone[blockIdx.y]) return;
    Ap += (int64_t)blockIdx.y * Np * Np;
    V += (int64_t)blockIdx.y * Np * Np;
    R += (int64_t)blockIdx.y * (nb / 2) * 2 * kPW * kPW;
    active += (int64_t)blockIdx.y * (nb / 2);
    extern __shared__ double sm[];
    double* sL = sm;                    // left factor (R_k^T, then the mirror staging buffer)
    double* sM = sm + kPW * kTLD;       // the tile (then T = R_k^T A_kl)
    double* sR = sm + 2 * kPW * kTLD;   // right factor R_l
    const int npairs = nb / 2;
    const int nA = npairs * (npairs + 1) / 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc[8][2];
    if ((int)blockIdx.x >= nA) {
        const int v = blockIdx.x - nA;
        const int r0 = (v / npairs) * kPW, l = v % npairs;
        if (!active[l]) return;
        int I, J;
        block_pair(nb, step, l, &I, &J);
        const double* Rl = R + (int64_t)l * 2 * kPW * kPW;
        for (int q = tid; q < kPW * kPW / 2; q += 256) {
            const int r = q / (kPW / 2), c = 2 * (q % (kPW / 2));
            bj_cp_async16(&sR[r * kTLD + c], &Rl[r * kPW + c]);
            bj_cp_async16(&sM[r * kTLD + c], &V[(int64_t)(r0 + r) * Np + pair_index(I, J, c)]);
        }
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        tile_product(sM, sR, acc);
        const int r = r0 + warp * 8 + (lane >> 2);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = j * 8 + 2 * (lane & 3);
            *reinterpret_cast<double2*>(&V[(int64_t)r * Np + pair_index(I, J, c)]) = make_double2(acc[j][0], acc[j][1]);
        }
        return;
    }
    int t = blockIdx.x, k = 0;
    while (t >= npairs - k) { t -= npairs - k; ++k; }
    const int l = k + t;
    if (!active[k] && !active[l]) return;
    int Ik, Jk, Il, Jl;
    block_pair(nb, step, k, &Ik, &Jk);
    block_pair(nb, step, l, &Il, &Jl);
    const double* RkT = R + (int64_t)k * 2 * kPW * kPW + kPW * kPW;
    const double* Rl = R + (int64_t)l * 2 * kPW * kPW;
    // asynchronous 16-byte copies: {R_k^T, A_kl} first, R_l lands while the first product runs
    for (int q = tid; q < kPW * kPW / 2; q += 256) {
        const int r = q / (kPW / 2), c = 2 * (q % (kPW / 2));
        bj_cp_async16(&sL[r * kTLD + c], &RkT[r * kPW + c]);
        bj_cp_async16(&sM[r * kTLD + c], &Ap[(int64_t)pair_index(Ik, Jk, r) * Np + pair_index(Il, Jl, c)]);
    }
    asm volatile("cp.async.commit_group;\n" ::);
    for (int q = tid; q < kPW * kPW / 2; q += 256) {
        const int r = q / (kPW / 2), c = 2 * (q % (kPW / 2));
        bj_cp_async16(&sR[r * kTLD + c], &Rl[r * kPW + c]);
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 1;\n" ::);
    __syncthreads();
    tile_product(sL, sM, acc);   // T = R_k^T A_kl
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    {
        const int r = warp * 8 + (lane >> 2);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<double2*>(&sM[r * kTLD + j * 8 + 2 * (lane & 3)]) = make_double2(acc[j][0], acc[j][1]);
    }
    __syncthreads();
    tile_product(sM, sR, acc);   // A_kl' = T R_l
    const int r = warp * 8 + (lane >> 2);
    const int gr = pair_index(Ik, Jk, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = j * 8 + 2 * (lane & 3);
        *reinterpret_cast<double2*>(&Ap[(int64_t)gr * Np + pair_index(Il, Jl, c)]) = make_double2(acc[j][0], acc[j][1]);
    }
    if (k == l) return;
    // mirror: A_lk = (A_kl')^T, transposed through shared memory (sL is free: product 1 is complete)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = j * 8 + 2 * (lane & 3);
        sL[r * TLD + c] = acc[j][0];
        sL[r * TLD + c + 1] = acc[j][1];
    }
    __syncthreads();
    for (int q = tid; q < kPW * kPW; q += 256) {
        const int rr = q / kPW, cc = q % kPW;   // element (rr, cc) of A_lk = element (cc, rr) of A_kl'
        Ap[(int64_t)pair_index(Il, Jl, rr) * Np + pair_index(Ik, Jk, cc)] = sL[cc * TLD + rr];
    }
}

// norms[0] += sum of squared off-diagonal entries, norms[1] += squared diagonal
__global__ void __launch_bounds__(256)
bj_norms_kernel(int Np, const double* __restrict__ Ap, double* __restrict__ norms,
                const int32_t* __restrict__ mdone) {
    __shared__ double red[8];
    if (mdone[blockIdx.y]) return;
    Ap += (int64_t)blockIdx.y * Np * Np;
    norms += 2 * blockIdx.y;
    double off = 0.0, dg = 0.0;
    const int64_t total = (int64_t)Np * Np;
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < total; q += (int64_t)gridDim.x * 256) {
        const double v = Ap[q];
        if (q / Np == q % Np) dg += v * v; else off += v * v;
    }
    off = bj_block_sum(off, red);
    dg = bj_block_sum(dg, red);
    if (threadIdx.x == 0) { atomicAdd(&norms[0], off); atomicAdd(&norms[1], dg); }
}

// rank of each of the first N diagonal entries (ascending, ties by index)
__global__ void __launch_bounds__(256)
bj_rank_kernel(int N, int Np, const double* __restrict__ Ap, int32_t* __restrict__ rank,
               double* __restrict__ eigval) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Ap += (int64_t)blockIdx.y * Np * Np;
    rank += (int64_t)blockIdx.y * N;
    eigval += (int64_t)blockIdx.y * N;
    const double v = Ap[(int64_t)i * Np + i];
    int r = 0;
    for (int k = 0; k < N; ++k) {
        const double u = Ap[(int64_t)k * Np + k];
        r += (u < v) || (u == v && k < i);
    }
    rank[i] = r;
    eigval[r] = v;
}

// modes[rank[c]][r] = V[r][c]   (32x32 transpose tiles)
__global__ void __launch_bounds__(256)
bj_export_kernel(int N, int Np, const double* __restrict__ V, const int32_t* __restrict__ rank,
                 double* __restrict__ modes) {
    __shared__ double tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    V += (int64_t)blockIdx.z * Np * Np;
    rank += (int64_t)blockIdx.z * N;
    modes += (int64_t)blockIdx.z * N * N;
    for (int rr = ty; rr < 32; rr += 8) {
        const int r = r0 + rr, c = c0 + tx;
        tile[rr][tx] = (r < N && c < N) ? V[(int64_t)r * Np + c] : 0.0;
    }
    __syncthreads();
    for (int cc = ty; cc < 32; cc += 8) {
        const int c = c0 + cc, r = r0 + tx;
        if (c < N && r < N) modes[(int64_t)rank[c] * N + r] = tile[tx][cc];
    }
}

static int padded_order(int N) {
    int nb = (int)ceil_div(N, kBW);
    if (nb % 2) nb += 1;
    if (nb < 2) nb = 2;
    return nb * kBW;
}

// mdone[g] = 1 once the off-diagonal norm of matrix g is negligible (or g is beyond the group)
__global__ void bj_flag_kernel(int G, int live, const double* __restrict__ norms, int32_t* __restrict__ mdone) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    if (g >= live) { mdone[g] = 1; return; }
    if (mdone[g]) return;
    const double off = norms[2 * g], dg = norms[2 * g + 1];
    mdone[g] = (off <= 1e-28 * (off + dg)) ? 1 : 0;
}

struct BjWork {
    double *Ap, *V, *R, *norms;
    int32_t *active, *rank, *mdone;
};

static size_t bj_bytes_per_matrix(int N) {
    const size_t Np = (size_t)padded_order(N);
    const size_t nb = Np / kBW;
    return sizeof(double) * (2 * Np * Np + (nb / 2) * 2 * kPW * kPW + 2) + sizeof(int32_t) * (nb / 2 + (size_t)N + 1);
}

// matrices of a batch that share the launches: as many as fit ~2 GB of workspace, at most 32
static int bj_group(int B, int N) {
    size_t g = ((size_t)2 << 30) / bj_bytes_per_matrix(N);
    if (g < 1) g = 1;
    if (g > 32) g = 32;
    return (int)(g < (size_t)B ? g : (size_t)B);
}

static void bj_carve(Arena& ar, BjWork* w, int N, int G) {
    const int Np = padded_order(N);
    const int nb = Np / kBW;
    w->Ap = ar.take<double>((size_t)G * Np * Np);
    w->V = ar.take<double>((size_t)G * Np * Np);
    w->R = ar.take<double>((size_t)G * (nb / 2) * 2 * kPW * kPW);
    w->norms = ar.take<double>((size_t)2 * G);
    w->active = ar.take<int32_t>((size_t)G * (nb / 2));
    w->rank = ar.take<int32_t>((size_t)G * N);
    w->mdone = ar.take<int32_t>((size_t)G);
}

size_t eig_full_block_workspace_bytes(int B, int N) {
    Arena ar(nullptr, 0);
    BjWork w;
    bj_carve(ar, &w, N, bj_group(B, N));
    return ar.off + 256;
}

int eig_full_block(int B, int N, double* A, double* eigval, double* modes, void* workspace, size_t workspace_bytes,
                   cudaStream_t st) {
    const int G = bj_group(B, N);
    Arena ar(workspace, workspace_bytes);
    BjWork w;
    bj_carve(ar, &w, N, G);
    if (!ar.ok()) return SCB_ERR_WORKSPACE;
    const int Np = padded_order(N);
    const int nb = Np / kBW;
    const int npairs = nb / 2;
    const size_t smem_pivot = sizeof(double) * 2 * kPW * (kPW + 1);
    const size_t smem_tile = sizeof(double) * 3 * kPW * kTLD;
    // per-device function attribute: set on every launch (cheap), no process-wide "configured" flag
    SCB_CUDA(cudaFuncSetAttribute(bj_pivot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pivot));
    SCB_CUDA(cudaFuncSetAttribute(bj_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tile));
    int inner_sweeps = 1;
    if (const char* env = getenv("SCB_BJ_INNER")) inner_sweeps = atoi(env) > 0 ? atoi(env) : inner_sweeps;
    int cross = 1;
    if (const char* env = getenv("SCB_BJ_CROSS")) cross = atoi(env);
    int32_t* h_done = nullptr;
    SCB_CUDA(cudaMallocHost(&h_done, G * sizeof(int32_t)));
    int status = SCB_OK;
    for (int s0 = 0; s0 < B && status == SCB_OK; s0 += G) {
        const int live = (B - s0 < G) ? B - s0 : G;   // matrices in this group
        bj_init_kernel<<<dim3((unsigned)ceil_div((int64_t)Np * Np, 256), (unsigned)live), 256, 0, st>>>(
            N, Np, A + (int64_t)s0 * N * N, w.Ap, w.V);
        SCB_CUDA(cudaMemsetAsync(w.mdone, 0, G * sizeof(int32_t), st));
        count_launches(3);  // init + rank + export
        bool converged = false;
        for (int sweep = 0; sweep < 60 && !converged; ++sweep) {
            for (int step = 0; step < nb - 1; ++step) {
                // step 0 of a sweep rotates every pair inside its 64x64 pivots (this covers the pairs inside each
                // diagonal block once per sweep); the other steps only rotate pairs ACROSS the two blocks
                bj_pivot_kernel<<<dim3(npairs, live), kPivotThreads, smem_pivot, st>>>(
                    Np, nb, step, w.Ap, w.R, w.active, inner_sweeps, (cross && step > 0) ? 1 : 0, w.mdone);
                bj_update_kernel<<<dim3(npairs * (npairs + 1) / 2 + (Np / kPW) * npairs, live), 256, smem_tile, st>>>(
                    Np, nb, step, w.Ap, w.V, w.R, w.active, w.mdone);
                count_launches(2);
            }
            cudaMemsetAsync(w.norms, 0, 2 * G * sizeof(double), st);
            const int norm_ctas = (4 * kNumSM + live - 1) / live;
            bj_norms_kernel<<<dim3(norm_ctas, live), 256, 0, st>>>(Np, w.Ap, w.norms, w.mdone);
            bj_flag_kernel<<<1, 32, 0, st>>>(G, live, w.norms, w.mdone);
            count_launches(2);
            if (cudaMemcpyAsync(h_done, w.mdone, G * sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaStreamSynchronize(st) != cudaSuccess) {
                set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__);
                status = SCB_ERR_CUDA;
                break;
            }
            converged = true;
            for (int g = 0; g < live; ++g) converged = converged && h_done[g] != 0;
        }
        if (status != SCB_OK) break;
        if (!converged) { status = SCB_ERR_NOT_CONVERGED; break; }
        bj_rank_kernel<<<dim3((unsigned)ceil_div(N, 256), (unsigned)live), 256, 0, st>>>(N, Np, w.Ap, w.rank,
                                                                                         eigval + (int64_t)s0 * N);
        bj_export_kernel<<<dim3((unsigned)ceil_div(N, 32), (unsigned)ceil_div(N, 32), (unsigned)live), 256, 0, st>>>(
            N, Np, w.V, w.rank, modes + (int64_t)s0 * N * N);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_last_cuda_error(e, __FILE__, __LINE__); status = SCB_ERR_CUDA; }
    }
    cudaFreeHost(h_done);
    return status;
}

}  // namespace scb
