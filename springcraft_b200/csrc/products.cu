// K4: fluctuation and covariance products built on the eigenpairs
// (nma.py:108-359, 422-473; anm.py:132-136).  Everything is a contraction
//     out = (U / lambda) . U^T
// over a set of modes.  MSF and the linear response are HBM-bound passes over
// the mode matrix; DCC and covariance are true dense contractions and run on
// the FP64 tensor cores (warp-level DMMA, mma.sync.m8n8k4.f64 -- tcgen05 has no
// FP64 kind).
#include "common.cuh"

namespace scb {

// ---------------------------------------------------------------------------
// MSF
// ---------------------------------------------------------------------------
// modes[B][m][N] (rows = modes), lam[B][m]
template <int D>
__global__ void __launch_bounds__(256)
msf_rows_kernel(int B, int n, int m, const double* __restrict__ lam, const double* __restrict__ modes,
                double scale, double* __restrict__ msf) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (int64_t)B * n) return;
    const int64_t s = q / n;
    const int i = (int)(q % n);
    const int64_t N = (int64_t)D * n;
    const double* M = modes + s * m * N;
    const double* L = lam + s * m;
    double acc = 0.0;
    for (int k = 0; k < m; ++k) {
        const double* v = M + k * N + (int64_t)D * i;
        double sq = __dmul_rn(v[0], v[0]);
        if (D == 3) sq = __dadd_rn(__dadd_rn(sq, __dmul_rn(v[1], v[1])), __dmul_rn(v[2], v[2]));
        acc = __dadd_rn(acc, sq / L[k]);  // nma.py:167-174, summed over modes in order
    }
    msf[q] = __dmul_rn(acc, scale);
}

// eigensolver layout X[B][N][b], modes = columns k0..k0+m-1
template <int D>
__global__ void __launch_bounds__(256)
msf_cols_kernel(int B, int n, int b, int k0, int m, const double* __restrict__ eigval,
                const double* __restrict__ X, double scale, double* __restrict__ msf) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (int64_t)B * n) return;
    const int64_t s = q / n;
    const int i = (int)(q % n);
    const int64_t N = (int64_t)D * n;
    const double* Xs = X + s * N * b + (int64_t)D * i * b;
    const double* L = eigval + s * b;
    double acc = 0.0;
    for (int k = k0; k < k0 + m; ++k) {
        double sq = __dmul_rn(Xs[k], Xs[k]);
        if (D == 3) sq = __dadd_rn(__dadd_rn(sq, __dmul_rn(Xs[b + k], Xs[b + k])), __dmul_rn(Xs[2 * b + k], Xs[2 * b + k]));
        acc = __dadd_rn(acc, sq / L[k]);
    }
    msf[q] = __dmul_rn(acc, scale);
}

// modes[s][k][r] = X[s][r][k0+k]   (32x32 smem transpose tiles)
__global__ void __launch_bounds__(256)
export_modes_kernel(int64_t N, int b, int k0, int m, const double* __restrict__ X, double* __restrict__ modes) {
    __shared__ double tile[32][33];
    const int s = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const double* Xs = X + (int64_t)s * N * b;
    double* Ms = modes + (int64_t)s * m * N;
    for (int rr = ty; rr < 32; rr += 8) {
        const int64_t r = r0 + rr;
        const int c = c0 + tx;
        tile[rr][tx] = (r < N && c < m) ? Xs[r * b + k0 + c] : 0.0;
    }
    __syncthreads();
    for (int cc = ty; cc < 32; cc += 8) {
        const int c = c0 + cc;
        const int64_t r = r0 + tx;
        if (c < m && r < N) Ms[(int64_t)c * N + r] = tile[tx][cc];
    }
}

// ---------------------------------------------------------------------------
// linear response  out = sum_k u_k (u_k . f) / lam_k
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lr_project_kernel(int64_t N, int m, const double* __restrict__ lam, const double* __restrict__ modes,
                  const double* __restrict__ f, double* __restrict__ t) {
    __shared__ double red[8];
    const int k = blockIdx.x;
    const double* u = modes + (int64_t)k * N;
    double acc = 0.0;
    for (int64_t r = threadIdx.x; r < N; r += 256) acc = fma(u[r], f[r], acc);
    acc = warp_sum(acc);
    if (lane_id() == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w];
        t[k] = v / lam[k];
    }
}

__global__ void __launch_bounds__(256)
lr_expand_kernel(int64_t N, int m, const double* __restrict__ modes, const double* __restrict__ t,
                 double* __restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    double acc = 0.0;
    for (int k = 0; k < m; ++k) acc = fma(modes[(int64_t)k * N + r], t[k], acc);
    out[r] = acc;
}

// ---------------------------------------------------------------------------
// DMMA GEMM:  C[M x Nc] = A[M x K] * Bt[Nc x K]^T  (+ fused DCC epilogue)
// ---------------------------------------------------------------------------
// operand preparation: P[i][k*D + a] = modes[k][D*i + a] * (inv ? 1/lam[k] : 1)
__global__ void __launch_bounds__(256)
dcc_prepare_kernel(int D, int n, int m, const double* __restrict__ lam, const double* __restrict__ modes,
                   double* __restrict__ Aop, double* __restrict__ Bop, int64_t ldk) {
    // 32x32 transpose tiles over (k, r) with r = D*i + a
    __shared__ double tile[32][33];
    const int64_t N = (int64_t)D * n;
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int kk = ty; kk < 32; kk += 8) {
        const int k = k0 + kk;
        const int64_t r = r0 + tx;
        tile[kk][tx] = (k < m && r < N) ? modes[(int64_t)k * N + r] : 0.0;
    }
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) {
        const int64_t r = r0 + rr;
        const int k = k0 + tx;
        if (r < N && k < m) {
            const int64_t i = r / D;
            const int a = (int)(r % D);
            const double v = tile[tx][rr];
            Bop[i * ldk + (int64_t)k * D + a] = v;
            Aop[i * ldk + (int64_t)k * D + a] = v / lam[k];
        }
    }
}

// d[i] = sum_kappa A[i][kappa] * B[i][kappa]   (diagonal of the product)
__global__ void __launch_bounds__(256)
rowdot_kernel(int n, int64_t K, int64_t ldk, const double* __restrict__ A, const double* __restrict__ Bm,
              double* __restrict__ d) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    double acc = 0.0;
    for (int64_t q = lane_id(); q < K; q += 32) acc = fma(A[i * ldk + q], Bm[i * ldk + q], acc);
    acc = warp_sum(acc);
    if (lane_id() == 0) d[i] = acc;
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(bytes));
}

constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 16, kGemmLD = 20;  // LD padded: conflict-free fragments

// A, Bt row-major with leading dimension ldk (K contiguous, ldk % 2 == 0, 16-byte aligned rows).
// C rows [row0, row0+Mrows) of the full product are written to out[(i-row0)*ldc + j].
// dnorm != nullptr: out = scale * C / sqrt(d_i d_j)   (nma.py:350-357) else out = scale * C.
__global__ void __launch_bounds__(256)
gemm_nt_dmma_kernel(int Mrows, int Nc, int64_t K, int64_t ldk, int row0, const double* __restrict__ A,
                    const double* __restrict__ Bt, const double* __restrict__ dnorm, double scale,
                    double* __restrict__ out, int64_t ldc) {
    extern __shared__ __align__(16) double gemm_smem[];
    double (*sA)[kGemmBM * kGemmLD] = reinterpret_cast<double (*)[kGemmBM * kGemmLD]>(gemm_smem);
    double (*sB)[kGemmBN * kGemmLD] =
        reinterpret_cast<double (*)[kGemmBN * kGemmLD]>(gemm_smem + 2 * kGemmBM * kGemmLD);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;  // 4 x 2 warps -> warp tile 32 x 64
    const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int stage, int64_t kk) {
        // 128 rows x 16 doubles = 1024 16-byte chunks per operand; 4 per thread
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = tid + q * 256;
            const int r = c >> 3, kc = (c & 7) * 2;
            const int64_t kg = kk + kc;
            {
                const int gi = row0 + m0 + r;
                const bool ok = (m0 + r) < Mrows && kg < K;
                cp_async16(&sA[stage][r * kGemmLD + kc], ok ? (const void*)(A + (int64_t)gi * ldk + kg) : (const void*)A, ok);
            }
            {
                const int gj = n0 + r;
                const bool ok = gj < Nc && kg < K;
                cp_async16(&sB[stage][r * kGemmLD + kc], ok ? (const void*)(Bt + (int64_t)gj * ldk + kg) : (const void*)Bt, ok);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };

    const int64_t nk = ceil_div(K, (int64_t)kGemmBK);
    load_stage(0, 0);
    for (int64_t it = 0; it < nk; ++it) {
        const int stage = (int)(it & 1);
        if (it + 1 < nk) {
            load_stage(stage ^ 1, (it + 1) * kGemmBK);
            asm volatile("cp.async.wait_group 1;\n" ::);
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
        const double* a_s = &sA[stage][(wm * 32) * kGemmLD];
        const double* b_s = &sB[stage][(wn * 64) * kGemmLD];
#pragma unroll
        for (int k4 = 0; k4 < kGemmBK; k4 += 4) {
            double af[4], bf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = a_s[(i * 8 + (lane >> 2)) * kGemmLD + k4 + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 8; ++j) bf[j] = b_s[(j * 8 + (lane >> 2)) * kGemmLD + k4 + (lane & 3)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
    // epilogue: C fragment (row = lane/4, cols = 2*(lane%4) + {0,1})
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = m0 + wm * 32 + i * 8 + (lane >> 2);
        if (r >= Mrows) continue;
        const double di = dnorm ? dnorm[row0 + r] : 1.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = n0 + wn * 64 + j * 8 + 2 * (lane & 3);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (c + e >= Nc) continue;
                double v = acc[i][j][e];
                if (dnorm) v = v / sqrt(di * dnorm[c + e]);
                out[(int64_t)r * ldc + c + e] = v * scale;
            }
        }
    }
}

// perturbation response scanning (nma.py:511-531): out[i][j] = sum_{a,b} cov[3i+a][3j+b]^2, optionally / out[i][i]
__global__ void __launch_bounds__(256)
prs_kernel(int n, const double* __restrict__ cov, int norm, double* __restrict__ out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (int64_t)n * n) return;
    const int i = (int)(q / n), j = (int)(q % n);
    const int64_t N = 3 * (int64_t)n;
    auto block_sq = [&](int r, int c) {
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const double v = cov[(3 * (int64_t)r + a) * N + 3 * (int64_t)c + b];
                acc += v * v;
            }
        return acc;
    };
    double v = block_sq(i, j);
    if (norm) v /= block_sq(i, i);
    out[q] = v;
}

// normal-mode trajectory (nma.py:402-419): disp[f][i][a] = shape(f) * amplitude * u[3i+a] / max_i |u_i|
__global__ void __launch_bounds__(256)
normal_mode_kernel(int n, int frames, const double* __restrict__ mode, double amplitude, int triangle,
                   double* __restrict__ out) {
    __shared__ double red[8];
    __shared__ double smax;
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double x = mode[3 * i], y = mode[3 * i + 1], z = mode[3 * i + 2];
        m = fmax(m, sqrt(x * x + y * y + z * z));
    }
    m = warp_max(m);
    if (lane_id() == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t = fmax(t, red[w]);
        smax = t;
    }
    __syncthreads();
    const double scale = amplitude / smax;
    const double kTwoPi = 6.283185307179586476925286766559;
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < (int64_t)frames * 3 * n; q += (int64_t)gridDim.x * 256) {
        const int f = (int)(q / (3 * n));
        const double time = (double)f / frames;
        double shape;
        if (triangle) shape = 2.0 * fabs(2.0 * (time - floor(time + 0.5))) - 1.0;
        else shape = sin(time * kTwoPi);
        out[q] = shape * (mode[q % (3 * n)] * scale);
    }
}

static int64_t padded_k(int64_t K) { return (K + 1) & ~int64_t(1); }

static int run_product(int D, int n, int m, const double* lam, const double* modes, int norm, double scale,
                       int row0, int row1, double* out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (!lam || !modes || !out || !workspace || n < 1 || m < 1 || row0 < 0 || row1 > n || row0 >= row1)
        return SCB_ERR_INVALID;
    const int64_t K = (int64_t)D * m, ldk = padded_k(K);
    Arena ar(workspace, workspace_bytes);
    double* Aop = ar.take<double>((size_t)n * ldk);
    double* Bop = ar.take<double>((size_t)n * ldk);
    double* dn = ar.take<double>(n);
    if (!ar.ok()) return SCB_ERR_WORKSPACE;
    if (ldk != K) {
        SCB_CUDA(cudaMemsetAsync(Aop, 0, sizeof(double) * (size_t)n * ldk, st));
        SCB_CUDA(cudaMemsetAsync(Bop, 0, sizeof(double) * (size_t)n * ldk, st));
    }
    const int64_t N = (int64_t)D * n;
    dim3 pg((unsigned)ceil_div(N, 32), (unsigned)ceil_div(m, 32));
    dcc_prepare_kernel<<<pg, 256, 0, st>>>(D, n, m, lam, modes, Aop, Bop, ldk);
    SCB_LAUNCH_CHECK();
    if (norm) {
        rowdot_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(n, K, ldk, Aop, Bop, dn);
        SCB_LAUNCH_CHECK();
    }
    const int Mrows = row1 - row0;
    dim3 grid((unsigned)ceil_div(n, kGemmBN), (unsigned)ceil_div(Mrows, kGemmBM));
    const size_t smem = sizeof(double) * 2 * (kGemmBM + kGemmBN) * kGemmLD;
    // per-device function attribute: set on every launch (cheap), no process-wide "configured" flag
    SCB_CUDA(cudaFuncSetAttribute(gemm_nt_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_nt_dmma_kernel<<<grid, 256, smem, st>>>(Mrows, n, K, ldk, row0, Aop, Bop, norm ? dn : nullptr, scale, out, n);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

}  // namespace scb

using namespace scb;

extern "C" int scb_msf(int D, int B, int n, int m, const double* lam, const double* modes, double scale,
                       double* msf, void* stream) {
    if (!lam || !modes || !msf || B < 1 || n < 1 || m < 1) return SCB_ERR_INVALID;
    const unsigned grid = (unsigned)ceil_div((int64_t)B * n, 256);
    if (D == 1) msf_rows_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(B, n, m, lam, modes, scale, msf);
    else if (D == 3) msf_rows_kernel<3><<<grid, 256, 0, as_stream(stream)>>>(B, n, m, lam, modes, scale, msf);
    else return SCB_ERR_INVALID;
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_msf_cols(int D, int B, int n, int b, int k0, int m, const double* eigval, const double* X,
                            double scale, double* msf, void* stream) {
    if (!eigval || !X || !msf || B < 1 || n < 1 || m < 1 || k0 < 0 || k0 + m > b) return SCB_ERR_INVALID;
    const unsigned grid = (unsigned)ceil_div((int64_t)B * n, 256);
    if (D == 1) msf_cols_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(B, n, b, k0, m, eigval, X, scale, msf);
    else if (D == 3) msf_cols_kernel<3><<<grid, 256, 0, as_stream(stream)>>>(B, n, b, k0, m, eigval, X, scale, msf);
    else return SCB_ERR_INVALID;
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_export_modes(int B, int N, int b, int k0, int m, const double* X, double* modes, void* stream) {
    if (!X || !modes || B < 1 || N < 1 || m < 1 || k0 < 0 || k0 + m > b) return SCB_ERR_INVALID;
    dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(m, 32), (unsigned)B);
    export_modes_kernel<<<grid, 256, 0, as_stream(stream)>>>(N, b, k0, m, X, modes);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" size_t scb_dcc_workspace_bytes(int D, int n, int m) {
    const int64_t ldk = padded_k((int64_t)D * m);
    return 2 * (((size_t)n * ldk * sizeof(double) + 255) & ~size_t(255)) + (((size_t)n * sizeof(double) + 255) & ~size_t(255)) + 256;
}

extern "C" int scb_dcc(int D, int n, int m, const double* lam, const double* modes, int norm, double scale,
                       int row0, int row1, double* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (D != 1 && D != 3) return SCB_ERR_INVALID;
    return run_product(D, n, m, lam, modes, norm, scale, row0, row1, out, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int scb_covariance(int N, int m, const double* lam, const double* modes, int row0, int row1,
                              double* out, void* workspace, size_t workspace_bytes, void* stream) {
    return run_product(1, N, m, lam, modes, 0, 1.0, row0, row1, out, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int scb_linear_response(int N, int m, const double* lam, const double* modes, const double* force,
                                   double* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!lam || !modes || !force || !out || !workspace || N < 1 || m < 1) return SCB_ERR_INVALID;
    if (workspace_bytes < sizeof(double) * (size_t)m) return SCB_ERR_WORKSPACE;
    double* t = static_cast<double*>(workspace);
    cudaStream_t st = as_stream(stream);
    lr_project_kernel<<<m, 256, 0, st>>>(N, m, lam, modes, force, t);
    SCB_LAUNCH_CHECK();
    lr_expand_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(N, m, modes, t, out);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

// ---- products of a covariance matrix that the CALLER assigned (anm.covariance = C, anm.py:138-148): the reference
// then serves dcc() and linear_response() from that matrix (nma.py:324-336, 473), not from an eigendecomposition
namespace scb {
// out[i][j] = scale * sum_a C[D i + a][D j + a]  (nma.py:329-336); raw (un-normalised) values
__global__ void __launch_bounds__(256)
dcc_from_cov_kernel(int D, int n, const double* __restrict__ cov, double scale, double* __restrict__ out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (int64_t)n * n) return;
    const int i = (int)(q / n), j = (int)(q % n);
    const int64_t N = (int64_t)D * n;
    double t = 0.0;
    for (int a = 0; a < D; ++a) t += cov[((int64_t)D * i + a) * N + (int64_t)D * j + a];
    out[q] = scale * t;
}
// out[i][j] /= sqrt(d_i d_j) with d = the diagonal BEFORE scaling by `scale` (normalisation first, nma.py:350-357)
__global__ void __launch_bounds__(256)
dcc_normalise_kernel(int n, const double* __restrict__ diag, double scale, double* __restrict__ out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (int64_t)n * n) return;
    const int i = (int)(q / n), j = (int)(q % n);
    out[q] = out[q] / sqrt((diag[i] / scale) * (diag[j] / scale));
}
__global__ void __launch_bounds__(256)
copy_diag_kernel(int n, const double* __restrict__ out, double* __restrict__ diag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) diag[i] = out[(int64_t)i * n + i];
}
// y = C f, one warp per row (C symmetric, row-major)
__global__ void __launch_bounds__(256)
symv_rows_kernel(int N, const double* __restrict__ cov, const double* __restrict__ f, double* __restrict__ y) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= N) return;
    double acc = 0.0;
    for (int c = lane_id(); c < N; c += 32) acc = fma(cov[(int64_t)row * N + c], f[c], acc);
    acc = warp_sum(acc);
    if (lane_id() == 0) y[row] = acc;
}
}  // namespace scb

extern "C" int scb_dcc_from_covariance(int D, int n, const double* cov, int norm, double scale, double* out,
                                       double* diag_scratch, void* stream) {
    if (!cov || !out || n < 1 || (D != 1 && D != 3) || (norm && !diag_scratch) || scale == 0.0) return SCB_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    const unsigned grid = (unsigned)ceil_div((int64_t)n * n, 256);
    scb::dcc_from_cov_kernel<<<grid, 256, 0, st>>>(D, n, cov, scale, out);
    SCB_LAUNCH_CHECK();
    if (norm) {
        scb::copy_diag_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(n, out, diag_scratch);
        SCB_LAUNCH_CHECK();
        scb::dcc_normalise_kernel<<<grid, 256, 0, st>>>(n, diag_scratch, scale, out);
        SCB_LAUNCH_CHECK();
    }
    return SCB_OK;
}

extern "C" int scb_symv(int N, const double* cov, const double* f, double* y, void* stream) {
    if (!cov || !f || !y || N < 1) return SCB_ERR_INVALID;
    scb::symv_rows_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, as_stream(stream)>>>(N, cov, f, y);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_prs(int n, const double* cov, int norm, double* out, void* stream) {
    if (!cov || !out || n < 1) return SCB_ERR_INVALID;
    prs_kernel<<<(unsigned)ceil_div((int64_t)n * n, 256), 256, 0, as_stream(stream)>>>(n, cov, norm, out);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_normal_mode(int n, int frames, const double* mode, double amplitude, int triangle, double* out,
                               void* stream) {
    if (!mode || !out || n < 1 || frames < 1) return SCB_ERR_INVALID;
    const int64_t total = (int64_t)frames * 3 * n;
    const unsigned grid = (unsigned)(ceil_div(total, 256) < 1024 ? ceil_div(total, 256) : 1024);
    normal_mode_kernel<<<grid, 256, 0, as_stream(stream)>>>(n, frames, mode, amplitude, triangle, out);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}
