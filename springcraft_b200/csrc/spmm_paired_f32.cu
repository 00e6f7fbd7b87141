// K3a'': single-precision twin of the row-paired SpMM, used for the EARLY outer iterations of the
// Chebyshev-filtered subspace iteration.
//
// The filter only has to enrich the block in the wanted directions; Rayleigh-Ritz, residuals and the
// final iterations stay in FP64.  A NumPy prototype on the C3 matrices shows identical outer-iteration
// counts when the first 5 of 8-9 filters run in FP32 (the FP32 filter stalls at a residual of ~2e-5,
// the solver switches a structure to the FP64 kernel once its residual is below 3e-6 * lambda_max (or stops halving), and the
// converged results are bit-for-bit products of FP64 iterations).  FP32 halves the record and X-row
// bytes (80-byte records, 128-byte X rows = one LDG.128 per lane and row), halves the registers
// (32 warps per SM instead of 16) and runs on the 2x wider FP32 pipe.
#include "paired.cuh"

namespace scb {

constexpr int kPair32Warps = 32;

// records: double -> float (whole capacity range; unused slots are never read)
template <int D>
__global__ void __launch_bounds__(256)
pair32_convert_kernel(int64_t cap, const PairEntry<D>* __restrict__ in, PairEntry32<D>* __restrict__ out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= cap) return;
    PairEntry32<D> e;
#pragma unroll
    for (int i = 0; i < 2 * D * D; ++i) e.blk[i] = (float)in[q].blk[i];
    e.col = in[q].col;
    e.pad = 0;
    out[q] = e;
}

// block vectors: per-structure precision conversion (skip[s] != 0 -> untouched)
__global__ void __launch_bounds__(256)
block_to_f32_kernel(int64_t per_struct, const double* __restrict__ in, float* __restrict__ out,
                    const int32_t* __restrict__ skip) {
    const int64_t s = blockIdx.y;
    if (skip && skip[s]) return;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= per_struct / 2) return;
    const double2 v = reinterpret_cast<const double2*>(in + s * per_struct)[q];
    reinterpret_cast<float2*>(out + s * per_struct)[q] = make_float2((float)v.x, (float)v.y);
}
__global__ void __launch_bounds__(256)
block_to_f64_kernel(int64_t per_struct, const float* __restrict__ in, double* __restrict__ out,
                    const int32_t* __restrict__ skip) {
    const int64_t s = blockIdx.y;
    if (skip && skip[s]) return;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= per_struct / 2) return;
    const float2 v = reinterpret_cast<const float2*>(in + s * per_struct)[q];
    reinterpret_cast<double2*>(out + s * per_struct)[q] = make_double2((double)v.x, (double)v.y);
}

template <int D, int BC>
__global__ void __launch_bounds__(kPair32Warps * 32, 1)
spmm_paired_f32_kernel(int n, int np, int pairs_per_cta, const int64_t* __restrict__ rowptr,
                       const int32_t* __restrict__ pcount, const PairEntry32<D>* __restrict__ pent,
                       const float* __restrict__ X, const float* __restrict__ W, float* __restrict__ Y,
                       const double* __restrict__ coef, int coef_stride, const int32_t* __restrict__ skip) {
    constexpr int R = 2 * D;
    using Entry = PairEntry32<D>;
    extern __shared__ __align__(128) unsigned char pair_smem[];
    Entry* stage_base = reinterpret_cast<Entry*>(pair_smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + kPair32Warps * 2 * kPairChunk);
    const int64_t s = blockIdx.y;
    if (skip && skip[s]) return;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int slot = lane >> 3, l8 = lane & 7;
    Entry* stage = stage_base + (size_t)warp * 2 * kPairChunk;
    uint64_t* bar = bars + 2 * warp;
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async;\n" ::: "memory");
    }
    __syncwarp();
    unsigned phase0 = 0, phase1 = 0;
    const int64_t N = (int64_t)D * n;
    float alpha = 1.f, cshift = 0.f, beta = 0.f;
    if (coef) {
        const double* cf = coef + s * coef_stride;
        alpha = (float)cf[0]; cshift = (float)cf[1]; beta = (float)cf[2];
    }
    const int t0 = blockIdx.x * pairs_per_cta;
    const int t1 = min(np, t0 + pairs_per_cta);
    constexpr int b = BC;               // compile-time: row offsets become immediates
    constexpr int ncg = BC >> 5;
    constexpr int rowlen = D * BC;
    if (coef && alpha == 0.f) {
        // idle filter step (see spmm_paired.cu): Y = X
        for (int t = t0 + warp; t < t1; t += kPair32Warps)
            for (int a = 0; a < R; ++a) {
                const int64_t r = (int64_t)D * 2 * t + a;
                if (r >= N) continue;
                const float* src = X + (s * N + r) * b;
                float* dst = Y + (s * N + r) * b;
                for (int q = lane; q < b; q += 32) dst[q] = src[q];
            }
        return;
    }
    for (int t = t0 + warp; t < t1; t += kPair32Warps) {
        const int64_t g = s * np + t;
        const int64_t base = rowptr[s * n + 2 * t] + 2 * g;
        const int cnt = pcount[g];
        const int nchunk = (cnt + kPairChunk - 1) / kPairChunk;
        for (int cg = 0; cg < ncg; ++cg) {
            const int c0 = cg * 32 + 4 * l8;  // this lane's 4 consecutive columns (16 bytes)
            const float* Xs = X + s * N * b + c0;
            // packed accumulators: acc2[a][0] = columns (c0, c0+1), acc2[a][1] = (c0+2, c0+3); FFMA2 (fma.rn.f32x2)
            float2 acc2[R][2];
#pragma unroll
            for (int a = 0; a < R; ++a) { acc2[a][0] = make_float2(0.f, 0.f); acc2[a][1] = make_float2(0.f, 0.f); }
            __syncwarp();
            if (nchunk > 0 && lane == 0) {
                const unsigned bytes = (unsigned)(min(kPairChunk, cnt) * sizeof(Entry));
                mbar_expect_tx(&bar[0], bytes);
                bulk_g2s(stage, pent + base, bytes, &bar[0]);
            }
            for (int ch = 0; ch < nchunk; ++ch) {
                const int e0 = ch * kPairChunk;
                const int m = min(kPairChunk, cnt - e0);
                const int buf = ch & 1;
                if (ch + 1 < nchunk && lane == 0) {
                    const unsigned bytes = (unsigned)(min(kPairChunk, cnt - e0 - kPairChunk) * sizeof(Entry));
                    mbar_expect_tx(&bar[buf ^ 1], bytes);
                    bulk_g2s(stage + (buf ^ 1) * kPairChunk, pent + base + e0 + kPairChunk, bytes, &bar[buf ^ 1]);
                }
                if (buf == 0) { mbar_wait(&bar[0], phase0); phase0 ^= 1; }
                else { mbar_wait(&bar[1], phase1); phase1 ^= 1; }
                const Entry* cur = stage + buf * kPairChunk;
#pragma unroll 2
                for (int q = slot; q < m; q += 4) {
                    const float* xr = Xs + cur[q].col * rowlen;  // 32-bit offset inside one structure
                    float2 x2[D][2];
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        const float4 u = __ldg(reinterpret_cast<const float4*>(xr + c * b));
                        x2[c][0] = make_float2(u.x, u.y);
                        x2[c][1] = make_float2(u.z, u.w);
                    }
                    float h[2 * D * D];
                    if (D == 3) {  // 18 floats: 4 x LDS.128 + 1 x LDS.64 (records are 16-byte aligned)
                        const float4* bp = reinterpret_cast<const float4*>(cur[q].blk);
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            const float4 u = bp[w];
                            h[4 * w] = u.x; h[4 * w + 1] = u.y; h[4 * w + 2] = u.z; h[4 * w + 3] = u.w;
                        }
                        const float2 u2 = *reinterpret_cast<const float2*>(cur[q].blk + 16);
                        h[16] = u2.x; h[17] = u2.y;
                    } else {
                        const float2 u2 = *reinterpret_cast<const float2*>(cur[q].blk);
                        h[0] = u2.x; h[1] = u2.y;
                    }
#pragma unroll
                    for (int a = 0; a < R; ++a)
#pragma unroll
                        for (int c = 0; c < D; ++c) {
                            const float2 hh = make_float2(h[a * D + c], h[a * D + c]);
                            acc2[a][0] = __ffma2_rn(hh, x2[c][0], acc2[a][0]);
                            acc2[a][1] = __ffma2_rn(hh, x2[c][1], acc2[a][1]);
                        }
                }
                __syncwarp();
            }
            float acc[R][4];
#pragma unroll
            for (int a = 0; a < R; ++a) {
                acc[a][0] = acc2[a][0].x; acc[a][1] = acc2[a][0].y; acc[a][2] = acc2[a][1].x; acc[a][3] = acc2[a][1].y;
            }
#pragma unroll
            for (int a = 0; a < R; ++a)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    float v = acc[a][cc];
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    acc[a][cc] = v;
                }
#pragma unroll
            for (int a = 0; a < R; ++a) {
                if ((a & 3) != slot) continue;
                const int64_t r = (int64_t)D * 2 * t + a;
                if (r >= N) continue;
                const int64_t idx = (s * N + r) * b + c0;
                float v[4] = {acc[a][0], acc[a][1], acc[a][2], acc[a][3]};
                if (coef) {
                    const float4 xo = *reinterpret_cast<const float4*>(X + idx);
                    v[0] = alpha * (v[0] - cshift * xo.x); v[1] = alpha * (v[1] - cshift * xo.y);
                    v[2] = alpha * (v[2] - cshift * xo.z); v[3] = alpha * (v[3] - cshift * xo.w);
                    if (W && beta != 0.f) {
                        const float4 w4 = *reinterpret_cast<const float4*>(W + idx);
                        v[0] -= beta * w4.x; v[1] -= beta * w4.y; v[2] -= beta * w4.z; v[3] -= beta * w4.w;
                    }
                }
                *reinterpret_cast<float4*>(Y + idx) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
}

size_t paired_entry32_bytes(int D) { return D == 3 ? sizeof(PairEntry32<3>) : sizeof(PairEntry32<1>); }

int build_paired32(int D, size_t cap, const void* pent, void* pent32, cudaStream_t st) {
    const unsigned grid = (unsigned)ceil_div((int64_t)cap, 256);
    if (D == 3)
        pair32_convert_kernel<3><<<grid, 256, 0, st>>>((int64_t)cap, static_cast<const PairEntry<3>*>(pent),
                                                       static_cast<PairEntry32<3>*>(pent32));
    else if (D == 1)
        pair32_convert_kernel<1><<<grid, 256, 0, st>>>((int64_t)cap, static_cast<const PairEntry<1>*>(pent),
                                                       static_cast<PairEntry32<1>*>(pent32));
    else
        return SCB_ERR_INVALID;
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

int block_to_f32(int B, int64_t per_struct, const double* in, float* out, const int32_t* skip, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(per_struct / 2, 256), (unsigned)B);
    block_to_f32_kernel<<<grid, 256, 0, st>>>(per_struct, in, out, skip);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}
int block_to_f64(int B, int64_t per_struct, const float* in, double* out, const int32_t* skip, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(per_struct / 2, 256), (unsigned)B);
    block_to_f64_kernel<<<grid, 256, 0, st>>>(per_struct, in, out, skip);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

template <int D, int BC>
static int launch_paired32(dim3 grid, int n, int np, int per_cta, const int64_t* rowptr,
                           const int32_t* pcount, const void* pent, const float* X, const float* W, float* Y,
                           const double* coef, int coef_stride, const int32_t* skip, cudaStream_t st) {
    const size_t smem = sizeof(PairEntry32<D>) * 2 * kPairChunk * kPair32Warps + sizeof(uint64_t) * 2 * kPair32Warps;
    // per-device function attribute: set on every launch (cheap), no process-wide "configured" flag
    SCB_CUDA(cudaFuncSetAttribute(spmm_paired_f32_kernel<D, BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    spmm_paired_f32_kernel<D, BC><<<grid, kPair32Warps * 32, smem, st>>>(
        n, np, per_cta, rowptr, pcount, static_cast<const PairEntry32<D>*>(pent), X, W, Y, coef, coef_stride, skip);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

int spmm_paired_f32(int D, int B, int n, const int64_t* rowptr, const int32_t* pcount, const void* pent32,
                    const float* X, const float* W, float* Y, int b, const double* coef, int coef_stride,
                    const int32_t* skip, cudaStream_t st) {
    if (b != 32 && b != 64) return SCB_ERR_UNSUPPORTED;
    const int np = (n + 1) / 2;
    const int64_t total = (int64_t)B * np;
    int per_cta = (int)ceil_div(total, 4 * kNumSM);
    per_cta = per_cta < kPair32Warps ? kPair32Warps : (per_cta > 256 ? 256 : per_cta);
    if (per_cta > np) per_cta = np;
    dim3 grid((unsigned)ceil_div(np, per_cta), (unsigned)B);
    if (D == 3 && b == 32) return launch_paired32<3, 32>(grid, n, np, per_cta, rowptr, pcount, pent32, X, W, Y, coef, coef_stride, skip, st);
    if (D == 3 && b == 64) return launch_paired32<3, 64>(grid, n, np, per_cta, rowptr, pcount, pent32, X, W, Y, coef, coef_stride, skip, st);
    if (D == 1 && b == 32) return launch_paired32<1, 32>(grid, n, np, per_cta, rowptr, pcount, pent32, X, W, Y, coef, coef_stride, skip, st);
    if (D == 1 && b == 64) return launch_paired32<1, 64>(grid, n, np, per_cta, rowptr, pcount, pent32, X, W, Y, coef, coef_stride, skip, st);
    return SCB_ERR_INVALID;
}

}  // namespace scb
