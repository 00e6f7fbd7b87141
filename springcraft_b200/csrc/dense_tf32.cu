// Residual-form Chebyshev filter for DENSE all-pairs operators on the 5th-generation tensor
// cores: tcgen05.mma kind::tf32 with TMA-fed shared-memory operands and the accumulator in tensor memory.
//
// The filter of the subspace iteration only has to produce the correction z = q(H) r / p(theta) of a Ritz pair
// (resident.cuh explains the residual form); z is proportional to the residual, so a low-precision operator and
// low-precision iterates perturb the new basis by O(eps_tf32 |r|), not O(eps_tf32 |x|): the FP64 tolerance is still
// reached, with H X, Rayleigh-Ritz and residuals in FP64 (NumPy prototype: scratch/c4_tf32_proto.py, same outer
// iteration count as an FP64 filter at degree 24).  FP64 has no tcgen05 kind; TF32 runs at ~30x the DMMA rate, so
// one filter step becomes a stream of the FP32 slab from HBM.
//
// Per filter step:   Znext^T[n][m] = A_n * ( (H Zcur)[m][n] - c Zcur^T[n][m] + R^T[n][m] ) - B_n * Zprev^T[n][m]
// Blocks are kept TRANSPOSED ([b][ld], the long dimension contiguous): that is the K-major B operand of the MMA and
// makes the epilogue's global accesses coalesced (TMEM lane = matrix row = consecutive addresses across a warp).
//
// Precision: a single TF32 product (10-bit mantissa) is not enough at full size -- with 60,000 closely spaced
// eigenvalues the error of the correction (~eps_tf32 |H| / gap) is O(1) and the solver needs 28 outer iterations
// instead of 9.  The kernel therefore evaluates the 3-term split product  H_hi z_hi + H_hi z_lo + H_lo z_hi  (hi = the
// TF32-representable part, lo = the remainder; FP32-class accuracy, "3xTF32"): two slab streams per step instead of
// one, still ~6x faster than the FP64 product.  SPLIT = 1 keeps the single product (tests, small systems).
//
// Kernel: one CTA per 128-row tile of the slab, 128 x b accumulator (b = 128) in TMEM.
//   warp 0   TMA producer: A tile 128 x 32 floats of the slab + B tile b x 32 floats of Zcur^T per stage, 128B swizzle
//   warp 1   TMEM allocation; one elected lane issues 4 x tcgen05.mma (K = 8 each) per stage, tcgen05.commit frees it
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns per instruction), recurrence, coalesced stores
#include <cuda.h>

#include "common.cuh"

namespace scb {

constexpr int kT32BM = 128;        // rows per CTA (UMMA M)
constexpr int kT32BN = 128;        // block columns (UMMA N)
constexpr int kT32BK = 32;         // floats per stage along K = one 128-byte swizzle row
constexpr int kT32Stages = 3;      // SPLIT 1: 32 KB per stage, 2 CTAs per SM; SPLIT 3: 64 KB per stage, 1 CTA per SM
constexpr int kT32Threads = 192;
constexpr uint32_t kT32TileBytes = kT32BM * kT32BK * 4;   // one 128 x 32 fp32 operand tile (A or B)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mb_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(s32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(s32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(s32(bar)), "r"(c_inner), "r"(c_outer)
        : "memory");
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);      // start address, 16-byte units
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M x N
constexpr uint32_t kT32Idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kT32BN >> 3) << 17) |
                               ((uint32_t)(kT32BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kT32Idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct T32Epilogue {
    // SPLIT 3: every z block is [2 b][ld], rows 0..b-1 the TF32-representable part, rows b..2b-1 the remainder
    const float* zcur;     // [b][ld]  Zcur^T   (NULL: plain product, out = H Z)
    const float* zprev;    // [b][ld]  Zprev^T
    const float* rhat;     // [b][ld]  normalised residual, transposed
    float* out[8];         // [b][ld]  Znext^T (may alias zprev) in the buffers of every rank (peer mapped); out[0] = own
    int nout;              // number of destinations (1 on one GPU)
    const float* cA;       // [b]
    const float* cB;       // [b]
    float cshift;
    int64_t ld;            // leading dimension of the transposed blocks (floats)
    int row0;              // first matrix row of the slab (offset of its rows inside the blocks)
    int rows;              // rows of the slab
};

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

template <int SPLIT>
__global__ void __launch_bounds__(kT32Threads, SPLIT == 1 ? 2 : 1)
dense_slab_tf32_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                       const __grid_constant__ CUtensorMap mapAlo, const __grid_constant__ CUtensorMap mapBlo, int K,
                       T32Epilogue ep) {
    constexpr uint32_t kT32StageBytes = (SPLIT == 1 ? 2 : 4) * kT32TileBytes;
    extern __shared__ __align__(1024) uint8_t t32_smem[];
    __shared__ uint64_t full_bar[kT32Stages], empty_bar[kT32Stages], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kT32BM;
    const int nkb = (K + kT32BK - 1) / kT32BK;
    // shared memory: the runtime only guarantees 16-byte alignment of dynamic smem; align to 1024 by hand
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(t32_smem) + 1023) & ~(uintptr_t)1023);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kT32Stages; ++s) { mb_init(&full_bar[s], 1); mb_init(&empty_bar[s], 1); }
        mb_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_smem)),
                     "r"((uint32_t)kT32BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kT32Stages;
                const uint32_t ph = (kb / kT32Stages) & 1;
                mb_wait(&empty_bar[s], ph ^ 1);
                uint8_t* a = tiles + (size_t)s * kT32StageBytes;
                uint8_t* b = a + kT32TileBytes;
                mb_expect_tx(&full_bar[s], kT32StageBytes);
                tma_load_2d(a, &mapA, &full_bar[s], kb * kT32BK, m0);
                tma_load_2d(b, &mapB, &full_bar[s], kb * kT32BK, 0);
                if (SPLIT == 3) {
                    tma_load_2d(b + kT32TileBytes, &mapAlo, &full_bar[s], kb * kT32BK, m0);
                    tma_load_2d(b + 2 * kT32TileBytes, &mapBlo, &full_bar[s], kb * kT32BK, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kT32Stages;
                const uint32_t ph = (kb / kT32Stages) & 1;
                mb_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = s32(tiles + (size_t)s * kT32StageBytes);
                const uint32_t b_addr = a_addr + kT32TileBytes;
                const uint64_t adesc = umma_desc_k128(a_addr), bdesc = umma_desc_k128(b_addr);
#pragma unroll
                for (int k = 0; k < kT32BK / 8; ++k)   // K = 8 tf32 = 32 bytes per instruction: +2 in 16-byte units
                    umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, (kb > 0 || k > 0) ? 1u : 0u);
                if (SPLIT == 3) {
                    const uint64_t alo = umma_desc_k128(b_addr + kT32TileBytes);
                    const uint64_t blo = umma_desc_k128(b_addr + 2 * kT32TileBytes);
#pragma unroll
                    for (int k = 0; k < kT32BK / 8; ++k) {
                        umma_tf32(tmem_base, adesc + 2 * k, blo + 2 * k, 1u);     // H_hi z_lo
                        umma_tf32(tmem_base, alo + 2 * k, bdesc + 2 * k, 1u);     // H_lo z_hi
                    }
                }
                umma_commit(&empty_bar[s]);            // frees the stage once these MMAs have read it
            }
            umma_commit(&tmem_full_bar);               // accumulator complete
        }
    } else {
        // ===== epilogue: warp w may read TMEM lanes 32 (w % 4) .. + 31
        const int q = warp & 3;
        mb_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int m = m0 + 32 * q + lane;              // row inside the slab
        const bool ok = m < ep.rows;
        const int64_t col = (int64_t)ep.row0 + m;      // position inside the transposed blocks
#pragma unroll 1
        for (int c = 0; c < kT32BN / 32; ++c) {
            float acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + 32 * c, acc);
            if (!ok) continue;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = 32 * c + j;
                const int64_t idx = (int64_t)n * ep.ld + col;
                float v = acc[j];
                const int64_t lo_off = (int64_t)kT32BN * ep.ld;     // SPLIT 3: the remainder rows of a block
                if (ep.zcur) {
                    float zc = ep.zcur[idx], zp = ep.zprev[idx];
                    if (SPLIT == 3) { zc += ep.zcur[idx + lo_off]; zp += ep.zprev[idx + lo_off]; }
                    v = ep.cA[n] * (fmaf(-ep.cshift, zc, v) + ep.rhat[idx]) - ep.cB[n] * zp;
                }
                // the all-gather of the row slabs is fused here: the result goes straight into the block of every rank
                if (SPLIT == 3) {
                    const float hi = tf32_hi(v);
                    const float lo = v - hi;
                    for (int p = 0; p < ep.nout; ++p) {
                        ep.out[p][idx] = hi;
                        ep.out[p][idx + lo_off] = lo;
                    }
                } else {
                    for (int p = 0; p < ep.nout; ++p) ep.out[p][idx] = v;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kT32BN));
    }
}

// ---- helpers around the filter ----------------------------------------------------------------------------
// FP64 row-major slab -> FP32 with a padded leading dimension (TMA needs 16-byte multiples)
// out_lo == NULL: plain cast; else out = the TF32-representable part, out_lo = the remainder (from the FP64 value)
__global__ void __launch_bounds__(256)
slab_to_f32_kernel(int64_t rows, int64_t N, int64_t ld, const double* __restrict__ in, float* __restrict__ out,
                   float* __restrict__ out_lo) {
    const int64_t r = blockIdx.y;
    for (int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x; c < ld; c += (int64_t)gridDim.x * 256) {
        const double v = c < N ? in[r * N + c] : 0.0;
        if (out_lo) {
            const float hi = tf32_hi((float)v);
            out[r * ld + c] = hi;
            out_lo[r * ld + c] = (float)(v - (double)hi);
        } else {
            out[r * ld + c] = (float)v;
        }
    }
}

// residual form, start of a filter: R = HX - X theta' (theta' = min(theta, lo)), normalised per column;
//   rhatT = (R / |r|)^T, z1T = rhatT * rho0 / e, z0T = 0, coefficient tables A_k, B_k (k = 1 .. deg-1)
__global__ void __launch_bounds__(256)
resform_prepare_kernel(int64_t N, int b, int64_t ld, const double* __restrict__ X, const double* __restrict__ HX,
                       const double* __restrict__ theta, const double* __restrict__ rn2, double lo, double ub,
                       float* __restrict__ rhatT, float* __restrict__ z1T, float* __restrict__ z0T, int split) {
    __shared__ float tile[32][33];
    const double ehalf = 0.5 * (ub - lo), cmid = 0.5 * (ub + lo);
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 warps
    for (int i = ty; i < 32; i += 8) {
        const int64_t r = r0 + i;
        const int c = c0 + tx;
        float v = 0.f;
        if (r < N) {
            const double th = fmin(theta[c], lo);
            const double nr = sqrt(fmax(rn2[c], 0.0));
            const double inv = nr > 0.0 ? 1.0 / nr : 0.0;
            v = (float)((HX[r * b + c] - th * X[r * b + c]) * inv);
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i;
        const int64_t r = r0 + tx;
        if (r < ld) {
            const float v = r < N ? tile[tx][i] : 0.f;
            const double x = (fmin(theta[c], lo) - cmid) / ehalf;
            const int64_t idx = (int64_t)c * ld + r;
            rhatT[idx] = v;
            const float z1 = (float)((double)v / (x * ehalf));
            z0T[idx] = 0.f;
            if (split) {                             // [2 b][ld]: TF32-representable part, then the remainder
                const float hi = tf32_hi(z1);
                const int64_t lo_off = (int64_t)b * ld;
                z1T[idx] = hi;
                z1T[idx + lo_off] = z1 - hi;
                z0T[idx + lo_off] = 0.f;
            } else {
                z1T[idx] = z1;
            }
        }
    }
}

__global__ void resform_coef_kernel(int b, int deg, const double* __restrict__ theta, double lo, double ub,
                                    float* __restrict__ cA, float* __restrict__ cB) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= b) return;
    const double ehalf = 0.5 * (ub - lo), cmid = 0.5 * (ub + lo);
    const double x = (fmin(theta[c], lo) - cmid) / ehalf;
    double rho = 1.0 / x;
    for (int k = 1; k < deg; ++k) {
        const double rn = 1.0 / (2.0 * x - rho);
        cA[(int64_t)k * b + c] = (float)(2.0 * rn / ehalf);
        cB[(int64_t)k * b + c] = (float)(rho * rn);
        rho = rn;
    }
}

// X[m][n] += |r_n| * zT[n][m]
__global__ void __launch_bounds__(256)
resform_finish_kernel(int64_t N, int b, int64_t ld, const double* __restrict__ rn2, const float* __restrict__ zT,
                      double* __restrict__ X, int split) {
    __shared__ float tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int64_t r = r0 + tx;
        float z = r < N ? zT[(int64_t)(c0 + i) * ld + r] : 0.f;
        if (split && r < N) z += zT[(int64_t)(b + c0 + i) * ld + r];
        tile[i][tx] = z;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int64_t r = r0 + i;
        const int c = c0 + tx;
        if (r < N) X[r * b + c] = fma(sqrt(fmax(rn2[c], 0.0)), (double)tile[tx][i], X[r * b + c]);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;   // immutable after the first lookup (same pointer for every device)
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 tensor [outer][inner] with row pitch ld floats, box 32 x box_outer, 128-byte swizzle
static int make_map(CUtensorMap* map, const float* base, int64_t inner, int64_t outer, int64_t ld, int box_outer) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return SCB_ERR_UNSUPPORTED;
    const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kT32BK, (cuuint32_t)box_outer};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SCB_OK : SCB_ERR_INVALID;
}

}  // namespace scb

using namespace scb;

extern "C" int64_t scb_tf32_ld(int64_t N) { return (N + 3) & ~(int64_t)3; }

extern "C" int scb_dense_slab_to_f32(int64_t N, int64_t rows, const double* slab, float* slab32, float* slab32_lo,
                                     void* stream) {
    if (!slab || !slab32 || N < 1 || rows < 1) return SCB_ERR_INVALID;
    const int64_t ld = scb_tf32_ld(N);
    dim3 grid((unsigned)ceil_div(ld, 256 * 8), (unsigned)rows);
    slab_to_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(rows, N, ld, slab, slab32, slab32_lo);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_resform_prepare(int64_t N, int b, int deg, const double* X, const double* HX, const double* theta,
                                   const double* rn2, double lo, double ub, float* rhatT, float* z1T, float* z0T,
                                   float* cA, float* cB, int split, void* stream) {
    if (!X || !HX || !theta || !rn2 || !rhatT || !z1T || !z0T || !cA || !cB) return SCB_ERR_INVALID;
    if (b % 32 != 0 || deg < 2 || !(ub > lo)) return SCB_ERR_INVALID;
    const int64_t ld = scb_tf32_ld(N);
    cudaStream_t st = as_stream(stream);
    dim3 grid((unsigned)ceil_div(ld, 32), (unsigned)(b / 32));
    resform_prepare_kernel<<<grid, 256, 0, st>>>(N, b, ld, X, HX, theta, rn2, lo, ub, rhatT, z1T, z0T, split);
    SCB_LAUNCH_CHECK();
    resform_coef_kernel<<<(unsigned)ceil_div(b, 128), 128, 0, st>>>(b, deg, theta, lo, ub, cA, cB);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_resform_finish(int64_t N, int b, const double* rn2, const float* zT, double* X, int split,
                                  void* stream) {
    if (!rn2 || !zT || !X || b % 32 != 0) return SCB_ERR_INVALID;
    dim3 grid((unsigned)ceil_div(N, 32), (unsigned)(b / 32));
    resform_finish_kernel<<<grid, 256, 0, as_stream(stream)>>>(N, b, scb_tf32_ld(N), rn2, zT, X, split);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

static int tf32_apply(int64_t N, int64_t row0, int64_t row1, const float* slab32, const float* slab32_lo, int b,
                      const float* zcurT, const float* zprevT, const float* rhatT, float* const* outs, int nout,
                      const float* cA, const float* cB, double cshift, int fused, cudaStream_t st) {
    if (!slab32 || !zcurT || !outs || nout < 1 || nout > 8 || N < 1 || row0 < 0 || row1 <= row0 || row1 > N)
        return SCB_ERR_INVALID;
    if (b != kT32BN) return SCB_ERR_UNSUPPORTED;
    if (fused && (!zprevT || !rhatT || !cA || !cB)) return SCB_ERR_INVALID;
    const int64_t ld = scb_tf32_ld(N);
    const int64_t rows = row1 - row0;
    const bool split = slab32_lo != nullptr;    // z blocks are then [2 b][ld] (hi rows, lo rows)
    CUtensorMap mapA, mapB, mapAlo, mapBlo;
    SCB_TRY(make_map(&mapA, slab32, N, rows, ld, kT32BM));
    SCB_TRY(make_map(&mapB, zcurT, N, b, ld, kT32BN));
    SCB_TRY(make_map(&mapAlo, split ? slab32_lo : slab32, N, rows, ld, kT32BM));
    SCB_TRY(make_map(&mapBlo, split ? zcurT + (int64_t)b * ld : zcurT, N, b, ld, kT32BN));
    T32Epilogue ep;
    ep.zcur = fused ? zcurT : nullptr; ep.zprev = zprevT; ep.rhat = rhatT;
    for (int p = 0; p < 8; ++p) ep.out[p] = p < nout ? outs[p] : nullptr;
    ep.nout = nout;
    ep.cA = cA; ep.cB = cB; ep.cshift = (float)cshift; ep.ld = ld; ep.row0 = (int)row0; ep.rows = (int)rows;
    const unsigned grid = (unsigned)ceil_div(rows, kT32BM);
    if (split) {
        const size_t smem = (size_t)kT32Stages * 4 * kT32TileBytes + 1024;
        SCB_CUDA(cudaFuncSetAttribute(dense_slab_tf32_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dense_slab_tf32_kernel<3><<<grid, kT32Threads, smem, st>>>(mapA, mapB, mapAlo, mapBlo, (int)N, ep);
    } else {
        const size_t smem = (size_t)kT32Stages * 2 * kT32TileBytes + 1024;
        SCB_CUDA(cudaFuncSetAttribute(dense_slab_tf32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dense_slab_tf32_kernel<1><<<grid, kT32Threads, smem, st>>>(mapA, mapB, mapAlo, mapBlo, (int)N, ep);
    }
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_dense_slab_tf32_apply(int64_t N, int64_t row0, int64_t row1, const float* slab32,
                                         const float* slab32_lo, int b, const float* zcurT, const float* zprevT,
                                         const float* rhatT, float* outT, const float* cA, const float* cB,
                                         double cshift, int fused, void* stream) {
    if (!outT) return SCB_ERR_INVALID;
    float* outs[1] = {outT};
    return tf32_apply(N, row0, row1, slab32, slab32_lo, b, zcurT, zprevT, rhatT, outs, 1, cA, cB, cshift, fused,
                      as_stream(stream));
}

extern "C" int scb_dense_slab_tf32_apply_allgather(int64_t N, int64_t row0, int64_t row1, const float* slab32,
                                                   const float* slab32_lo, int b, const float* zcurT,
                                                   const float* zprevT, const float* rhatT, float* const* out_all,
                                                   int world, const float* cA, const float* cB, double cshift,
                                                   int fused, void* stream) {
    return tf32_apply(N, row0, row1, slab32, slab32_lo, b, zcurT, zprevT, rhatT, out_all, world, cA, cB, cshift, fused,
                      as_stream(stream));
}
