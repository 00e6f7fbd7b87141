// K3a': SpMM on a row-PAIRED block format ("BSR 2D x D"), the tuned operator of
// the lowest-k eigensolver.
//
// ncu on the one-row-per-warp kernel (profiles/r1a) shows L1/TEX at 94 % of
// peak with the FP64 pipe 23 % busy: with one column per lane every contact
// costs ~6 L1 wavefronts of X rows plus 5 broadcast wavefronts for the block.
// Here two consecutive block rows (residues 2t, 2t+1 -- chain neighbours share
// ~85 % of their contacts) are merged into one list of (2D x D) blocks, the
// diagonal blocks folded in, and the warp is split into 4 contact slots x 8
// lanes x 4 columns:
//   * an X row segment is loaded once (LDG.128 x2 per lane, each instruction
//     covering 128 contiguous bytes per slot = full 32-byte sectors) and used
//     for both block rows and 4 columns;
//   * block values are read as LDS.128 with 4 distinct addresses per
//     instruction (conflict free: 160-byte records);
//   * 72 DFMA per lane and merged contact.
// A merged contact is one 160-byte record {2 DxD blocks, column index}; a chunk
// of 16 records is staged global->shared by ONE TMA bulk copy
// (cp.async.bulk + mbarrier, double buffered per warp), which keeps the
// staging traffic off the LSU/L1 wavefront budget (profiles/r1c: the LDGSTS
// staging cost ~30 % of the L1 wavefronts).
#include "paired.cuh"

namespace scb {

// ---------------------------------------------------------------------------
// format conversion: CSR/BSR (D x D) -> paired: a two-pointer merge of the two
// sorted column lists of a row pair (+ the two diagonal columns).
// Capacity offsets need no scan: pair (s,t) owns [rowptr[r0] + 2*g, ...) with
// g its global pair index and r0 its first row, at most cnt0+cnt1+2 entries.
// ---------------------------------------------------------------------------
// Two launches: (1) pair_index_kernel -- the merge itself, one thread per row pair, writes only the 16 bytes of
// every record that are not block values: {col, source of the upper block, source of the lower block} (sources are
// offsets inside the row, kSrcNone = zero block, kSrcDiag = the diagonal block);  (2) pair_fill_kernel -- all lanes
// of a warp copy the block values of a pair's records (18 doubles per record, contiguous on both sides).  The
// one-launch version (a thread wrote whole 160-byte records) spent 8.2 ms per C3 batch on strided 8-byte stores.
constexpr int32_t kSrcNone = -1, kSrcDiag = -2;

template <int D>
__global__ void __launch_bounds__(128)
pair_index_kernel(int n, int np, int64_t total_pairs, const int64_t* __restrict__ rowptr,
                  const int32_t* __restrict__ col, int32_t* __restrict__ pcount, PairEntry<D>* __restrict__ pent) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_pairs) return;
    const int64_t s = g / np;
    const int t = (int)(g % np);
    const int i0 = 2 * t, i1 = 2 * t + 1;
    const bool has1 = i1 < n;
    const int64_t r0 = s * n + i0, r1 = r0 + 1;
    const int64_t a0 = rowptr[r0], ae = rowptr[r0 + 1];
    const int64_t b0 = has1 ? rowptr[r1] : 0, be = has1 ? rowptr[r1 + 1] : 0;
    int64_t a = a0, b = b0;
    int64_t out = a0 + 2 * g;
    const int64_t out0 = out;
    bool d0 = false, d1 = !has1;  // diagonal columns already emitted?
    const int BIG = 0x7fffffff;
    while (true) {
        // next candidate column of each row (diagonal entry inserted in order)
        int ca = (a < ae) ? col[a] : BIG;
        bool a_is_diag = false;
        if (!d0 && i0 < ca) { ca = i0; a_is_diag = true; }
        int cb = BIG;
        bool b_is_diag = false;
        if (has1) {
            cb = (b < be) ? col[b] : BIG;
            if (!d1 && i1 < cb) { cb = i1; b_is_diag = true; }
        }
        const int c = min(ca, cb);
        if (c == BIG) break;
        int4 meta = make_int4(c, kSrcNone, kSrcNone, 0);
        if (ca == c) {
            meta.y = a_is_diag ? kSrcDiag : (int)(a - a0);
            if (a_is_diag) d0 = true; else ++a;
        }
        if (cb == c) {
            meta.z = b_is_diag ? kSrcDiag : (int)(b - b0);
            if (b_is_diag) d1 = true; else ++b;
        }
        *reinterpret_cast<int4*>(&pent[out].col) = meta;   // col + pad[3]: one aligned 16-byte store
        ++out;
    }
    pcount[g] = (int)(out - out0);
}

template <int D>
__global__ void __launch_bounds__(256)
pair_fill_kernel(int n, int np, int64_t total_pairs, const int64_t* __restrict__ rowptr,
                 const double* __restrict__ offdiag, const double* __restrict__ diag,
                 const int32_t* __restrict__ pcount, PairEntry<D>* __restrict__ pent) {
    constexpr int DD = D * D, V = 2 * DD;   // values per record
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= total_pairs) return;
    const int lane = threadIdx.x & 31;
    const int64_t s = g / np;
    const int t = (int)(g % np);
    const int64_t r0 = s * n + 2 * t, r1 = r0 + 1;
    const int64_t a0 = rowptr[r0];
    const int64_t b0 = (2 * t + 1 < n) ? rowptr[r1] : 0;
    PairEntry<D>* ent = pent + a0 + 2 * g;
    const int total = pcount[g] * V;
    // four independent (header -> block value -> store) chains in flight per lane: the copy is latency-bound otherwise
    for (int base = lane; base < total; base += 128) {
        int e[4], q[4], src[4];
        bool low[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + 32 * u;
            e[u] = idx / V;
            q[u] = idx - e[u] * V;
            low[u] = q[u] >= DD;
            src[u] = kSrcNone;
            if (idx < total) src[u] = low[u] ? ent[e[u]].pad[1] : ent[e[u]].pad[0];
        }
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int qq = low[u] ? q[u] - DD : q[u];
            v[u] = 0.0;
            if (src[u] == kSrcDiag) v[u] = diag[(low[u] ? r1 : r0) * DD + qq];
            else if (src[u] != kSrcNone) v[u] = offdiag[((low[u] ? b0 : a0) + src[u]) * DD + qq];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (base + 32 * u < total) ent[e[u]].blk[q[u]] = v[u];
    }
}

// ---------------------------------------------------------------------------
// Y = alpha (H X - c X) - beta W   on the paired format; b = 32 * ncg columns
// ---------------------------------------------------------------------------
template <int D, int BC>
__global__ void __launch_bounds__(kPairWarps * 32, 1)
spmm_paired_kernel(int n, int np, int pairs_per_cta, const int64_t* __restrict__ rowptr,
                   const int32_t* __restrict__ pcount, const PairEntry<D>* __restrict__ pent,
                   const double* __restrict__ X, const double* __restrict__ W, double* __restrict__ Y,
                   const double* __restrict__ coef, int coef_stride, const int32_t* __restrict__ done) {
    constexpr int R = 2 * D;             // rows per pair (6 or 2)
    using Entry = PairEntry<D>;
    extern __shared__ __align__(128) unsigned char pair_smem[];
    Entry* stage_base = reinterpret_cast<Entry*>(pair_smem);                                  // [warps][2][chunk]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + kPairWarps * 2 * kPairChunk);   // [warps][2]
    const int64_t s = blockIdx.y;
    if (done && done[s]) return;
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
    unsigned phase0 = 0, phase1 = 0;     // parity of the next completion of each buffer
    const int64_t N = (int64_t)D * n;
    double alpha = 1.0, cshift = 0.0, beta = 0.0;
    if (coef) {
        const double* cf = coef + s * coef_stride;
        alpha = cf[0]; cshift = cf[1]; beta = cf[2];
    }
    const int t0 = blockIdx.x * pairs_per_cta;
    const int t1 = min(np, t0 + pairs_per_cta);
    constexpr int b = BC;               // compile-time: row offsets become immediates
    constexpr int ncg = BC >> 5;
    constexpr int rowlen = D * BC;            // doubles between the X rows of consecutive nodes
    if (coef && alpha == 0.0) {
        // idle filter step of a structure that runs fewer degrees: carry its rows forward, Y = X
        for (int t = t0 + warp; t < t1; t += kPairWarps)
            for (int a = 0; a < R; ++a) {
                const int64_t r = (int64_t)D * 2 * t + a;
                if (r >= N) continue;
                const double* src = X + (s * N + r) * b;
                double* dst = Y + (s * N + r) * b;
                for (int q = lane; q < b; q += 32) dst[q] = src[q];
            }
        return;
    }
    for (int t = t0 + warp; t < t1; t += kPairWarps) {
        const int64_t g = s * np + t;
        const int64_t base = rowptr[s * n + 2 * t] + 2 * g;
        const int cnt = pcount[g];
        const int nchunk = (cnt + kPairChunk - 1) / kPairChunk;
        for (int cg = 0; cg < ncg; ++cg) {
            const int c0 = cg * 32 + 2 * l8;  // this lane's columns: c0, c0+1, c0+16, c0+17
            const double* Xs = X + s * N * b + c0;
            double acc[R][4];
#pragma unroll
            for (int a = 0; a < R; ++a)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) acc[a][cc] = 0.0;
            __syncwarp();  // every lane is done with both buffers
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
                    // one elected lane arms the barrier and launches the bulk copy of the next chunk
                    const unsigned bytes = (unsigned)(min(kPairChunk, cnt - e0 - kPairChunk) * sizeof(Entry));
                    mbar_expect_tx(&bar[buf ^ 1], bytes);
                    bulk_g2s(stage + (buf ^ 1) * kPairChunk, pent + base + e0 + kPairChunk, bytes, &bar[buf ^ 1]);
                }
                if (buf == 0) { mbar_wait(&bar[0], phase0); phase0 ^= 1; }
                else { mbar_wait(&bar[1], phase1); phase1 ^= 1; }
                const Entry* cur = stage + buf * kPairChunk;
                // this slot owns contacts slot, slot+4, slot+8, slot+12 of the chunk
#pragma unroll 2
                for (int q = slot; q < m; q += 4) {
                    const double* xr = Xs + cur[q].col * rowlen;  // 32-bit offset inside one structure
                    double x[D][4];
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        const double2 u = __ldg(reinterpret_cast<const double2*>(xr + c * b));
                        const double2 v = __ldg(reinterpret_cast<const double2*>(xr + c * b + 16));
                        x[c][0] = u.x; x[c][1] = u.y; x[c][2] = v.x; x[c][3] = v.y;
                    }
                    const double2* bp = reinterpret_cast<const double2*>(cur[q].blk);
#pragma unroll
                    for (int a = 0; a < R; ++a) {
                        double h[D];
                        if (D == 3) {
                            const int o = a * 3;  // elements o .. o+2 of the 18-double record
                            const double2 u = bp[o >> 1];
                            const double2 v = bp[(o >> 1) + 1];
                            if (o & 1) { h[0] = u.y; h[1] = v.x; h[2] = v.y; }
                            else { h[0] = u.x; h[1] = u.y; h[2] = v.x; }
                        } else {
                            const double2 u = bp[0];
                            h[0] = a ? u.y : u.x;
                        }
#pragma unroll
                        for (int c = 0; c < D; ++c)
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) acc[a][cc] = fma(h[c], x[c][cc], acc[a][cc]);
                    }
                }
                __syncwarp();  // buffer `buf` may be refilled by the next bulk copy
            }
            // combine the four contact slots
#pragma unroll
            for (int a = 0; a < R; ++a)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    double v = acc[a][cc];
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    acc[a][cc] = v;
                }
            // slot q writes rows q, q+4 of the pair (each row: 8 lanes x 16 B contiguous, twice)
#pragma unroll
            for (int a = 0; a < R; ++a) {
                if ((a & 3) != slot) continue;
                const int64_t r = (int64_t)D * 2 * t + a;
                if (r >= N) continue;
                const int64_t idx = (s * N + r) * b + c0;
                double v[4] = {acc[a][0], acc[a][1], acc[a][2], acc[a][3]};
                if (coef) {
                    const double2* xp = reinterpret_cast<const double2*>(X + idx);
                    const double2 u = xp[0], w2 = xp[8];
                    const double xo[4] = {u.x, u.y, w2.x, w2.y};
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) v[cc] = alpha * (v[cc] - cshift * xo[cc]);
                    if (W && beta != 0.0) {
                        const double2* wp = reinterpret_cast<const double2*>(W + idx);
                        const double2 p = wp[0], q2 = wp[8];
                        v[0] -= beta * p.x; v[1] -= beta * p.y; v[2] -= beta * q2.x; v[3] -= beta * q2.y;
                    }
                }
                double2* yp = reinterpret_cast<double2*>(Y + idx);
                yp[0] = make_double2(v[0], v[1]);
                yp[8] = make_double2(v[2], v[3]);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
size_t paired_capacity(int B, int n, int64_t P) { return (size_t)P + 2 * (size_t)B * ((n + 1) / 2); }
size_t paired_entry_bytes(int D) { return D == 3 ? sizeof(PairEntry<3>) : sizeof(PairEntry<1>); }

int build_paired(int D, int B, int n, const int64_t* rowptr, const int32_t* col, const double* offdiag,
                 const double* diag, int32_t* pcount, void* pent, cudaStream_t st) {
    const int np = (n + 1) / 2;
    const int64_t total = (int64_t)B * np;
    const unsigned grid = (unsigned)ceil_div(total, 128);
    const unsigned grid2 = (unsigned)ceil_div(total, 8);
    if (D == 3) {
        pair_index_kernel<3><<<grid, 128, 0, st>>>(n, np, total, rowptr, col, pcount, static_cast<PairEntry<3>*>(pent));
        SCB_LAUNCH_CHECK();
        pair_fill_kernel<3><<<grid2, 256, 0, st>>>(n, np, total, rowptr, offdiag, diag, pcount, static_cast<PairEntry<3>*>(pent));
    } else if (D == 1) {
        pair_index_kernel<1><<<grid, 128, 0, st>>>(n, np, total, rowptr, col, pcount, static_cast<PairEntry<1>*>(pent));
        SCB_LAUNCH_CHECK();
        pair_fill_kernel<1><<<grid2, 256, 0, st>>>(n, np, total, rowptr, offdiag, diag, pcount, static_cast<PairEntry<1>*>(pent));
    } else {
        return SCB_ERR_INVALID;
    }
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

template <int D, int BC>
static int launch_paired(dim3 grid, int n, int np, int per_cta, const int64_t* rowptr, const int32_t* pcount,
                         const void* pent, const double* X, const double* W, double* Y, const double* coef,
                         int coef_stride, const int32_t* done, cudaStream_t st) {
    const size_t smem = sizeof(PairEntry<D>) * 2 * kPairChunk * kPairWarps + sizeof(uint64_t) * 2 * kPairWarps;
    // per-device function attribute: set on every launch (cheap), no process-wide "configured" flag
    SCB_CUDA(cudaFuncSetAttribute(spmm_paired_kernel<D, BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    spmm_paired_kernel<D, BC><<<grid, kPairWarps * 32, smem, st>>>(n, np, per_cta, rowptr, pcount,
                                                             static_cast<const PairEntry<D>*>(pent), X, W, Y, coef,
                                                             coef_stride, done);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

int spmm_paired(int D, int B, int n, const int64_t* rowptr, const int32_t* pcount, const void* pent,
                const double* X, const double* W, double* Y, int b, const double* coef, int coef_stride,
                const int32_t* done, cudaStream_t st) {
    if (b != 32 && b != 64) return SCB_ERR_UNSUPPORTED;
    const int np = (n + 1) / 2;
    const int64_t total = (int64_t)B * np;
    int per_cta = (int)ceil_div(total, 4 * kNumSM);
    per_cta = per_cta < kPairWarps ? kPairWarps : (per_cta > 256 ? 256 : per_cta);
    if (per_cta > np) per_cta = np;
    dim3 grid((unsigned)ceil_div(np, per_cta), (unsigned)B);
    if (D == 3 && b == 32) return launch_paired<3, 32>(grid, n, np, per_cta, rowptr, pcount, pent, X, W, Y, coef, coef_stride, done, st);
    if (D == 3 && b == 64) return launch_paired<3, 64>(grid, n, np, per_cta, rowptr, pcount, pent, X, W, Y, coef, coef_stride, done, st);
    if (D == 1 && b == 32) return launch_paired<1, 32>(grid, n, np, per_cta, rowptr, pcount, pent, X, W, Y, coef, coef_stride, done, st);
    if (D == 1 && b == 64) return launch_paired<1, 64>(grid, n, np, per_cta, rowptr, pcount, pent, X, W, Y, coef, coef_stride, done, st);
    return SCB_ERR_INVALID;
}

}  // namespace scb

using namespace scb;

extern "C" size_t scb_paired_bytes(int D, int B, int n, int64_t P, size_t* count_off, size_t* entry_off) {
    Arena ar(nullptr, 0);
    const size_t cap = paired_capacity(B, n, P);
    const size_t o0 = ar.off;
    ar.take<int32_t>((size_t)B * ((n + 1) / 2));
    const size_t o1 = ar.off;
    ar.take<char>(cap * paired_entry_bytes(D));
    if (count_off) *count_off = o0;
    if (entry_off) *entry_off = o1;
    return ar.off + 256;
}

extern "C" int scb_paired_build(int D, int B, int n, int64_t P, const int64_t* rowptr, const int32_t* col,
                                const double* offdiag, const double* diag, void* paired, void* stream) {
    if (!rowptr || !col || !offdiag || !diag || !paired) return SCB_ERR_INVALID;
    size_t o0, o1;
    scb_paired_bytes(D, B, n, P, &o0, &o1);
    char* base = static_cast<char*>(paired);
    return build_paired(D, B, n, rowptr, col, offdiag, diag, (int32_t*)(base + o0), base + o1, as_stream(stream));
}

extern "C" int scb_spmm_paired(int D, int B, int n, int64_t P, const int64_t* rowptr, const void* paired,
                               const double* X, double* Y, int b, void* stream) {
    if (!rowptr || !paired || !X || !Y) return SCB_ERR_INVALID;
    size_t o0, o1;
    scb_paired_bytes(D, B, n, P, &o0, &o1);
    const char* base = static_cast<const char*>(paired);
    return spmm_paired(D, B, n, rowptr, (const int32_t*)(base + o0), base + o1, X, nullptr, Y, b, nullptr, 0, nullptr,
                       as_stream(stream));
}
