// Resident operator format: BSR (3x3) -> rank-one row-pair records, grouped and interleaved per warp
// (see resident.cuh).  One CTA per structure.
#include "resident.cuh"

namespace scb {

// H_ij = t d d^T, t <= 0  ->  v = sqrt(-t) d  (H_ij = -v v^T), taken from the block row with the largest diagonal
__device__ __forceinline__ bool block_to_v(const double* __restrict__ blk, float v[3]) {
    const double d0 = -blk[0], d1 = -blk[4], d2 = -blk[8];
    const bool bad = (d0 < 0.0) || (d1 < 0.0) || (d2 < 0.0);   // positive t (negative force constant)
    int a = 0;
    double m = d0;
    if (d1 > m) { m = d1; a = 1; }
    if (d2 > m) { m = d2; a = 2; }
    if (!(m > 0.0)) { v[0] = v[1] = v[2] = 0.f; return bad; }
    const double va = sqrt(m);
    const double inv = -1.0 / va;
    v[0] = (float)(a == 0 ? va : blk[3 * a + 0] * inv);
    v[1] = (float)(a == 1 ? va : blk[3 * a + 1] * inv);
    v[2] = (float)(a == 2 ? va : blk[3 * a + 2] * inv);
    return bad;
}


// cursor over the merged column list of a row pair that stops only at columns of one parity
struct ParityCursor {
    int64_t a, ae, b, be;
    int parity;
};
__device__ __forceinline__ void cursor_init(ParityCursor& c, const int64_t* __restrict__ rowptr, int64_t r0, bool has1,
                                            int parity) {
    c.a = rowptr[r0]; c.ae = rowptr[r0 + 1];
    c.b = has1 ? rowptr[r0 + 1] : 0; c.be = has1 ? rowptr[r0 + 2] : 0;
    c.parity = parity;
}
// next merged contact of the cursor's parity -> record; false when the list is exhausted
__device__ __forceinline__ bool cursor_next(ParityCursor& c, const int32_t* __restrict__ col,
                                            const double* __restrict__ offdiag, ResRec& e, bool& bad) {
    while (c.a < c.ae || c.b < c.be) {
        const int ca = c.a < c.ae ? col[c.a] : 0x7fffffff;
        const int cb = c.b < c.be ? col[c.b] : 0x7fffffff;
        const int m = min(ca, cb);
        if ((m & 1) != c.parity) { c.a += (ca == m); c.b += (cb == m); continue; }
        e.col = m;
        e.pad = 0;
        if (ca == m) { bad |= block_to_v(offdiag + c.a * 9, e.v0); ++c.a; }
        else { e.v0[0] = e.v0[1] = e.v0[2] = 0.f; }
        if (cb == m) { bad |= block_to_v(offdiag + c.b * 9, e.v1); ++c.b; }
        else { e.v1[0] = e.v1[1] = e.v1[2] = 0.f; }
        return true;
    }
    return false;
}

constexpr int kBuildThreads = 256;
constexpr int kBuildMaxPairs = 768;

__global__ void __launch_bounds__(kBuildThreads)
resident_build_kernel(int n, int np, int G, int rpw, int64_t rec_pad, const int64_t* __restrict__ rowptr,
                      const int32_t* __restrict__ col, const double* __restrict__ offdiag,
                      const double* __restrict__ diag, ResRec* __restrict__ rec, int32_t* __restrict__ gstart,
                      uint16_t* __restrict__ order, float* __restrict__ diag32, int32_t* __restrict__ flag) {
    __shared__ int cnt[kBuildMaxPairs];
    __shared__ int cnt_even[kBuildMaxPairs];
    __shared__ int rnk[kBuildMaxPairs];
    __shared__ int gs[kResMaxWarps + 1];
    __shared__ int bad_any;
    const int64_t s = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid == 0) bad_any = 0;
    const int64_t row_base = s * n;
    // ---- 1. records per row pair = size of the union of the two column lists
    for (int t = tid; t < np; t += kBuildThreads) {
        const int64_t r0 = row_base + 2 * t;
        const bool has1 = 2 * t + 1 < n;
        int64_t a = rowptr[r0], ae = rowptr[r0 + 1];
        int64_t b = has1 ? rowptr[r0 + 1] : 0, be = has1 ? rowptr[r0 + 2] : 0;
        int c = 0, ce = 0;
        while (a < ae || b < be) {
            const int ca = a < ae ? col[a] : 0x7fffffff;
            const int cb = b < be ? col[b] : 0x7fffffff;
            const int m = min(ca, cb);
            a += (ca == m);
            b += (cb == m);
            ++c;
            ce += !(m & 1);
        }
        cnt[t] = c;
        cnt_even[t] = ce;
    }
    __syncthreads();
    // ---- 2. descending order by count (ties by index): rank by counting
    uint16_t* ord = order + s * (int64_t)G * rpw;
    for (int t = tid; t < np; t += kBuildThreads) {
        const int c = cnt[t];
        int r = 0;
        for (int u = 0; u < np; ++u) r += (cnt[u] > c) || (cnt[u] == c && u < t);
        rnk[t] = r;
        ord[r] = (uint16_t)t;
    }
    for (int r = np + tid; r < G * rpw; r += kBuildThreads) ord[r] = 0xFFFF;
    __syncthreads();
    // ---- 3. iterations per group (longest list of the group, rounded up to an even count)
    if (tid == 0) {
        int acc = 0;
        for (int g = 0; g < G; ++g) {
            gs[g] = acc;
            const int first = ord[g * rpw];   // ranks are dense: the first slot of a group always holds a row pair
            int it = cnt[first];
            it = (it + 1) & ~1;
            acc += it;
        }
        gs[G] = acc;
    }
    __syncthreads();
    int32_t* gst = gstart + s * (int64_t)(G + 1);
    for (int g = tid; g <= G; g += kBuildThreads) gst[g] = gs[g];
    // ---- 4. fill: (group, iteration, slot) interleaved records
    const int64_t start = ((rowptr[row_base] + rec_pad * s) + 7) & ~(int64_t)7;
    ResRec* base = rec + start;
    bool bad = false;
    ResRec zrec;
    zrec.col = 0; zrec.pad = 0;
    zrec.v0[0] = zrec.v0[1] = zrec.v0[2] = 0.f;
    zrec.v1[0] = zrec.v1[1] = zrec.v1[2] = 0.f;
    if (rpw == 8) {
        // 16 columns per CTA: a quarter warp (one shared-memory wavefront of the z gather) serves the row pairs of
        // slots 2m and 2m+1, and node j sits at bank offset 16 (j mod 2).  Order the two record lists so that, step by
        // step, the two columns have different parity whenever the lists allow it: the gather is then conflict free.
        for (int r2 = tid; r2 < G * 4; r2 += kBuildThreads) {
            const int g = r2 >> 2, m = r2 & 3;
            const int iters = gs[g + 1] - gs[g];
            const int rA = g * 8 + 2 * m, rB = rA + 1;
            const int tA = rA < np ? (int)ord[rA] : -1, tB = rB < np ? (int)ord[rB] : -1;
            ParityCursor cur[2][2];     // [A|B][even|odd]
            int rem[2][2] = {{0, 0}, {0, 0}};
            if (tA >= 0) {
                cursor_init(cur[0][0], rowptr, row_base + 2 * tA, 2 * tA + 1 < n, 0);
                cursor_init(cur[0][1], rowptr, row_base + 2 * tA, 2 * tA + 1 < n, 1);
                rem[0][0] = cnt_even[tA]; rem[0][1] = cnt[tA] - cnt_even[tA];
            }
            if (tB >= 0) {
                cursor_init(cur[1][0], rowptr, row_base + 2 * tB, 2 * tB + 1 < n, 0);
                cursor_init(cur[1][1], rowptr, row_base + 2 * tB, 2 * tB + 1 < n, 1);
                rem[1][0] = cnt_even[tB]; rem[1][1] = cnt[tB] - cnt_even[tB];
            }
            ResRec* outA = base + (int64_t)gs[g] * 8 + 2 * m;
            for (int i = 0; i < iters; ++i) {
                const bool actA = rem[0][0] + rem[0][1] > 0, actB = rem[1][0] + rem[1][1] > 0;
                int pA, pB;
                if (actA && actB) {
                    const bool f1 = rem[0][0] > 0 && rem[1][1] > 0;   // A even, B odd
                    const bool f2 = rem[0][1] > 0 && rem[1][0] > 0;   // A odd, B even
                    bool first;
                    if (f1 && f2) first = (rem[0][0] - rem[0][1]) + (rem[1][1] - rem[1][0]) >= 0;
                    else first = f1 || !f2;
                    pA = first ? 0 : 1; pB = first ? 1 : 0;
                    if (!f1 && !f2) {                                  // forced: same parity on both sides
                        pA = rem[0][0] > 0 ? 0 : 1;
                        pB = rem[1][0] > 0 ? 0 : 1;
                    }
                } else if (actA) {
                    pA = rem[0][0] >= rem[0][1] ? 0 : 1; pB = pA ^ 1;
                } else {
                    pB = rem[1][0] >= rem[1][1] ? 0 : 1; pA = pB ^ 1;
                }
                ResRec e = zrec;
                e.col = pA;                                            // padding record: harmless column of the free parity
                if (actA) { cursor_next(cur[0][pA], col, offdiag, e, bad); --rem[0][pA]; }
                outA[(int64_t)i * 8] = e;
                e = zrec;
                e.col = pB;
                if (actB) { cursor_next(cur[1][pB], col, offdiag, e, bad); --rem[1][pB]; }
                outA[(int64_t)i * 8 + 1] = e;
            }
        }
    } else
    for (int r = tid; r < G * rpw; r += kBuildThreads) {
        const int g = r / rpw, slot = r % rpw;
        const int iters = gs[g + 1] - gs[g];
        ResRec* out = base + (int64_t)gs[g] * rpw + slot;
        int i = 0;
        if (r < np) {
            const int t = ord[r];
            const int64_t r0 = row_base + 2 * t;
            const bool has1 = 2 * t + 1 < n;
            int64_t a = rowptr[r0], ae = rowptr[r0 + 1];
            int64_t b = has1 ? rowptr[r0 + 1] : 0, be = has1 ? rowptr[r0 + 2] : 0;
            while (a < ae || b < be) {
                const int ca = a < ae ? col[a] : 0x7fffffff;
                const int cb = b < be ? col[b] : 0x7fffffff;
                const int m = min(ca, cb);
                ResRec e;
                e.col = m;
                e.pad = 0;
                if (ca == m) { bad |= block_to_v(offdiag + a * 9, e.v0); ++a; }
                else { e.v0[0] = e.v0[1] = e.v0[2] = 0.f; }
                if (cb == m) { bad |= block_to_v(offdiag + b * 9, e.v1); ++b; }
                else { e.v1[0] = e.v1[1] = e.v1[2] = 0.f; }
                out[(int64_t)i * rpw] = e;
                ++i;
            }
        }
        for (; i < iters; ++i) out[(int64_t)i * rpw] = zrec;
    }
    // ---- 5. diagonal blocks in single precision, 12 floats per node
    for (int q = tid; q < n * 12; q += kBuildThreads) {
        const int i = q / 12, k = q % 12;
        diag32[(row_base + i) * 12 + k] = k < 9 ? (float)diag[(row_base + i) * 9 + k] : 0.f;
    }
    if (bad) bad_any = 1;
    __syncthreads();
    if (tid == 0 && bad_any) atomicExch(flag, 1);
}

int resident_cols(int n, int b) {
    if (b != 32 && b != 64) return 0;
    const int np = (n + 1) / 2;
    const int64_t N = 3 * (int64_t)n;
    for (int cols = 16; cols >= 4; cols >>= 1) {
        const int rpw = 128 / cols;
        const int G = (np + rpw - 1) / rpw;
        const size_t smem = (size_t)N * cols * 12 + 2 * kResDegreeCap * cols * 4 + 1024;
        if (G <= kResMaxWarps && smem <= 220 * 1024) return cols;
    }
    return 0;
}

static int64_t resident_rec_pad(int n, int cols) {
    const int np = (n + 1) / 2;
    const int rpw = 128 / cols;
    const int G = (np + rpw - 1) / rpw;
    // groups are padded to their longest list (<= n records, rounded up to even): at most rpw * (n + 1) extra
    // records per structure beyond its contact count (see resident.cuh), + empty slots of the last group, + alignment
    return (int64_t)rpw * (n + 2) + (int64_t)G * rpw + 8;
}

size_t resident_capacity(int B, int n, int64_t P, int cols) {
    return (size_t)P + (size_t)resident_rec_pad(n, cols) * B + 16;
}

int resident_build(int B, int n, const int64_t* rowptr, const int32_t* col, const double* offdiag,
                   const double* diag, const ResLayout& L, cudaStream_t st) {
    const int np = (n + 1) / 2;
    if (np > kBuildMaxPairs || L.G > kResMaxWarps) return SCB_ERR_UNSUPPORTED;
    SCB_CUDA(cudaMemsetAsync(L.flag, 0, sizeof(int32_t), st));
    resident_build_kernel<<<B, kBuildThreads, 0, st>>>(n, np, L.G, L.rpw, L.rec_pad, rowptr, col, offdiag, diag, L.rec,
                                                       L.gstart, L.order, L.diag32, L.flag);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

int64_t resident_rec_pad_of(int n, int cols) { return resident_rec_pad(n, cols); }

}  // namespace scb
