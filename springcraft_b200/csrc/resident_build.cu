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

// ---------------------------------------------------------------------------
// 16 columns per CTA (4 lanes per row pair, 8 slots per warp).  A quarter warp = the two slots 2m, 2m+1 is one
// shared-memory wavefront of the z gather, and node j sits at bank offset 16 (j mod 2), so the unit of the layout
// is the QUARTER:
//   * a LONG row pair is split over the two slots of a quarter: slot 2m walks its even columns, slot 2m+1 its odd
//     columns (conflict free by construction); the filter adds the two partial sums with a warp shuffle.  Splitting
//     the longest lists evens out the work of the warps, which meet at a barrier after every filter step.
//   * two SHORT row pairs share a quarter; their lists are ordered so that, step by step, the two columns have
//     different parity whenever the lists allow it.
// Quarters are sorted by length; 4 consecutive quarters form the group of one warp.
// ---------------------------------------------------------------------------
constexpr int kBuildMaxQuarters = kResMaxWarps * 4;

__global__ void __launch_bounds__(kBuildThreads)
resident_build16_kernel(int n, int np, int G, int nsplit, int64_t rec_pad, const int64_t* __restrict__ rowptr,
                        const int32_t* __restrict__ col, const double* __restrict__ offdiag,
                        const double* __restrict__ diag, ResRec* __restrict__ rec, int32_t* __restrict__ gstart,
                        uint16_t* __restrict__ order, float* __restrict__ diag32, int32_t* __restrict__ flag) {
    __shared__ int cnt[kBuildMaxPairs];
    __shared__ int cnt_even[kBuildMaxPairs];
    __shared__ uint16_t ord0[kBuildMaxPairs];
    __shared__ int qa[kBuildMaxQuarters], qb[kBuildMaxQuarters], qlen[kBuildMaxQuarters], qsorted[kBuildMaxQuarters];
    __shared__ int gs[kResMaxWarps + 1];
    __shared__ int bad_any;
    const int64_t s = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid == 0) bad_any = 0;
    const int64_t row_base = s * n;
    const int Q = G * 4;
    // ---- 1. records per row pair (union of the two column lists), and how many of them have an even column
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
    // ---- 2. row pairs in descending order of their record count (ties by index)
    for (int t = tid; t < np; t += kBuildThreads) {
        const int c = cnt[t];
        int r = 0;
        for (int u = 0; u < np; ++u) r += (cnt[u] > c) || (cnt[u] == c && u < t);
        ord0[r] = (uint16_t)t;
    }
    __syncthreads();
    // ---- 3. quarters: the nsplit longest row pairs alone (split), the others two by two
    for (int q = tid; q < Q; q += kBuildThreads) {
        int a = -1, b = -1, len = 0;
        if (q < nsplit) {
            a = ord0[q];
            b = -2;                                   // marks a split quarter
            len = max(cnt_even[a], cnt[a] - cnt_even[a]);
        } else {
            const int r = nsplit + 2 * (q - nsplit);
            if (r < np) { a = ord0[r]; len = cnt[a]; }
            if (r + 1 < np) b = ord0[r + 1];
        }
        qa[q] = a; qb[q] = b; qlen[q] = len;
    }
    __syncthreads();
    for (int q = tid; q < Q; q += kBuildThreads) {
        const int c = qlen[q];
        int r = 0;
        for (int u = 0; u < Q; ++u) r += (qlen[u] > c) || (qlen[u] == c && u < q);
        qsorted[r] = q;
    }
    __syncthreads();
    // ---- 4. iterations per group = longest quarter of the group, rounded up to an even count
    if (tid == 0) {
        int acc = 0;
        for (int g = 0; g < G; ++g) {
            gs[g] = acc;
            acc += (qlen[qsorted[4 * g]] + 1) & ~1;
        }
        gs[G] = acc;
    }
    __syncthreads();
    int32_t* gst = gstart + s * (int64_t)(G + 1);
    for (int g = tid; g <= G; g += kBuildThreads) gst[g] = gs[g];
    // slot table: row pair id, bit 15 = half of a split row pair (even slot: even columns, owner of the rows)
    uint16_t* ord = order + s * (int64_t)G * 8;
    for (int r = tid; r < Q; r += kBuildThreads) {
        const int q = qsorted[r];
        uint16_t oa = 0xFFFF, ob = 0xFFFF;
        if (qb[q] == -2) { oa = (uint16_t)(qa[q] | 0x8000); ob = oa; }
        else {
            if (qa[q] >= 0) oa = (uint16_t)qa[q];
            if (qb[q] >= 0) ob = (uint16_t)qb[q];
        }
        ord[2 * r] = oa;
        ord[2 * r + 1] = ob;
    }
    // ---- 5. fill: (group, iteration, slot) interleaved records
    const int64_t start = ((2 * rowptr[row_base] + rec_pad * s) + 7) & ~(int64_t)7;
    ResRec* base = rec + start;
    bool bad = false;
    ResRec zrec;
    zrec.col = 0; zrec.pad = 0;
    zrec.v0[0] = zrec.v0[1] = zrec.v0[2] = 0.f;
    zrec.v1[0] = zrec.v1[1] = zrec.v1[2] = 0.f;
    for (int r = tid; r < Q; r += kBuildThreads) {
        const int q = qsorted[r];
        const int g = r >> 2, m = r & 3;
        const int iters = gs[g + 1] - gs[g];
        const bool split = qb[q] == -2;
        const int tA = qa[q], tB = split ? qa[q] : qb[q];
        ParityCursor cur[2][2];     // [slot A|B][even|odd]
        int rem[2][2] = {{0, 0}, {0, 0}};
        if (tA >= 0) {
            cursor_init(cur[0][0], rowptr, row_base + 2 * tA, 2 * tA + 1 < n, 0);
            cursor_init(cur[0][1], rowptr, row_base + 2 * tA, 2 * tA + 1 < n, 1);
            rem[0][0] = cnt_even[tA];
            rem[0][1] = split ? 0 : cnt[tA] - cnt_even[tA];
        }
        if (tB >= 0) {
            cursor_init(cur[1][0], rowptr, row_base + 2 * tB, 2 * tB + 1 < n, 0);
            cursor_init(cur[1][1], rowptr, row_base + 2 * tB, 2 * tB + 1 < n, 1);
            rem[1][0] = split ? 0 : cnt_even[tB];
            rem[1][1] = cnt[tB] - cnt_even[tB];
        }
        ResRec* out = base + (int64_t)gs[g] * 8 + 2 * m;
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
            e.col = pA;                                            // padding: a harmless column of the free parity
            if (actA) { cursor_next(cur[0][pA], col, offdiag, e, bad); --rem[0][pA]; }
            out[(int64_t)i * 8] = e;
            e = zrec;
            e.col = pB;
            if (actB) { cursor_next(cur[1][pB], col, offdiag, e, bad); --rem[1][pB]; }
            out[(int64_t)i * 8 + 1] = e;
        }
    }
    // ---- 6. diagonal blocks in single precision, 12 floats per node
    for (int q = tid; q < n * 12; q += kBuildThreads) {
        const int i = q / 12, k = q % 12;
        diag32[(row_base + i) * 12 + k] = k < 9 ? (float)diag[(row_base + i) * 9 + k] : 0.f;
    }
    if (bad) bad_any = 1;
    __syncthreads();
    if (tid == 0 && bad_any) atomicExch(flag, 1);
}

// other widths (8 or 4 columns per CTA, larger structures): slots sorted by length, no splitting
__global__ void __launch_bounds__(kBuildThreads)
resident_build_kernel(int n, int np, int G, int rpw, int64_t rec_pad, const int64_t* __restrict__ rowptr,
                      const int32_t* __restrict__ col, const double* __restrict__ offdiag,
                      const double* __restrict__ diag, ResRec* __restrict__ rec, int32_t* __restrict__ gstart,
                      uint16_t* __restrict__ order, float* __restrict__ diag32, int32_t* __restrict__ flag) {
    __shared__ int cnt[kBuildMaxPairs];
    __shared__ int gs[kResMaxWarps + 1];
    __shared__ int bad_any;
    const int64_t s = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid == 0) bad_any = 0;
    const int64_t row_base = s * n;
    for (int t = tid; t < np; t += kBuildThreads) {
        const int64_t r0 = row_base + 2 * t;
        const bool has1 = 2 * t + 1 < n;
        int64_t a = rowptr[r0], ae = rowptr[r0 + 1];
        int64_t b = has1 ? rowptr[r0 + 1] : 0, be = has1 ? rowptr[r0 + 2] : 0;
        int c = 0;
        while (a < ae || b < be) {
            const int ca = a < ae ? col[a] : 0x7fffffff;
            const int cb = b < be ? col[b] : 0x7fffffff;
            const int m = min(ca, cb);
            a += (ca == m);
            b += (cb == m);
            ++c;
        }
        cnt[t] = c;
    }
    __syncthreads();
    uint16_t* ord = order + s * (int64_t)G * rpw;
    for (int t = tid; t < np; t += kBuildThreads) {
        const int c = cnt[t];
        int r = 0;
        for (int u = 0; u < np; ++u) r += (cnt[u] > c) || (cnt[u] == c && u < t);
        ord[r] = (uint16_t)t;
    }
    for (int r = np + tid; r < G * rpw; r += kBuildThreads) ord[r] = 0xFFFF;
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int g = 0; g < G; ++g) {
            gs[g] = acc;
            acc += (cnt[ord[g * rpw]] + 1) & ~1;   // ranks are dense: the first slot of a group holds its longest list
        }
        gs[G] = acc;
    }
    __syncthreads();
    int32_t* gst = gstart + s * (int64_t)(G + 1);
    for (int g = tid; g <= G; g += kBuildThreads) gst[g] = gs[g];
    const int64_t start = ((rowptr[row_base] + rec_pad * s) + 7) & ~(int64_t)7;
    ResRec* base = rec + start;
    bool bad = false;
    ResRec zrec;
    zrec.col = 0; zrec.pad = 0;
    zrec.v0[0] = zrec.v0[1] = zrec.v0[2] = 0.f;
    zrec.v1[0] = zrec.v1[1] = zrec.v1[2] = 0.f;
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

// groups (= warps) per structure and, for 16-column CTAs, how many of the longest row pairs are split in two
void resident_shape(int n, int cols, int* G, int* nsplit, int* rec_mul, int64_t* rec_pad) {
    const int np = (n + 1) / 2;
    const int rpw = 128 / cols;
    if (cols == 16) {
        int g = (2 * np + rpw - 1) / rpw;
        if (g > kResMaxWarps) g = kResMaxWarps;
        const int gmin = (np + rpw - 1) / rpw;
        if (g < gmin) g = gmin;
        const int Q = g * 4;
        int s = 2 * Q - np - 1;                       // s + ceil((np - s) / 2) <= Q
        s = s < 0 ? 0 : (s > np ? np : s);
        *G = g; *nsplit = s; *rec_mul = 2;
        // every slot of a group is padded to the group's longest list: <= 2 x (records of the previous group) + the
        // first group (<= n + 1 per slot) + the even rounding + alignment
        *rec_pad = (int64_t)rpw * (n + 2) + (int64_t)g * rpw + 8;
    } else {
        const int g = (np + rpw - 1) / rpw;
        *G = g; *nsplit = 0; *rec_mul = 1;
        *rec_pad = (int64_t)rpw * (n + 2) + (int64_t)g * rpw + 8;
    }
}

size_t resident_capacity(int B, int n, int64_t P, int cols) {
    int G, nsplit, mul;
    int64_t pad;
    resident_shape(n, cols, &G, &nsplit, &mul, &pad);
    return (size_t)mul * (size_t)P + (size_t)pad * B + 16;
}

int resident_build(int B, int n, const int64_t* rowptr, const int32_t* col, const double* offdiag,
                   const double* diag, const ResLayout& L, cudaStream_t st) {
    const int np = (n + 1) / 2;
    if (np > kBuildMaxPairs || L.G > kResMaxWarps) return SCB_ERR_UNSUPPORTED;
    SCB_CUDA(cudaMemsetAsync(L.flag, 0, sizeof(int32_t), st));
    if (L.cols == 16)
        resident_build16_kernel<<<B, kBuildThreads, 0, st>>>(n, np, L.G, L.nsplit, L.rec_pad, rowptr, col, offdiag, diag,
                                                             L.rec, L.gstart, L.order, L.diag32, L.flag);
    else
        resident_build_kernel<<<B, kBuildThreads, 0, st>>>(n, np, L.G, L.rpw, L.rec_pad, rowptr, col, offdiag, diag,
                                                           L.rec, L.gstart, L.order, L.diag32, L.flag);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

}  // namespace scb
