// Structure-resident Chebyshev filter in residual form (see resident.cuh).
//
// One CTA per (structure, group of COLS block columns).  The CTA keeps three N x COLS single-precision arrays in
// shared memory for the whole filter: z_k (gather source), z_{k-1} (overwritten by z_{k+1}) and the normalised
// residual.  A warp owns one group of row pairs; a lane owns (row pair, 4 columns) and walks the pair's records:
//     s = v . z_col      (3 FMA per row and column)        acc += v s   (3 FMA)      [H_ij = -v v^T]
// then z_{k+1} = A_k ((H - c) z_k + r) - B_k z_{k-1} with per-column coefficients A_k = 2 rho_k / e,
// B_k = rho_{k-1} rho_k, rho_k = T_k(x)/T_{k+1}(x) at x = (theta - c)/e.  After `degree` steps the lane writes
// X + |r| z (FP64), deflated against the analytic null space, back to global memory.
#include <stdlib.h>

#include "resident.cuh"

namespace scb {

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// one record against the lane's 4 columns: acc[h][a] (+)= v_h[a] * (v_h . z)
template <int COLS>
__device__ __forceinline__ void rec_apply(const float4 ra, const float4 rb, const float* __restrict__ ZA, int c0,
                                          float2 (&acc)[2][3][2]) {
    const int col = __float_as_int(ra.x);
    const float4* zp = reinterpret_cast<const float4*>(ZA + col * (3 * COLS) + c0);
    const float4 z0 = zp[0], z1 = zp[COLS / 4], z2 = zp[2 * (COLS / 4)];
    const float2 z0l = f2(z0.x, z0.y), z0h = f2(z0.z, z0.w);
    const float2 z1l = f2(z1.x, z1.y), z1h = f2(z1.z, z1.w);
    const float2 z2l = f2(z2.x, z2.y), z2h = f2(z2.z, z2.w);
    {
        const float2 vx = f2(ra.y, ra.y), vy = f2(ra.z, ra.z), vz = f2(ra.w, ra.w);
        float2 sl = __fmul2_rn(vx, z0l), sh = __fmul2_rn(vx, z0h);
        sl = __ffma2_rn(vy, z1l, sl); sh = __ffma2_rn(vy, z1h, sh);
        sl = __ffma2_rn(vz, z2l, sl); sh = __ffma2_rn(vz, z2h, sh);
        acc[0][0][0] = __ffma2_rn(vx, sl, acc[0][0][0]); acc[0][0][1] = __ffma2_rn(vx, sh, acc[0][0][1]);
        acc[0][1][0] = __ffma2_rn(vy, sl, acc[0][1][0]); acc[0][1][1] = __ffma2_rn(vy, sh, acc[0][1][1]);
        acc[0][2][0] = __ffma2_rn(vz, sl, acc[0][2][0]); acc[0][2][1] = __ffma2_rn(vz, sh, acc[0][2][1]);
    }
    {
        const float2 vx = f2(rb.x, rb.x), vy = f2(rb.y, rb.y), vz = f2(rb.z, rb.z);
        float2 sl = __fmul2_rn(vx, z0l), sh = __fmul2_rn(vx, z0h);
        sl = __ffma2_rn(vy, z1l, sl); sh = __ffma2_rn(vy, z1h, sh);
        sl = __ffma2_rn(vz, z2l, sl); sh = __ffma2_rn(vz, z2h, sh);
        acc[1][0][0] = __ffma2_rn(vx, sl, acc[1][0][0]); acc[1][0][1] = __ffma2_rn(vx, sh, acc[1][0][1]);
        acc[1][1][0] = __ffma2_rn(vy, sl, acc[1][1][0]); acc[1][1][1] = __ffma2_rn(vy, sh, acc[1][1][1]);
        acc[1][2][0] = __ffma2_rn(vz, sl, acc[1][2][0]); acc[1][2][1] = __ffma2_rn(vz, sh, acc[1][2][1]);
    }
}

// split row pairs: add the partial sums of the two slots of a quarter (lanes l and l ^ LPP); warp-uniform call
template <int LPP>
__device__ __forceinline__ void merge_split(float2 (&acc)[2][3][2], bool split) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float px = __shfl_xor_sync(0xffffffffu, acc[h][a][u].x, LPP);
                const float py = __shfl_xor_sync(0xffffffffu, acc[h][a][u].y, LPP);
                if (split) { acc[h][a][u].x += px; acc[h][a][u].y += py; }
            }
}

// warp -> group: the longest groups go to the highest warp ids (issue priority), snaked over the 4 sub-partitions
__device__ __forceinline__ int warp_group(int warp, int nwarps) {
    const int kq = warp >> 2, smsp = warp & 3;
    const int k = (nwarps >> 2) - 1 - kq;
    const int r = (k & 1) ? smsp : 3 - smsp;
    return 4 * k + r;
}

// acc (= sum v (v.z) over the records, i.e. -(H_offdiag z)) -> (H z) rows of the lane's two nodes, 4 columns
template <int COLS>
__device__ __forceinline__ void walk_records(const float4* __restrict__ rp, int iters, int stride4,
                                             const float* __restrict__ ZA, int c0, float2 (&acc)[2][3][2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int a = 0; a < 3; ++a) { acc[h][a][0] = f2(0.f, 0.f); acc[h][a][1] = f2(0.f, 0.f); }
    if (iters <= 0) return;
    float4 r0a = __ldg(rp), r0b = __ldg(rp + 1);
    float4 r1a = __ldg(rp + stride4), r1b = __ldg(rp + stride4 + 1);
    for (int it = 0; it < iters; it += 2) {
        float4 n0a = r0a, n0b = r0b, n1a = r1a, n1b = r1b;
        if (it + 2 < iters) {
            const float4* np = rp + (size_t)(it + 2) * stride4;
            n0a = __ldg(np); n0b = __ldg(np + 1);
            n1a = __ldg(np + stride4); n1b = __ldg(np + stride4 + 1);
        }
        rec_apply<COLS>(r0a, r0b, ZA, c0, acc);
        rec_apply<COLS>(r1a, r1b, ZA, c0, acc);
        r0a = n0a; r0b = n0b; r1a = n1a; r1b = n1b;
    }
}

// (H z)[row 3i+a] for node i from the negated record sum and the diagonal block
__device__ __forceinline__ void add_diag(const float* __restrict__ dg, const float4 (&zo)[3], const float2 (&acc)[3][2],
                                         float4 (&hz)[3]) {
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(dg));
    const float4 d1 = __ldg(reinterpret_cast<const float4*>(dg) + 1);
    const float4 d2 = __ldg(reinterpret_cast<const float4*>(dg) + 2);
    const float D[9] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        hz[a].x = fmaf(D[3 * a + 2], zo[2].x, fmaf(D[3 * a + 1], zo[1].x, fmaf(D[3 * a], zo[0].x, -acc[a][0].x)));
        hz[a].y = fmaf(D[3 * a + 2], zo[2].y, fmaf(D[3 * a + 1], zo[1].y, fmaf(D[3 * a], zo[0].y, -acc[a][0].y)));
        hz[a].z = fmaf(D[3 * a + 2], zo[2].z, fmaf(D[3 * a + 1], zo[1].z, fmaf(D[3 * a], zo[0].z, -acc[a][1].x)));
        hz[a].w = fmaf(D[3 * a + 2], zo[2].w, fmaf(D[3 * a + 1], zo[1].w, fmaf(D[3 * a], zo[0].w, -acc[a][1].y)));
    }
}

// what a lane needs to know about one of its warp's groups
struct LaneCtx {
    const float4* rp;      // first record of the lane's slot
    const float* dg0;      // diagonal block of the first node
    int iters;             // warp iterations of the group
    int i0;                // first node of the row pair
    int nh;                // nodes owned in this group: 0 (empty slot / second half of a split pair), 1 (last pair of an odd n) or 2
    bool split;            // the slot holds half of a split row pair: partial sums are added across the quarter
};

template <int COLS>
__device__ __forceinline__ LaneCtx lane_ctx(int g, int G, int slot, int64_t s, int n, const ResRec* __restrict__ rec,
                                            int64_t start, const int32_t* __restrict__ gst,
                                            const uint16_t* __restrict__ order, const float* __restrict__ diag32) {
    constexpr int RPW = 32 / (COLS / 4);
    LaneCtx c;
    c.rp = nullptr; c.dg0 = diag32; c.iters = 0; c.i0 = 0; c.nh = 0; c.split = false;
    if (g < 0 || g >= G) return c;
    const int praw = order[s * (int64_t)G * RPW + g * RPW + slot];
    const int g0 = gst[g];
    c.iters = gst[g + 1] - g0;
    c.rp = reinterpret_cast<const float4*>(rec + start + (int64_t)g0 * RPW + slot);
    if (praw != 0xFFFF) {
        const int p = praw & 0x7FFF;
        c.split = (praw & 0x8000) != 0;
        c.i0 = 2 * p;
        c.dg0 = diag32 + (s * n + c.i0) * 12;
        if (!c.split || !(slot & 1)) c.nh = (c.i0 + 1 < n) ? 2 : 1;   // the even slot of a split pair owns the rows
    }
    return c;
}

// groups of a warp: gpw == 1: one group, longest groups on the highest warp ids; gpw == 2: the k-th longest and the
// k-th shortest group share a warp (balanced totals: the CTA waits for its slowest warp at every filter step)
__device__ __forceinline__ void warp_groups(int warp, int nwarps, int G, int gpw, int (&gsel)[2]) {
    if (gpw == 1) {
        gsel[0] = warp_group(warp, nwarps);
        gsel[1] = -1;
        return;
    }
    const int k = nwarps - 1 - warp;           // heaviest pairs on the highest warp ids
    const int half = (G + 1) / 2;
    gsel[0] = k < half ? k : -1;
    gsel[1] = (k < half && G - 1 - k > k) ? G - 1 - k : -1;
}

template <int COLS, int GPW>
__global__ void __launch_bounds__(kResMaxWarps * 32, 1)
resident_filter_kernel(int n, int b, int G, int rec_mul, int64_t rec_pad, const int64_t* __restrict__ rowptr,
                       const ResRec* __restrict__ rec, const int32_t* __restrict__ gstart,
                       const uint16_t* __restrict__ order, const float* __restrict__ diag32,
                       const double* X, const double* __restrict__ HX, const double* __restrict__ theta,
                       const double* __restrict__ rn2, const EigState* __restrict__ state,
                       const int32_t* __restrict__ done, const double* __restrict__ Zr, int nz, double* Xout,
                       unsigned long long* __restrict__ app_counter, int kwant, double tol) {
    constexpr int LPP = COLS / 4;        // lanes per row pair
    constexpr int RPW = 32 / LPP;        // row pairs per warp (group size)
    extern __shared__ __align__(16) float res_smem[];
    const int64_t s = blockIdx.y;
    if (done && done[s]) return;
    const int N = 3 * n;
    float* ZA = res_smem;
    float* ZB = ZA + (size_t)N * COLS;
    float* RH = ZB + (size_t)N * COLS;
    float* cA = RH + (size_t)N * COLS;               // [kResDegreeCap][COLS]
    float* cB = cA + kResDegreeCap * COLS;
    __shared__ double s_th[COLS], s_nrm[COLS], s_inv[COLS];
    __shared__ float s_sc0[COLS];
    __shared__ double s_p[8 * COLS];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nwarps = blockDim.x >> 5;
    const int cg = blockIdx.x;
    const EigState e = state[s];
    int deg = e.degree_next;
    deg = deg < 2 ? 2 : (deg > kResDegreeCap ? kResDegreeCap : deg);
    // A column group whose columns are all WANTED modes that already meet the tolerance (with a margin) needs no
    // further correction: the lowest modes converge first, so in the last outer iterations the CTA of columns 0..15
    // retires while the group holding the highest wanted modes and the guard vectors keeps filtering.
    if ((cg + 1) * COLS <= kwant) {
        const double scale = fmax(fabs(theta[s * b + kwant - 1]), 1e-6 * e.ub);
        const double lim = 0.1 * tol * scale;
        bool all_done = true;
        for (int c = 0; c < COLS; ++c) all_done = all_done && (rn2[s * b + cg * COLS + c] <= lim * lim);
        if (all_done) return;
    }
    // bench accounting: operator applications per structure (integer adds commute: the total is reproducible)
    if (app_counter && tid == 0 && cg == 0) atomicAdd(app_counter, (unsigned long long)(deg - 1));
    const double ehalf = 0.5 * (e.ub - e.lo), cmid = 0.5 * (e.ub + e.lo);
    if (tid < COLS) {
        const int gc = cg * COLS + tid;
        const double th = fmin(theta[s * b + gc], e.lo);
        const double nr = sqrt(fmax(rn2[s * b + gc], 0.0));
        s_th[tid] = th;
        s_nrm[tid] = nr;
        s_inv[tid] = nr > 0.0 ? 1.0 / nr : 0.0;
        const double x = (th - cmid) / ehalf;        // <= -1
        double rho = 1.0 / x;
        s_sc0[tid] = (float)(rho / ehalf);
        for (int k = 1; k < deg; ++k) {
            const double rn = 1.0 / (2.0 * x - rho);
            cA[k * COLS + tid] = (float)(2.0 * rn / ehalf);
            cB[k * COLS + tid] = (float)(rho * rn);
            rho = rn;
        }
    }
    __syncthreads();

    int gsel[2];
    warp_groups(warp, nwarps, G, GPW, gsel);
    const int slot = lane / LPP, q = lane % LPP;
    const int c0 = 4 * q;                             // first local column of the lane
    const int gc0 = cg * COLS + c0;                   // ... in the block
    const double* Xs = X + s * (int64_t)N * b;
    const double* Hs = HX + s * (int64_t)N * b;
    const int64_t start = ((rec_mul * rowptr[s * n] + rec_pad * s) + 7) & ~(int64_t)7;
    const int32_t* gst = gstart + s * (int64_t)(G + 1);

    LaneCtx ctx[GPW];
    bool wsplit[GPW];
#pragma unroll
    for (int gi = 0; gi < GPW; ++gi) {
        ctx[gi] = lane_ctx<COLS>(gsel[gi], G, slot, s, n, rec, start, gst, order, diag32);
        wsplit[gi] = __any_sync(0xffffffffu, ctx[gi].split);
    }

    // ---- prologue: residual of the lane's rows -> RH, z_1 -> ZA, z_0 = 0 -> ZB
#pragma unroll
    for (int gi = 0; gi < GPW; ++gi) {
        const LaneCtx c = ctx[gi];
        for (int h = 0; h < c.nh; ++h) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int row = 3 * (c.i0 + h) + a;
                const double2 xa = *reinterpret_cast<const double2*>(Xs + (int64_t)row * b + gc0);
                const double2 xb = *reinterpret_cast<const double2*>(Xs + (int64_t)row * b + gc0 + 2);
                const double2 ha = *reinterpret_cast<const double2*>(Hs + (int64_t)row * b + gc0);
                const double2 hb = *reinterpret_cast<const double2*>(Hs + (int64_t)row * b + gc0 + 2);
                float4 r;
                r.x = (float)((ha.x - s_th[c0 + 0] * xa.x) * s_inv[c0 + 0]);
                r.y = (float)((ha.y - s_th[c0 + 1] * xa.y) * s_inv[c0 + 1]);
                r.z = (float)((hb.x - s_th[c0 + 2] * xb.x) * s_inv[c0 + 2]);
                r.w = (float)((hb.y - s_th[c0 + 3] * xb.y) * s_inv[c0 + 3]);
                const float4 sc = *reinterpret_cast<const float4*>(s_sc0 + c0);
                *reinterpret_cast<float4*>(RH + row * COLS + c0) = r;
                *reinterpret_cast<float4*>(ZA + row * COLS + c0) = make_float4(r.x * sc.x, r.y * sc.y, r.z * sc.z, r.w * sc.w);
                *reinterpret_cast<float4*>(ZB + row * COLS + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    __syncthreads();

    constexpr int stride4 = RPW * 2;                  // float4 per iteration of the group
    const float cm = (float)cmid;
    for (int k = 1; k < deg; ++k) {
        const float4 A4 = *reinterpret_cast<const float4*>(cA + k * COLS + c0);
        const float4 B4 = *reinterpret_cast<const float4*>(cB + k * COLS + c0);
#pragma unroll
        for (int gi = 0; gi < GPW; ++gi) {
            if (gsel[gi] < 0) continue;
            const LaneCtx c = ctx[gi];
            float2 acc[2][3][2];
            walk_records<COLS>(c.rp, c.iters, stride4, ZA, c0, acc);
            if (wsplit[gi]) merge_split<LPP>(acc, c.split);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h >= c.nh) break;
                float4 zo[3], hz[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) zo[a] = *reinterpret_cast<const float4*>(ZA + (3 * (c.i0 + h) + a) * COLS + c0);
                add_diag(c.dg0 + 12 * h, zo, acc[h], hz);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const int off = (3 * (c.i0 + h) + a) * COLS + c0;
                    const float4 w = *reinterpret_cast<const float4*>(ZB + off);
                    const float4 r = *reinterpret_cast<const float4*>(RH + off);
                    float4 o;
                    o.x = A4.x * (fmaf(-cm, zo[a].x, hz[a].x) + r.x) - B4.x * w.x;
                    o.y = A4.y * (fmaf(-cm, zo[a].y, hz[a].y) + r.y) - B4.y * w.y;
                    o.z = A4.z * (fmaf(-cm, zo[a].z, hz[a].z) + r.z) - B4.z * w.z;
                    o.w = A4.w * (fmaf(-cm, zo[a].w, hz[a].w) + r.w) - B4.w * w.w;
                    *reinterpret_cast<float4*>(ZB + off) = o;
                }
            }
        }
        __syncthreads();
        float* t = ZA; ZA = ZB; ZB = t;
    }

    // ---- epilogue: X + |r| z  (FP64), deflated against the analytic null space (nz <= 8 vectors)
    double* Xo = Xout + s * (int64_t)N * b;
    const double* Zs = (Zr && nz > 0) ? Zr + s * (int64_t)N * nz : nullptr;
    double pz[8][4];
#pragma unroll
    for (int z = 0; z < 8; ++z)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) pz[z][cc] = 0.0;
#pragma unroll 1
    for (int gi = 0; gi < GPW; ++gi) {
        const LaneCtx c = ctx[gi];
        for (int h = 0; h < c.nh; ++h) {
#pragma unroll 1
            for (int a = 0; a < 3; ++a) {
                const int row = 3 * (c.i0 + h) + a;
                const float4 zf = *reinterpret_cast<const float4*>(ZA + row * COLS + c0);
                const double2 xa = *reinterpret_cast<const double2*>(Xs + (int64_t)row * b + gc0);
                const double2 xb = *reinterpret_cast<const double2*>(Xs + (int64_t)row * b + gc0 + 2);
                double xn[4];
                xn[0] = fma(s_nrm[c0 + 0], (double)zf.x, xa.x);
                xn[1] = fma(s_nrm[c0 + 1], (double)zf.y, xa.y);
                xn[2] = fma(s_nrm[c0 + 2], (double)zf.z, xb.x);
                xn[3] = fma(s_nrm[c0 + 3], (double)zf.w, xb.y);
                *reinterpret_cast<double2*>(Xo + (int64_t)row * b + gc0) = make_double2(xn[0], xn[1]);
                *reinterpret_cast<double2*>(Xo + (int64_t)row * b + gc0 + 2) = make_double2(xn[2], xn[3]);
                if (Zs) {
#pragma unroll
                    for (int z = 0; z < 8; ++z)
                        if (z < nz) {
                            const double zv = Zs[(int64_t)row * nz + z];
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) pz[z][cc] = fma(zv, xn[cc], pz[z][cc]);
                        }
                }
            }
        }
    }
    if (!Zs) return;
    // fixed-order reduction: lanes of a warp (same q), then warps in index order
    double* red = reinterpret_cast<double*>(ZB);      // [nwarps][LPP][8][4]  (ZB is free after the last barrier)
#pragma unroll
    for (int z = 0; z < 8; ++z)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            double v = pz[z][cc];
#pragma unroll
            for (int o = LPP; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            pz[z][cc] = v;
        }
    if (slot == 0) {
#pragma unroll
        for (int z = 0; z < 8; ++z)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) red[((warp * LPP + q) * 8 + z) * 4 + cc] = pz[z][cc];
    }
    __syncthreads();
    if (tid < 8 * COLS) {
        const int z = tid / COLS, c = tid % COLS;
        double t = 0.0;
        for (int w = 0; w < nwarps; ++w) t += red[((w * LPP + c / 4) * 8 + z) * 4 + (c & 3)];
        s_p[z * COLS + c] = t;
    }
    __syncthreads();
#pragma unroll 1
    for (int gi = 0; gi < GPW; ++gi) {
        const LaneCtx c = ctx[gi];
        for (int h = 0; h < c.nh; ++h) {
#pragma unroll 1
            for (int a = 0; a < 3; ++a) {
                const int row = 3 * (c.i0 + h) + a;
                double2 xa = *reinterpret_cast<const double2*>(Xo + (int64_t)row * b + gc0);
                double2 xb = *reinterpret_cast<const double2*>(Xo + (int64_t)row * b + gc0 + 2);
#pragma unroll
                for (int z = 0; z < 8; ++z)
                    if (z < nz) {
                        const double zv = Zs[(int64_t)row * nz + z];
                        xa.x = fma(-zv, s_p[z * COLS + c0 + 0], xa.x);
                        xa.y = fma(-zv, s_p[z * COLS + c0 + 1], xa.y);
                        xb.x = fma(-zv, s_p[z * COLS + c0 + 2], xb.x);
                        xb.y = fma(-zv, s_p[z * COLS + c0 + 3], xb.y);
                    }
                *reinterpret_cast<double2*>(Xo + (int64_t)row * b + gc0) = xa;
                *reinterpret_cast<double2*>(Xo + (int64_t)row * b + gc0 + 2) = xb;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Spectrum upper bound: `steps` steps of column-wise Lanczos in single precision with the block of one
// structure resident in shared memory (every column is an independent Lanczos run from a random vector).
// ---------------------------------------------------------------------------
// sum over the whole CTA of per-lane partial sums for the lane's 4 columns (fixed order: lanes, then warps)
template <int COLS>
__device__ __forceinline__ float4 cta_colsum(float4 v, float* red, int warp, int nwarps, int slot, int q) {
    constexpr int LPP = COLS / 4;
#pragma unroll
    for (int o = LPP; o < 32; o <<= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
        v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
    }
    if (slot == 0) *reinterpret_cast<float4*>(red + warp * COLS + 4 * q) = v;
    __syncthreads();
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w = 0; w < nwarps; ++w) {
        const float4 u = *reinterpret_cast<const float4*>(red + w * COLS + 4 * q);
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    __syncthreads();
    return t;
}

constexpr int kResLanczosMax = 16;

template <int COLS>
__global__ void __launch_bounds__(kResMaxWarps * 32, 1)
resident_lanczos_kernel(int n, int b, int G, int rec_mul, int64_t rec_pad, const int64_t* __restrict__ rowptr,
                        const ResRec* __restrict__ rec, const int32_t* __restrict__ gstart,
                        const uint16_t* __restrict__ order, const float* __restrict__ diag32, int steps,
                        uint64_t seed, double* __restrict__ est) {
    constexpr int LPP = COLS / 4;
    constexpr int RPW = 32 / LPP;
    extern __shared__ __align__(16) float res_smem[];
    const int64_t s = blockIdx.y;
    const int N = 3 * n;
    float* VA = res_smem;                              // v_j (gather source)
    float* VB = VA + (size_t)N * COLS;                 // v_{j+1}
    float* VP = VB + (size_t)N * COLS;                 // v_{j-1}
    float* red = VP + (size_t)N * COLS;                // [nwarps][COLS]
    __shared__ float s_alpha[kResLanczosMax][COLS], s_beta2[kResLanczosMax][COLS];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nwarps = blockDim.x >> 5;
    const int cg = blockIdx.x;
    const int g = warp_group(warp, nwarps);
    const int slot = lane / LPP, q = lane % LPP;
    const int c0 = 4 * q, gc0 = cg * COLS + c0;
    const int64_t start = ((rec_mul * rowptr[s * n] + rec_pad * s) + 7) & ~(int64_t)7;
    const int32_t* gst = gstart + s * (int64_t)(G + 1);
    const LaneCtx ctx = lane_ctx<COLS>(g, G, slot, s, n, rec, start, gst, order, diag32);
    const bool wsplit = __any_sync(0xffffffffu, ctx.split);
    const int i0 = ctx.i0, nh = ctx.nh;

    // random start, normalised per column
    float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int h = 0; h < nh; ++h)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int row = 3 * (i0 + h) + a;
            const uint64_t idx = ((uint64_t)s * N + row) * b + gc0;
            float4 v;
            v.x = (float)uniform_pm1(seed, idx); v.y = (float)uniform_pm1(seed, idx + 1);
            v.z = (float)uniform_pm1(seed, idx + 2); v.w = (float)uniform_pm1(seed, idx + 3);
            *reinterpret_cast<float4*>(VA + row * COLS + c0) = v;
            *reinterpret_cast<float4*>(VP + row * COLS + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
            part.x = fmaf(v.x, v.x, part.x); part.y = fmaf(v.y, v.y, part.y);
            part.z = fmaf(v.z, v.z, part.z); part.w = fmaf(v.w, v.w, part.w);
        }
    {
        const float4 t = cta_colsum<COLS>(part, red, warp, nwarps, slot, q);
        const float4 inv = make_float4(rsqrtf(fmaxf(t.x, 1e-30f)), rsqrtf(fmaxf(t.y, 1e-30f)),
                                       rsqrtf(fmaxf(t.z, 1e-30f)), rsqrtf(fmaxf(t.w, 1e-30f)));
        for (int h = 0; h < nh; ++h)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float4* vp = reinterpret_cast<float4*>(VA + (3 * (i0 + h) + a) * COLS + c0);
                float4 v = *vp;
                v.x *= inv.x; v.y *= inv.y; v.z *= inv.z; v.w *= inv.w;
                *vp = v;
            }
    }
    __syncthreads();

    constexpr int stride4 = RPW * 2;
    const float4* rp = ctx.rp;
    const int iters = ctx.iters;
    const float* dg0 = ctx.dg0;
    float4 bprev = make_float4(0.f, 0.f, 0.f, 0.f);      // beta_{j-1} per column

    for (int j = 0; j < steps; ++j) {
        float4 w[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int a = 0; a < 3; ++a) w[h][a] = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 pa = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < G) {
            float2 acc[2][3][2];
            walk_records<COLS>(rp, iters, stride4, VA, c0, acc);
            if (wsplit) merge_split<LPP>(acc, ctx.split);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h >= nh) break;
                float4 vo[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) vo[a] = *reinterpret_cast<const float4*>(VA + (3 * (i0 + h) + a) * COLS + c0);
                add_diag(dg0 + 12 * h, vo, acc[h], w[h]);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    pa.x = fmaf(vo[a].x, w[h][a].x, pa.x); pa.y = fmaf(vo[a].y, w[h][a].y, pa.y);
                    pa.z = fmaf(vo[a].z, w[h][a].z, pa.z); pa.w = fmaf(vo[a].w, w[h][a].w, pa.w);
                }
            }
        }
        const float4 alpha = cta_colsum<COLS>(pa, red, warp, nwarps, slot, q);
        float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h >= nh) break;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int off = (3 * (i0 + h) + a) * COLS + c0;
                const float4 vo = *reinterpret_cast<const float4*>(VA + off);
                const float4 vp = *reinterpret_cast<const float4*>(VP + off);
                float4 x = w[h][a];
                x.x = x.x - alpha.x * vo.x - bprev.x * vp.x; x.y = x.y - alpha.y * vo.y - bprev.y * vp.y;
                x.z = x.z - alpha.z * vo.z - bprev.z * vp.z; x.w = x.w - alpha.w * vo.w - bprev.w * vp.w;
                w[h][a] = x;
                pb.x = fmaf(x.x, x.x, pb.x); pb.y = fmaf(x.y, x.y, pb.y);
                pb.z = fmaf(x.z, x.z, pb.z); pb.w = fmaf(x.w, x.w, pb.w);
            }
        }
        const float4 beta2 = cta_colsum<COLS>(pb, red, warp, nwarps, slot, q);
        if (warp == 0 && slot == 0) {
            *reinterpret_cast<float4*>(&s_alpha[j][c0]) = alpha;
            *reinterpret_cast<float4*>(&s_beta2[j][c0]) = beta2;
        }
        if (j + 1 < steps) {
            const float4 inv = make_float4(rsqrtf(fmaxf(beta2.x, 1e-30f)), rsqrtf(fmaxf(beta2.y, 1e-30f)),
                                           rsqrtf(fmaxf(beta2.z, 1e-30f)), rsqrtf(fmaxf(beta2.w, 1e-30f)));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h >= nh) break;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const int off = (3 * (i0 + h) + a) * COLS + c0;
                    *reinterpret_cast<float4*>(VP + off) = *reinterpret_cast<const float4*>(VA + off);
                    const float4 x = w[h][a];
                    *reinterpret_cast<float4*>(VB + off) = make_float4(x.x * inv.x, x.y * inv.y, x.z * inv.z, x.w * inv.w);
                }
            }
            bprev = make_float4(sqrtf(fmaxf(beta2.x, 0.f)), sqrtf(fmaxf(beta2.y, 0.f)), sqrtf(fmaxf(beta2.z, 0.f)),
                                sqrtf(fmaxf(beta2.w, 0.f)));
            __syncthreads();
            float* t = VA; VA = VB; VB = t;
        }
    }
    __syncthreads();
    // largest eigenvalue of each column's tridiagonal (bisection on the Sturm count), max over the columns
    if (warp == 0) {
        double th = 0.0;
        if (lane < COLS) {
            const int c = lane;
            double lo = 0.0, hi = 0.0;
            for (int j = 0; j < steps; ++j) {
                const double a = s_alpha[j][c];
                const double bl = j > 0 ? sqrt(fmax((double)s_beta2[j - 1][c], 0.0)) : 0.0;
                const double br = j < steps - 1 ? sqrt(fmax((double)s_beta2[j][c], 0.0)) : 0.0;
                hi = (j == 0) ? a + bl + br : fmax(hi, a + bl + br);
                lo = (j == 0) ? a - bl - br : fmin(lo, a - bl - br);
            }
            for (int it = 0; it < 40; ++it) {
                const double x = 0.5 * (lo + hi);
                int below = 0;
                double d = 1.0;
                for (int j = 0; j < steps; ++j) {
                    const double a = s_alpha[j][c];
                    const double b2 = j > 0 ? fmax((double)s_beta2[j - 1][c], 0.0) : 0.0;
                    d = (a - x) - (j > 0 ? b2 / d : 0.0);
                    if (d == 0.0) d = -1e-300;
                    below += d < 0.0;
                }
                if (below >= steps) hi = x; else lo = x;
            }
            th = hi;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) th = fmax(th, __shfl_xor_sync(0xffffffffu, th, o));
        if (lane == 0 && th > 0.0) atomic_max_nonneg(&est[s], th);   // max is order independent: reproducible
    }
}

__global__ void lanczos_apply_kernel(int B, const double* __restrict__ est, double factor, EigState* st) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B) return;
    const double e = factor * est[s];
    if (e > 0.0 && e < st[s].ub) st[s].ub = e;
}

static size_t filter_smem(int n, int cols) {
    return sizeof(float) * ((size_t)3 * 3 * n * cols + 2 * (size_t)kResDegreeCap * cols);
}

template <int COLS>
static int launch_filter(int B, int n, int b, const int64_t* rowptr, const ResLayout& L, const double* X,
                         const double* HX, const double* theta, const double* rn2, const EigState* state,
                         const int32_t* done, const double* Z, int nz, double* Xout, int kwant, double tol,
                         cudaStream_t st) {
    const size_t smem = filter_smem(n, COLS);
    // groups per warp: 2 pairs the k-th longest with the k-th shortest group (balanced warps, half the threads);
    // measured on the C3 batch: 1 group per warp (20 warps) 310 ms per step, 2 groups (10 warps) 317 ms
    int gpw = 1;
    if (const char* env = getenv("SCB_RES_GPW")) gpw = atoi(env) == 2 ? 2 : 1;
    if (L.G < 8) gpw = 1;
    dim3 grid((unsigned)(b / COLS), (unsigned)B);
    // the shared-memory opt-in is per device and cheap: set it on every launch (no process-wide "configured" flag)
    if (gpw == 1) {
        SCB_CUDA(cudaFuncSetAttribute(resident_filter_kernel<COLS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        resident_filter_kernel<COLS, 1><<<grid, 32 * 4 * ((L.G + 3) / 4), smem, st>>>(
            n, b, L.G, L.rec_mul, L.rec_pad, rowptr, L.rec, L.gstart, L.order, L.diag32, X, HX, theta, rn2, state, done, Z, nz,
            Xout, L.app_counter, kwant, tol);
    } else {
        SCB_CUDA(cudaFuncSetAttribute(resident_filter_kernel<COLS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        resident_filter_kernel<COLS, 2><<<grid, 32 * ((L.G + 1) / 2), smem, st>>>(
            n, b, L.G, L.rec_mul, L.rec_pad, rowptr, L.rec, L.gstart, L.order, L.diag32, X, HX, theta, rn2, state, done, Z, nz,
            Xout, L.app_counter, kwant, tol);
    }
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

template <int COLS>
static int launch_lanczos(int B, int n, int b, const int64_t* rowptr, const ResLayout& L, int steps, uint64_t seed,
                          cudaStream_t st) {
    const int nwarps = 4 * ((L.G + 3) / 4);
    const size_t smem = sizeof(float) * ((size_t)3 * 3 * n * COLS + (size_t)nwarps * COLS);
    SCB_CUDA(cudaFuncSetAttribute(resident_lanczos_kernel<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // one column group per structure: COLS independent Lanczos runs bound the spectrum as well as b of them
    (void)b;
    dim3 grid(1u, (unsigned)B);
    resident_lanczos_kernel<COLS><<<grid, 32 * nwarps, smem, st>>>(n, b, L.G, L.rec_mul, L.rec_pad, rowptr, L.rec,
                                                                  L.gstart, L.order, L.diag32, steps, seed, L.est);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

int resident_lanczos(int B, int n, int b, const int64_t* rowptr, const ResLayout& L, int steps, uint64_t seed,
                     double ub_factor, EigState* state, cudaStream_t st) {
    if (steps < 2 || steps > kResLanczosMax) return SCB_ERR_INVALID;
    SCB_CUDA(cudaMemsetAsync(L.est, 0, sizeof(double) * (size_t)B, st));
    int status = SCB_ERR_UNSUPPORTED;
    if (L.cols == 16) status = launch_lanczos<16>(B, n, b, rowptr, L, steps, seed, st);
    else if (L.cols == 8) status = launch_lanczos<8>(B, n, b, rowptr, L, steps, seed, st);
    else if (L.cols == 4) status = launch_lanczos<4>(B, n, b, rowptr, L, steps, seed, st);
    if (status != SCB_OK) return status;
    lanczos_apply_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, st>>>(B, L.est, ub_factor, state);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

int resident_filter(int B, int n, int b, const int64_t* rowptr, const ResLayout& L, const double* X,
                    const double* HX, const double* theta, const double* rn2, const EigState* state,
                    const int32_t* done, const double* Z, int nz, double* Xout, int kwant, double tol,
                    cudaStream_t st) {
    if (nz > 8) return SCB_ERR_UNSUPPORTED;
    if (L.cols == 16) return launch_filter<16>(B, n, b, rowptr, L, X, HX, theta, rn2, state, done, Z, nz, Xout, kwant, tol, st);
    if (L.cols == 8) return launch_filter<8>(B, n, b, rowptr, L, X, HX, theta, rn2, state, done, Z, nz, Xout, kwant, tol, st);
    if (L.cols == 4) return launch_filter<4>(B, n, b, rowptr, L, X, HX, theta, rn2, state, done, Z, nz, Xout, kwant, tol, st);
    return SCB_ERR_UNSUPPORTED;
}

}  // namespace scb
