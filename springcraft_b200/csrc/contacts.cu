// K1: contact search -> CSR neighbour list.
// Replaces springcraft/interaction.py:149-178, 193-213 (+ biotite CellList).
//
// Two variants with identical results:
//   * tiled all-against-all (batched ensembles, small/medium n): coordinates of
//     a j-tile staged in shared memory, one warp per row, warp-ballot compaction
//     -> columns come out ascending, i.e. in the reference's np.where order.
//   * cell list (one large structure): counting sort into cells of edge >= cutoff,
//     27-cell stencil per atom, ballot compaction into a per-row scratch list,
//     rank-by-counting to restore ascending column order.
// Distance criterion everywhere: fp64 ((dx*dx)+(dy*dy))+(dz*dz) <= cutoff^2 with
// every product/sum rounded separately (no FMA contraction).
#include "common.cuh"

namespace scb {

struct PatchView {
    int n_off, n_on;
    const uint8_t* dead;
    const int32_t* off;
    const int32_t* on;
};

static PatchView make_patch_view(const scb_patch* p) {
    PatchView v{0, 0, nullptr, nullptr, nullptr};
    if (p) {
        v.n_off = p->pair_off ? p->n_pair_off : 0;
        v.n_on = p->pair_on ? p->n_pair_on : 0;
        v.dead = p->dead;
        v.off = p->pair_off;
        v.on = p->pair_on;
    }
    return v;
}

__device__ __forceinline__ double sq_dist_rn(double dx, double dy, double dz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// does atom i appear in any switched off / on pair?  (warp-cooperative)
__device__ __forceinline__ bool row_is_patched(const PatchView& pv, int i, unsigned lane) {
    bool hit = false;
    for (int q = lane; q < pv.n_off; q += 32) hit |= (pv.off[2 * q] == i) | (pv.off[2 * q + 1] == i);
    for (int q = lane; q < pv.n_on; q += 32) hit |= (pv.on[2 * q] == i) | (pv.on[2 * q + 1] == i);
    return __any_sync(0xffffffffu, hit);
}

// final adjacency value: on(i,j) || (base && !dead_i && !dead_j && !off(i,j))
// (order shutdown -> pair_off -> pair_on, interaction.py:201-213)
__device__ __forceinline__ bool apply_patch(const PatchView& pv, bool base, int i, int j,
                                            bool dead_i, bool special) {
    if (pv.dead) base = base && !dead_i && !pv.dead[j];
    if (special) {
        for (int q = 0; q < pv.n_off; ++q) {
            int a = pv.off[2 * q], b = pv.off[2 * q + 1];
            if ((a == i && b == j) || (a == j && b == i)) base = false;
        }
        for (int q = 0; q < pv.n_on; ++q) {
            int a = pv.on[2 * q], b = pv.on[2 * q + 1];
            if ((a == i && b == j) || (a == j && b == i)) base = true;
        }
    }
    return base;
}

// ---------------------------------------------------------------------------
// tiled all-against-all
// ---------------------------------------------------------------------------
constexpr int kRowsPerCta = 64;
constexpr int kContactThreads = 256;
constexpr int kRowsPerWarp = kRowsPerCta / (kContactThreads / 32);
constexpr int kTileJ = 2048;  // 3 * 2048 * 8 B = 48 KB static shared memory

template <bool FILL>
__global__ void __launch_bounds__(kContactThreads)
contacts_tiled_kernel(const double* __restrict__ xyz, int n, double cutoff_sq, PatchView pv,
                      const int64_t* __restrict__ rowptr, int32_t* __restrict__ rowcount,
                      int32_t* __restrict__ col) {
    __shared__ double sx[kTileJ], sy[kTileJ], sz[kTileJ];
    const int s = blockIdx.y;
    const double* X = xyz + (size_t)s * 3 * n;
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int row_base = blockIdx.x * kRowsPerCta + warp * kRowsPerWarp;
    const bool all_pairs = cutoff_sq < 0.0;

    int cnt[kRowsPerWarp];
    double xi[kRowsPerWarp], yi[kRowsPerWarp], zi[kRowsPerWarp];
    bool special[kRowsPerWarp], dead_i[kRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
        const int i = row_base + r;
        cnt[r] = 0;
        special[r] = false;
        dead_i[r] = false;
        xi[r] = yi[r] = zi[r] = 0.0;
        if (i < n) {
            xi[r] = X[i];
            yi[r] = X[n + i];
            zi[r] = X[2 * n + i];
            if (pv.n_off + pv.n_on > 0) special[r] = row_is_patched(pv, i, lane);
            if (pv.dead) dead_i[r] = pv.dead[i] != 0;
        }
    }

    for (int j0 = 0; j0 < n; j0 += kTileJ) {
        const int tn = min(kTileJ, n - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < tn; t += kContactThreads) {
            sx[t] = X[j0 + t];
            sy[t] = X[n + j0 + t];
            sz[t] = X[2 * n + j0 + t];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
            const int i = row_base + r;
            if (i >= n) continue;  // warp-uniform
            int64_t out0 = 0;
            if (FILL) out0 = rowptr[(int64_t)s * n + i];
            for (int jj = 0; jj < tn; jj += 32) {
                const int t = jj + (int)lane;
                const int j = j0 + t;
                bool c = false;
                if (t < tn && j != i) {
                    bool base = all_pairs;
                    if (!all_pairs) {
                        const double sq = sq_dist_rn(sx[t] - xi[r], sy[t] - yi[r], sz[t] - zi[r]);
                        base = sq <= cutoff_sq;
                    }
                    c = apply_patch(pv, base, i, j, dead_i[r], special[r]);
                }
                const unsigned mask = __ballot_sync(0xffffffffu, c);
                if (FILL && c) col[out0 + cnt[r] + __popc(mask & ((1u << lane) - 1u))] = j;
                cnt[r] += __popc(mask);
            }
        }
    }
    if (!FILL && lane == 0) {
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
            const int i = row_base + r;
            if (i < n) rowcount[(int64_t)s * n + i] = cnt[r];
        }
    }
}

// ---------------------------------------------------------------------------
// cell list (B == 1)
// ---------------------------------------------------------------------------
struct CellGrid {
    double x0, y0, z0, inv_edge;
    int nx, ny, nz;
};

__device__ __forceinline__ int cell_coord(double v, double v0, double inv_edge, int nmax) {
    int c = (int)floor((v - v0) * inv_edge);
    return min(max(c, 0), nmax - 1);
}

__global__ void cell_bounds_kernel(const double* __restrict__ xyz, int n, double* __restrict__ bounds) {
    // bounds[0..2] = min, bounds[3..5] = max ; single CTA
    __shared__ double red[6][32];
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        for (int a = 0; a < 3; ++a) {
            double v = xyz[(size_t)a * n + i];
            mn[a] = fmin(mn[a], v);
            mx[a] = fmax(mx[a], v);
        }
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    for (int a = 0; a < 3; ++a) {
        double lo = -warp_max(-mn[a]), hi = warp_max(mx[a]);
        if (lane == 0) { red[a][warp] = lo; red[3 + a][warp] = hi; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double lo = 1e300, hi = -1e300;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            lo = fmin(lo, red[threadIdx.x][w]);
            hi = fmax(hi, red[3 + threadIdx.x][w]);
        }
        bounds[threadIdx.x] = lo;
        bounds[3 + threadIdx.x] = hi;
    }
}

__global__ void cell_assign_kernel(const double* __restrict__ xyz, int n, CellGrid g,
                                   int32_t* __restrict__ cell_of, int32_t* __restrict__ cell_count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx = cell_coord(xyz[i], g.x0, g.inv_edge, g.nx);
    int cy = cell_coord(xyz[(size_t)n + i], g.y0, g.inv_edge, g.ny);
    int cz = cell_coord(xyz[2 * (size_t)n + i], g.z0, g.inv_edge, g.nz);
    int c = (cz * g.ny + cy) * g.nx + cx;
    cell_of[i] = c;
    atomicAdd(&cell_count[c], 1);
}

// members of a cell are stored in ascending atom index (deterministic): each atom
// finds its slot by counting lower-indexed atoms of the same cell among the
// cell's members -- done with a stable "scatter then sort inside cell" pass.
__global__ void cell_scatter_kernel(int n, const int32_t* __restrict__ cell_of,
                                    const int64_t* __restrict__ cell_start,
                                    int32_t* __restrict__ cell_fill, int32_t* __restrict__ members) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cell_of[i];
    int slot = atomicAdd(&cell_fill[c], 1);
    members[cell_start[c] + slot] = i;
}

__global__ void cell_sort_kernel(int ncell, const int64_t* __restrict__ cell_start,
                                 int32_t* __restrict__ members) {
    // insertion sort of each (small) cell by atom index -> deterministic order
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    int64_t b = cell_start[c], e = cell_start[c + 1];
    for (int64_t p = b + 1; p < e; ++p) {
        int v = members[p];
        int64_t q = p - 1;
        while (q >= b && members[q] > v) { members[q + 1] = members[q]; --q; }
        members[q + 1] = v;
    }
}

constexpr int kCellRowCap = 1024;  // max cutoff neighbours per row in the cell-list path

template <bool FILL>
__global__ void __launch_bounds__(256)
contacts_cell_kernel(const double* __restrict__ xyz, int n, double cutoff_sq, PatchView pv, CellGrid g,
                     const int32_t* __restrict__ cell_of, const int64_t* __restrict__ cell_start,
                     const int32_t* __restrict__ members, const int64_t* __restrict__ rowptr,
                     int32_t* __restrict__ rowcount, int32_t* __restrict__ col,
                     int32_t* __restrict__ overflow) {
    __shared__ int32_t list[8][FILL ? kCellRowCap : 1];
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 8 + warp;
    if (i >= n) return;
    const double xi = xyz[i], yi = xyz[(size_t)n + i], zi = xyz[2 * (size_t)n + i];
    const bool special = (pv.n_off + pv.n_on > 0) ? row_is_patched(pv, i, lane) : false;
    const bool dead_i = pv.dead ? pv.dead[i] != 0 : false;
    const int c = cell_of[i];
    const int cx = c % g.nx, cy = (c / g.nx) % g.ny, cz = c / (g.nx * g.ny);
    int cnt = 0;
    for (int dz = -1; dz <= 1; ++dz) {
        const int z = cz + dz;
        if (z < 0 || z >= g.nz) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = cy + dy;
            if (y < 0 || y >= g.ny) continue;
            // the three x-neighbours are contiguous cells -> one member range
            const int xa = max(cx - 1, 0), xb = min(cx + 1, g.nx - 1);
            const int64_t b = cell_start[(z * g.ny + y) * g.nx + xa];
            const int64_t e = cell_start[(z * g.ny + y) * g.nx + xb + 1];
            for (int64_t p0 = b; p0 < e; p0 += 32) {
                const int64_t p = p0 + lane;
                bool hit = false;
                int j = -1;
                if (p < e) {
                    j = members[p];
                    if (j != i) {
                        const double sq = sq_dist_rn(xyz[j] - xi, xyz[(size_t)n + j] - yi,
                                                     xyz[2 * (size_t)n + j] - zi);
                        hit = apply_patch(pv, sq <= cutoff_sq, i, j, dead_i, special);
                    }
                }
                const unsigned mask = __ballot_sync(0xffffffffu, hit);
                if (FILL && hit) {
                    int slot = cnt + __popc(mask & ((1u << lane) - 1u));
                    if (slot < kCellRowCap) list[warp][slot] = j;
                }
                cnt += __popc(mask);
            }
        }
    }
    // switched-on pairs may lie outside the 27-cell stencil
    if (special) {
        for (int q = 0; q < pv.n_on; ++q) {
            int a = pv.on[2 * q], bq = pv.on[2 * q + 1];
            int j = (a == i) ? bq : ((bq == i) ? a : -1);
            if (j < 0 || j == i) continue;
            // already found inside the stencil?  (only if within cutoff => base true;
            // on-pairs are always contacts, so test membership by distance+stencil)
            const int cj = cell_of[j];
            const int jx = cj % g.nx, jy = (cj / g.nx) % g.ny, jz = cj / (g.nx * g.ny);
            const bool in_stencil = abs(jx - cx) <= 1 && abs(jy - cy) <= 1 && abs(jz - cz) <= 1;
            bool dup = false;
            for (int q2 = 0; q2 < q; ++q2) {  // the same pair listed twice
                int a2 = pv.on[2 * q2], b2 = pv.on[2 * q2 + 1];
                if ((a2 == a && b2 == bq) || (a2 == bq && b2 == a)) dup = true;
            }
            if (!in_stencil && !dup) {
                if (FILL && lane == 0 && cnt < kCellRowCap) list[warp][cnt] = j;
                cnt += 1;
            }
        }
    }
    if (!FILL) {
        if (lane == 0) {
            rowcount[i] = cnt;
            if (cnt > kCellRowCap) atomicExch(overflow, 1);
        }
        return;
    }
    __syncwarp();
    cnt = min(cnt, kCellRowCap);
    // rank by counting -> ascending columns
    const int64_t out0 = rowptr[i];
    for (int e = lane; e < cnt; e += 32) {
        const int v = list[warp][e];
        int rank = 0;
        for (int f = 0; f < cnt; ++f) rank += list[warp][f] < v;
        col[out0 + rank] = v;
    }
}

// ---------------------------------------------------------------------------
// exclusive scan  rowcount[int32] -> rowptr[int64]
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 16;
constexpr int kScanChunk = kScanThreads * kScanPerThread;

__global__ void __launch_bounds__(kScanThreads)
scan_chunk_sums_kernel(const int32_t* __restrict__ in, int64_t nrows, int64_t* __restrict__ sums) {
    __shared__ int64_t red[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanChunk;
    int64_t acc = 0;
    for (int k = 0; k < kScanPerThread; ++k) {
        int64_t idx = base + (int64_t)k * kScanThreads + threadIdx.x;
        if (idx < nrows) acc += in[idx];
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane_id() == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += red[w];
        sums[blockIdx.x] = t;
    }
}

__global__ void scan_sums_kernel(int64_t* sums, int64_t nchunks) {
    // single warp, exclusive scan in place
    const unsigned lane = lane_id();
    int64_t carry = 0;
    for (int64_t b = 0; b < nchunks; b += 32) {
        int64_t idx = b + lane;
        int64_t v = idx < nchunks ? sums[idx] : 0;
        int64_t inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)lane >= o) inc += t;
        }
        if (idx < nchunks) sums[idx] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

__global__ void __launch_bounds__(kScanThreads)
scan_final_kernel(const int32_t* __restrict__ in, int64_t nrows, const int64_t* __restrict__ sums,
                  int64_t* __restrict__ out) {
    // thread t owns kScanPerThread CONTIGUOUS elements of the chunk
    __shared__ int64_t warp_tot[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanPerThread;
    int32_t v[kScanPerThread];
    int64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) {
        v[k] = (base + k < nrows) ? in[base + k] : 0;
        tsum += v[k];
    }
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    int64_t inc = tsum;
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int64_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_tot[w];
    int64_t run = sums[blockIdx.x] + woff + inc - tsum;
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) {
        if (base + k < nrows) out[base + k] = run;
        run += v[k];
        if (base + k == nrows - 1) out[nrows] = run;
    }
}

__global__ void pairs_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                             int64_t nrows, int n, int64_t* __restrict__ pairs) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int64_t b = rowptr[row], e = rowptr[row + 1];
    const int64_t i = row % n;
    for (int64_t p = b + lane_id(); p < e; p += 32) {
        pairs[2 * p] = i;
        pairs[2 * p + 1] = col[p];
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct CellScratch {
    // one allocation, freed by the stream-ordered allocator
    double* bounds;
    int32_t* cell_of;
    int32_t* cell_count;
    int32_t* cell_fill;
    int64_t* cell_start;
    int32_t* members;
    int64_t* scan_tmp;
    int32_t* overflow;
    void* block;
};

static int build_cells(const double* xyz, int n, double cutoff_sq, cudaStream_t st, CellGrid* grid,
                       CellScratch* cs) {
    // bounding box (one tiny D2H sync: the grid dimensions size the allocations)
    double* bounds_d = nullptr;
    SCB_CUDA(cudaMallocAsync(&bounds_d, 6 * sizeof(double), st));
    cell_bounds_kernel<<<1, 1024, 0, st>>>(xyz, n, bounds_d);
    SCB_LAUNCH_CHECK();
    double hb[6];
    SCB_CUDA(cudaMemcpyAsync(hb, bounds_d, sizeof(hb), cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaStreamSynchronize(st));
    SCB_CUDA(cudaFreeAsync(bounds_d, st));
    const double cutoff = sqrt(cutoff_sq);
    // edge slightly above the cutoff so that rounding in the cell assignment can
    // never push a pair within the cutoff outside the 27-cell stencil
    double edge = cutoff * (1.0 + 1e-9) + 1e-9;
    CellGrid g;
    g.x0 = hb[0]; g.y0 = hb[1]; g.z0 = hb[2];
    auto dim = [&](double lo, double hi) {
        double ext = hi - lo;
        int64_t c = (int64_t)floor(ext / edge) + 1;
        return (int)(c < 1 ? 1 : c);
    };
    g.nx = dim(hb[0], hb[3]); g.ny = dim(hb[1], hb[4]); g.nz = dim(hb[2], hb[5]);
    // bound the number of cells (sparse, elongated inputs): coarsen the grid
    while ((int64_t)g.nx * g.ny * g.nz > (int64_t)8 * n + 4096) {
        edge *= 1.26;
        g.nx = dim(hb[0], hb[3]); g.ny = dim(hb[1], hb[4]); g.nz = dim(hb[2], hb[5]);
    }
    g.inv_edge = 1.0 / edge;
    *grid = g;
    const int64_t ncell = (int64_t)g.nx * g.ny * g.nz;
    const size_t scan_bytes = scb_scan_scratch_bytes(ncell);
    size_t bytes = 0;
    auto carve = [&](size_t b) { size_t o = bytes; bytes += (b + 255) & ~size_t(255); return o; };
    size_t o_cell_of = carve(sizeof(int32_t) * n);
    size_t o_count = carve(sizeof(int32_t) * ncell);
    size_t o_fill = carve(sizeof(int32_t) * ncell);
    size_t o_start = carve(sizeof(int64_t) * (ncell + 1));
    size_t o_members = carve(sizeof(int32_t) * n);
    size_t o_scan = carve(scan_bytes);
    size_t o_ovf = carve(sizeof(int32_t));
    char* blk = nullptr;
    SCB_CUDA(cudaMallocAsync(&blk, bytes, st));
    cs->block = blk;
    cs->cell_of = (int32_t*)(blk + o_cell_of);
    cs->cell_count = (int32_t*)(blk + o_count);
    cs->cell_fill = (int32_t*)(blk + o_fill);
    cs->cell_start = (int64_t*)(blk + o_start);
    cs->members = (int32_t*)(blk + o_members);
    cs->scan_tmp = (int64_t*)(blk + o_scan);
    cs->overflow = (int32_t*)(blk + o_ovf);
    SCB_CUDA(cudaMemsetAsync(cs->cell_count, 0, (o_start - o_count), st));  // count + fill
    SCB_CUDA(cudaMemsetAsync(cs->overflow, 0, sizeof(int32_t), st));
    cell_assign_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(xyz, n, g, cs->cell_of, cs->cell_count);
    SCB_LAUNCH_CHECK();
    SCB_TRY(scb_contacts_scan(cs->cell_count, ncell, cs->cell_start, cs->scan_tmp, st));
    cell_scatter_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(n, cs->cell_of, cs->cell_start,
                                                                   cs->cell_fill, cs->members);
    SCB_LAUNCH_CHECK();
    cell_sort_kernel<<<(unsigned)ceil_div(ncell, 128), 128, 0, st>>>((int)ncell, cs->cell_start, cs->members);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

template <bool FILL>
static int run_contacts(const double* xyz, int B, int n, double cutoff_sq, const scb_patch* patch,
                        int use_cell_list, const int64_t* rowptr, int32_t* rowcount, int32_t* col,
                        void* stream) {
    if (!xyz || B < 1 || n < 1) return SCB_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    PatchView pv = make_patch_view(patch);
    if (use_cell_list && cutoff_sq >= 0.0 && B == 1) {
        CellGrid g;
        CellScratch cs{};
        SCB_TRY(build_cells(xyz, n, cutoff_sq, st, &g, &cs));
        contacts_cell_kernel<FILL><<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(
            xyz, n, cutoff_sq, pv, g, cs.cell_of, cs.cell_start, cs.members, rowptr, rowcount, col, cs.overflow);
        SCB_LAUNCH_CHECK();
        int status = SCB_OK;
        if (!FILL) {
            int32_t ovf = 0;
            SCB_CUDA(cudaMemcpyAsync(&ovf, cs.overflow, sizeof(ovf), cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaStreamSynchronize(st));
            if (ovf) status = SCB_ERR_UNSUPPORTED;  // > kCellRowCap neighbours: use the tiled path
        }
        SCB_CUDA(cudaFreeAsync(cs.block, st));
        return status;
    }
    dim3 grid((unsigned)ceil_div(n, kRowsPerCta), (unsigned)B);
    contacts_tiled_kernel<FILL><<<grid, kContactThreads, 0, st>>>(xyz, n, cutoff_sq, pv, rowptr, rowcount, col);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

}  // namespace scb

using namespace scb;

extern "C" int scb_contacts_count(const double* xyz, int B, int n, double cutoff_sq, const scb_patch* patch,
                                  int use_cell_list, int32_t* rowcount, void* stream) {
    if (!rowcount) return SCB_ERR_INVALID;
    int s = run_contacts<false>(xyz, B, n, cutoff_sq, patch, use_cell_list, nullptr, rowcount, nullptr, stream);
    if (s == SCB_ERR_UNSUPPORTED)  // dense neighbourhoods: fall back to the tiled kernel (same result)
        s = run_contacts<false>(xyz, B, n, cutoff_sq, patch, 0, nullptr, rowcount, nullptr, stream);
    return s;
}

extern "C" int scb_contacts_fill(const double* xyz, int B, int n, double cutoff_sq, const scb_patch* patch,
                                 int use_cell_list, const int64_t* rowptr, int32_t* col, void* stream) {
    if (!rowptr || !col) return SCB_ERR_INVALID;
    return run_contacts<true>(xyz, B, n, cutoff_sq, patch, use_cell_list, rowptr, nullptr, col, stream);
}

extern "C" size_t scb_scan_scratch_bytes(int64_t nrows) {
    return (size_t)(ceil_div(nrows, kScanChunk) + 1) * sizeof(int64_t);
}

extern "C" int scb_contacts_scan(const int32_t* rowcount, int64_t nrows, int64_t* rowptr, void* scratch,
                                 void* stream) {
    if (!rowcount || !rowptr || !scratch || nrows < 1) return SCB_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    const int64_t nchunks = ceil_div(nrows, kScanChunk);
    int64_t* sums = static_cast<int64_t*>(scratch);
    scan_chunk_sums_kernel<<<(unsigned)nchunks, kScanThreads, 0, st>>>(rowcount, nrows, sums);
    SCB_LAUNCH_CHECK();
    scan_sums_kernel<<<1, 32, 0, st>>>(sums, nchunks);
    SCB_LAUNCH_CHECK();
    scan_final_kernel<<<(unsigned)nchunks, kScanThreads, 0, st>>>(rowcount, nrows, sums, rowptr);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}

extern "C" int scb_pairs_materialize(const int64_t* rowptr, const int32_t* col, int B, int n, int64_t* pairs,
                                     void* stream) {
    if (!rowptr || !col || !pairs) return SCB_ERR_INVALID;
    const int64_t nrows = (int64_t)B * n;
    pairs_kernel<<<(unsigned)ceil_div(nrows, 8), 256, 0, as_stream(stream)>>>(rowptr, col, nrows, n, pairs);
    SCB_LAUNCH_CHECK();
    return SCB_OK;
}
