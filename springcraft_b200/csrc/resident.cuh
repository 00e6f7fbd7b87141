// Structure-resident lowest-k path (ANM, D = 3) for ensembles of small structures.
//
// Every off-diagonal ANM block is rank one, H_ij = t d d^T with t = -fc/sq <= 0 (interaction.py:96-101), so a
// contact is stored as v = sqrt(-t) d (H_ij = -v v^T): 3 floats instead of 9 and 6 FMA per column instead of 9.
// Two consecutive block rows (residues 2p, 2p+1) share one record {col, v_top, v_bottom} (32 bytes).  The row
// pairs of a structure are sorted by record count and cut into groups of kResRowPairs(COLS); a group is the unit of
// work of one warp and its records are interleaved by iteration so that one warp iteration reads one contiguous
// block of records (group lists are padded with zero records to the longest list of the group, an even count).
//
// The Chebyshev filter runs in RESIDUAL FORM: for a Ritz pair (theta, x) with r = Hx - theta x,
//     p(H) x = p(theta) x + q(H) r,        q(l) = (p(l) - p(theta)) / (l - theta),
// so the filtered vector is x + z with z = q(H) r / p(theta) obeying the Chebyshev recurrence with a source term.
// z is proportional to the residual: evaluating it in FP32 perturbs the new basis by ~1e-6 |r| instead of
// ~1e-6 |x|, so every filter application can run in single precision down to the FP64 tolerance, with the block of
// one structure held in shared memory for the whole multi-degree filter (one launch per outer iteration).
#pragma once
#include "subspace.cuh"

namespace scb {

struct ResRec {            // one merged contact of a row pair
    int32_t col;           // node index of the contact inside the structure
    float v0[3];           // sqrt(-t) d of row 2p   (0 when that row has no contact with `col`)
    float v1[3];           // same for row 2p+1
    int32_t pad;
};
static_assert(sizeof(ResRec) == 32, "record layout");

constexpr int kResMaxWarps = 24;     // warps per CTA (one group each)
constexpr int kResDegreeCap = 64;

struct ResLayout {         // device buffers of the resident operator (carved from the solver workspace)
    ResRec* rec;           // [capacity] interleaved records; structure s starts at rowptr[s*n] + rec_pad*s
    int32_t* gstart;       // [B][G+1] prefix of group iteration counts
    uint16_t* order;       // [B][G*rpw] row pair of (group, slot); 0xFFFF = empty slot; bit 15 = half of a split pair
    float* diag32;         // [B][n][12] diagonal blocks (9 used, row-major)
    int32_t* flag;         // [1] != 0: some block is not of the form -v v^T (positive t) -> path unusable
    double* est;           // [B] Lanczos estimate of the largest eigenvalue
    unsigned long long* app_counter;   // [1] structure x operator applications of the filter launches (profiling)
    int G;                 // groups (= warps) per structure
    int nsplit;            // 16-column layout: the nsplit longest row pairs are split over two slots
    int rec_mul;           // structure s starts at record rec_mul * rowptr[s*n] + rec_pad * s (rounded up to 8)
    int rpw;               // row pairs per group
    int cols;              // block columns per CTA (16, 8 or 4)
    int64_t rec_pad;       // extra record capacity per structure (padding of the groups)
};

// columns per CTA for a structure of n nodes (0: does not fit -> streaming kernels)
int resident_cols(int n, int b);
void resident_shape(int n, int cols, int* G, int* nsplit, int* rec_mul, int64_t* rec_pad);
size_t resident_capacity(int B, int n, int64_t P, int cols);
int resident_build(int B, int n, const int64_t* rowptr, const int32_t* col, const double* offdiag,
                   const double* diag, const ResLayout& L, cudaStream_t st);
// X <- deflate(X + |r| z) for every active structure: one CTA per (structure, column group)
int resident_filter(int B, int n, int b, const int64_t* rowptr, const ResLayout& L, const double* X,
                    const double* HX, const double* theta, const double* rn2, const EigState* state,
                    const int32_t* done, const double* Z, int nz, double* Xout, int kwant, double tol,
                    cudaStream_t st);
// upper spectrum bound from `steps` steps of column-wise Lanczos (FP32, structure-resident): state[s].ub
int resident_lanczos(int B, int n, int b, const int64_t* rowptr, const ResLayout& L, int steps, uint64_t seed,
                     double ub_factor, EigState* state, cudaStream_t st);

// filter-kernel profile of this process (scb_profile): enabled flag and accumulators
bool profile_enabled();
void profile_add(double filter_ms, long long launches, unsigned long long applications);

}  // namespace scb
