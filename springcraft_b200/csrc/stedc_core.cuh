// Scalar core of the divide-and-conquer eigensolver for symmetric tridiagonal
// matrices (eig_full_tridiag.cu): the deflation scan and the secular-equation
// root finder of one rank-one merge  D + rho z z^T.  Together with the
// Householder reduction this replaces LAPACK dsyevd behind np.linalg.eigh
// (reference nma.py:61, anm.py:135).
//
// Everything here is __host__ __device__ and free of CUDA intrinsics, so that
// tests/native/stedc_host.cpp can run the SAME code on the CPU (g++) against
// LAPACK; the kernels call it with a warp as the lane context.
//
// Method (Cuppen; Gu & Eisenstat for the stable eigenvectors): with the
// sub-problems T1 = Q1 D1 Q1^T, T2 = Q2 D2 Q2^T and the coupling element beta,
//   T = diag(Q1,Q2) (D + rho z z^T) diag(Q1,Q2)^T,  z = (last row of Q1, first row of Q2)/sqrt(2), rho = 2 beta.
// Entries with negligible z or (after a Givens rotation) coinciding d deflate;
// the others give the roots of  1/rho + sum_i z_i^2 / (d_i - lambda) = 0, found
// per root with the two-pole rational ("middle way") iteration inside a
// bracket, in coordinates shifted to the nearer pole so that all differences
// d_i - lambda_j are accurate.  z is then recomputed from those differences,
// which makes the eigenvector matrix numerically orthogonal.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SCB_HD __host__ __device__ inline
#else
#define SCB_HD inline
#endif

namespace scb {
namespace stedc {

constexpr double kEps = 1.1102230246251565e-16;  // 2^-53, LAPACK's dlamch('E')

// lane context of a single thread (host code, or one thread per root)
struct OneLane {
    SCB_HD int lane() const { return 0; }
    SCB_HD int lanes() const { return 1; }
    SCB_HD double sum(double v) const { return v; }
};

struct Rotation {
    int p, q;      // rows (eigenvectors) p and q of the sub-problem, local indices
    double c, s;   // row_p' = c row_p + s row_q ; row_q' = c row_q - s row_p
};

// Deflation scan of one merge (sequential; follows the order of LAPACK dlaed2).
//   n      order of the merged problem
//   order  order[t] = local index of the t-th smallest d (ties by index)
//   d, z   eigenvalues of the two sub-problems and the normalised coupling vector, local indexing; both are
//          modified (rotations move weight between close eigenvalues)
//   rho    2 |beta| > 0
// Output: k survivors  nd[0..k) (ascending d), deflated entries dfl[0..n-k), rotations rot[0..*nrot).
//   n1, mixed  optional: the first n1 entries belong to the first sub-problem; mixed[l] (zeroed by the caller) is
//          set for survivors that absorbed weight from the OTHER sub-problem through a rotation -- their
//          eigenvector rows are no longer confined to one half of the merged range
SCB_HD int deflation_scan(int n, const int* order, double* d, double* z, double rho, int* nd, int* dfl,
                          Rotation* rot, int* nrot, int n1 = 0, unsigned char* mixed = nullptr) {
    double dmax = 0.0, zmax = 0.0;
    for (int i = 0; i < n; ++i) {
        dmax = fmax(dmax, fabs(d[i]));
        zmax = fmax(zmax, fabs(z[i]));
    }
    const double tol = 8.0 * kEps * fmax(dmax, zmax);
    int k = 0, nf = 0, nr = 0;
    if (rho * zmax <= tol) {
        for (int t = 0; t < n; ++t) dfl[nf++] = order[t];
        *nrot = 0;
        return 0;
    }
    int pj = -1;
    for (int t = 0; t < n; ++t) {
        const int nj = order[t];
        if (rho * fabs(z[nj]) <= tol) {
            dfl[nf++] = nj;
            continue;
        }
        if (pj < 0) {
            pj = nj;
            continue;
        }
        // |gap * c * s| <= tol with c = z_nj / r, s = -z_pj / r, r^2 = z_pj^2 + z_nj^2, tested without the root
        const double zp = z[pj], zn = z[nj];
        const double r2 = zp * zp + zn * zn;
        const double gap = d[nj] - d[pj];
        if (fabs(gap * zp * zn) <= tol * r2) {
            const double r = sqrt(r2);
            const double c = zn / r, s = -zp / r;
            z[nj] = r;
            z[pj] = 0.0;
            rot[nr].p = pj; rot[nr].q = nj; rot[nr].c = c; rot[nr].s = s;
            ++nr;
            if (mixed && (((pj < n1) != (nj < n1)) || mixed[pj])) mixed[nj] = 1;
            const double dp = d[pj] * c * c + d[nj] * s * s;
            d[nj] = d[pj] * s * s + d[nj] * c * c;
            d[pj] = dp;
            dfl[nf++] = pj;
        } else {
            nd[k++] = pj;
        }
        pj = nj;
    }
    if (pj >= 0) nd[k++] = pj;
    *nrot = nr;
    return k;
}

// One root of the secular equation.  dl[0..k) ascending and distinct, w[0..k) non-zero, rho > 0.
// Returns lambda_j and writes delta[i] = dl[i] - lambda_j (accurate) for the lanes' share i = lane, lane+lanes, ...
// Every lane of the context calls it with the same arguments and gets the same result.
// pos (optional): delta[pos[i]] is written instead of delta[i] (columns of the eigenvector matrix grouped by child).
template <class Ctx>
SCB_HD double secular_root(const Ctx& cx, int k, int j, const double* dl, const double* w, double rho,
                           double* delta, const int* pos = nullptr) {
    const int l0 = cx.lane(), nl = cx.lanes();
    if (k == 1) {
        const double t = rho * w[0] * w[0];
        if (l0 == 0) delta[pos ? pos[0] : 0] = -t;
        return dl[0] + t;
    }
    const double rhoinv = 1.0 / rho;
    const bool last = (j == k - 1);
    const int js = last ? k - 2 : j;          // psi = terms 0..js, phi = terms js+1..k-1
    int org;                                  // index of the pole the coordinates are shifted to
    double p1, p2, lo, hi, tau;
    if (!last) {
        const double gapw = dl[j + 1] - dl[j];
        const double half = 0.5 * gapw;
        double rest = 0.0;
        for (int i = l0; i < k; i += nl) {
            if (i == j || i == j + 1) continue;
            rest += w[i] * w[i] / ((dl[i] - dl[j]) - half);
        }
        const double c0 = rhoinv + cx.sum(rest);
        const double A = w[j] * w[j], Bw = w[j + 1] * w[j + 1];
        const double gmid = c0 - A / half + Bw / half;
        if (gmid >= 0.0) {   // root in the left half: shift to dl[j]
            org = j; p1 = 0.0; p2 = gapw; lo = 0.0; hi = half;
            const double a = c0 * gapw + A + Bw, b = A * gapw;
            const double disc = sqrt(fabs(a * a - 4.0 * b * c0));
            tau = (a > 0.0) ? 2.0 * b / (a + disc) : (a - disc) / (2.0 * c0);
        } else {             // right half: shift to dl[j+1]
            org = j + 1; p1 = -gapw; p2 = 0.0; lo = -half; hi = 0.0;
            const double a = c0 * gapw - A - Bw, b = Bw * gapw;
            const double disc = sqrt(fabs(a * a + 4.0 * b * c0));
            tau = (a < 0.0) ? 2.0 * b / (a - disc) : -(a + disc) / (2.0 * c0);
        }
    } else {
        org = k - 1; p1 = dl[k - 2] - dl[k - 1]; p2 = 0.0;
        double nrm = 0.0;
        for (int i = l0; i < k; i += nl) nrm += w[i] * w[i];
        nrm = cx.sum(nrm);
        lo = 0.0; hi = rho * nrm;
        const double mid = 0.5 * hi;
        double rest = 0.0;
        for (int i = l0; i < k - 2; i += nl) rest += w[i] * w[i] / ((dl[i] - dl[k - 1]) - mid);
        const double c0 = rhoinv + cx.sum(rest);
        const double A = w[k - 2] * w[k - 2], Bw = w[k - 1] * w[k - 1];
        if (c0 > 0.0) {
            const double bb = c0 * p1 + A + Bw;
            const double disc = sqrt(fabs(bb * bb - 4.0 * c0 * Bw * p1));
            tau = (bb >= 0.0) ? (bb + disc) / (2.0 * c0) : 2.0 * Bw * p1 / (bb - disc);
        } else {
            tau = mid;
        }
    }
    if (!(tau > lo && tau < hi)) tau = 0.5 * (lo + hi);
    if (last && !(tau > 0.0)) tau = hi;

    const double dorg = dl[org];
    for (int it = 0; it < 120; ++it) {
        double psi = 0.0, dpsi = 0.0, phi = 0.0, dphi = 0.0;
        for (int i = l0; i < k; i += nl) {
            const double inv = 1.0 / ((dl[i] - dorg) - tau);
            const double t = w[i] * w[i] * inv;
            if (i <= js) { psi += t; dpsi += t * inv; }
            else { phi += t; dphi += t * inv; }
        }
        psi = cx.sum(psi); dpsi = cx.sum(dpsi); phi = cx.sum(phi); dphi = cx.sum(dphi);
        const double g = rhoinv + psi + phi;
        const double errb = 8.0 * (fabs(phi) + fabs(psi)) + rhoinv + fabs(tau) * (dpsi + dphi);
        if (fabs(g) <= kEps * errb) break;
        if (g < 0.0) lo = tau; else hi = tau;
        if (hi - lo <= 2.0 * kEps * fmax(fabs(lo), fabs(hi))) break;
        const double D1 = p1 - tau, D2 = p2 - tau;
        double eta;
        if (!last) {
            const double c = g - D1 * dpsi - D2 * dphi;
            const double a = (D1 + D2) * g - D1 * D2 * (dpsi + dphi);
            const double b = D1 * D2 * g;
            const double disc = sqrt(fabs(a * a - 4.0 * b * c));
            if (c == 0.0) eta = b / a;
            else if (a <= 0.0) eta = (a - disc) / (2.0 * c);
            else eta = 2.0 * b / (a + disc);
        } else {
            double c = g - D1 * dpsi - D2 * dphi;
            const double a = (D1 + D2) * g - D1 * D2 * (dpsi + dphi);
            const double b = D1 * D2 * g;
            if (c < 0.0) c = fabs(c);
            const double disc = sqrt(fabs(a * a - 4.0 * b * c));
            if (c == 0.0) eta = hi - tau;
            else if (a >= 0.0) eta = (a + disc) / (2.0 * c);
            else eta = 2.0 * b / (a - disc);
        }
        // g increases with tau: a step in the wrong direction is replaced by a Newton step
        if (g * eta >= 0.0) eta = -g / (dpsi + dphi);
        double next = tau + eta;
        if (!(next > lo && next < hi)) next = 0.5 * (lo + hi);   // also catches NaN
        if (next == tau) break;
        tau = next;
    }
    for (int i = l0; i < k; i += nl) delta[pos ? pos[i] : i] = (dl[i] - dorg) - tau;
    return dorg + tau;
}

}  // namespace stedc
}  // namespace scb
