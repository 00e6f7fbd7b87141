// Test-only CPU driver for springcraft_b200/csrc/stedc_core.cuh: runs the divide-and-conquer merge tree with the
// SAME deflation scan and secular root finder the CUDA kernels call, single threaded, so that the numerical core of
// the full-spectrum solver can be checked against LAPACK without a GPU (tests/test_host_logic.py).  Not shipped.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../springcraft_b200/csrc/stedc_core.cuh"

using namespace scb::stedc;

namespace {

void node_range(int N, int level, int t, int* lo, int* hi) {
    int a = 0, b = N;
    for (int bit = level - 1; bit >= 0; --bit) {
        const int mid = a + (b - a) / 2;
        if ((t >> bit) & 1) a = mid; else b = mid;
    }
    *lo = a; *hi = b;
}

// cyclic Jacobi on a small dense symmetric matrix (leaf solver)
void leaf_jacobi(int n, std::vector<double>& A, std::vector<double>& V) {
    V.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) (i == j ? dg : off) += A[(size_t)i * n + j] * A[(size_t)i * n + j];
        if (off <= 1e-32 * dg || off == 0.0) break;
        for (int p = 0; p < n; ++p) for (int q = p + 1; q < n; ++q) {
            const double apq = A[(size_t)p * n + q];
            if (apq == 0.0) continue;
            const double dd = A[(size_t)q * n + q] - A[(size_t)p * n + p], h = 2.0 * apq;
            const double r = std::sqrt(dd * dd + h * h);
            double t = std::fabs(h) / (std::fabs(dd) + r);
            if ((dd < 0.0) != (h < 0.0)) t = -t;
            const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
            for (int i = 0; i < n; ++i) {
                const double x = A[(size_t)i * n + p], y = A[(size_t)i * n + q];
                A[(size_t)i * n + p] = c * x - s * y; A[(size_t)i * n + q] = s * x + c * y;
            }
            for (int i = 0; i < n; ++i) {
                const double x = A[(size_t)p * n + i], y = A[(size_t)q * n + i];
                A[(size_t)p * n + i] = c * x - s * y; A[(size_t)q * n + i] = s * x + c * y;
            }
            for (int i = 0; i < n; ++i) {
                const double x = V[(size_t)i * n + p], y = V[(size_t)i * n + q];
                V[(size_t)i * n + p] = c * x - s * y; V[(size_t)i * n + q] = s * x + c * y;
            }
        }
    }
}

}  // namespace

// d[N], e[N-1] -> lam[N] ascending, Zt[N][N] (row m = eigenvector m).  leaf = largest leaf order.  Returns the
// number of merge levels.  stats[0] = total deflated entries, stats[1] = rotations, stats[2] = total merged order.
extern "C" int stedc_host(int N, const double* d_in, const double* e_in, double* lam, double* Zt, int leaf,
                          long long* stats) {
    std::vector<double> d(d_in, d_in + N), e(N > 1 ? N - 1 : 0), sgn(N, 1.0);
    for (int i = 0; i + 1 < N; ++i) {
        e[i] = std::fabs(e_in[i]);
        sgn[i + 1] = sgn[i] * (e_in[i] < 0.0 ? -1.0 : 1.0);
    }
    int L = 0;
    while (((N + (1 << L) - 1) >> L) > leaf) ++L;
    // tear: every boundary of every level
    for (int level = 1; level <= L; ++level)
        for (int t = 1; t < (1 << level); t += 2) {
            int lo, hi;
            node_range(N, level, t, &lo, &hi);
            d[lo - 1] -= e[lo - 1];
            d[lo] -= e[lo - 1];
        }
    std::vector<double> Za((size_t)N * N, 0.0), Zb((size_t)N * N, 0.0), D(N), Dn(N), U((size_t)N * N);
    for (int t = 0; t < (1 << L); ++t) {
        int lo, hi;
        node_range(N, L, t, &lo, &hi);
        const int n = hi - lo;
        std::vector<double> A((size_t)n * n, 0.0), V;
        for (int i = 0; i < n; ++i) {
            A[(size_t)i * n + i] = d[lo + i];
            if (i + 1 < n) A[(size_t)i * n + i + 1] = A[(size_t)(i + 1) * n + i] = e[lo + i];
        }
        leaf_jacobi(n, A, V);
        for (int m = 0; m < n; ++m) {
            D[lo + m] = A[(size_t)m * n + m];
            for (int i = 0; i < n; ++i) Za[(size_t)(lo + m) * N + lo + i] = V[(size_t)i * n + m];
        }
    }
    stats[0] = stats[1] = stats[2] = 0;
    std::vector<int> order(N), nd(N), dfl(N);
    std::vector<Rotation> rot(N);
    std::vector<double> dd(N), z(N), dl(N), w(N), zh(N);
    for (int level = L - 1; level >= 0; --level) {
        for (int t = 0; t < (1 << level); ++t) {
            int lo, hi, l1, h1;
            node_range(N, level, t, &lo, &hi);
            node_range(N, level + 1, 2 * t, &l1, &h1);
            const int mid = h1, n = hi - lo;
            const double rho = 2.0 * e[mid - 1];
            for (int l = 0; l < n; ++l) {
                dd[l] = D[lo + l];
                z[l] = Za[(size_t)(lo + l) * N + (lo + l < mid ? mid - 1 : mid)] * M_SQRT1_2;
            }
            for (int l = 0; l < n; ++l) {   // rank by counting (what the kernel does)
                int r = 0;
                for (int q = 0; q < n; ++q) r += (dd[q] < dd[l]) || (dd[q] == dd[l] && q < l);
                order[r] = l;
            }
            int nrot = 0;
            const int n1 = mid - lo;
            std::vector<unsigned char> mixed(n, 0);
            const int k = deflation_scan(n, order.data(), dd.data(), z.data(), rho, nd.data(), dfl.data(), rot.data(), &nrot,
                                         n1, mixed.data());
            stats[0] += n - k; stats[1] += nrot; stats[2] += n;
            for (int r = 0; r < nrot; ++r) {
                double* x = &Za[(size_t)(lo + rot[r].p) * N + lo];
                double* y = &Za[(size_t)(lo + rot[r].q) * N + lo];
                for (int i = 0; i < n; ++i) {
                    const double a = x[i], b = y[i];
                    x[i] = rot[r].c * a + rot[r].s * b;
                    y[i] = rot[r].c * b - rot[r].s * a;
                }
            }
            // columns of U grouped [first child | mixed by a rotation across the children | second child], as on the GPU
            std::vector<int> pos(k), ndg(k);
            int k1 = 0, km = 0;
            for (int q = 0; q < k; ++q) { if (mixed[nd[q]]) ++km; else if (nd[q] < n1) ++k1; }
            {
                int c1 = 0, cm = k1, c2 = k1 + km;
                for (int q = 0; q < k; ++q) {
                    const int l = nd[q];
                    const int pq = mixed[l] ? cm++ : (l < n1 ? c1++ : c2++);
                    pos[q] = pq; ndg[pq] = l;
                }
            }
            for (int l = 0; l < k; ++l) { dl[l] = dd[nd[l]]; w[l] = z[nd[l]]; }
            OneLane cx;
            for (int j = 0; j < k; ++j)
                Dn[lo + j] = secular_root(cx, k, j, dl.data(), w.data(), rho, &U[(size_t)j * N], pos.data());
            for (int i = 0; i < k; ++i) {
                double p = U[(size_t)i * N + pos[i]];
                for (int j = 0; j < k; ++j) if (j != i) p *= U[(size_t)j * N + pos[i]] / (dl[i] - dl[j]);
                zh[i] = std::copysign(std::sqrt(std::fabs(p)), w[i]);
            }
            for (int j = 0; j < k; ++j) {
                double s = 0.0;
                for (int i = 0; i < k; ++i) { const double u = zh[i] / U[(size_t)j * N + pos[i]]; U[(size_t)j * N + pos[i]] = u; s += u * u; }
                s = 1.0 / std::sqrt(s);
                for (int i = 0; i < k; ++i) U[(size_t)j * N + i] *= s;
            }
            // two half products: columns [0, n1) from the groups [first | mixed], columns [n1, n) from [mixed | second]
            for (int j = 0; j < k; ++j) {
                double* out = &Zb[(size_t)(lo + j) * N + lo];
                std::fill(out, out + n, 0.0);
                for (int g = 0; g < k1 + km; ++g) {
                    const double u = U[(size_t)j * N + g];
                    const double* src = &Za[(size_t)(lo + ndg[g]) * N + lo];
                    for (int i = 0; i < n1; ++i) out[i] += u * src[i];
                }
                for (int g = k1; g < k; ++g) {
                    const double u = U[(size_t)j * N + g];
                    const double* src = &Za[(size_t)(lo + ndg[g]) * N + lo];
                    for (int i = n1; i < n; ++i) out[i] += u * src[i];
                }
            }
            for (int q = 0; q < n - k; ++q) {
                Dn[lo + k + q] = dd[dfl[q]];
                std::memcpy(&Zb[(size_t)(lo + k + q) * N + lo], &Za[(size_t)(lo + dfl[q]) * N + lo], sizeof(double) * n);
            }
        }
        Za.swap(Zb);
        D.swap(Dn);
    }
    for (int m = 0; m < N; ++m) {
        int r = 0;
        for (int q = 0; q < N; ++q) r += (D[q] < D[m]) || (D[q] == D[m] && q < m);
        lam[r] = D[m];
        for (int i = 0; i < N; ++i) Zt[(size_t)r * N + i] = sgn[i] * Za[(size_t)m * N + i];
    }
    return L;
}
