// gw.cuh — Golub-Welsch for the TINY rules only (Hermite n <= 20, Laguerre n < 15: the reference's
// hermite_gw, src/gausshermite.jl:325-341, and gausslaguerre_GW, src/gausslaguerre.jl:535-542).
// The reference calls LAPACK's symmetric-tridiagonal eigen-solver; here one thread runs the implicit
// QL iteration with Wilkinson shifts on the n x n Jacobi matrix and carries only the FIRST ROW of the
// eigenvector matrix (all the weights need).  O(n^2) work for n <= 20: nothing to parallelise.
#ifndef FGQ_GW_CUH_
#define FGQ_GW_CUH_

namespace fgq {

constexpr int GW_MAX = 32;

// d[0..n): diagonal, e[0..n-1): sub-diagonal (e[n-1] unused).  On return d holds the eigenvalues in
// ascending order and z[i] the first component of the i-th normalised eigenvector.  Returns false if
// an eigenvalue needs more than 60 iterations.
__device__ inline bool tridiag_ql_first_row(int n, double* d, double* e, double* z) {
    for (int i = 0; i < n; ++i) z[i] = (i == 0) ? 1.0 : 0.0;
    e[n - 1] = 0.0;
    for (int l = 0; l < n; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) {
                const double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 60) return false;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + copysign(r, g));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    r = hypot(f, g);
                    e[i + 1] = r;
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    f = z[i + 1];
                    z[i + 1] = s * z[i] + c * f;
                    z[i] = c * z[i] - s * f;
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    // ascending order (insertion sort; n <= 32)
    for (int i = 1; i < n; ++i) {
        const double di = d[i], zi = z[i];
        int j = i - 1;
        while (j >= 0 && d[j] > di) {
            d[j + 1] = d[j];
            z[j + 1] = z[j];
            --j;
        }
        d[j + 1] = di;
        z[j + 1] = zi;
    }
    return true;
}

}  // namespace fgq
#endif
