// Symmetric eigen-decomposition by cyclic Jacobi rotations (host, double): A = V diag(w) V^T.
// Used once per dense Bingham target (d <= 64) so that x^T A x = sum_j w_j (v_j . x)^2 can be evaluated from
// the rows of X V alone (kernels_logp_tc.cuh) -- the decomposition the reference's Bingham object also holds
// (numpy eigh, geosss/distributions.py:67-76).  Plain C++: also compiled into the test-only host build.
#pragma once
#include <cmath>
#include <vector>

namespace geosss {

// a: (n, n) row-major symmetric (copied); w: n eigenvalues; V: (n, n) row-major, COLUMN j is the eigenvector of w[j].
inline void eigh_jacobi(int n, std::vector<double> a, std::vector<double>& w, std::vector<double>& V) {
    V.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i) {
            diag += a[(size_t)i * n + i] * a[(size_t)i * n + i];
            for (int j = i + 1; j < n; ++j) off += a[(size_t)i * n + j] * a[(size_t)i * n + j];
        }
        if (off <= 1e-32 * (diag + off) || off == 0.0) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[(size_t)p * n + q];
                if (apq == 0.0) continue;
                const double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {  // columns p, q
                    const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
                    a[(size_t)k * n + p] = c * akp - s * akq;
                    a[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {  // rows p, q
                    const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
                    a[(size_t)p * n + k] = c * apk - s * aqk;
                    a[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
                    V[(size_t)k * n + p] = c * vkp - s * vkq;
                    V[(size_t)k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    w.resize(n);
    for (int i = 0; i < n; ++i) w[i] = a[(size_t)i * n + i];
}

}  // namespace geosss
