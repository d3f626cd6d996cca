// Target log-densities, "literal" form: same operations and operation order as the reference
// (float64 gives <= 1e-12 relative agreement, the float32 instantiation is the plain fp32
// version of the same code).  One chain's point is read through a strided view (SVec) so the
// same functions serve the batched log_prob kernels and the lane-per-chain samplers.
//
//   VmfMixP   geosss/distributions.py:156-157 (component), :218-221 (mixture + scipy logsumexp)
//   BinghamP  geosss/distributions.py:78-86, Fisher term :113-114
//   CurvedP   geosss/distributions.py:272-275 -> spherical_curve.py:95-102 -> :10-32
//   CpdP      geosss/registration.py:47-53, :101-109, :211-259 (+ GMM parent :111-119)
#pragma once
#include "common.cuh"

namespace geosss {

// scipy.special.logsumexp (>= 1.15) separates the maximal elements: with m maxima and
// s = sum over the others of exp(a - a_max),  out = log1p(s / m) + log(m) + a_max.
template <typename T> GEOSSS_HD T lse_finish(T a_max, T s_others, int n_max) {
    if (!(a_max > Const<T>::neg_inf())) return a_max;  // all -inf (or NaN) stays what it is
    if (a_max == INFINITY) return a_max;
    T m = (T)n_max;
    T s = (s_others == T(0)) ? s_others : s_others / m;
    return m_log1p(s) + m_log(m) + a_max;
}

// ------------------------------------------------------------------ vMF mixture
template <typename T> struct VmfMixP {
    const T* mu;     // (K, d) row-major
    const T* logw;   // (K)  log of the normalised weights
    const T* logi0;  // (K)  log(i0(|mu_k|)), +inf when i0 overflows, as in the reference
    int K, d, single;

    template <class Vec> GEOSSS_HD T comp(int k, Vec v) const {
        T dot = 0;
        const T* m = mu + (int64_t)k * d;
        for (int j = 0; j < d; ++j) dot += v[j] * m[j];
        return (dot - Const<T>::log_two_pi()) - logi0[k];
    }
    template <class Vec> GEOSSS_HD T logp(Vec v) const {
        if (single) return comp(0, v);
        T a_max = Const<T>::neg_inf();
        int n_max = 0;
        for (int k = 0; k < K; ++k) {
            T a = comp(k, v) + logw[k];
            if (a > a_max) { a_max = a; n_max = 1; }
            else if (a == a_max) ++n_max;
            else if (a != a) return a;  // NaN propagates
        }
        if (n_max == 0) n_max = 1;
        T s = 0;
        if (a_max > Const<T>::neg_inf() && a_max < INFINITY)
            for (int k = 0; k < K; ++k) {
                T a = comp(k, v) + logw[k];
                if (a != a_max) s += m_exp(a - a_max);
            }
        return lse_finish(a_max, s, n_max);
    }
    // distributions.py:159-160 (component: mu) and :223-227 (mixture: responsibilities times mu).  The
    // reference forms exp(p_k) / exp(logsumexp p) without a shift, which is 0/0 far from every mode;
    // the shifted softmax here is the same number wherever that is finite.
    template <class Vec, class Out> GEOSSS_HD void grad(Vec v, Out g) const {
        if (single) {
            for (int j = 0; j < d; ++j) g[j] = mu[j];
            return;
        }
        T a_max = Const<T>::neg_inf();
        for (int k = 0; k < K; ++k) {
            T a = comp(k, v) + logw[k];
            if (a > a_max) a_max = a;
        }
        for (int j = 0; j < d; ++j) g[j] = T(0);
        T tot = 0;
        for (int k = 0; k < K; ++k) {
            T e = m_exp((comp(k, v) + logw[k]) - a_max);
            tot += e;
            const T* m = mu + (int64_t)k * d;
            for (int j = 0; j < d; ++j) g[j] += e * m[j];
        }
        for (int j = 0; j < d; ++j) g[j] = g[j] / tot;
    }
};

// ------------------------------------------------------------------ Bingham (+ Fisher term)
template <typename T> struct BinghamP {
    const T* A;  // (d, d) row-major, symmetric
    const T* b;  // (d) or nullptr
    int d;

    template <class Vec> GEOSSS_HD T logp(Vec v) const {
        T tot = 0;
        for (int j = 0; j < d; ++j) {
            T col = 0;
            for (int i = 0; i < d; ++i) col += v[i] * A[(int64_t)i * d + j];  // (x @ A)_j
            tot += col * v[j];
        }
        if (b) {
            T lin = 0;
            for (int j = 0; j < d; ++j) lin += v[j] * b[j];
            tot += lin;
        }
        return tot;
    }
    // 2 A x (distributions.py:88-89); BinghamFisher inherits it, so b does not enter (kept as upstream)
    template <class Vec, class Out> GEOSSS_HD void grad(Vec v, Out g) const {
        for (int i = 0; i < d; ++i) {
            T row = 0;
            for (int j = 0; j < d; ++j) row += (T(2) * A[(int64_t)i * d + j]) * v[j];
            g[i] = row;
        }
    }
};

// ------------------------------------------------------------------ curved vMF
// x . y_seg depends on x only through G = knots @ x; per segment
//   t = clip(atan2(G_b - G_a cos(ang), G_a sin(ang)), 0, ang)
//   x.y = (sin(ang - t) G_a + sin(t) G_b) / (sin(ang) + 1e-10)
// and the segment with the smallest arccos(clip(x.y)) wins, first index on ties.
template <typename T> struct CurvedP {
    const T* knots;   // (M, d)
    const T* ang;     // (M-1) arccos(clip(a.b))
    const T* cosang;  // (M-1)
    const T* sinang;  // (M-1)
    T kappa;
    int M, d;

    template <class Vec> GEOSSS_HD T knot_dot(int m, Vec v) const {
        T g = 0;
        const T* kn = knots + (int64_t)m * d;
        for (int j = 0; j < d; ++j) g += kn[j] * v[j];
        return g;
    }
    GEOSSS_HD void segment(int j, T ga, T gb, T& dot, T& dist) const {
        T t = m_atan2(gb - ga * cosang[j], ga * sinang[j]);
        t = m_clip(t, T(0), ang[j]);
        dot = (m_sin(ang[j] - t) * ga + m_sin(t) * gb) / (sinang[j] + T(1e-10));
        dist = m_acos(m_clip(dot, T(-1), T(1)));
    }
    template <class Vec> GEOSSS_HD T logp(Vec v) const {
        T ga = knot_dot(0, v);
        T best_dot = 0, best_dist = 0;
        for (int j = 0; j + 1 < M; ++j) {
            T gb = knot_dot(j + 1, v);
            T dot, dist;
            segment(j, ga, gb, dot, dist);
            if (j == 0 || dist < best_dist) { best_dist = dist; best_dot = dot; }
            ga = gb;
        }
        return kappa * best_dot;
    }
    // kappa * nearest curve point (distributions.py:277-278, spherical_curve.py:27-32,95-102)
    template <class Vec, class Out> GEOSSS_HD void grad(Vec v, Out g) const {
        T ga = knot_dot(0, v);
        T best_dist = 0, best_t = 0;
        int best = 0;
        for (int j = 0; j + 1 < M; ++j) {
            T gb = knot_dot(j + 1, v);
            T t = m_clip(m_atan2(gb - ga * cosang[j], ga * sinang[j]), T(0), ang[j]);
            T dot = (m_sin(ang[j] - t) * ga + m_sin(t) * gb) / (sinang[j] + T(1e-10));
            T dist = m_acos(m_clip(dot, T(-1), T(1)));
            if (j == 0 || dist < best_dist) { best_dist = dist; best_t = t; best = j; }
            ga = gb;
        }
        const T ca = m_sin(ang[best] - best_t), cb = m_sin(best_t), den = sinang[best] + T(1e-10);
        const T* a = knots + (int64_t)best * d;
        const T* b = a + d;
        for (int j = 0; j < d; ++j) g[j] = kappa * ((ca * a[j] + cb * b[j]) / den);
    }
};

// ------------------------------------------------------------------ CPD registration on S^3
constexpr int CPD_KMAX = 64;  // largest k of the k-nearest-neighbour truncation

template <typename T> struct CpdP {
    const T* tgt;    // (nt, 3)
    const T* wt;     // (nt)
    const T* src;    // (ns, 3)
    const T* logws;  // (ns) log source weights
    T sigma2, log_const, log_outlier, beta;
    int nt, ns, k, outlier;  // k <= 0 or k >= ns: all source points; outlier 0: GMM parent
    T shift[3];              // translation added after the rotation (pointcloud.py:260-264)
    int dim;                 // 3, or 2: rotate then drop z (RotationProjection, pointcloud.py:273-293)
};

// scipy Rotation.from_quat(q).as_matrix(): scalar-last, q normalised first (pointcloud.py:115)
template <typename T> GEOSSS_HD void quat_to_matrix(T qx, T qy, T qz, T qw, T R[9]) {
    T n = m_sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
    T x = qx / n, y = qy / n, z = qz / n, w = qw / n;
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

// moved_k = R s_k + t  (positions @ R.T + translation, pointcloud.py:258-264), k = first, first+step, ...
// 2-D targets: positions @ R[:-1].T + t (pointcloud.py:288-293); z is held at 0 like the zero-padded
// target rows, so the same 3-D distance code serves both.
template <typename T> GEOSSS_HD void cpd_move_source(const CpdP<T>& p, const T R[9], T* moved, int first, int step) {
    for (int k = first; k < p.ns; k += step) {
        T s0 = p.src[3 * k], s1 = p.src[3 * k + 1], s2 = p.src[3 * k + 2];
        moved[3 * k]     = (s0 * R[0] + s1 * R[1] + s2 * R[2]) + p.shift[0];
        moved[3 * k + 1] = (s0 * R[3] + s1 * R[4] + s2 * R[5]) + p.shift[1];
        moved[3 * k + 2] = p.dim == 2 ? T(0) : (s0 * R[6] + s1 * R[7] + s2 * R[8]) + p.shift[2];
    }
}

template <typename T> GEOSSS_HD T cpd_term(const CpdP<T>& p, T d2, T lw) {
    T log_phi = (T(-0.5) * d2) / p.sigma2;
    return p.outlier ? (lw + log_phi) + p.log_const : (log_phi + p.log_const) + lw;
}

// log-likelihood of target row l given the moved source cloud, times the row weight
template <typename T> GEOSSS_HD T cpd_row(const CpdP<T>& p, const T* moved, int l) {
    const T t0 = p.tgt[3 * l], t1 = p.tgt[3 * l + 1], t2 = p.tgt[3 * l + 2];
    const bool dense = p.k <= 0 || p.k >= p.ns;
    T a_max = p.outlier ? p.log_outlier : Const<T>::neg_inf();
    int n_max = p.outlier ? 1 : 0;
    T s = 0;
    if (dense) {
        for (int k = 0; k < p.ns; ++k) {
            T e0 = t0 - moved[3 * k], e1 = t1 - moved[3 * k + 1], e2 = t2 - moved[3 * k + 2];
            T a = cpd_term(p, e0 * e0 + e1 * e1 + e2 * e2, p.logws[k]);
            if (a > a_max) { a_max = a; n_max = 1; } else if (a == a_max) ++n_max;
        }
        if (a_max > Const<T>::neg_inf() && a_max < INFINITY) {
            for (int k = 0; k < p.ns; ++k) {
                T e0 = t0 - moved[3 * k], e1 = t1 - moved[3 * k + 1], e2 = t2 - moved[3 * k + 2];
                T a = cpd_term(p, e0 * e0 + e1 * e1 + e2 * e2, p.logws[k]);
                if (a != a_max) s += m_exp(a - a_max);
            }
            if (p.outlier && p.log_outlier != a_max) s += m_exp(p.log_outlier - a_max);
        }
    } else {
        // k nearest moved source points (scipy KDTree.query, registration.py:107-108)
        T bd[CPD_KMAX], bl[CPD_KMAX];
        int cnt = 0;
        const int kk = p.k;
        for (int k = 0; k < p.ns; ++k) {
            T e0 = t0 - moved[3 * k], e1 = t1 - moved[3 * k + 1], e2 = t2 - moved[3 * k + 2];
            T d2 = e0 * e0 + e1 * e1 + e2 * e2;
            if (cnt < kk || d2 < bd[cnt - 1]) {
                int i = cnt < kk ? cnt++ : kk - 1;
                while (i > 0 && bd[i - 1] > d2) { bd[i] = bd[i - 1]; bl[i] = bl[i - 1]; --i; }
                bd[i] = d2; bl[i] = p.logws[k];
            }
        }
        for (int i = 0; i < cnt; ++i) {
            T a = cpd_term(p, bd[i], bl[i]);
            bd[i] = a;
            if (a > a_max) { a_max = a; n_max = 1; } else if (a == a_max) ++n_max;
        }
        if (a_max > Const<T>::neg_inf() && a_max < INFINITY) {
            for (int i = 0; i < cnt; ++i)
                if (bd[i] != a_max) s += m_exp(bd[i] - a_max);
            if (p.outlier && p.log_outlier != a_max) s += m_exp(p.log_outlier - a_max);
        }
    }
    if (n_max == 0) n_max = 1;
    return lse_finish(a_max, s, n_max) * p.wt[l];
}

// ---- gradient of the registration log-likelihood (registration.py:55-59, :261-293) -------------
// Row l adds  w_l / sigma^2 * sum_k gamma_lk (y_l - R x_k) x_k^T  to G (3x3, row-major), with
// gamma_lk = exp(clip(log p_lk - log p_l, -20, 0)) and log p_l the row's logsumexp (outlier column
// included for CPD).  G = -grad_R of the reference; the two minus signs of upstream (grad_R and the
// Jacobian) cancel in cpd_grad_finish.
template <typename T> GEOSSS_HD void cpd_row_grad(const CpdP<T>& p, const T* moved, int l, T G[9]) {
    const T t0 = p.tgt[3 * l], t1 = p.tgt[3 * l + 1], t2 = p.tgt[3 * l + 2];
    const bool dense = p.k <= 0 || p.k >= p.ns;
    T a_max = p.outlier ? p.log_outlier : Const<T>::neg_inf();
    int n_max = p.outlier ? 1 : 0;
    T s = 0;
    T bd[CPD_KMAX];
    int bi[CPD_KMAX];
    int cnt = 0;
    if (dense) {
        for (int k = 0; k < p.ns; ++k) {
            T e0 = t0 - moved[3 * k], e1 = t1 - moved[3 * k + 1], e2 = t2 - moved[3 * k + 2];
            T a = cpd_term(p, e0 * e0 + e1 * e1 + e2 * e2, p.logws[k]);
            if (a > a_max) { a_max = a; n_max = 1; } else if (a == a_max) ++n_max;
        }
        for (int k = 0; k < p.ns; ++k) {
            T e0 = t0 - moved[3 * k], e1 = t1 - moved[3 * k + 1], e2 = t2 - moved[3 * k + 2];
            T a = cpd_term(p, e0 * e0 + e1 * e1 + e2 * e2, p.logws[k]);
            if (a != a_max) s += m_exp(a - a_max);
        }
    } else {
        const int kk = p.k;
        for (int k = 0; k < p.ns; ++k) {
            T e0 = t0 - moved[3 * k], e1 = t1 - moved[3 * k + 1], e2 = t2 - moved[3 * k + 2];
            T d2 = e0 * e0 + e1 * e1 + e2 * e2;
            if (cnt < kk || d2 < bd[cnt - 1]) {
                int i = cnt < kk ? cnt++ : kk - 1;
                while (i > 0 && bd[i - 1] > d2) { bd[i] = bd[i - 1]; bi[i] = bi[i - 1]; --i; }
                bd[i] = d2; bi[i] = k;
            }
        }
        for (int i = 0; i < cnt; ++i) {
            T a = cpd_term(p, bd[i], p.logws[bi[i]]);
            bd[i] = a;
            if (a > a_max) { a_max = a; n_max = 1; } else if (a == a_max) ++n_max;
        }
        for (int i = 0; i < cnt; ++i)
            if (bd[i] != a_max) s += m_exp(bd[i] - a_max);
    }
    if (p.outlier && p.log_outlier != a_max) s += m_exp(p.log_outlier - a_max);
    if (n_max == 0) n_max = 1;
    const T log_pl = lse_finish(a_max, s, n_max);
    const T wl = p.wt[l];
    const int n = dense ? p.ns : cnt;
    for (int i = 0; i < n; ++i) {
        const int k = dense ? i : bi[i];
        const T e0 = t0 - moved[3 * k], e1 = t1 - moved[3 * k + 1], e2 = t2 - moved[3 * k + 2];
        const T a = dense ? cpd_term(p, e0 * e0 + e1 * e1 + e2 * e2, p.logws[k]) : bd[i];
        const T c = (wl * m_exp(m_clip(a - log_pl, T(-20), T(0)))) / p.sigma2;
        const T s0 = p.src[3 * k], s1 = p.src[3 * k + 1], s2 = p.src[3 * k + 2];
        G[0] += c * e0 * s0; G[1] += c * e0 * s1; G[2] += c * e0 * s2;
        G[3] += c * e1 * s0; G[4] += c * e1 * s1; G[5] += c * e1 * s2;
        G[6] += c * e2 * s0; G[7] += c * e2 * s1; G[8] += c * e2 * s2;
    }
}

// grad_q[c] = beta * sum_ji G[j][i] dR[j][i]/dq_c for R(q) = N(q) / r, r = |q|^2 + 1e-300 (the quotient-rule
// form of pointcloud.py:135-203; q is not assumed to be a unit vector)
template <typename T, class Out> GEOSSS_HD void cpd_grad_finish(const CpdP<T>& p, T x, T y, T z, T w, const T G[9], Out g) {
    const T r = ((w * w + x * x) + y * y) + z * z + Const<T>::tiny_r();
    const T Nm[9] = {w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z};
    const T dN[4][9] = {{x, y, z, y, -x, -w, z, w, -x},
                        {-y, x, w, x, y, z, -w, z, -y},
                        {-z, -w, x, w, -z, y, x, y, z},
                        {w, -z, y, z, w, -x, -y, x, w}};
    const T q[4] = {x, y, z, w};
    T GN = 0;
    for (int i = 0; i < 9; ++i) GN += G[i] * Nm[i];
    for (int c = 0; c < 4; ++c) {
        T acc = 0;
        for (int i = 0; i < 9; ++i) acc += G[i] * (T(2) * dN[c][i]);
        g[c] = p.beta * (acc / r - GN * (T(2) * q[c]) / (r * r));
    }
}

}  // namespace geosss
