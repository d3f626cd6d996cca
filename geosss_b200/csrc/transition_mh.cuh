// One transition of the baseline samplers (random-walk MH, independence, their mixture, spherical HMC)
// and the two draw sources they read.  Host-compilable (tests/hostsim); the kernels are in kernels_mh.cuh.
//
//   geosss/mcmc.py:78-112   AdaptiveStepsize          (x 1.02 / x 0.98 during the first `burnin` transitions)
//   geosss/mcmc.py:115-173  MetropolisHastings        (radius sqrt(2 Gamma(d/2)), perturb, project back)
//   geosss/mcmc.py:176-179  IndependenceSampler
//   geosss/mcmc.py:182-234  MixtureRWMHIndependenceSampler
//   geosss/mcmc.py:237-332  SphericalHMC              (tangent momenta, geodesic leapfrog)
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace geosss {

enum { MH_RWMH = 2, MH_INDEPENDENCE = 3, MH_MIXTURE = 4, MH_HMC = 5 };

// Sequential reader of the chain's Philox stream (restated in oracle/philox.py: PhiloxSeqRNG): the words
// of blocks 0, 1, ... of the transition are consumed in order; a uniform takes one word, normals come in
// Box-Muller pairs (the second is kept), the radius draw Gamma(d/2) is half a chi-square of d normals.
template <typename T> struct PhiloxSeq {
    uint32_t k0, k1, c0, c1, s_lo, s_hi;
    int word;
    U4 blk;
    bool has_spare;
    T spare;

    GEOSSS_HD void init(uint64_t seed, uint64_t chain) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        c0 = (uint32_t)chain; c1 = (uint32_t)(chain >> 32);
        word = 0; has_spare = false;
    }
    GEOSSS_HD void begin_step(uint64_t step) {
        s_lo = (uint32_t)step; s_hi = (uint32_t)(step >> 32) & 0xFFFFu;
        word = 0; has_spare = false;
    }
    GEOSSS_HD uint32_t next_word() {
        const int i = word & 3;
        if (i == 0) blk = philox4x32_10(U4{c0, c1, s_lo, (s_hi << 16) | (uint32_t)(word >> 2)}, k0, k1);
        ++word;
        return i == 0 ? blk.x : i == 1 ? blk.y : i == 2 ? blk.z : blk.w;
    }
    GEOSSS_HD T normal() {
        if (has_spare) { has_spare = false; return spare; }
        const uint32_t wa = next_word(), wb = next_word();
        T z0;
        box_muller<T>(wa, wb, z0, spare);
        has_spare = true;
        return z0;
    }
    GEOSSS_HD T uniform() { return u01(next_word(), T(0)); }
    GEOSSS_HD T gamma_half(int d) {
        T acc = 0;
        for (int j = 0; j < d; ++j) { const T z = normal(); acc = acc + z * z; }
        return T(0.5) * acc;
    }
    GEOSSS_HD bool ok() const { return true; }
};

// the reference's own draws in consumption order (oracle/tape.py); the gamma variate itself is on the tape
template <typename T> struct TapeSeq {
    TapeDraws<T> t;
    GEOSSS_HD void begin_step(uint64_t) {}
    GEOSSS_HD T normal() { return t.next(); }
    GEOSSS_HD T uniform() { return t.next(); }
    GEOSSS_HD T gamma_half(int) { return t.next(); }
    GEOSSS_HD bool ok() const { return t.ok(); }
};

struct MhParams {
    double alpha;         // mixture: probability of a random-walk proposal
    int n_leap;           // HMC leapfrog steps
    int64_t adapt_total;  // burn-in length of the step-size adaptation
};

struct MhChain {
    double stepsize;      // float64 as upstream, whatever the state dtype
    int64_t adapt_count;
    unsigned long long n_walk;
};

// One transition.  x, v: current position / momentum (v is only used by HMC); xp, vp, g: scratch.
// f.logp(view), f.grad(view, out).  lp: log p(x), updated on acceptance.  Returns true if accepted.
template <typename T, class Draws, class Vec, class F>
GEOSSS_HD bool mh_transition(int kind, int d, Vec x, Vec v, Vec xp, Vec vp, Vec g, Draws& dr, const F& f, MhChain& c,
                             const MhParams& a, T& lp) {
    bool walk = kind == MH_RWMH;
    if (kind == MH_MIXTURE) {
        walk = dr.uniform() < (T)a.alpha;
        c.n_walk += walk ? 1ull : 0ull;
    }
    const T eps = (T)c.stepsize;
    T delta, lp_new;
    if (kind == MH_HMC) {
        // momentum refresh: z - x (x . z)  (mcmc.py:272-278)
        T xz = 0;
        for (int j = 0; j < d; ++j) { const T z = dr.normal(); v[j] = z; xz += x[j] * z; }
        T kin0 = 0;
        for (int j = 0; j < d; ++j) {
            const T p = v[j] - x[j] * xz;
            v[j] = p; vp[j] = p; xp[j] = x[j];
            kin0 += p * p;
        }
        for (int t = -1; t < a.n_leap; ++t) {
            if (t >= 0) {  // geodesic flow for time eps (mcmc.py:302-306)
                T vv = 0;
                for (int j = 0; j < d; ++j) vv += vp[j] * vp[j];
                const T nv = m_sqrt(vv);
                const T cs = m_cos(eps * nv), sn = m_sin(eps * nv);
                for (int j = 0; j < d; ++j) {
                    const T y = xp[j];
                    xp[j] = y * cs + (vp[j] / nv) * sn;
                    vp[j] = vp[j] * cs - (y * nv) * sn;
                }
            }
            // kick: half steps first and last (mcmc.py:298, 308-311)
            f.grad(xp, g);
            T xg = 0;
            for (int j = 0; j < d; ++j) xg += xp[j] * g[j];
            const T w = (t < 0 || t == a.n_leap - 1) ? T(0.5) * eps : eps;
            for (int j = 0; j < d; ++j) vp[j] += w * (g[j] - xp[j] * xg);
        }
        T xx = 0, kin1 = 0;
        for (int j = 0; j < d; ++j) { xx += xp[j] * xp[j]; kin1 += vp[j] * vp[j]; }
        const T nx = m_sqrt(xx) + Const<T>::tiny();
        for (int j = 0; j < d; ++j) xp[j] = xp[j] / nx;
        lp_new = f.logp(xp);
        delta = (T(0.5) * kin0 - lp) - (T(0.5) * kin1 - lp_new);
    } else {
        if (walk) {
            const T r = m_sqrt(T(2) * dr.gamma_half(d));
            for (int j = 0; j < d; ++j) xp[j] = r * x[j] + eps * dr.normal();
        } else {
            for (int j = 0; j < d; ++j) xp[j] = dr.normal();
        }
        T yy = 0;
        for (int j = 0; j < d; ++j) yy += xp[j] * xp[j];
        const T ny = m_sqrt(yy) + Const<T>::tiny();
        for (int j = 0; j < d; ++j) xp[j] = xp[j] / ny;
        lp_new = f.logp(xp);
        delta = lp_new - lp;
    }
    const bool accepted = m_log(dr.uniform()) < delta;
    if ((kind != MH_MIXTURE || walk) && c.adapt_count < a.adapt_total) {
        c.stepsize *= accepted ? 1.02 : 0.98;
        ++c.adapt_count;
    }
    if (accepted) {
        for (int j = 0; j < d; ++j) x[j] = xp[j];
        if (kind == MH_HMC)
            for (int j = 0; j < d; ++j) v[j] = vp[j];
        lp = lp_new;
    }
    return accepted;
}

}  // namespace geosss
