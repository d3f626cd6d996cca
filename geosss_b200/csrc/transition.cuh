// One geodesic slice-sampling transition, "literal" form (reference op order).
//
//   geosss/mcmc.py:357-374  RejectionSphericalSliceSampler.__next__
//   geosss/mcmc.py:382-401  ShrinkageSphericalSliceSampler.__next__
//   geosss/sphere.py:10-33  radial / orthogonal / spherical projection
//
// Kept from the reference because float64 tape parity depends on it: the state is never
// renormalised, the pole is normalised as x / (|x| + 1e-100), log p(x) is re-evaluated at
// every transition, acceptance is a strict '>', the shrink side is chosen by theta < 0.
// Added: a proposal cap (the reference loops for ever); past it the chain stays at x.
#pragma once
#include "common.cuh"

namespace geosss {

// x, u, y: views of length d (u receives the direction, y is proposal scratch).
// Returns the number of rejected proposals; *lp_new receives log p(new state).
// lp_x_io (optional, throughput kernels only): log p(x) carried from the previous transition -- NaN means
// "not known yet"; it is updated when a proposal is accepted and left alone when the chain stays.  log p is a
// pure function and the accepted y IS the next x, so the carried value equals the re-evaluated one bit for bit.
template <typename T, int KIND, class Draws, class Vec, class LogP>
GEOSSS_HD int transition_literal(int d, Vec x, Vec u, Vec y, Draws& dr, const LogP& logp, T* lp_new, T* lp_x_io = nullptr) {
    // direction: u = unit(z - (z.n) n),  n = x / (|x| + tiny)
    dr.fill_normals(u, d);
    T xx = 0;
    for (int j = 0; j < d; ++j) xx += x[j] * x[j];
    const T nx = m_sqrt(xx) + Const<T>::tiny();
    T zn = 0;
    for (int j = 0; j < d; ++j) zn += u[j] * (x[j] / nx);
    T pp = 0;
    for (int j = 0; j < d; ++j) {
        T p = u[j] - zn * (x[j] / nx);
        u[j] = p;
        pp += p * p;
    }
    const T np_ = m_sqrt(pp) + Const<T>::tiny();
    for (int j = 0; j < d; ++j) u[j] = u[j] / np_;

    T lp_x;
    if (lp_x_io && *lp_x_io == *lp_x_io) lp_x = *lp_x_io;
    else {
        lp_x = logp(x);
        if (lp_x_io) *lp_x_io = lp_x;
    }
    const T level = lp_x + m_log(dr.level_uniform());

    T lo, hi;
    if (KIND == SHRINK) {
        hi = T(0) + (Const<T>::two_pi() - T(0)) * dr.bracket_uniform();
        lo = hi - Const<T>::two_pi();
    } else {
        lo = T(0);
        hi = Const<T>::two_pi();
    }
    const int cap = dr.max_proposals();
    for (int r = 0; r < cap; ++r) {
        const T theta = lo + (hi - lo) * dr.proposal_uniform(r);
        const T c = m_cos(theta), s = m_sin(theta);
        for (int j = 0; j < d; ++j) y[j] = c * x[j] + s * u[j];
        const T ly = logp(y);
        if (ly > level) {
            for (int j = 0; j < d; ++j) x[j] = y[j];
            *lp_new = ly;
            if (lp_x_io) *lp_x_io = ly;
            return r;
        }
        if (KIND == SHRINK) {
            if (theta < T(0)) lo = theta; else hi = theta;
            // Bracket collapsed onto theta = 0: every further proposal IS x, and log p(x) > level can
            // only fail because "log p(x) + log U" rounded onto log p(x) (in
            // fp32 with |log p| ~ 2400 this happens once per ~1e4 transitions).  The reference would
            // spin for ever; the chain stays at x, which is always inside the slice.
            if (!(hi - lo > Const<T>::bracket_floor())) { *lp_new = level; return r + 1; }
        }
        if (!dr.ok()) break;
    }
    *lp_new = level;  // not used for decisions; the chain stays where it is
    return cap;
}

}  // namespace geosss
