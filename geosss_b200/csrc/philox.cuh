// Per-chain counter-based random stream (Philox4x32-10, Salmon et al. 2011) and the two
// draw sources the samplers read from: the Philox stream (production) and a flat float64
// tape (parity mode: the reference's own draws, recorded by oracle/tape.py).
//
// Replaces numpy's Generator calls in geosss/mcmc.py:362-369, 387-395.  PCG64 + ziggurat
// cannot be reproduced by a counter-based generator, so the layout below is ours; it is
// restated in oracle/philox.py, which the tests use to feed the SAME draws to the reference
// algorithm on the CPU.
//
//   key     = (seed lo, seed hi)
//   counter = (chain lo, chain hi, step lo, (step hi16 << 16) | block)
//   block b < nb = ceil(d/4): normals 4b..4b+3 (Box-Muller on words (0,1) and (2,3))
//   aux words (slice-level uniform, shrink-bracket uniform): words 2,3 of block nb-1 when
//     d % 4 is 1 or 2 (that Box-Muller pair would be unused), else words 0,1 of block nb
//   proposals start at block pb (= nb, resp. nb+1): proposal j = word j%4 of block pb + j/4,
//     so one Philox block feeds one batch of four proposals
// chain is the GLOBAL chain id, so a chain's trajectory does not depend on how chains are
// split over GPUs, blocks or launches.
#pragma once
#include "common.cuh"

namespace geosss {

struct U4 { uint32_t x, y, z, w; };

GEOSSS_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

GEOSSS_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(0xD2511F53u, c.x, hi0, lo0);
        mulhilo(0xCD9E8D57u, c.z, hi1, lo1);
        c = U4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// uint32 -> uniform strictly inside (0,1).  fp32: top 23 bits, exact; fp64: all 32 bits.
GEOSSS_HD float u01(uint32_t w, float) { return fmaf((float)(w >> 9), 1.1920928955078125e-07f, 5.9604644775390625e-08f); }
GEOSSS_HD double u01(uint32_t w, double) { return ((double)w + 0.5) * 2.3283064365386963e-10; }

template <typename T> GEOSSS_HD void box_muller(uint32_t wa, uint32_t wb, T& z0, T& z1) {
    T u1 = u01(wa, T(0)), u2 = u01(wb, T(0));
    T r = m_sqrt(T(-2) * m_log(u1));
    T a = Const<T>::two_pi() * u2;
    z0 = r * m_cos(a);
    z1 = r * m_sin(a);
}

// ------------------------------------------------------------------------------------------
GEOSSS_HD bool philox_aux_in_last(int d) { return (d & 3) == 1 || (d & 3) == 2; }

template <typename T> struct PhiloxDraws {
    uint32_t k0, k1, c0, c1, s_lo, s_hi;
    uint32_t aux0, aux1;  // level / bracket words of the current transition
    int nb, pb, d_;
    int cached;  // index of the proposal block held in w (-1: none)
    U4 w;

    GEOSSS_HD void init(uint64_t seed, uint64_t chain, int d) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        c0 = (uint32_t)chain; c1 = (uint32_t)(chain >> 32);
        d_ = d;
        nb = (d + 3) >> 2;
        pb = philox_aux_in_last(d) ? nb : nb + 1;
        cached = -1;
    }
    GEOSSS_HD U4 block(int b) const { return philox4x32_10(U4{c0, c1, s_lo, (s_hi << 16) | (uint32_t)b}, k0, k1); }
    GEOSSS_HD void begin_step(uint64_t step) {
        s_lo = (uint32_t)step; s_hi = (uint32_t)(step >> 32) & 0xFFFFu;
        cached = -1;
        if (!philox_aux_in_last(d_)) { U4 a = block(nb); aux0 = a.x; aux1 = a.y; }
    }
    template <class Vec> GEOSSS_HD void fill_normals(Vec z, int d) {
        for (int b = 0; b < nb; ++b) {
            U4 r = block(b);
            T n0, n1, n2, n3;
            box_muller<T>(r.x, r.y, n0, n1);
            int j = 4 * b;
            z[j] = n0;
            if (j + 1 < d) z[j + 1] = n1;
            if (j + 2 < d) {
                box_muller<T>(r.z, r.w, n2, n3);
                z[j + 2] = n2;
                if (j + 3 < d) z[j + 3] = n3;
            } else { aux0 = r.z; aux1 = r.w; }
        }
    }
    GEOSSS_HD T level_uniform() { return u01(aux0, T(0)); }
    GEOSSS_HD T bracket_uniform() { return u01(aux1, T(0)); }
    GEOSSS_HD T proposal_uniform(int j) {
        int b = pb + (j >> 2);
        if (b != cached) { w = block(b); cached = b; }
        int i = j & 3;
        return u01(i == 0 ? w.x : i == 1 ? w.y : i == 2 ? w.z : w.w, T(0));
    }
    GEOSSS_HD bool ok() const { return true; }
    // most proposals a transition may draw before the 16-bit block index would wrap
    GEOSSS_HD int max_proposals() const { return MAX_PROPOSALS; }
};

// Flat float64 tape in consumption order: d normals, level uniform, [bracket uniform,]
// one uniform per proposal; then the next transition.
template <typename T> struct TapeDraws {
    const double* t;
    int64_t cur, end;
    bool good;

    GEOSSS_HD void init(const double* tape, int64_t begin, int64_t end_) { t = tape; cur = begin; end = end_; good = true; }
    GEOSSS_HD void begin_step(uint64_t) {}
    GEOSSS_HD T next() {
        if (cur >= end) { good = false; return T(0.5); }
        return (T)t[cur++];
    }
    template <class Vec> GEOSSS_HD void fill_normals(Vec z, int d) {
        for (int j = 0; j < d; ++j) z[j] = next();
    }
    GEOSSS_HD T level_uniform() { return next(); }
    GEOSSS_HD T bracket_uniform() { return next(); }
    GEOSSS_HD T proposal_uniform(int) { return next(); }
    GEOSSS_HD bool ok() const { return good; }
    GEOSSS_HD int max_proposals() const { return MAX_PROPOSALS; }
};

}  // namespace geosss
