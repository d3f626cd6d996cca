// TEST-ONLY host build of the device math (targets.cuh, transition.cuh, transition_mh.cuh, philox.cuh).
//
// Compiled with g++ by tests/test_hostsim.py so that the per-chain arithmetic the CUDA kernels
// run can be checked against the golden vectors on a machine without a GPU.  It is not part of
// the product: geosss_b200 never loads it, and it contains no kernels -- chains are looped
// serially and the warp-cooperative CPD evaluation is emulated by summing rows in order.
#include <cstdint>
#include <cstring>
#include <vector>

#include "eigh_jacobi.h"
#include "philox.cuh"
#include "targets.cuh"
#include "transition.cuh"
#include "transition_mh.cuh"

using namespace geosss;

namespace {

template <typename T> std::vector<T> cast(const double* p, size_t n) {
    std::vector<T> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = (T)p[i];
    return v;
}

// parameter pack mirrors the C ABI's create calls; arrays are float64 on input
struct Spec {
    int kind, d;
    // vmf
    int K, single; const double *mu, *logw, *logi0;
    // bingham
    const double *A, *b;
    // curved
    int M; double kappa; const double *knots, *ang, *cosang, *sinang;
    // cpd
    int nt, ns, k, outlier; double sigma2, log_const, log_outlier, beta; const double *tgt, *wt, *src, *logws;
    double shift[3]; int dim;
};

template <typename T> struct Built {
    std::vector<T> a0, a1, a2, a3;
    VmfMixP<T> vmf; BinghamP<T> bing; CurvedP<T> curved; CpdP<T> cpd;
    explicit Built(const Spec& s) {
        switch (s.kind) {
            case TGT_VMF_MIX:
                a0 = cast<T>(s.mu, (size_t)s.K * s.d); a1 = cast<T>(s.logw, s.K); a2 = cast<T>(s.logi0, s.K);
                vmf = {a0.data(), a1.data(), a2.data(), s.K, s.d, s.single};
                break;
            case TGT_BINGHAM:
                a0 = cast<T>(s.A, (size_t)s.d * s.d);
                if (s.b) a1 = cast<T>(s.b, s.d);
                bing = {a0.data(), s.b ? a1.data() : nullptr, s.d};
                break;
            case TGT_CURVED_VMF:
                a0 = cast<T>(s.knots, (size_t)s.M * s.d); a1 = cast<T>(s.ang, s.M - 1);
                a2 = cast<T>(s.cosang, s.M - 1); a3 = cast<T>(s.sinang, s.M - 1);
                curved = {a0.data(), a1.data(), a2.data(), a3.data(), (T)s.kappa, s.M, s.d};
                break;
            default:
                a0 = cast<T>(s.tgt, (size_t)s.nt * 3); a1 = cast<T>(s.wt, s.nt);
                a2 = cast<T>(s.src, (size_t)s.ns * 3); a3 = cast<T>(s.logws, s.ns);
                cpd = {a0.data(), a1.data(), a2.data(), a3.data(), (T)s.sigma2, (T)s.log_const, (T)s.log_outlier,
                       (T)s.beta, s.nt, s.ns, s.k, s.outlier, {(T)s.shift[0], (T)s.shift[1], (T)s.shift[2]}, s.dim};
        }
    }
};

template <typename T> struct HostLogP {
    const Spec* s; const Built<T>* b; mutable std::vector<T> moved;
    template <class Vec> T operator()(Vec v) const {
        switch (s->kind) {
            case TGT_VMF_MIX: return b->vmf.logp(v);
            case TGT_BINGHAM: return b->bing.logp(v);
            case TGT_CURVED_VMF: return b->curved.logp(v);
            default: {
                T R[9];
                quat_to_matrix<T>(v[0], v[1], v[2], v[3], R);
                moved.resize((size_t)3 * b->cpd.ns);
                cpd_move_source(b->cpd, R, moved.data(), 0, 1);
                // emulate the warp reduction: 32 per-lane partial sums in double, then a butterfly
                double part[32] = {0};
                for (int l = 0; l < b->cpd.nt; ++l) part[l & 31] += (double)cpd_row(b->cpd, moved.data(), l);
                for (int o = 16; o > 0; o >>= 1)
                    for (int i = 0; i < 32; ++i) if ((i & o) == 0) { double t = part[i] + part[i ^ o]; part[i] = part[i ^ o] = t; }
                return b->cpd.beta * (T)part[0];
            }
        }
    }
};

// value + gradient as the baseline-sampler kernels see them (kernels_mh.cuh: LaneTarget / CpdWarpTarget)
template <typename T> struct HostTarget {
    HostLogP<T> lp;
    template <class Vec> T logp(Vec v) const { return lp(v); }
    template <class Vec, class Out> void grad(Vec v, Out g) const {
        const Built<T>* b = lp.b;
        switch (lp.s->kind) {
            case TGT_VMF_MIX: b->vmf.grad(v, g); break;
            case TGT_BINGHAM: b->bing.grad(v, g); break;
            case TGT_CURVED_VMF: b->curved.grad(v, g); break;
            default: {
                T R[9];
                quat_to_matrix<T>(v[0], v[1], v[2], v[3], R);
                lp.moved.resize((size_t)3 * b->cpd.ns);
                cpd_move_source(b->cpd, R, lp.moved.data(), 0, 1);
                T part[32][9] = {};
                for (int l = 0; l < b->cpd.nt; ++l) cpd_row_grad(b->cpd, lp.moved.data(), l, part[l & 31]);
                T G[9];
                for (int i = 0; i < 9; ++i) {
                    double s[32];
                    for (int k = 0; k < 32; ++k) s[k] = (double)part[k][i];
                    for (int o = 16; o > 0; o >>= 1)
                        for (int k = 0; k < 32; ++k) if ((k & o) == 0) { double t = s[k] + s[k ^ o]; s[k] = s[k ^ o] = t; }
                    G[i] = (T)s[0];
                }
                cpd_grad_finish(b->cpd, v[0], v[1], v[2], v[3], G, g);
            }
        }
    }
};

template <typename T> void grad_all(const Spec& s, const double* x, double* out, int64_t N) {
    Built<T> b(s);
    HostTarget<T> f{{&s, &b, {}}};
    std::vector<T> row(s.d), g(s.d);
    for (int64_t i = 0; i < N; ++i) {
        for (int j = 0; j < s.d; ++j) row[j] = (T)x[i * s.d + j];
        f.grad(SVec<T>{row.data(), 1}, SVec<T>{g.data(), 1});
        for (int j = 0; j < s.d; ++j) out[i * s.d + j] = (double)g[j];
    }
}

// baseline samplers: per transition the position, momentum, step size and whether it accepted
template <typename T, class Draws>
int run_mh_chain(const Spec& s, int kind, Draws& dr, const double* x0, int64_t n_steps, uint64_t step0, double stepsize,
                 MhParams mp, double* samples, double* momenta, double* stepsizes, int32_t* accepted, int64_t* n_walk) {
    Built<T> b(s);
    HostTarget<T> f{{&s, &b, {}}};
    const int d = s.d;
    std::vector<T> buf((size_t)5 * d, T(0));
    SVec<T> x{buf.data(), 1}, v{buf.data() + d, 1}, xp{buf.data() + 2 * d, 1}, vp{buf.data() + 3 * d, 1}, g{buf.data() + 4 * d, 1};
    for (int j = 0; j < d; ++j) x[j] = (T)x0[j];
    MhChain ch{stepsize, 0, 0ull};
    T lp = f.logp(x);
    for (int64_t t = 0; t < n_steps; ++t) {
        dr.begin_step(step0 + t);
        const bool ok = mh_transition<T>(kind, d, x, v, xp, vp, g, dr, f, ch, mp, lp);
        if (!dr.ok()) return 1;
        for (int j = 0; j < d; ++j) { samples[t * d + j] = (double)x[j]; momenta[t * d + j] = (double)v[j]; }
        stepsizes[t] = ch.stepsize;
        accepted[t] = ok ? 1 : 0;
    }
    *n_walk = (int64_t)ch.n_walk;
    return 0;
}

template <typename T> int run_mh_any(const Spec& s, int kind, const double* tape, int64_t tape_len, uint64_t seed, uint64_t chain,
                                     const double* x0, int64_t n_steps, uint64_t step0, double stepsize, MhParams mp, double* samples,
                                     double* momenta, double* stepsizes, int32_t* accepted, int64_t* n_walk) {
    if (tape) {
        TapeSeq<T> dr;
        dr.t.init(tape, 0, tape_len);
        return run_mh_chain<T>(s, kind, dr, x0, n_steps, step0, stepsize, mp, samples, momenta, stepsizes, accepted, n_walk);
    }
    PhiloxSeq<T> dr;
    dr.init(seed, chain);
    return run_mh_chain<T>(s, kind, dr, x0, n_steps, step0, stepsize, mp, samples, momenta, stepsizes, accepted, n_walk);
}

template <typename T> void logp_all(const Spec& s, const double* x, double* out, int64_t N) {
    Built<T> b(s);
    HostLogP<T> f{&s, &b, {}};
    std::vector<T> row(s.d);
    for (int64_t i = 0; i < N; ++i) {
        for (int j = 0; j < s.d; ++j) row[j] = (T)x[i * s.d + j];
        out[i] = (double)f(SVec<T>{row.data(), 1});
    }
}

template <typename T, int KIND, class Draws>
int run_chain(const Spec& s, const Built<T>& b, Draws& dr, const double* x0, int64_t n_steps, uint64_t step0,
              double* samples, int32_t* nrej) {
    HostLogP<T> f{&s, &b, {}};
    const int d = s.d;
    std::vector<T> xs(d), us(d), ys(d);
    for (int j = 0; j < d; ++j) xs[j] = (T)x0[j];
    SVec<T> x{xs.data(), 1}, u{us.data(), 1}, y{ys.data(), 1};
    for (int64_t t = 0; t < n_steps; ++t) {
        dr.begin_step(step0 + t);
        T lp;
        int r = transition_literal<T, KIND>(d, x, u, y, dr, f, &lp);
        if (!dr.ok()) return 1;
        for (int j = 0; j < d; ++j) samples[t * d + j] = (double)xs[j];
        nrej[t] = r;
    }
    return 0;
}

template <typename T> int run_any(const Spec& s, int kind, const double* tape, int64_t tape_len, uint64_t seed, uint64_t chain,
                                  const double* x0, int64_t n_steps, uint64_t step0, double* samples, int32_t* nrej) {
    Built<T> b(s);
    if (tape) {
        TapeDraws<T> dr;
        dr.init(tape, 0, tape_len);
        return kind == SHRINK ? run_chain<T, SHRINK>(s, b, dr, x0, n_steps, step0, samples, nrej)
                              : run_chain<T, REJECT>(s, b, dr, x0, n_steps, step0, samples, nrej);
    }
    PhiloxDraws<T> dr;
    dr.init(seed, chain, s.d);
    return kind == SHRINK ? run_chain<T, SHRINK>(s, b, dr, x0, n_steps, step0, samples, nrej)
                          : run_chain<T, REJECT>(s, b, dr, x0, n_steps, step0, samples, nrej);
}

}  // namespace

extern "C" {

void hs_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    U4 r = philox4x32_10(U4{ctr[0], ctr[1], ctr[2], ctr[3]}, key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

void hs_logp(const Spec* s, int dtype, const double* x, double* out, int64_t N) {
    if (dtype == F64) logp_all<double>(*s, x, out, N); else logp_all<float>(*s, x, out, N);
}

// one chain; tape != NULL replays recorded draws, else the Philox stream (seed, chain) is used
int hs_run(const Spec* s, int dtype, int kind, const double* tape, int64_t tape_len, uint64_t seed, uint64_t chain,
           const double* x0, int64_t n_steps, uint64_t step0, double* samples, int32_t* nrej) {
    return dtype == F64 ? run_any<double>(*s, kind, tape, tape_len, seed, chain, x0, n_steps, step0, samples, nrej)
                        : run_any<float>(*s, kind, tape, tape_len, seed, chain, x0, n_steps, step0, samples, nrej);
}

void hs_grad(const Spec* s, int dtype, const double* x, double* out, int64_t N) {
    if (dtype == F64) grad_all<double>(*s, x, out, N); else grad_all<float>(*s, x, out, N);
}

// one chain of a baseline sampler (kind 2..5); outputs per transition
int hs_run_mh(const Spec* s, int dtype, int kind, const double* tape, int64_t tape_len, uint64_t seed, uint64_t chain,
              const double* x0, int64_t n_steps, uint64_t step0, double stepsize, double alpha, int n_leap, int64_t adapt_total,
              double* samples, double* momenta, double* stepsizes, int32_t* accepted, int64_t* n_walk) {
    MhParams mp{alpha, n_leap, adapt_total};
    return dtype == F64 ? run_mh_any<double>(*s, kind, tape, tape_len, seed, chain, x0, n_steps, step0, stepsize, mp, samples, momenta,
                                             stepsizes, accepted, n_walk)
                        : run_mh_any<float>(*s, kind, tape, tape_len, seed, chain, x0, n_steps, step0, stepsize, mp, samples, momenta,
                                            stepsizes, accepted, n_walk);
}

// symmetric eigen-decomposition used for the dense Bingham tensor-core path (csrc/eigh_jacobi.h)
void hs_eigh(int n, const double* a, double* w, double* V) {
    std::vector<double> A(a, a + (size_t)n * n), ww, VV;
    eigh_jacobi(n, A, ww, VV);
    std::memcpy(w, ww.data(), sizeof(double) * n);
    std::memcpy(V, VV.data(), sizeof(double) * n * n);
}

int hs_spec_size(void) { return (int)sizeof(Spec); }
}
