// Baseline samplers of the reference's four-way comparison, batched: random-walk Metropolis-Hastings
// with burn-in step-size adaptation, the independence sampler, their mixture, and spherical HMC.
//
//   geosss/mcmc.py:78-112   AdaptiveStepsize          (x 1.02 / x 0.98 during the first `burnin` transitions)
//   geosss/mcmc.py:115-173  MetropolisHastings        (radius sqrt(2 Gamma(d/2)), perturb, project back)
//   geosss/mcmc.py:176-179  IndependenceSampler
//   geosss/mcmc.py:182-234  MixtureRWMHIndependenceSampler
//   geosss/mcmc.py:237-332  SphericalHMC              (tangent momenta, geodesic leapfrog)
//
// Literal path: one lane per chain (one warp per chain for the registration target), reference
// operation order, float64 for draw-tape parity or float32.  log p of the current state is carried
// instead of being re-evaluated (same value, the targets are pure functions).
#pragma once
#include "kernels_literal.cuh"
#include "transition_mh.cuh"

namespace geosss {

template <typename T, bool TAPE> struct SeqOf;
template <typename T> struct SeqOf<T, false> {
    using type = PhiloxSeq<T>;
    static __device__ __forceinline__ void init(type& dr, const RunArgs& a, int64_t c) { dr.init(a.seed, a.chain0 + (uint64_t)c); }
};
template <typename T> struct SeqOf<T, true> {
    using type = TapeSeq<T>;
    static __device__ __forceinline__ void init(type& dr, const RunArgs& a, int64_t c) { dr.t.init(a.tape, a.tape_off[c], a.tape_off[c + 1]); }
};

template <class P> struct LaneTarget {
    P p;
    template <class Vec> __device__ __forceinline__ auto logp(Vec v) const { return p.logp(v); }
    template <class Vec, class Out> __device__ __forceinline__ void grad(Vec v, Out g) const { p.grad(v, g); }
};

// lane-per-chain; ws: (5, d, N) column-major scratch (x, v, xp, vp, g)
template <typename T, bool TAPE, class P>
__global__ void __launch_bounds__(LANE_BLOCK) k_mh_lane(P p, int kind, RunArgs a) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.N) return;
    const int d = p.d;
    extern __shared__ __align__(16) unsigned char lane_smem_raw[];  // the five vectors of a lane, when the launch found room (RunArgs::ws_smem)
    T* ws = a.ws_smem ? reinterpret_cast<T*>(lane_smem_raw) + threadIdx.x : static_cast<T*>(a.ws) + c;
    const int64_t stride = a.ws_smem ? (int64_t)blockDim.x : a.N, plane = (int64_t)d * stride;
    SVec<T> x{ws, stride}, v{ws + plane, stride}, xp{ws + 2 * plane, stride}, vp{ws + 3 * plane, stride}, g{ws + 4 * plane, stride};
    T* st = static_cast<T*>(a.state) + c * d;
    T* mom = static_cast<T*>(a.momentum) + c * d;
    for (int j = 0; j < d; ++j) { x[j] = st[j]; v[j] = mom[j]; }

    using D = SeqOf<T, TAPE>;
    typename D::type dr;
    D::init(dr, a, c);
    LaneTarget<P> f{p};
    MhChain ch{a.stepsize[c], a.adapt_count[c], a.n_walk[c]};
    const MhParams mp{a.alpha, a.n_leap, a.adapt_total};
    unsigned long long acc_total = a.n_reject[c];
    unsigned int acc_window = 0;
    T lp = f.logp(x);
    T* samples = static_cast<T*>(a.samples);
    T* momenta = static_cast<T*>(a.momenta_out);
    T* lpo = static_cast<T*>(a.logp);
    bool good = true;
    int64_t since_keep = 0, kept = 0;
    for (int64_t t = 0; t < a.n_steps; ++t) {
        dr.begin_step(a.step0 + (uint64_t)t);
        const bool ok = mh_transition<T>(kind, d, x, v, xp, vp, g, dr, f, ch, mp, lp);
        if (!dr.ok()) { good = false; break; }
        acc_total += ok ? 1u : 0u;
        acc_window += ok ? 1u : 0u;
        if (++since_keep == a.thin) {
            const int64_t k = kept++;
            since_keep = 0;
            if (samples) {
                T* o = samples + (k * a.ss_step + c * a.ss_chain);
                for (int j = 0; j < d; ++j) o[j] = x[j];
            }
            if (momenta) {
                T* o = momenta + (k * a.ss_step + c * a.ss_chain);
                for (int j = 0; j < d; ++j) o[j] = v[j];
            }
            if (lpo) lpo[k * a.sl_step + c * a.sl_chain] = lp;
            if (a.nrej) a.nrej[k * a.sl_step + c * a.sl_chain] = acc_window;
            if (a.fout) {
                const T* dir = static_cast<const T*>(a.fdir);
                T f_ = 0;
                for (int j = 0; j < d; ++j) f_ += x[j] * dir[j];
                static_cast<T*>(a.fout)[k * a.sl_step + c * a.sl_chain] = f_;
            }
            acc_window = 0;
        }
    }
    for (int j = 0; j < d; ++j) { st[j] = x[j]; mom[j] = v[j]; }
    a.n_reject[c] = acc_total;
    a.stepsize[c] = ch.stepsize; a.adapt_count[c] = ch.adapt_count; a.n_walk[c] = ch.n_walk;
    if (a.status) a.status[c] = good ? 0 : 1;
}

// Warp-cooperative registration target: value and gradient of one quaternion, identical on all lanes.
template <typename T> struct CpdWarpTarget {
    CpdP<T> p;
    T* moved;
    template <class Vec> __device__ __forceinline__ T logp(Vec q) const { return CpdWarpLogP<T>{p, moved}(q); }
    template <class Vec, class Out> __device__ __forceinline__ void grad(Vec q, Out g) const {
        const int lane = threadIdx.x & 31;
        T R[9];
        quat_to_matrix<T>(q[0], q[1], q[2], q[3], R);
        __syncwarp();
        cpd_move_source(p, R, moved, lane, 32);
        __syncwarp();
        T G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int l = lane; l < p.nt; l += 32) cpd_row_grad(p, moved, l, G);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            double s = (double)G[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            G[i] = (T)__shfl_sync(0xffffffffu, s, 0);
        }
        cpd_grad_finish(p, q[0], q[1], q[2], q[3], G, g);
    }
};

template <typename T, bool TAPE>
__global__ void __launch_bounds__(CPD_WARPS * 32) k_mh_cpd(CpdP<T> p, int kind, RunArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* moved = reinterpret_cast<T*>(smem_raw) + (size_t)warp * 3 * p.ns;
    const int64_t c = (int64_t)blockIdx.x * CPD_WARPS + warp;
    if (c >= a.N) return;
    T xs[4], vs[4], xps[4], vps[4], gs[4];
    Reg4<T> x{xs}, v{vs}, xp{xps}, vp{vps}, g{gs};
    T* st = static_cast<T*>(a.state) + c * 4;
    T* mom = static_cast<T*>(a.momentum) + c * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) { xs[j] = st[j]; vs[j] = mom[j]; }

    using D = SeqOf<T, TAPE>;
    typename D::type dr;
    D::init(dr, a, c);
    CpdWarpTarget<T> f{p, moved};
    MhChain ch{a.stepsize[c], a.adapt_count[c], a.n_walk[c]};
    const MhParams mp{a.alpha, a.n_leap, a.adapt_total};
    unsigned long long acc_total = a.n_reject[c];
    unsigned int acc_window = 0;
    T lp = f.logp(x);
    T* samples = static_cast<T*>(a.samples);
    T* momenta = static_cast<T*>(a.momenta_out);
    T* lpo = static_cast<T*>(a.logp);
    bool good = true;
    int64_t since_keep = 0, kept = 0;
    for (int64_t t = 0; t < a.n_steps; ++t) {
        dr.begin_step(a.step0 + (uint64_t)t);
        const bool ok = mh_transition<T>(kind, 4, x, v, xp, vp, g, dr, f, ch, mp, lp);
        if (!dr.ok()) { good = false; break; }
        acc_total += ok ? 1u : 0u;
        acc_window += ok ? 1u : 0u;
        if (++since_keep == a.thin) {
            const int64_t k = kept++;
            since_keep = 0;
            if (lane == 0) {
                if (samples) {
                    T* o = samples + (k * a.ss_step + c * a.ss_chain);
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = xs[j];
                }
                if (momenta) {
                    T* o = momenta + (k * a.ss_step + c * a.ss_chain);
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = vs[j];
                }
                if (lpo) lpo[k * a.sl_step + c * a.sl_chain] = lp;
                if (a.nrej) a.nrej[k * a.sl_step + c * a.sl_chain] = acc_window;
            }
            acc_window = 0;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { st[j] = xs[j]; mom[j] = vs[j]; }
        a.n_reject[c] = acc_total;
        a.stepsize[c] = ch.stepsize; a.adapt_count[c] = ch.adapt_count; a.n_walk[c] = ch.n_walk;
        if (a.status) a.status[c] = good ? 0 : 1;
    }
}

// ------------------------------------------------------------------ batched gradient
template <typename T, class P>
__global__ void __launch_bounds__(LANE_BLOCK) k_grad_lane(P p, const T* __restrict__ x, T* __restrict__ out, int64_t N) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    SVec<const T> v{x + i * p.d, 1};
    SVec<T> g{out + i * p.d, 1};
    p.grad(v, g);
}

template <typename T>
__global__ void __launch_bounds__(CPD_WARPS * 32) k_grad_cpd(CpdP<T> p, const T* __restrict__ q, T* __restrict__ out, int64_t N) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    T* moved = reinterpret_cast<T*>(smem_raw) + (size_t)warp * 3 * p.ns;
    int64_t i = (int64_t)blockIdx.x * CPD_WARPS + warp;
    if (i >= N) return;
    CpdWarpTarget<T> f{p, moved};
    SVec<const T> v{q + i * 4, 1};
    T gs[4];
    Reg4<T> g{gs};
    f.grad(v, g);
    if ((threadIdx.x & 31) == 0)
        for (int j = 0; j < 4; ++j) out[i * 4 + j] = gs[j];
}

// Host-side launchers, defined (and explicitly instantiated) in mh.cu.  Return a cudaError_t value, or -1
// when the registration cloud does not fit shared memory.
template <typename T, class P> int launch_mh(const P& p, int kind, bool tape, const RunArgs& a, cudaStream_t st);
template <typename T, class P> int launch_grad(const P& p, const T* x, T* out, int64_t N, cudaStream_t st);

}  // namespace geosss
