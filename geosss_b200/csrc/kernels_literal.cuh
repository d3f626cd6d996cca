// Kernels of the literal path: batched log_prob and lane-per-chain / warp-per-chain samplers
// that follow the reference operation for operation, in float64 (parity) or float32.
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "targets.cuh"
#include "transition.cuh"

namespace geosss {

constexpr int LANE_BLOCK = 128;  // threads per block, lane-per-chain kernels
constexpr int CPD_WARPS = 4;     // warps (= chains or points) per block, CPD kernels

// ------------------------------------------------------------------ batched log_prob
// x: (N, d) row-major.  One lane per row; parameters are read through L1 as warp broadcasts.
template <typename T, class P>
__global__ void __launch_bounds__(LANE_BLOCK) k_logp_lane(P p, const T* __restrict__ x, T* __restrict__ out, int64_t N) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    SVec<const T> v{x + i * p.d, 1};
    out[i] = p.logp(v);
}

// Warp-cooperative CPD log-likelihood: every lane passes the same quaternion and receives the
// same value.  moved: per-warp shared scratch, 3*ns elements.
template <typename T> struct CpdWarpLogP {
    CpdP<T> p;
    T* moved;
    template <class Vec> __device__ __forceinline__ T operator()(Vec q) const {
        const int lane = threadIdx.x & 31;
        T R[9];
        quat_to_matrix<T>(q[0], q[1], q[2], q[3], R);
        __syncwarp();
        cpd_move_source(p, R, moved, lane, 32);
        __syncwarp();
        double acc = 0.0;
        for (int l = lane; l < p.nt; l += 32) acc += (double)cpd_row(p, moved, l);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        acc = __shfl_sync(0xffffffffu, acc, 0);
        return p.beta * (T)acc;
    }
};

template <typename T>
__global__ void __launch_bounds__(CPD_WARPS * 32) k_logp_cpd(CpdP<T> p, const T* __restrict__ q, T* __restrict__ out, int64_t N) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    T* moved = reinterpret_cast<T*>(smem_raw) + (size_t)warp * 3 * p.ns;
    int64_t i = (int64_t)blockIdx.x * CPD_WARPS + warp;
    if (i >= N) return;  // whole warp leaves together
    CpdWarpLogP<T> f{p, moved};
    SVec<const T> v{q + i * 4, 1};
    T val = f(v);
    if ((threadIdx.x & 31) == 0) out[i] = val;
}

// ------------------------------------------------------------------ sampler drivers
struct RunArgs {
    int64_t N;            // chains owned by this sampler
    int64_t n_steps;      // transitions to run
    int64_t thin;         // keep every thin-th state
    uint64_t seed;        // Philox key
    uint64_t chain0;      // global id of chain 0
    uint64_t step0;       // transitions already done
    const double* tape;   // tape mode: flat draws ...
    const int64_t* tape_off;  // ... and (N+1) offsets into it
    void* state;          // (N, d) row-major, in/out
    void* ws;             // (3, d, N) column-major scratch
    int ws_smem;          // 1: the lane kernels keep that scratch in shared memory instead ((planes, d, block) columns; set at launch when it fits)
    void* samples;        // (n_keep, N, d) or null
    void* logp;           // (n_keep, N) or null
    uint32_t* nrej;       // (n_keep, N) rejections of each kept transition (thin == 1) or null
    void* zbuf;           // great-circle kernels with on-the-fly scores: (ceil(d/4), N) float4 scratch holding the transition's normals
                          // between the prologue and the accepted move (null: they are regenerated from the Philox counter)
    const void* fdir;     // (d) direction v of the scalar test function f(x) = x . v, sampler dtype, or null
    void* fout;           // (n_keep, N) trace of f at the kept states (scalar layout like logp) or null
    unsigned long long* n_reject;  // (N) running totals
    int* status;          // (N) tape mode: 0 ok, 1 tape exhausted; may be null
    int quorum;           // fast path: lanes that must wait before a warp runs the prologue (32 = lock-step)
    // output layout in elements: kept state k of chain c goes to samples[k * ss_step + c * ss_chain + j],
    // scalars to [k * sl_step + c * sl_chain]  (step-major: (N d, d, N, 1); chain-major: (d, n_keep d, 1, n_keep))
    int64_t ss_step, ss_chain, sl_step, sl_chain;
    // baseline samplers (kernels_mh.cuh); n_reject then counts ACCEPTED proposals
    double* stepsize;              // (N) per-chain step size
    int64_t* adapt_count;          // (N) adaptation steps already taken
    unsigned long long* n_walk;    // (N) random-walk proposals of the mixture sampler
    void* momentum;                // (N, d) HMC momenta, in/out
    void* momenta_out;             // like samples, or null
    int64_t adapt_total;           // burn-in length of the step-size adaptation
    double alpha;                  // mixture: probability of a random-walk proposal
    int n_leap;                    // HMC leapfrog steps
};

// dynamic shared memory for the lane kernels' scratch: `planes` vectors of d elements per lane; 0 when that would leave fewer than
// two blocks per SM (the kernels then use the global scratch)
template <typename T> inline size_t lane_scratch_smem(int planes, int d) {
    const size_t bytes = (size_t)planes * d * LANE_BLOCK * sizeof(T);
    return bytes <= 100 * 1024 ? bytes : 0;
}

template <typename T, int KIND, bool TAPE> struct DrawsOf;
template <typename T, int KIND> struct DrawsOf<T, KIND, false> {
    using type = PhiloxDraws<T>;
    static __device__ __forceinline__ void init(type& dr, const RunArgs& a, int64_t c, int d) { dr.init(a.seed, a.chain0 + (uint64_t)c, d); }
};
template <typename T, int KIND> struct DrawsOf<T, KIND, true> {
    using type = TapeDraws<T>;
    static __device__ __forceinline__ void init(type& dr, const RunArgs& a, int64_t c, int) { dr.init(a.tape, a.tape_off[c], a.tape_off[c + 1]); }
};

template <class P> struct LaneLogP {
    P p;
    template <class Vec> __device__ __forceinline__ auto operator()(Vec v) const { return p.logp(v); }
};

// lane-per-chain; all n_steps transitions of a chain run inside one launch
template <typename T, int KIND, bool TAPE, class P>
__global__ void __launch_bounds__(LANE_BLOCK) k_sampler_lane(P p, RunArgs a) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.N) return;
    const int d = p.d;
    // x / u / y of a lane: column layout, in shared memory when the launch found room for it (the global (3, d, N) scratch otherwise)
    extern __shared__ __align__(16) unsigned char lane_smem_raw[];
    T* ws = a.ws_smem ? reinterpret_cast<T*>(lane_smem_raw) + threadIdx.x : static_cast<T*>(a.ws) + c;
    const int64_t stride = a.ws_smem ? (int64_t)blockDim.x : a.N, plane = (int64_t)d * stride;
    SVec<T> x{ws, stride}, u{ws + plane, stride}, y{ws + 2 * plane, stride};
    T* st = static_cast<T*>(a.state) + c * d;
    for (int j = 0; j < d; ++j) x[j] = st[j];

    using D = DrawsOf<T, KIND, TAPE>;
    typename D::type dr;
    D::init(dr, a, c, d);
    LaneLogP<P> f{p};
    unsigned long long rej_total = a.n_reject[c];
    unsigned int rej_window = 0;
    T* samples = static_cast<T*>(a.samples);
    T* lpo = static_cast<T*>(a.logp);
    bool good = true;
    int64_t since_keep = 0, kept = 0;
    for (int64_t t = 0; t < a.n_steps; ++t) {
        dr.begin_step(a.step0 + (uint64_t)t);
        T lp;
        int r = transition_literal<T, KIND>(d, x, u, y, dr, f, &lp);
        if (!dr.ok()) { good = false; break; }
        rej_total += (unsigned)r;
        rej_window += (unsigned)r;
        if (++since_keep == a.thin) {
            const int64_t k = kept++;
            since_keep = 0;
            if (samples) {
                T* o = samples + (k * a.ss_step + c * a.ss_chain);
                for (int j = 0; j < d; ++j) o[j] = x[j];
            }
            if (lpo) lpo[k * a.sl_step + c * a.sl_chain] = lp;
            if (a.nrej) a.nrej[k * a.sl_step + c * a.sl_chain] = rej_window;
            if (a.fout) {
                const T* v = static_cast<const T*>(a.fdir);
                T f_ = 0;
                for (int j = 0; j < d; ++j) f_ += x[j] * v[j];
                static_cast<T*>(a.fout)[k * a.sl_step + c * a.sl_chain] = f_;
            }
            rej_window = 0;
        }
    }
    for (int j = 0; j < d; ++j) st[j] = x[j];
    a.n_reject[c] = rej_total;
    if (a.status) a.status[c] = good ? 0 : 1;
}

// warp-per-chain CPD sampler: all lanes carry the same 4-vector state and the same draws, the
// log-likelihood is evaluated cooperatively (target rows over lanes).
template <typename T> struct Reg4 {
    T* v;
    __device__ __forceinline__ T& operator[](int j) const { return v[j]; }
};

template <typename T, int KIND, bool TAPE>
__global__ void __launch_bounds__(CPD_WARPS * 32) k_sampler_cpd(CpdP<T> p, RunArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* moved = reinterpret_cast<T*>(smem_raw) + (size_t)warp * 3 * p.ns;
    const int64_t c = (int64_t)blockIdx.x * CPD_WARPS + warp;
    if (c >= a.N) return;
    T xs[4], us[4], ys[4];
    Reg4<T> x{xs}, u{us}, y{ys};
    T* st = static_cast<T*>(a.state) + c * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) xs[j] = st[j];

    using D = DrawsOf<T, KIND, TAPE>;
    typename D::type dr;
    D::init(dr, a, c, 4);
    CpdWarpLogP<T> f{p, moved};
    unsigned long long rej_total = a.n_reject[c];
    unsigned int rej_window = 0;
    T* samples = static_cast<T*>(a.samples);
    T* lpo = static_cast<T*>(a.logp);
    bool good = true;
    int64_t since_keep = 0, kept = 0;
    for (int64_t t = 0; t < a.n_steps; ++t) {
        dr.begin_step(a.step0 + (uint64_t)t);
        T lp;
        int r = transition_literal<T, KIND>(4, x, u, y, dr, f, &lp);
        if (!dr.ok()) { good = false; break; }
        rej_total += (unsigned)r;
        rej_window += (unsigned)r;
        if (++since_keep == a.thin) {
            const int64_t k = kept++;
            since_keep = 0;
            if (lane == 0) {
                if (samples) {
                    T* o = samples + (k * a.ss_step + c * a.ss_chain);
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = xs[j];
                }
                if (lpo) lpo[k * a.sl_step + c * a.sl_chain] = lp;
                if (a.nrej) a.nrej[k * a.sl_step + c * a.sl_chain] = rej_window;
            }
            rej_window = 0;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) st[j] = xs[j];
        a.n_reject[c] = rej_total;
        if (a.status) a.status[c] = good ? 0 : 1;
    }
}

// Float32 throughput variant (dense sum with outlier column only): one pass over the pairs with the
// fixed shift M0 = max(log_outlier, log_const + max_k log w_k) >= every term, so nothing can overflow and
// the outlier term keeps the row sum away from zero; exp on the MUFU pipe in base 2 with all constants
// folded into one FFMA per pair:  2^(lw2_k + c2 + d2 * q2 - M2).
struct CpdWarpLogPFast {
    CpdP<float> p;
    float4* mv;                // shared (ns): rotated source point (x, y, z) and w = (log w_k + log_const - M0) log2(e)
    float q2, out2m, m0;       // -0.5 log2(e) / sigma^2, (log_outlier - M0) log2(e), M0
    template <class Vec> __device__ __forceinline__ float operator()(Vec q) const {
        const int lane = threadIdx.x & 31;
        float R[9];
        quat_to_matrix<float>(q[0], q[1], q[2], q[3], R);
        __syncwarp();
        for (int k = lane; k < p.ns; k += 32) {  // moved_k = R s_k (pointcloud.py:258-262); w stays
            const float s0 = p.src[3 * k], s1 = p.src[3 * k + 1], s2 = p.src[3 * k + 2];
            float* m = reinterpret_cast<float*>(mv + k);
            m[0] = (s0 * R[0] + s1 * R[1] + s2 * R[2]) + p.shift[0];
            m[1] = (s0 * R[3] + s1 * R[4] + s2 * R[5]) + p.shift[1];
            m[2] = p.dim == 2 ? 0.0f : (s0 * R[6] + s1 * R[7] + s2 * R[8]) + p.shift[2];
        }
        __syncwarp();
        double acc = 0.0;
        const float base = exp2f(out2m);
        // two target rows per lane and iteration: one 16-byte broadcast load of a source point feeds both.
        // (The expanded form |t|^2 + |m|^2 - 2 t.m needs 3 fewer FP ops per pair but loses 3e-3 abs on
        // log p ~ -2400 to cancellation -- measured -- so differences are formed directly.)
        for (int l = lane; l < p.nt; l += 64) {
            const int lb = l + 32;
            const bool has_b = lb < p.nt;
            const int lbb = has_b ? lb : l;
            const float a0 = p.tgt[3 * l], a1 = p.tgt[3 * l + 1], a2 = p.tgt[3 * l + 2];
            const float b0 = p.tgt[3 * lbb], b1 = p.tgt[3 * lbb + 1], b2 = p.tgt[3 * lbb + 2];
            float sa = base, sb = base;
#pragma unroll 2
            for (int k = 0; k < p.ns; ++k) {
                const float4 m = mv[k];
                const float e0 = a0 - m.x, e1 = a1 - m.y, e2 = a2 - m.z;
                const float f0 = b0 - m.x, f1 = b1 - m.y, f2 = b2 - m.z;
                sa += fast_ex2(fmaf(fmaf(e0, e0, fmaf(e1, e1, e2 * e2)), q2, m.w));
                sb += fast_ex2(fmaf(fmaf(f0, f0, fmaf(f1, f1, f2 * f2)), q2, m.w));
            }
            acc += (double)((m0 + 0.6931471805599453f * fast_lg2(sa)) * p.wt[l]);
            if (has_b) acc += (double)((m0 + 0.6931471805599453f * fast_lg2(sb)) * p.wt[lb]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        acc = __shfl_sync(0xffffffffu, acc, 0);
        return p.beta * (float)acc;
    }
    // per-warp set-up; smem per warp: ns float4
    __device__ __forceinline__ void init(const CpdP<float>& p_, float* warp_smem, float max_logw) {
        p = p_; mv = reinterpret_cast<float4*>(warp_smem);
        constexpr float L2E = 1.4426950408889634f;
        const int lane = threadIdx.x & 31;
        m0 = fmaxf(p.log_outlier, p.log_const + max_logw);
        q2 = -0.5f * L2E / p.sigma2;
        out2m = (p.log_outlier - m0) * L2E;
        for (int k = lane; k < p.ns; k += 32) mv[k] = make_float4(0.f, 0.f, 0.f, (p.logws[k] + p.log_const - m0) * L2E);
        __syncwarp();
    }
};

template <int KIND>
__global__ void __launch_bounds__(CPD_WARPS * 32) k_sampler_cpd_fast(CpdP<float> p, float max_logw, RunArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * CPD_WARPS + warp;
    if (c >= a.N) return;
    CpdWarpLogPFast f;
    f.init(p, reinterpret_cast<float*>(smem_raw) + (size_t)warp * 4 * p.ns, max_logw);
    float xs[4], us[4], ys[4];
    Reg4<float> x{xs}, u{us}, y{ys};
    float* st = static_cast<float*>(a.state) + c * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) xs[j] = st[j];
    PhiloxDraws<float> dr;
    dr.init(a.seed, a.chain0 + (uint64_t)c, 4);
    unsigned long long rej_total = a.n_reject[c];
    unsigned int rej_window = 0;
    float* samples = static_cast<float*>(a.samples);
    float* lpo = static_cast<float*>(a.logp);
    int64_t since_keep = 0, kept = 0;
    for (int64_t t = 0; t < a.n_steps; ++t) {
        dr.begin_step(a.step0 + (uint64_t)t);
        float lp;
        int r = transition_literal<float, KIND>(4, x, u, y, dr, f, &lp);
        rej_total += (unsigned)r;
        rej_window += (unsigned)r;
        if (++since_keep == a.thin) {
            const int64_t k = kept++;
            since_keep = 0;
            if (lane == 0) {
                if (samples) {
                    float* o = samples + (k * a.ss_step + c * a.ss_chain);
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = xs[j];
                }
                if (lpo) lpo[k * a.sl_step + c * a.sl_chain] = lp;
                if (a.nrej) a.nrej[k * a.sl_step + c * a.sl_chain] = rej_window;
            }
            rej_window = 0;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) st[j] = xs[j];
        a.n_reject[c] = rej_total;
    }
}

static __global__ void __launch_bounds__(CPD_WARPS * 32) k_logp_cpd_fast(CpdP<float> p, float max_logw, const float* __restrict__ q,
                                                                  float* __restrict__ out, int64_t N) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    int64_t i = (int64_t)blockIdx.x * CPD_WARPS + warp;
    if (i >= N) return;
    CpdWarpLogPFast f;
    f.init(p, reinterpret_cast<float*>(smem_raw) + (size_t)warp * 4 * p.ns, max_logw);
    SVec<const float> v{q + i * 4, 1};
    float val = f(v);
    if ((threadIdx.x & 31) == 0) out[i] = val;
}

}  // namespace geosss
