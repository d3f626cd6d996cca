// Float32 throughput kernels for targets whose log-density along a great circle has a cheap
// closed form: Bingham(-Fisher) and the curved vMF.  Same persistent, quorum-scheduled bracket
// loop as k_vmf_fast (kernels_fast.cuh), but for runtime d up to a few hundred: the chain state x
// and the direction u live in shared memory in column layout (element j of lane t at
// [j * BLOCK + t], conflict-free), target parameters are staged into shared memory once per block.
//
// With y(theta) = c x + s u (x, u orthonormal):
//   Bingham   geosss/distributions.py:78-86,113-114
//             y^T A y - x^T A x = s (s (Quu - Qxx) + 2 c Qxu),   b.y - b.x = (c - 1) bx + s bu
//             with Qxx = x^T A x, Qxu = x^T A u, Quu = u^T A u: O(1) per proposal, and the slice test
//             "log p(y) - log p(x) > log U" is cancellation-free.
//   curved vMF distributions.py:272-275, spherical_curve.py:10-32,95-102
//             knots.y = c (knots.x) + s (knots.u); per segment (a, b) with angle ang the nearest
//             point on the arc has x.y = hypot(p1, p2) when its angle t* = atan2(p2, p1) lies in
//             [0, ang] (p1 = G_a, p2 = (G_b - G_a cos ang) / sin ang), else G_a (t* < 0) or G_b: the
//             clip(atan2) / sin / arccos chain of the reference reduces to two sign tests and a sqrt.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "kernels_fast.cuh"
#include "kernels_literal.cuh"
#include "philox.cuh"

namespace geosss {

constexpr int GC_BLOCK = 64;        // chains per block with stored normals (d < 64, dense Bingham)
constexpr int GC_BLOCK_WIDE = 128;  // on-the-fly scores (d >= 64): the staged parameters are amortised over more chains (C3: +6 %)
constexpr int gc_block(bool store) { return store ? GC_BLOCK : GC_BLOCK_WIDE; }
constexpr int GC_MAX_KNOTS = 16;

struct GcParams {
    const float* a0;  // Bingham: A (d,d) or its diagonal (d);  curved: knots (M,d)
    const float* a1;  // Bingham: b (d) or null;                 curved: cos(ang) (M-1)
    const float* a2;  //                                         curved: sin(ang) (M-1)
    float kappa;
    int d, M, diag;
};

// ------------------------------------------------------------------ policies
struct BinghamGc {
    struct Coef { float qxx, qxu, quu, bx, bu; };
    __device__ __forceinline__ bool needs_vector() const { return !diag; }  // dense A: finish() reads the stored normals
    const float* A;  // shared: diagonal (d) or dense (d, dp) with rows padded to dp = 4 ceil(d/4) floats
    const float* b;  // shared or null
    int d, dp, diag;

    static __host__ __device__ size_t smem_floats(const GcParams& p) {
        return (size_t)(p.diag ? p.d : p.d * ((p.d + 3) & ~3)) + (p.a1 ? p.d : 0) + 4;
    }
    __device__ void stage(const GcParams& p, float* sm) {
        d = p.d; diag = p.diag; dp = (d + 3) & ~3;
        const int nA = diag ? dp : d * dp;
        if (diag) { for (int i = threadIdx.x; i < dp; i += blockDim.x) sm[i] = i < d ? p.a0[i] : 0.f; }
        else for (int i = threadIdx.x; i < nA; i += blockDim.x) { const int r = i / dp, q = i - r * dp; sm[i] = q < d ? p.a0[r * d + q] : 0.f; }
        if (p.a1) for (int i = threadIdx.x; i < d; i += blockDim.x) sm[nA + i] = p.a1[i];
        A = sm;
        b = p.a1 ? sm + nA : nullptr;
    }
    __device__ __forceinline__ void begin(Coef& c) const { c.qxx = c.qxu = c.quu = c.bx = c.bu = 0.f; }
    __device__ __forceinline__ void accumulate(Coef& c, int j, float xj, float uj) const {
        if (diag) {
            const float l = A[j];
            c.qxx = fmaf(l * xj, xj, c.qxx); c.qxu = fmaf(l * xj, uj, c.qxu); c.quu = fmaf(l * uj, uj, c.quu);
        }
        if (b) { c.bx = fmaf(b[j], xj, c.bx); c.bu = fmaf(b[j], uj, c.bu); }
    }
    // dense A: (A x)_i and (A u)_i for four rows i at a time.  A is symmetric, so A[i..i+3][j] is read as
    // the contiguous A[j][i..i+3] (one 16-byte broadcast load); per j that is 3 loads for 8 FMAs.
    template <class V> __device__ __forceinline__ void finish(Coef& c, V xs, V us) const {
        if (diag) return;
        for (int i = 0; i < dp; i += 4) {
            float ax[4] = {0.f, 0.f, 0.f, 0.f}, au[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = 0; j < d; ++j) {
                const float4 a4 = *reinterpret_cast<const float4*>(A + j * dp + i);
                const float xj = xs[j], uj = us[j];
                ax[0] = fmaf(a4.x, xj, ax[0]); au[0] = fmaf(a4.x, uj, au[0]);
                ax[1] = fmaf(a4.y, xj, ax[1]); au[1] = fmaf(a4.y, uj, au[1]);
                ax[2] = fmaf(a4.z, xj, ax[2]); au[2] = fmaf(a4.z, uj, au[2]);
                ax[3] = fmaf(a4.w, xj, ax[3]); au[3] = fmaf(a4.w, uj, au[3]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) if (i + q < d) {
                const float xi = xs[i + q], ui = us[i + q];
                c.qxx = fmaf(xi, ax[q], c.qxx); c.qxu = fmaf(xi, au[q], c.qxu); c.quu = fmaf(ui, au[q], c.quu);
            }
        }
    }
    // coefficients were accumulated against the raw normals z; the direction is u = inp (z - zx x)
    __device__ __forceinline__ void fixup(Coef& c, float zx, float inp) const {
        c.quu = inp * inp * fmaf(zx, fmaf(zx, c.qxx, -2.f * c.qxu), c.quu);
        c.qxu = inp * fmaf(-zx, c.qxx, c.qxu);
        c.bu = inp * fmaf(-zx, c.bx, c.bu);
    }
    // another lane's coefficients (cooperative reject batches)
    __device__ __forceinline__ Coef fetch(const Coef& c, int src) const {
        Coef o;
        o.qxx = __shfl_sync(0xffffffffu, c.qxx, src); o.qxu = __shfl_sync(0xffffffffu, c.qxu, src);
        o.quu = __shfl_sync(0xffffffffu, c.quu, src); o.bx = __shfl_sync(0xffffffffu, c.bx, src);
        o.bu = __shfl_sync(0xffffffffu, c.bu, src);
        return o;
    }
    __device__ __forceinline__ float delta(const Coef& c, float cs, float sn) const {
        float dlt = sn * fmaf(sn, c.quu - c.qxx, 2.f * cs * c.qxu);
        if (b) dlt += fmaf(cs - 1.f, c.bx, sn * c.bu);
        return dlt;
    }
    __device__ __forceinline__ float logp_x(const Coef& c) const { return c.qxx + c.bx; }
};

// QUADS = knot quads held per chain (scores in registers: 8 QUADS floats).  EXACT: the target has exactly QUADS quads, so the
// quad loops carry no run-time guards (the sampler kernels are instantiated per quad count: less code, no dead quad);
// CurvedGc = <4, false> serves any M <= 16 with guards (standalone log_prob).
template <int QUADS, bool EXACT> struct CurvedGcT {
    static constexpr int MAXK = 4 * QUADS;
    struct Coef { float gx[MAXK], gu[MAXK], best0; };
    __device__ __forceinline__ bool needs_vector() const { return false; }
    const float* knotsT;  // shared (d, MP): knotsT[j * MP + m], MP = 4 ceil(M/4), zero for m >= M (float4 loads)
    const float* ca;      // shared (M-1)
    const float* isa;     // shared (M-1) 1 / (sin(ang) + 1e-10)
    const float* sa;      // shared (M-1)
    float kappa;
    int d, M;

    static __host__ __device__ size_t smem_floats(const GcParams& p) { return (size_t)((p.M + 3) & ~3) * p.d + 3 * (size_t)(p.M - 1) + 4; }
    __device__ void stage(const GcParams& p, float* sm) {
        d = p.d; M = p.M; kappa = p.kappa;
        const int MP = (M + 3) & ~3;
        for (int i = threadIdx.x; i < MP * d; i += blockDim.x) {
            const int j = i / MP, m = i - j * MP;
            sm[i] = m < M ? p.a0[m * d + j] : 0.f;
        }
        float* c_ = sm + MP * d;
        float* is_ = c_ + (M - 1);
        float* s_ = is_ + (M - 1);
        for (int i = threadIdx.x; i < M - 1; i += blockDim.x) { c_[i] = p.a1[i]; s_[i] = p.a2[i]; is_[i] = 1.f / (p.a2[i] + 1e-10f); }
        knotsT = sm; ca = c_; isa = is_; sa = s_;
    }
    __device__ __forceinline__ void begin(Coef& c) const {
#pragma unroll
        for (int m = 0; m < MAXK; ++m) { c.gx[m] = 0.f; c.gu[m] = 0.f; }
    }
    __device__ __forceinline__ void accumulate(Coef& c, int j, float xj, float uj) const {
        const int nq = EXACT ? QUADS : (M + 3) >> 2;  // warp-uniform
        const float4* col = reinterpret_cast<const float4*>(knotsT + j * (nq << 2));
#pragma unroll
        for (int q = 0; q < QUADS; ++q) if (EXACT || q < nq) {
            const float4 k = col[q];
            c.gx[4 * q] = fmaf(k.x, xj, c.gx[4 * q]);         c.gu[4 * q] = fmaf(k.x, uj, c.gu[4 * q]);
            c.gx[4 * q + 1] = fmaf(k.y, xj, c.gx[4 * q + 1]); c.gu[4 * q + 1] = fmaf(k.y, uj, c.gu[4 * q + 1]);
            c.gx[4 * q + 2] = fmaf(k.z, xj, c.gx[4 * q + 2]); c.gu[4 * q + 2] = fmaf(k.z, uj, c.gu[4 * q + 2]);
            c.gx[4 * q + 3] = fmaf(k.w, xj, c.gx[4 * q + 3]); c.gu[4 * q + 3] = fmaf(k.w, uj, c.gu[4 * q + 3]);
        }
    }
    // knots . x only (standalone log_prob: no direction), knot quads as a compile-time count: the caller selects the
    // instantiation once per row (warp-uniform), so the padded quads cost nothing (predicated-off FMAs still issue)
    template <int NQ> __device__ __forceinline__ void accumulate_x_n(Coef& c, int j, float xj) const {
        const float4* col = reinterpret_cast<const float4*>(knotsT + j * (NQ << 2));
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float4 k = col[q];
            c.gx[4 * q] = fmaf(k.x, xj, c.gx[4 * q]);         c.gx[4 * q + 1] = fmaf(k.y, xj, c.gx[4 * q + 1]);
            c.gx[4 * q + 2] = fmaf(k.z, xj, c.gx[4 * q + 2]); c.gx[4 * q + 3] = fmaf(k.w, xj, c.gx[4 * q + 3]);
        }
    }
    __device__ __forceinline__ void accumulate_x(Coef& c, int j, float xj) const {
        const int nq = EXACT ? QUADS : (M + 3) >> 2;
        const float4* col = reinterpret_cast<const float4*>(knotsT + j * (nq << 2));
#pragma unroll
        for (int q = 0; q < QUADS; ++q) if (EXACT || q < nq) {
            const float4 k = col[q];
            c.gx[4 * q] = fmaf(k.x, xj, c.gx[4 * q]);         c.gx[4 * q + 1] = fmaf(k.y, xj, c.gx[4 * q + 1]);
            c.gx[4 * q + 2] = fmaf(k.z, xj, c.gx[4 * q + 2]); c.gx[4 * q + 3] = fmaf(k.w, xj, c.gx[4 * q + 3]);
        }
    }
    // coefficients were accumulated against the raw normals z; the direction is u = inp (z - zx x)
    __device__ __forceinline__ void fixup(Coef& c, float zx, float inp) const {
#pragma unroll
        for (int m = 0; m < MAXK; ++m) c.gu[m] = inp * fmaf(-zx, c.gx[m], c.gu[m]);
    }
    // another lane's coefficients (cooperative reject batches)
    __device__ __forceinline__ Coef fetch(const Coef& c, int src) const {
        Coef o;
        const int nq = EXACT ? QUADS : (M + 3) >> 2;
#pragma unroll
        for (int m = 0; m < MAXK; ++m) {
            if (EXACT || (m >> 2) < nq) { o.gx[m] = __shfl_sync(0xffffffffu, c.gx[m], src); o.gu[m] = __shfl_sync(0xffffffffu, c.gu[m], src); }
            else { o.gx[m] = 0.f; o.gu[m] = 0.f; }
        }
        o.best0 = __shfl_sync(0xffffffffu, c.best0, src);
        return o;
    }
    // max over segments of x(theta) . nearest-point, first segment wins ties (argmin of arccos)
    __device__ __forceinline__ float best_dot(const Coef& c, float cs, float sn) const {
        float ga = fmaf(cs, c.gx[0], sn * c.gu[0]);
        float best = 0.f, best_clipped = -INFINITY;
#pragma unroll
        for (int m = 1; m < MAXK; ++m) if (m < M) {
            const float gb = fmaf(cs, c.gx[m], sn * c.gu[m]);
            const float p2 = (gb - ga * ca[m - 1]) * isa[m - 1];
            const bool inside = p2 >= 0.f && fmaf(ga, sa[m - 1], -p2 * ca[m - 1]) >= 0.f;
            // branch-free: the square root is one MUFU instruction; as a branch the compiler ran it at 7-10 active lanes per segment
            // (profiles/r2_ncu_full_k_gc_fast_c3_2e20.txt; C3 1.76 -> 1.85e9 samples/s)
            const float hyp = fast_sqrt(fmaf(ga, ga, p2 * p2));
            const float end = p2 < 0.f ? ga : gb;
            const float dot = inside ? hyp : end;
            // the reference ranks segments by arccos(clip(dot, -1, 1)) (spherical_curve.py:32,102): dots above
            // 1 (|x| > 1) tie at distance 0 and the first such segment wins, but its unclipped dot is returned
            const float dc = fminf(dot, 1.f);
            if (dc > best_clipped) { best_clipped = dc; best = dot; }
            ga = gb;
        }
        return best;
    }
    template <class V> __device__ __forceinline__ void finish(Coef& c, V, V) const { c.best0 = best_dot(c, 1.f, 0.f); }
    __device__ __forceinline__ float delta(const Coef& c, float cs, float sn) const { return kappa * (best_dot(c, cs, sn) - c.best0); }
    __device__ __forceinline__ float logp_x(const Coef& c) const { return kappa * c.best0; }
};
using CurvedGc = CurvedGcT<GC_MAX_KNOTS / 4, false>;


// vMF mixture of any dimension that fits shared memory and up to GC_MAX_K components (the register-
// resident k_vmf_fast covers d <= 16, K <= 8): scores mu_k . y = c A_k + s B_k as in kernels_fast.cuh,
// log p(y) - log p(x) through a max-shifted base-2 log-sum-exp.
constexpr int GC_MAX_K = 16;
// QUADS / EXACT as for CurvedGcT: the sampler kernels are instantiated per component-quad count (2 QUADS x 4 scores in
// registers instead of 32, no guarded quads).
template <int QUADS, bool EXACT> struct VmfGcT {
    static constexpr int MAXK = 4 * QUADS;
    struct Coef { float A[MAXK], B[MAXK], lp2; };
    __device__ __forceinline__ bool needs_vector() const { return false; }
    const float* muT;  // shared (d, KP): muT[j * KP + k] * log2(e), KP = 4 ceil(K/4), zero padded
    const float* c2;   // shared (KP): constants * log2(e), -inf for the padding
    int d, K;

    static __host__ __device__ size_t smem_floats(const GcParams& p) { return (size_t)((p.M + 3) & ~3) * (p.d + 1) + 4; }
    __device__ void stage(const GcParams& p, float* sm) {  // p.M carries K, a0 = mu (K,d), a1 = consts (K)
        d = p.d; K = p.M;
        const int KP = (K + 3) & ~3;
        constexpr float L2E = 1.4426950408889634f;
        for (int i = threadIdx.x; i < KP * d; i += blockDim.x) {
            const int j = i / KP, k = i - j * KP;
            sm[i] = k < K ? p.a0[k * d + j] * L2E : 0.f;
        }
        float* c_ = sm + KP * d;
        for (int i = threadIdx.x; i < KP; i += blockDim.x) c_[i] = i < K ? p.a1[i] * L2E : -INFINITY;
        muT = sm; c2 = c_;
    }
    __device__ __forceinline__ void begin(Coef& c) const {
#pragma unroll
        for (int k = 0; k < MAXK; ++k) { c.A[k] = 0.f; c.B[k] = 0.f; }
    }
    __device__ __forceinline__ void accumulate(Coef& c, int j, float xj, float uj) const {
        const int nq = EXACT ? QUADS : (K + 3) >> 2;
        const float4* col = reinterpret_cast<const float4*>(muT + j * (nq << 2));
#pragma unroll
        for (int q = 0; q < QUADS; ++q) if (EXACT || q < nq) {
            const float4 m = col[q];
            c.A[4 * q] = fmaf(m.x, xj, c.A[4 * q]);         c.B[4 * q] = fmaf(m.x, uj, c.B[4 * q]);
            c.A[4 * q + 1] = fmaf(m.y, xj, c.A[4 * q + 1]); c.B[4 * q + 1] = fmaf(m.y, uj, c.B[4 * q + 1]);
            c.A[4 * q + 2] = fmaf(m.z, xj, c.A[4 * q + 2]); c.B[4 * q + 2] = fmaf(m.z, uj, c.B[4 * q + 2]);
            c.A[4 * q + 3] = fmaf(m.w, xj, c.A[4 * q + 3]); c.B[4 * q + 3] = fmaf(m.w, uj, c.B[4 * q + 3]);
        }
    }
    __device__ __forceinline__ void fixup(Coef& c, float zx, float inp) const {
#pragma unroll
        for (int k = 0; k < MAXK; ++k) c.B[k] = inp * fmaf(-zx, c.A[k], c.B[k]);
    }
    __device__ __forceinline__ float lse2(const Coef& c, float cs, float sn) const {
        const int KP = (K + 3) & ~3;
        float e[MAXK], m = -INFINITY;
#pragma unroll
        for (int k = 0; k < MAXK; ++k) if (EXACT || k < KP) { e[k] = fmaf(cs, c.A[k], fmaf(sn, c.B[k], c2[k])); m = fmaxf(m, e[k]); }
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < MAXK; ++k) if (EXACT || k < KP) acc += fast_ex2(e[k] - m);
        return m + fast_lg2(acc);
    }
    template <class V> __device__ __forceinline__ void finish(Coef& c, V, V) const { c.lp2 = lse2(c, 1.f, 0.f); }
    __device__ __forceinline__ float delta(const Coef& c, float cs, float sn) const { return 0.6931471805599453f * (lse2(c, cs, sn) - c.lp2); }
    __device__ __forceinline__ float logp_x(const Coef& c) const { return 0.6931471805599453f * c.lp2; }
    __device__ __forceinline__ Coef fetch(const Coef& c, int src) const {
        Coef o;
        const int nq = EXACT ? QUADS : (K + 3) >> 2;
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
            if (EXACT || (k >> 2) < nq) { o.A[k] = __shfl_sync(0xffffffffu, c.A[k], src); o.B[k] = __shfl_sync(0xffffffffu, c.B[k], src); }
            else { o.A[k] = 0.f; o.B[k] = 0.f; }
        }
        o.lp2 = __shfl_sync(0xffffffffu, c.lp2, src);
        return o;
    }
};
using VmfGc = VmfGcT<GC_MAX_K / 4, false>;


// column view of one lane's vector in shared memory
template <int BLOCK> struct SCol {
    float* p;
    __device__ __forceinline__ float& operator[](int j) const { return p[j * BLOCK]; }
};

// ------------------------------------------------------------------ kernel
// STORE: the d normals of the current transition are kept in shared memory next to x (needed for small d,
// where |z - (z.x) x|^2 is formed from the stored vector, and for dense Bingham).  Otherwise only x is
// resident -- half the shared memory, twice the resident warps -- the scores are accumulated against the
// raw normals on the fly (u = inp (z - zx x) is linear in z) and the accepted move regenerates the normals
// from the chain's Philox counter: x' = (c - s inp zx) x + (s inp) z.
template <class Policy, int KIND, bool STORE>
__global__ void __launch_bounds__(gc_block(STORE)) k_gc_fast(GcParams P, RunArgs a) {
    constexpr int BLOCK = gc_block(STORE);
    extern __shared__ __align__(16) float gc_smem[];
    const int d = P.d;
    SCol<BLOCK> xs{gc_smem + threadIdx.x}, us{gc_smem + (size_t)d * BLOCK + threadIdx.x};
    Policy pol;
    pol.stage(P, gc_smem + (STORE ? 2 : 1) * (size_t)d * BLOCK);
    __syncthreads();

    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < a.N;
    const int nbn = (d + 3) >> 2;
    const bool aux_in_last = philox_aux_in_last(d);
    const int pb = aux_in_last ? nbn : nbn + 1;
    const int quorum = a.quorum;
    const int n_steps = (int)a.n_steps;
    constexpr float LN2 = 0.6931471805599453f;

    float* st = static_cast<float*>(a.state) + (live ? c : 0) * d;
    for (int j = 0; j < d; ++j) { xs[j] = live ? st[j] : (j == 0 ? 1.f : 0.f); if (STORE) us[j] = 0.f; }

    PhiloxDraws<float> dr;
    dr.init(a.seed, a.chain0 + (uint64_t)c, d);
    unsigned long long rej_total = live ? a.n_reject[c] : 0ull;
    unsigned int rej_window = 0;
    float* samples = static_cast<float*>(a.samples);
    float* lpo = static_cast<float*>(a.logp);
    int since_keep = 0;
    int64_t kept = 0;

    // Normals of the current transition for the accepted move x' = alpha x + beta z when they are not kept in shared memory:
    // column layout (block b, chain c) -> zb[b * N], so that a warp's 32 lanes store / load 512 contiguous bytes.  Half of an
    // on-the-fly transition was Philox + Box-Muller because the normals were drawn twice; 2 x 4 d bytes of (mostly L2 / HBM)
    // traffic per transition replace the second draw (C3: 800 B x 2e9 /s = 1.6 TB/s, a quarter of the HBM bandwidth).
    float4* zb = (!STORE && a.zbuf && live) ? static_cast<float4*>(a.zbuf) + c : nullptr;

    typename Policy::Coef co;
    pol.begin(co);
    float lo = 0.f, hi = 2.f, lnu = 0.f, dl_acc = 0.f, lp_x = 0.f;
    float sn = 0.f, cs = 1.f, u_scale = 1.f, zx_keep = 0.f;  // direction u = u_scale (z - zx_keep x), z = the transition's normals
    int t = live ? 0 : n_steps;
    int r = 0, batch = 0;
    bool inloop = false, pending = false, moved = false;

    auto finish_transition = [&]() {
        // apply the accepted move and store.  x' = c x + s u is a unit vector up to fp32 rounding (x, u orthonormal), so the
        // renormalisation only has to stop the drift: it runs on every 16th transition of the chain, by the GLOBAL step index
        // (split-invariant, and warp-uniform in lock-step) -- one pass over x less on the other fifteen.
        // Only in the kernels with on-the-fly scores (d >= 64: C3 +7 %): in the stored-normals instantiation the same edit made the
        // dense-Bingham path 17 % slower (same work, different code generation; A/B on one box), so those keep renormalising always.
        const bool renorm = STORE || ((a.step0 + (uint64_t)(t - 1)) & 15u) == 0u;
        float xx = 0.f;
        if (moved) {
            const float beta = sn * u_scale, alpha = fmaf(-beta, zx_keep, cs);  // x' = alpha x + beta z
            if (STORE) {
                for (int j = 0; j < d; ++j) { const float v = fmaf(alpha, xs[j], beta * us[j]); xs[j] = v; xx = fmaf(v, v, xx); }
            } else {
                for (int b = 0; b < nbn; ++b) {
                    float n0, n1, n2, n3;
                    if (zb) {  // the normals of this transition, parked in global memory by the prologue (coalesced 16-byte loads)
                        const float4 n4 = zb[(size_t)b * a.N];
                        n0 = n4.x; n1 = n4.y; n2 = n4.z; n3 = n4.w;
                    } else {   // regenerated: dr still holds the counters of the transition being finished
                        const U4 w = dr.block(b);
                        box_muller_fast(w.x, w.y, n0, n1);
                        box_muller_fast(w.z, w.w, n2, n3);
                    }
                    const int j = 4 * b;
                    float v = fmaf(alpha, xs[j], beta * n0); xs[j] = v; xx = fmaf(v, v, xx);
                    if (j + 1 < d) { v = fmaf(alpha, xs[j + 1], beta * n1); xs[j + 1] = v; xx = fmaf(v, v, xx); }
                    if (j + 2 < d) { v = fmaf(alpha, xs[j + 2], beta * n2); xs[j + 2] = v; xx = fmaf(v, v, xx); }
                    if (j + 3 < d) { v = fmaf(alpha, xs[j + 3], beta * n3); xs[j + 3] = v; xx = fmaf(v, v, xx); }
                }
            }
        } else if (renorm) {
            for (int j = 0; j < d; ++j) xx = fmaf(xs[j], xs[j], xx);
        }
        if (renorm) {
            const float inx = rsqrtf(xx);
            for (int j = 0; j < d; ++j) xs[j] *= inx;
        }
        rej_total += (unsigned)r;
        rej_window += (unsigned)r;
        if (++since_keep == (int)a.thin) {
            const int64_t k = kept++;
            since_keep = 0;
            if (samples) {
                float* o = samples + (k * a.ss_step + c * a.ss_chain);
                if ((d & 1) == 0) for (int j = 0; j < d; j += 2) *reinterpret_cast<float2*>(o + j) = make_float2(xs[j], xs[j + 1]);
                else for (int j = 0; j < d; ++j) o[j] = xs[j];
            }
            if (lpo) lpo[k * a.sl_step + c * a.sl_chain] = lp_x + (moved ? dl_acc : 0.f);
            if (a.nrej) a.nrej[k * a.sl_step + c * a.sl_chain] = rej_window;
            if (a.fout) {  // scalar test function x . v (see test_function, kernels_fast.cuh)
                const float* v = static_cast<const float*>(a.fdir);
                float f = 0.f;
                for (int j = 0; j < d; ++j) f = fmaf(xs[j], __ldg(v + j), f);
                static_cast<float*>(a.fout)[k * a.sl_step + c * a.sl_chain] = f;
            }
            rej_window = 0;
        }
        pending = false;
    };

    while (true) {
        const bool need = !inloop && t < n_steps;
        const unsigned m_need = __ballot_sync(0xffffffffu, need);
        const unsigned m_loop = __ballot_sync(0xffffffffu, inloop);
        if ((m_need | m_loop) == 0u) break;

        if (m_loop == 0u || __popc(m_need) >= quorum) {
            if (need) {
                if (pending) finish_transition();
                else {
                    // first transition of a launch: normalise a state that is not unit yet (a user's
                    // initial state); a state written by finish_transition is left bit-for-bit alone so
                    // that splitting a run over launches does not change the chain
                    float xx = 0.f;
                    for (int j = 0; j < d; ++j) xx = fmaf(xs[j], xs[j], xx);
                    if (fabsf(xx - 1.f) > 1e-5f) {
                        const float inx = rsqrtf(xx);
                        for (int j = 0; j < d; ++j) xs[j] *= inx;
                    }
                }
                const uint64_t step = a.step0 + (uint64_t)t;
                dr.s_lo = (uint32_t)step;
                dr.s_hi = (uint32_t)(step >> 32) & 0xFFFFu;
                // one pass over the d normals: z.x, |z|^2 and the target's scores against z
                uint32_t aux0 = 0, aux1 = 0;
                float zx = 0.f, zz = 0.f;
                pol.begin(co);
                const bool on_the_fly = !pol.needs_vector();
                auto take = [&](int j, float n) {
                    const float xj = xs[j];
                    if (STORE) us[j] = n;
                    zx = fmaf(n, xj, zx);
                    zz = fmaf(n, n, zz);
                    if (on_the_fly) pol.accumulate(co, j, xj, n);
                };
                auto take_block = [&](int b, const U4& w) {
                    float n0, n1, n2, n3;
                    box_muller_fast(w.x, w.y, n0, n1);
                    box_muller_fast(w.z, w.w, n2, n3);
                    if (!STORE && zb) zb[(size_t)b * a.N] = make_float4(n0, n1, n2, n3);
                    const int j = 4 * b;
                    take(j, n0);
                    if (j + 1 < d) take(j + 1, n1);
                    if (j + 2 < d) take(j + 2, n2);
                    if (j + 3 < d) take(j + 3, n3);
                    if (b == nbn - 1) { aux0 = w.z; aux1 = w.w; }
                };
                // Stored normals (d < 64, dense Bingham): two Philox blocks in flight per iteration -- one block is a
                // chain of 10 dependent rounds and a scheduler holds ~3 warps (measured: Bingham d=50 diagonal +4 %,
                // dense +9 %).  With the scores accumulated on the fly (curved vMF d=100) the same change costs 16 %.
                int b = 0;
                if (STORE) {
                    for (; b + 1 < nbn; b += 2) {
                        const U4 w0 = dr.block(b), w1 = dr.block(b + 1);
                        take_block(b, w0);
                        take_block(b + 1, w1);
                    }
                }
                for (; b < nbn; ++b) take_block(b, dr.block(b));
                if (!aux_in_last) { const U4 w = dr.block(nbn); aux0 = w.x; aux1 = w.y; }
                float pp;
                if (STORE) {
                    if (d < 32) {  // |z - zx x|^2 from the stored vector: no cancellation for small d
                        pp = 0.f;
                        for (int j = 0; j < d; ++j) { const float p = fmaf(-zx, xs[j], us[j]); pp = fmaf(p, p, pp); }
                    } else {
                        pp = fmaf(-zx, zx, zz);  // x is unit and zx^2 ~ 1 << |z|^2 ~ d: one pass over x and z less (d = 50: -7 % instructions)
                    }
                    if (!on_the_fly) for (int j = 0; j < d; ++j) pol.accumulate(co, j, xs[j], us[j]);
                } else {
                    pp = fmaf(-zx, zx, zz);  // x is unit; d >= 64, where zx^2 << |z|^2
                }
                u_scale = rsqrtf(pp);
                zx_keep = zx;
                pol.finish(co, xs, us);
                pol.fixup(co, zx, u_scale);
                lp_x = pol.logp_x(co);
                lnu = LN2 * fast_lg2(u01(aux0, 0.f));
                if (KIND == SHRINK) { hi = 2.0f * u01(aux1, 0.f); lo = hi - 2.0f; }
                else { lo = 0.f; hi = 2.0f; }
                r = 0; batch = 0; moved = false;
                ++t;
                inloop = true;
            }
        }

        // cooperative batches for the reject sampler: see k_vmf_fast (kernels_fast.cuh)
        bool own_batch = inloop;
        if (KIND == REJECT) {
            const unsigned m_now = __ballot_sync(0xffffffffu, inloop);
            const int nW = __popc(m_now);
            if (nW > 0 && nW <= 16) {
                own_batch = false;
                const int lane = threadIdx.x & 31;
                const int lg = 31 - __clz(32 / nW);
                const int g = 1 << lg;
                const int grp = lane >> lg, sub = lane & (g - 1);
                const bool helping = grp < nW;
                const int owner = __fns(m_now, 0, (helping ? grp : 0) + 1);
                const typename Policy::Coef co_o = pol.fetch(co, owner);
                const float lnu_o = __shfl_sync(0xffffffffu, lnu, owner);
                const int batch_o = __shfl_sync(0xffffffffu, batch, owner);
                const int t_o = __shfl_sync(0xffffffffu, t, owner);
                int jacc = 0x7fffffff;
                float sn_h = 0.f, cs_h = 1.f, dl_h = 0.f;
                if (helping) {
                    const uint64_t chain_o = a.chain0 + (uint64_t)(c - lane + owner);
                    const uint64_t step_o = a.step0 + (uint64_t)(t_o - 1);
                    const U4 w = philox4x32_10(U4{(uint32_t)chain_o, (uint32_t)(chain_o >> 32), (uint32_t)step_o,
                                                  (((uint32_t)(step_o >> 32) & 0xFFFFu) << 16) | (uint32_t)(pb + batch_o + sub)},
                                               dr.k0, dr.k1);
                    const uint32_t wi[FAST_BATCH] = {w.x, w.y, w.z, w.w};
                    float sb_[FAST_BATCH], cb_[FAST_BATCH], dl[FAST_BATCH];
#pragma unroll
                    for (int i = 0; i < FAST_BATCH; ++i) {
                        fast_sincospi(fmaf(2.0f, u01(wi[i], 0.f), 0.f), sb_[i], cb_[i]);
                        dl[i] = pol.delta(co_o, cb_[i], sb_[i]);
                    }
#pragma unroll
                    for (int i = FAST_BATCH - 1; i >= 0; --i)
                        if (dl[i] > lnu_o) { jacc = FAST_BATCH * sub + i; sn_h = sb_[i]; cs_h = cb_[i]; dl_h = dl[i]; }
                }
                int jmin = jacc;
                for (int o = g >> 1; o > 0; o >>= 1) jmin = min(jmin, __shfl_xor_sync(0xffffffffu, jmin, o));
                const int rank = __popc(m_now & ((1u << lane) - 1u));
                const int src = (inloop ? rank : 0) << lg;
                const int jmine = __shfl_sync(0xffffffffu, jmin, src);
                const int wl = src + (jmine == 0x7fffffff ? 0 : jmine / FAST_BATCH);
                const float sn_w = __shfl_sync(0xffffffffu, sn_h, wl);
                const float cs_w = __shfl_sync(0xffffffffu, cs_h, wl);
                const float dl_w = __shfl_sync(0xffffffffu, dl_h, wl);
                if (inloop) {
                    if (jmine != 0x7fffffff) {
                        sn = sn_w; cs = cs_w; dl_acc = dl_w;
                        r += jmine;
                        moved = true; inloop = false; pending = true;
                    } else {
                        r += FAST_BATCH * g;
                        batch += g;
                        if (r >= MAX_PROPOSALS || !(lp_x == lp_x)) { inloop = false; pending = true; }
                    }
                }
            }
        }

        if (own_batch) {
            const U4 w = dr.block(pb + batch);
            const uint32_t wi[FAST_BATCH] = {w.x, w.y, w.z, w.w};
            float tt[FAST_BATCH], sb_[FAST_BATCH], cb_[FAST_BATCH], dl[FAST_BATCH];
            float lo_ = lo, hi_ = hi;
#pragma unroll
            for (int i = 0; i < FAST_BATCH; ++i) {
                tt[i] = fmaf(hi_ - lo_, u01(wi[i], 0.f), lo_);
                if (KIND == SHRINK) { if (tt[i] < 0.f) lo_ = tt[i]; else hi_ = tt[i]; }
            }
#pragma unroll
            for (int i = 0; i < FAST_BATCH; ++i) {
                fast_sincospi(tt[i], sb_[i], cb_[i]);
                dl[i] = pol.delta(co, cb_[i], sb_[i]);
            }
            int first = FAST_BATCH;
#pragma unroll
            for (int i = FAST_BATCH - 1; i >= 0; --i) if (dl[i] > lnu) first = i;
            const int cap = KIND == SHRINK ? FAST_SHRINK_CAP : MAX_PROPOSALS;
            if (first < FAST_BATCH) {
#pragma unroll
                for (int i = 0; i < FAST_BATCH; ++i) if (i == first) { sn = sb_[i]; cs = cb_[i]; dl_acc = dl[i]; }
                r += first;
                moved = true; inloop = false; pending = true;
            } else {
                r += FAST_BATCH;
                lo = lo_; hi = hi_;
                ++batch;
                if (r >= cap || !(lp_x == lp_x)) { inloop = false; pending = true; }  // cap / NaN density: stay at x
            }
        }
    }

    if (live) {
        if (pending) finish_transition();
        for (int j = 0; j < d; ++j) st[j] = xs[j];
        a.n_reject[c] = rej_total;
    }
}

}  // namespace geosss
