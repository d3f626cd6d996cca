// Float32 throughput kernels ("fast path") for the vMF-mixture target.
//
// Same Markov kernel as geosss/mcmc.py:357-401, restructured for the GPU (SURVEY.md section 7,
// "algebraic shortcuts along the great circle"): for y(theta) = cos(theta) x + sin(theta) u the
// component scores mu_k . y are linear in (cos, sin), so one transition computes
// a_k = mu_k.x and b_k = mu_k.u once (2Kd FMAs) and every proposal costs O(K):
//     accept  <=>  log p(y) > log p(x) + log U  <=>  sum_k 2^(c A_k + s B_k + G_k) > 1
// with A,B scaled by log2(e) and G_k = C_k - (log2 p(x) + log2 U).  Comparing the SUM with 1
// never forms "log p(x) + log U", so the level cannot round onto log p(x) (the bracket-collapse
// hazard of SURVEY.md section 7), and no log is needed per proposal.
//
// Deliberate deviations from the literal path, all distribution-preserving: x is renormalised
// at every transition (the reference lets |x| drift), the bracket is kept in half-turns and
// (cos, sin) come from the MUFU pipe followed by a Newton renormalisation (fast_sincospi), log p(x)
// is not re-evaluated from scratch but from the fresh a_k, and a proposal cap keeps a chain at x
// instead of looping for ever.
//
// One lane per chain; state, direction and the 2K scores live in registers (D, K are template
// parameters, D padded to a multiple of 4 with zeros); mixture parameters are read from the
// kernel-parameter constant bank.  Draws come from the per-chain Philox stream of philox.cuh.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "kernels_literal.cuh"  // RunArgs
#include "philox.cuh"

namespace geosss {

constexpr int FAST_MAX_K = 8;
constexpr int FAST_MAX_D = 16;
constexpr int FAST_BLOCK = 64;
constexpr int FAST_SHRINK_CAP = 200;  // a shrink bracket is far below fp32 resolution long before this

struct VmfFastParams {
    float mu[FAST_MAX_K * FAST_MAX_D];  // (K, D) zero padded, times log2(e)
    float c[FAST_MAX_K];                // (log w_k - log 2pi - log i0(kappa_k)) * log2(e)
    int d;                              // true dimension (<= D)
};

// (sin, cos) of pi * t for |t| <= 2 on the MUFU pipe (2 MUFU + 7 FMA-pipe instructions instead of the
// ~25 of sincospif): |error| ~ 5e-7 rad on the angle, then one Newton step puts (c, s) back on the unit
// circle to fp32 rounding so that y = c x + s u keeps unit norm.  The angle error is a smooth 1e-6-relative
// perturbation of a uniformly drawn proposal angle.
__device__ __forceinline__ void fast_sincospi(float t, float& sn, float& cs) {
    const float a = 3.14159265358979f * t;
    float s0 = __sinf(a), c0 = __cosf(a);
    const float f = fmaf(-0.5f, fmaf(s0, s0, c0 * c0), 1.5f);
    sn = s0 * f;
    cs = c0 * f;
}

// Box-Muller on the MUFU pipe: |error| ~ 1e-6 per normal, irrelevant for a proposal direction.
__device__ __forceinline__ void box_muller_fast(uint32_t wa, uint32_t wb, float& z0, float& z1) {
    const float u1 = u01(wa, 0.f), u2 = u01(wb, 0.f);
    const float r = fast_sqrt(fmaxf(-1.3862943611198906f * fast_lg2(u1), 0.f));  // sqrt(-2 ln u1)
    float sn, cs;
    __sincosf(6.283185307179586f * u2, &sn, &cs);
    z0 = r * cs;
    z1 = r * sn;
}

// Stores the first d of x[0..D) to o, vectorised when d allows (o is 8- resp. 16-byte aligned
// whenever d is even resp. a multiple of 4, because rows are d floats apart).
template <int DT> __device__ __forceinline__ void store_row(float* __restrict__ o, const float (&x)[DT]) {
    if (DT % 4 == 0) {
#pragma unroll
        for (int j = 0; j + 3 < DT; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
    } else if (DT % 2 == 0) {
#pragma unroll
        for (int j = 0; j + 1 < DT; j += 2) *reinterpret_cast<float2*>(o + j) = make_float2(x[j], x[j + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < DT; ++j) o[j] = x[j];
    }
}

// f(x) = x . v, the scalar test function whose trace the diagnostics use (first coordinate: scripts/ess_mixture_vMF.py:233;
// x . mode: scripts/bingham.py:50-52); v is read through the uniform path (same address for every lane)
template <int DT> __device__ __forceinline__ float test_function(const float (&x)[DT], const float* __restrict__ v) {
    float f = 0.f;
#pragma unroll
    for (int j = 0; j < DT; ++j) f = fmaf(x[j], __ldg(v + j), f);
    return f;
}

constexpr int FAST_BATCH = 4;  // proposals evaluated together = words of one Philox block

// log2 p at the accepted point: level2 + log2(S) is exact for it; when S overflowed fp32 the
// mixture is evaluated directly with a max shift.
template <int K>
__device__ __forceinline__ float kept_log2p(bool moved, float level2, float S, float lp2, float cs, float sn,
                                            const float (&A)[K], const float (&B)[K], const float* cst) {
    float l2 = moved ? level2 + fast_lg2(S) : lp2;
    if (!(l2 < INFINITY)) {
        float mm = -INFINITY, e[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { e[k] = fmaf(cs, A[k], sn * B[k]) + cst[k]; mm = fmaxf(mm, e[k]); }
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) ss += fast_ex2(e[k] - mm);
        l2 = mm + fast_lg2(ss);
    }
    return l2;
}

// Persistent bracket loop with quorum scheduling.
//
// A transition has two parts with very different shapes: a fixed-cost PROLOGUE (renormalise x,
// d normals -> direction u, scores A_k = mu_k.x, B_k = mu_k.u, level) and a BRACKET LOOP whose
// trip count differs from chain to chain (geosss/mcmc.py:394-401; p99/mean of the rejection count
// is ~3 for shrink and ~7 for reject).  Run in lock-step, the 32 chains of a
// warp would all wait for the slowest one.  Here every lane owns its chain's step counter and the
// warp keeps iterating
//     vote;  if >= quorum lanes wait for a new transition (or nobody is in a loop): those lanes
//            apply their accepted move, store the sample and run the prologue;
//            lanes inside a bracket loop evaluate their next batch of proposals;
// so that lanes drift apart by a few transitions instead of idling.  The quorum trades idle lanes
// in the prologue against idle lanes in the loop (quorum 32 = lock-step).  Per-chain results do not
// depend on the schedule: every draw is addressed by (seed, chain, step, block).
//
// A batch is four proposals.  The angles of a batch depend only on the uniforms and on the SIGNS
// of the earlier angles (mcmc.py:400), never on the density, so all four are known up front; their
// densities are evaluated with four-way ILP and the first accepted one wins -- identical to
// evaluating them one after the other.
//
// The true dimension DT is a template parameter (the parameter block keeps rows of D = 4 ceil(DT/4) floats):
// the unused Box-Muller pair of the last Philox block, the padded FMAs of the projections and of the 2K dot
// products and the tail predicates all disappear at compile time (d = 10: 52 of the prologue's ~580 instructions).
template <int K, int DT, int KIND>
__global__ void __launch_bounds__(FAST_BLOCK) k_vmf_fast(const __grid_constant__ VmfFastParams P, RunArgs a) {
    constexpr int D = (DT + 3) / 4 * 4;
    constexpr int d = DT;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < a.N;  // lanes past the end still vote
    constexpr float LN2 = 0.6931471805599453f;
    constexpr bool aux_in_last = (DT & 3) == 1 || (DT & 3) == 2;  // philox_aux_in_last
    constexpr int nbn = D / 4;                       // normal blocks
    constexpr int pb = aux_in_last ? nbn : nbn + 1;  // first proposal block
    const int quorum = a.quorum;
    const int n_steps = (int)a.n_steps;

    float x[DT], u[DT];
    float* st = static_cast<float*>(a.state) + (live ? c : 0) * d;
#pragma unroll
    for (int j = 0; j < DT; ++j) { x[j] = live ? st[j] : 0.0f; u[j] = 0.f; }

    PhiloxDraws<float> dr;
    dr.init(a.seed, a.chain0 + (uint64_t)c, d);
    unsigned long long rej_total = live ? a.n_reject[c] : 0ull;
    unsigned int rej_window = 0;
    float* samples = static_cast<float*>(a.samples);
    float* lpo = static_cast<float*>(a.logp);
    int since_keep = 0;
    int64_t kept = 0;

    float A[K], B[K], G[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { A[k] = 0.f; B[k] = 0.f; G[k] = 0.f; }
    float lo = 0.f, hi = 2.f, level2 = 0.f, lp2 = 0.f;
    bool have_lp = false;  // lp2 = log2 p(x) is carried from the accepted proposal (level2 + log2 S) instead of re-evaluated
    float sn = 0.f, cs = 1.f, S = 0.f;  // the accepted proposal, applied at the next prologue
    int t = live ? 0 : n_steps;         // transitions started by this lane
    int r = 0, batch = 0, cap = 0;
    bool inloop = false, pending = false, moved = false;

    while (true) {
        const bool need = !inloop && t < n_steps;
        const unsigned m_need = __ballot_sync(0xffffffffu, need);
        const unsigned m_loop = __ballot_sync(0xffffffffu, inloop);
        if ((m_need | m_loop) == 0u) break;

        if (m_loop == 0u || __popc(m_need) >= quorum) {
            if (need) {
                // ---- finish the previous transition: move, store ----
                if (pending) {
                    if (moved) {
#pragma unroll
                        for (int j = 0; j < DT; ++j) x[j] = fmaf(cs, x[j], sn * u[j]);
                        lp2 = kept_log2p<K>(true, level2, S, lp2, cs, sn, A, B, P.c);
                    }
                    rej_total += (unsigned)r;
                    rej_window += (unsigned)r;
                    if (++since_keep == (int)a.thin) {
                        const int64_t k = kept++;
                        since_keep = 0;
                        if (samples) store_row<DT>(samples + (k * a.ss_step + c * a.ss_chain), x);
                        if (lpo) lpo[k * a.sl_step + c * a.sl_chain] = lp2 * LN2;
                        if (a.nrej) a.nrej[k * a.sl_step + c * a.sl_chain] = rej_window;
                        if (a.fout) static_cast<float*>(a.fout)[k * a.sl_step + c * a.sl_chain] = test_function<DT>(x, static_cast<const float*>(a.fdir));
                        rej_window = 0;
                    }
                    pending = false;
                }
                // ---- prologue of transition t ----
                const uint64_t step = a.step0 + (uint64_t)t;
                dr.s_lo = (uint32_t)step;
                dr.s_hi = (uint32_t)(step >> 32) & 0xFFFFu;
                float xx = 0.f;
#pragma unroll
                for (int j = 0; j < DT; ++j) xx = fmaf(x[j], x[j], xx);
                const float inx = rsqrtf(xx);
#pragma unroll
                for (int j = 0; j < DT; ++j) x[j] *= inx;

                uint32_t aux0 = 0, aux1 = 0;
#pragma unroll
                for (int b = 0; b < nbn; ++b) {
                    U4 w = dr.block(b);
                    float z0, z1;
                    box_muller_fast(w.x, w.y, z0, z1);
                    u[4 * b] = z0;
                    if (4 * b + 1 < DT) u[4 * b + 1] = z1;
                    if (4 * b + 2 < DT) {
                        box_muller_fast(w.z, w.w, z0, z1);
                        u[4 * b + 2] = z0;
                        if (4 * b + 3 < DT) u[4 * b + 3] = z1;
                    }
                    if (b == nbn - 1) { aux0 = w.z; aux1 = w.w; }
                }
                if (!aux_in_last) { U4 w = dr.block(nbn); aux0 = w.x; aux1 = w.y; }
                float zx = 0.f;
#pragma unroll
                for (int j = 0; j < DT; ++j) zx = fmaf(u[j], x[j], zx);
                float pp = 0.f;
#pragma unroll
                for (int j = 0; j < DT; ++j) { u[j] = fmaf(-zx, x[j], u[j]); pp = fmaf(u[j], u[j], pp); }
                const float inp = rsqrtf(pp);
#pragma unroll
                for (int j = 0; j < DT; ++j) u[j] *= inp;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float sa = 0.f, sb = 0.f;
#pragma unroll
                    for (int j = 0; j < DT; ++j) { sa = fmaf(P.mu[k * D + j], x[j], sa); sb = fmaf(P.mu[k * D + j], u[j], sb); }
                    A[k] = sa; B[k] = sb;
                }
                if (!have_lp) {  // first transition of the launch: evaluate the mixture at x
                    float mx = -INFINITY;
#pragma unroll
                    for (int k = 0; k < K; ++k) mx = fmaxf(mx, A[k] + P.c[k]);
                    float s0 = 0.f;
#pragma unroll
                    for (int k = 0; k < K; ++k) s0 += fast_ex2(A[k] + P.c[k] - mx);
                    lp2 = mx + fast_lg2(s0);
                    have_lp = true;
                }
                level2 = lp2 + fast_lg2(u01(aux0, 0.f));
#pragma unroll
                for (int k = 0; k < K; ++k) G[k] = P.c[k] - level2;
                if (KIND == SHRINK) { hi = 2.0f * u01(aux1, 0.f); lo = hi - 2.0f; }  // half-turns: theta = pi * tt
                else { lo = 0.f; hi = 2.0f; }
                cap = (lp2 > -INFINITY) ? (KIND == SHRINK ? FAST_SHRINK_CAP : MAX_PROPOSALS) : 0;
                r = 0; batch = 0; moved = false;
                ++t;
                if (cap > 0) inloop = true; else pending = true;  // density is -inf/NaN everywhere: stay
            }
        }

        // ---- cooperative batches (reject sampler): when at most 16 chains of the warp are still looking
        // for an acceptable proposal, the idle lanes help.  Reject proposals are i.i.d. uniform on the
        // circle and proposal j of a chain is a pure function of (seed, chain, step, j), so g = 32/nW lanes
        // can evaluate 4g consecutive proposals of one chain at once; the lowest accepted index wins,
        // which is exactly what the sequential loop of geosss/mcmc.py:368-374 returns.
        bool own_batch = inloop;
        if (KIND == REJECT) {
            const unsigned m_now = __ballot_sync(0xffffffffu, inloop);
            const int nW = __popc(m_now);
            if (nW > 0 && nW <= 16) {
                own_batch = false;
                const int lane = threadIdx.x & 31;
                const int lg = 31 - __clz(32 / nW);  // helpers per chain g = 2^lg
                const int g = 1 << lg;
                const int grp = lane >> lg, sub = lane & (g - 1);
                const bool helping = grp < nW;
                const int owner = __fns(m_now, 0, (helping ? grp : 0) + 1);
                float Ao[K], Bo[K], Go[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    Ao[k] = __shfl_sync(0xffffffffu, A[k], owner);
                    Bo[k] = __shfl_sync(0xffffffffu, B[k], owner);
                    Go[k] = __shfl_sync(0xffffffffu, G[k], owner);
                }
                const int batch_o = __shfl_sync(0xffffffffu, batch, owner);
                const int t_o = __shfl_sync(0xffffffffu, t, owner);
                int jacc = 0x7fffffff;
                float sn_h = 0.f, cs_h = 1.f, S_h = 0.f;
                if (helping) {
                    const uint64_t chain_o = a.chain0 + (uint64_t)(c - lane + owner);
                    const uint64_t step_o = a.step0 + (uint64_t)(t_o - 1);
                    const U4 w = philox4x32_10(U4{(uint32_t)chain_o, (uint32_t)(chain_o >> 32), (uint32_t)step_o,
                                                  (((uint32_t)(step_o >> 32) & 0xFFFFu) << 16) | (uint32_t)(pb + batch_o + sub)},
                                               dr.k0, dr.k1);
                    const uint32_t wi[FAST_BATCH] = {w.x, w.y, w.z, w.w};
                    float sb_[FAST_BATCH], cb_[FAST_BATCH], Sb[FAST_BATCH];
#pragma unroll
                    for (int i = 0; i < FAST_BATCH; ++i) {
                        fast_sincospi(fmaf(2.0f, u01(wi[i], 0.f), 0.f), sb_[i], cb_[i]);
                        float acc = 0.f;
#pragma unroll
                        for (int k = 0; k < K; ++k) acc += fast_ex2(fmaf(cb_[i], Ao[k], fmaf(sb_[i], Bo[k], Go[k])));
                        Sb[i] = acc;
                    }
#pragma unroll
                    for (int i = FAST_BATCH - 1; i >= 0; --i)
                        if (Sb[i] > 1.0f) { jacc = FAST_BATCH * sub + i; sn_h = sb_[i]; cs_h = cb_[i]; S_h = Sb[i]; }
                }
                int jmin = jacc;
                for (int o = g >> 1; o > 0; o >>= 1) jmin = min(jmin, __shfl_xor_sync(0xffffffffu, jmin, o));
                // every owner reads its group's verdict (group index = rank of the owner among looping lanes)
                const int rank = __popc(m_now & ((1u << lane) - 1u));
                const int src = (inloop ? rank : 0) << lg;
                const int jmine = __shfl_sync(0xffffffffu, jmin, src);
                const int wl = src + (jmine == 0x7fffffff ? 0 : jmine / FAST_BATCH);
                const float sn_w = __shfl_sync(0xffffffffu, sn_h, wl);
                const float cs_w = __shfl_sync(0xffffffffu, cs_h, wl);
                const float S_w = __shfl_sync(0xffffffffu, S_h, wl);
                if (inloop) {
                    if (jmine != 0x7fffffff) {
                        sn = sn_w; cs = cs_w; S = S_w;
                        r += jmine;
                        moved = true; inloop = false; pending = true;
                    } else {
                        r += FAST_BATCH * g;
                        batch += g;
                        if (r >= cap) { inloop = false; pending = true; }
                    }
                }
            }
        }

        if (own_batch) {
            const U4 w = dr.block(pb + batch);
            const uint32_t wi[FAST_BATCH] = {w.x, w.y, w.z, w.w};
            float tt[FAST_BATCH], sb_[FAST_BATCH], cb_[FAST_BATCH], Sb[FAST_BATCH];
            float lo_ = lo, hi_ = hi;
#pragma unroll
            for (int i = 0; i < FAST_BATCH; ++i) {
                tt[i] = fmaf(hi_ - lo_, u01(wi[i], 0.f), lo_);
                if (KIND == SHRINK) { if (tt[i] < 0.f) lo_ = tt[i]; else hi_ = tt[i]; }
            }
#pragma unroll
            for (int i = 0; i < FAST_BATCH; ++i) {
                fast_sincospi(tt[i], sb_[i], cb_[i]);
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) acc += fast_ex2(fmaf(cb_[i], A[k], fmaf(sb_[i], B[k], G[k])));
                Sb[i] = acc;
            }
            int first = FAST_BATCH;
#pragma unroll
            for (int i = FAST_BATCH - 1; i >= 0; --i) if (Sb[i] > 1.0f) first = i;
            if (first < FAST_BATCH) {
#pragma unroll
                for (int i = 0; i < FAST_BATCH; ++i) if (i == first) { sn = sb_[i]; cs = cb_[i]; S = Sb[i]; }
                r += first;
                moved = true; inloop = false; pending = true;
            } else {
                r += FAST_BATCH;
                lo = lo_; hi = hi_;
                ++batch;
                if (r >= cap) { inloop = false; pending = true; }  // cap reached: stay at x
            }
        }
    }

    // flush the last transition of every lane (converged)
    if (live) {
        if (pending) {
            if (moved) {
#pragma unroll
                for (int j = 0; j < DT; ++j) x[j] = fmaf(cs, x[j], sn * u[j]);
                lp2 = kept_log2p<K>(true, level2, S, lp2, cs, sn, A, B, P.c);
            }
            rej_total += (unsigned)r;
            rej_window += (unsigned)r;
            if (++since_keep == (int)a.thin) {
                const int64_t k = kept++;
                if (samples) store_row<DT>(samples + (k * a.ss_step + c * a.ss_chain), x);
                if (lpo) lpo[k * a.sl_step + c * a.sl_chain] = lp2 * LN2;
                if (a.nrej) a.nrej[k * a.sl_step + c * a.sl_chain] = rej_window;
                if (a.fout) static_cast<float*>(a.fout)[k * a.sl_step + c * a.sl_chain] = test_function<DT>(x, static_cast<const float*>(a.fdir));
            }
        }
#pragma unroll
        for (int j = 0; j < DT; ++j) st[j] = x[j];
        a.n_reject[c] = rej_total;
    }
}

// ------------------------------------------------------------------ host-side dispatch
inline int fast_supported(int target_kind, int d, int K, int single) {
    (void)single;
    if (target_kind != TGT_VMF_MIX) return -1;
    if (d < 2 || d > FAST_MAX_D || K > FAST_MAX_K || K < 1) return -1;
    return 0;
}

// (K, D) zero-padded parameter block in log2 units
inline VmfFastParams make_vmf_fast_params(int d, int K, const double* mu, const double* consts) {
    const int Dp = (d + 3) / 4 * 4;
    VmfFastParams P;
    const double L2E = 1.4426950408889634;
    for (int k = 0; k < FAST_MAX_K; ++k) {
        for (int j = 0; j < FAST_MAX_D; ++j) P.mu[k * FAST_MAX_D + j] = 0.f;
        P.c[k] = 0.f;
    }
    for (int k = 0; k < K; ++k) {
        for (int j = 0; j < d; ++j) P.mu[k * Dp + j] = (float)(mu[(size_t)k * d + j] * L2E);
        P.c[k] = (float)(consts[k] * L2E);
    }
    P.d = d;
    return P;
}

// returns 0, a negative GEOSSS_E_* code, or a positive cudaError_t (defined in vmf_fast.cu: the K x D x kind
// instantiations of k_vmf_fast are compiled in their own translation unit)
int launch_fast(int target_kind, int kind, int d, int K, const double* mu, const double* consts, const RunArgs& a, cudaStream_t st);

}  // namespace geosss
