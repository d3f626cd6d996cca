// Registration likelihood with a cutoff: cell lists in the SOURCE frame.
//
// Replaces, for the float32 throughput path, the per-evaluation KD-tree of the reference
// (geosss/registration.py:101-109: a tree of the rotated source is built for every log_prob call and
// queried for the k nearest points of each target row) and the dense 214 x 214 pair sum of
// CpdWarpLogPFast (kernels_literal.cuh).
//
// |y_l - (R x_k + t)|^2 = |R^T (y_l - t) - x_k|^2, so instead of rotating the source cloud the TARGET rows
// are rotated into the source frame, where the source points never move.  A uniform grid over the source
// cloud (cell edge h >= the cutoff radius rc) is built once on the host; every cell carries the ready-made
// list of all source points of its 27-cell neighbourhood as float4 (x, y, z, folded log-weight), so a
// target row costs one cell lookup and one contiguous run of 16-byte shared-memory loads.  Every source
// point within rc of the row is in that run; the others -- dropped -- each contribute less than
// CPD_CELL_EPS of the outlier floor of the row sum (registration.py:237-243: the outlier column keeps the
// row sum >= omega / V), so log p changes by less than nt * ns * eps relative -- far below float32
// rounding.  With the protein clouds of scripts/protein_reg3d3d.py (214 x 214, sigma = 1, rc = 7.1) a row
// meets 26 candidates on average instead of 214.
//
// Mapping: warp per chain (all lanes carry the same quaternion and draws, as in the dense kernel);
// four lanes per target row, eight rows per warp pass; the lists, the cell table and the target rows
// sit in shared memory once per block and serve all of its warps for all steps of the launch.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "kernels_literal.cuh"
#include "philox.cuh"
#include "targets.cuh"
#include "transition.cuh"
#include "transition_mh.cuh"

namespace geosss {

constexpr int CPD_CELL_WARPS = 16;            // warps (= chains or points) per block
constexpr double CPD_CELL_LOG_EPS = -17.5;    // log of the largest dropped term relative to the outlier floor (2.5e-8)
constexpr int CPD_CELL_COUNT_SHIFT = 20;      // cell word = first entry | (entries << 20)

struct CpdCellGrid {
    float org[3];      // lower corner of the padded grid, source frame
    float inv_h;       // 1 / cell edge
    int n[3];          // cells per axis, padding included
    int n_cells, n_entries, nt;
    const uint32_t* cell;  // (n_cells) device
    const float4* entry;   // (n_entries) device: (x, y, z) in cell units from org, (log w_k + log_const - M0) log2(e)
    const float4* tgt4;    // (nt) device: target row and its weight
    float q2, base, m0, beta;  // -0.5 log2(e) h^2 / sigma^2, exp(log_outlier - M0), M0, beta
    float shift[3];
    float rc2;  // (cutoff / h)^2: squared cutoff radius in cell units
    int k;      // k-nearest truncation of the reference (registration.py:107-108); <= 0: every source point
    size_t smem_bytes() const { return ((size_t)n_entries + nt) * sizeof(float4) + (size_t)n_cells * sizeof(uint32_t); }
};

struct CpdCellShared {
    const float4* entry;
    const float4* tgt4;
    const uint32_t* cell;
};

// cooperative copy of the lists into shared memory (whole block), then __syncthreads
__device__ __forceinline__ CpdCellShared cpd_cell_stage(const CpdCellGrid& g, unsigned char* smem_raw) {
    float4* e = reinterpret_cast<float4*>(smem_raw);
    float4* t4 = e + g.n_entries;
    uint32_t* cl = reinterpret_cast<uint32_t*>(t4 + g.nt);
    for (int i = threadIdx.x; i < g.n_entries; i += blockDim.x) e[i] = g.entry[i];
    for (int i = threadIdx.x; i < g.nt; i += blockDim.x) t4[i] = g.tgt4[i];
    for (int i = threadIdx.x; i < g.n_cells; i += blockDim.x) cl[i] = g.cell[i];
    __syncthreads();
    return CpdCellShared{e, t4, cl};
}

// All coordinates are in CELL UNITS from the grid corner: entries hold (x_k - org) / h, the row is
// f = (R^T / h)(y - t) - org / h, so floor(f) is the cell and |f - e|^2 h^2 the squared distance (h^2 is folded
// into q2).  A row outside the padded grid is clamped onto a border cell: the candidates it then meets are all
// farther than h -- legitimate, negligible terms of the dense sum.
template <bool KNN> struct CpdWarpLogPCell {
    CpdCellGrid g;
    CpdCellShared s;

    // Four lanes per row, eight rows per pass: measured on B200 against eight lanes per row (more passes, each
    // paying the cell lookup; half-empty candidate iterations) and against two row groups in flight per lane
    // (registers: one block per SM instead of two) -- profiles/r1_ncu_full_k_logp_cpd_cell.txt.
    template <class Vec> __device__ __forceinline__ float operator()(Vec q) const {
        const int lane = threadIdx.x & 31;
        const int grp = lane >> 2, sub = lane & 3;
        float R[9];
        quat_to_matrix<float>(q[0], q[1], q[2], q[3], R);
        // M = R^T / h, c = -(R^T t + org) / h
        float M[9], c[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) M[3 * i + jj] = R[3 * jj + i] * g.inv_h;
            c[i] = -fmaf(M[3 * i], g.shift[0], fmaf(M[3 * i + 1], g.shift[1], fmaf(M[3 * i + 2], g.shift[2], g.org[i] * g.inv_h)));
        }
        const float hx = (float)g.n[0] - 1.f, hy = (float)g.n[1] - 1.f, hz = (float)g.n[2] - 1.f;
        double acc = 0.0;
#pragma unroll 1
        for (int l0 = 0; l0 < g.nt; l0 += 8) {
            const int l = l0 + grp;
            const float4 y = s.tgt4[l < g.nt ? l : g.nt - 1];
            const float fx = fmaf(M[0], y.x, fmaf(M[1], y.y, fmaf(M[2], y.z, c[0])));
            const float fy = fmaf(M[3], y.x, fmaf(M[4], y.y, fmaf(M[5], y.z, c[1])));
            const float fz = fmaf(M[6], y.x, fmaf(M[7], y.y, fmaf(M[8], y.z, c[2])));
            const int cx = (int)fminf(fmaxf(fx, 0.f), hx), cy = (int)fminf(fmaxf(fy, 0.f), hy), cz = (int)fminf(fmaxf(fz, 0.f), hz);
            const uint32_t cw = s.cell[(cz * g.n[1] + cy) * g.n[0] + cx];
            const int cnt = (int)(cw >> CPD_CELL_COUNT_SHIFT);
            const float4* e = s.entry + (cw & ((1u << CPD_CELL_COUNT_SHIFT) - 1u));
            float sum = 0.f;
            int inside = 0;  // candidates within the cutoff radius (KNN only)
#pragma unroll 2
            for (int k = sub; k < cnt; k += 4) {
                const float4 m = e[k];
                const float e0 = fx - m.x, e1 = fy - m.y, e2 = fz - m.z;
                const float d2 = fmaf(e0, e0, fmaf(e1, e1, e2 * e2));
                if (KNN) inside += d2 < g.rc2 ? 1 : 0;
                sum += fast_ex2(fmaf(d2, g.q2, m.w));
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            if (KNN) {
                // The reference sums the k nearest source points only.  With at most k points inside the cutoff
                // the other members of that set are beyond it (negligible, like the candidates summed here), so
                // the two sums agree.  With more than k inside, the k nearest are all candidates: keep exactly
                // those (rank by distance, ties by list position).  Rare: never at the shipped sigma = 1, k = 20.
                inside += __shfl_xor_sync(0xffffffffu, inside, 1);
                inside += __shfl_xor_sync(0xffffffffu, inside, 2);
                const bool crowded = inside > g.k;
                if (__any_sync(0xffffffffu, crowded)) {
                    const int n = crowded ? cnt : 0;
                    float exact = 0.f;
                    for (int i = sub; i < n; i += 4) {
                        const float4 mi = e[i];
                        const float a0 = fx - mi.x, a1 = fy - mi.y, a2 = fz - mi.z;
                        const float di = fmaf(a0, a0, fmaf(a1, a1, a2 * a2));
                        int rank = 0;
                        for (int j = 0; j < n; ++j) {
                            const float4 mj = e[j];
                            const float b0 = fx - mj.x, b1 = fy - mj.y, b2 = fz - mj.z;
                            const float dj = fmaf(b0, b0, fmaf(b1, b1, b2 * b2));
                            rank += (dj < di || (dj == di && j < i)) ? 1 : 0;
                        }
                        if (rank < g.k) exact += fast_ex2(fmaf(di, g.q2, mi.w));
                    }
                    exact += __shfl_xor_sync(0xffffffffu, exact, 1);
                    exact += __shfl_xor_sync(0xffffffffu, exact, 2);
                    if (crowded) sum = exact;
                }
            }
            const float row = (g.m0 + 0.6931471805599453f * fast_lg2(g.base + sum)) * y.w;
            if (sub == 0 && l < g.nt) acc += (double)row;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const float probe = R[0] + R[4] + R[8];
        return probe == probe ? g.beta * (float)acc : probe;  // NaN quaternion -> NaN, as the dense sum gives
    }

    // Gradient with respect to the quaternion (registration.py:261-293) on the same lists.  With e = t_l - (R x_k + shift) =
    // h R (f - m) and x_k = org + h m (m: list entry, f: the row in cell units) the 3 x 3 matrix of cpd_row_grad is
    //   G = sum_l sum_k c_lk e x_k^T = h R (V org^T + h W),  V = sum c (f - m),  W = sum c (f - m) m^T,
    // c_lk = w_l gamma_lk / sigma^2 with the responsibilities gamma_lk = term_k / (outlier + sum of the row's terms): 12
    // accumulators per lane.  Candidates beyond the cutoff carry responsibilities below float32 resolution (upstream clips
    // them at e^-20; their share of G is ~1e-7 relative).  The quotient-rule Jacobian of R(q) is cpd_grad_finish's.
    template <class Vec, class Out> __device__ __forceinline__ void grad(Vec q, Out out) const {
        const int lane = threadIdx.x & 31;
        const int grp = lane >> 2, sub = lane & 3;
        float R[9];
        quat_to_matrix<float>(q[0], q[1], q[2], q[3], R);
        float M[9], c[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) M[3 * i + jj] = R[3 * jj + i] * g.inv_h;
            c[i] = -fmaf(M[3 * i], g.shift[0], fmaf(M[3 * i + 1], g.shift[1], fmaf(M[3 * i + 2], g.shift[2], g.org[i] * g.inv_h)));
        }
        const float hx = (float)g.n[0] - 1.f, hy = (float)g.n[1] - 1.f, hz = (float)g.n[2] - 1.f;
        float V[3] = {0.f, 0.f, 0.f}, W[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int l0 = 0; l0 < g.nt; l0 += 8) {
            const int l = l0 + grp;
            const float4 y = s.tgt4[l < g.nt ? l : g.nt - 1];
            const float fx = fmaf(M[0], y.x, fmaf(M[1], y.y, fmaf(M[2], y.z, c[0])));
            const float fy = fmaf(M[3], y.x, fmaf(M[4], y.y, fmaf(M[5], y.z, c[1])));
            const float fz = fmaf(M[6], y.x, fmaf(M[7], y.y, fmaf(M[8], y.z, c[2])));
            const int cx = (int)fminf(fmaxf(fx, 0.f), hx), cy = (int)fminf(fmaxf(fy, 0.f), hy), cz = (int)fminf(fmaxf(fz, 0.f), hz);
            const uint32_t cw = s.cell[(cz * g.n[1] + cy) * g.n[0] + cx];
            const int cnt = (int)(cw >> CPD_CELL_COUNT_SHIFT);
            const float4* e = s.entry + (cw & ((1u << CPD_CELL_COUNT_SHIFT) - 1u));
            float sum = 0.f, v[3] = {0.f, 0.f, 0.f}, w[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            int inside = 0;
            for (int k = sub; k < cnt; k += 4) {
                const float4 m = e[k];
                const float e0 = fx - m.x, e1 = fy - m.y, e2 = fz - m.z;
                const float d2 = fmaf(e0, e0, fmaf(e1, e1, e2 * e2));
                if (KNN) inside += d2 < g.rc2 ? 1 : 0;
                const float tk = fast_ex2(fmaf(d2, g.q2, m.w));
                sum += tk;
                const float a0 = tk * e0, a1 = tk * e1, a2 = tk * e2;
                v[0] += a0; v[1] += a1; v[2] += a2;
                w[0] = fmaf(a0, m.x, w[0]); w[1] = fmaf(a0, m.y, w[1]); w[2] = fmaf(a0, m.z, w[2]);
                w[3] = fmaf(a1, m.x, w[3]); w[4] = fmaf(a1, m.y, w[4]); w[5] = fmaf(a1, m.z, w[5]);
                w[6] = fmaf(a2, m.x, w[6]); w[7] = fmaf(a2, m.y, w[7]); w[8] = fmaf(a2, m.z, w[8]);
            }
            if (KNN) {  // more than k candidates inside the cutoff: keep exactly the k nearest, as operator() does
                inside += __shfl_xor_sync(0xffffffffu, inside, 1);
                inside += __shfl_xor_sync(0xffffffffu, inside, 2);
                const bool crowded = inside > g.k;
                if (__any_sync(0xffffffffu, crowded)) {
                    const int n = crowded ? cnt : 0;
                    if (crowded) { sum = 0.f; v[0] = v[1] = v[2] = 0.f; for (int i = 0; i < 9; ++i) w[i] = 0.f; }
                    for (int i = sub; i < n; i += 4) {
                        const float4 mi = e[i];
                        const float e0 = fx - mi.x, e1 = fy - mi.y, e2 = fz - mi.z;
                        const float di = fmaf(e0, e0, fmaf(e1, e1, e2 * e2));
                        int rank = 0;
                        for (int j = 0; j < n; ++j) {
                            const float4 mj = e[j];
                            const float b0 = fx - mj.x, b1 = fy - mj.y, b2 = fz - mj.z;
                            const float dj = fmaf(b0, b0, fmaf(b1, b1, b2 * b2));
                            rank += (dj < di || (dj == di && j < i)) ? 1 : 0;
                        }
                        if (rank < g.k) {
                            const float tk = fast_ex2(fmaf(di, g.q2, mi.w));
                            sum += tk;
                            const float a0 = tk * e0, a1 = tk * e1, a2 = tk * e2;
                            v[0] += a0; v[1] += a1; v[2] += a2;
                            w[0] = fmaf(a0, mi.x, w[0]); w[1] = fmaf(a0, mi.y, w[1]); w[2] = fmaf(a0, mi.z, w[2]);
                            w[3] = fmaf(a1, mi.x, w[3]); w[4] = fmaf(a1, mi.y, w[4]); w[5] = fmaf(a1, mi.z, w[5]);
                            w[6] = fmaf(a2, mi.x, w[6]); w[7] = fmaf(a2, mi.y, w[7]); w[8] = fmaf(a2, mi.z, w[8]);
                        }
                    }
                }
            }
            float tot = sum;
            tot += __shfl_xor_sync(0xffffffffu, tot, 1);
            tot += __shfl_xor_sync(0xffffffffu, tot, 2);
            const float sc = l < g.nt ? y.w / (g.base + tot) : 0.f;  // w_l / (outlier + sum): this lane's terms become responsibilities
#pragma unroll
            for (int i = 0; i < 3; ++i) V[i] = fmaf(sc, v[i], V[i]);
#pragma unroll
            for (int i = 0; i < 9; ++i) W[i] = fmaf(sc, w[i], W[i]);
        }
        // Q = V org^T + h W, reduced over the warp in double like the dense kernel, then G = (h / sigma^2) R Q
        const float h = 1.0f / g.inv_h;
        const float inv_s2 = -2.0f * g.q2 * g.inv_h * g.inv_h * 0.6931471805599453f;  // q2 = -0.5 log2(e) h^2 / sigma^2
        float Q[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
                double t = (double)fmaf(V[i], g.org[jj], h * W[3 * i + jj]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                Q[3 * i + jj] = (float)t;
            }
        float G[9];
        const float k0 = h * inv_s2;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj)
                G[3 * i + jj] = k0 * fmaf(R[3 * i], Q[jj], fmaf(R[3 * i + 1], Q[3 + jj], R[3 * i + 2] * Q[6 + jj]));
        CpdP<float> pb{};
        pb.beta = g.beta;
        cpd_grad_finish<float>(pb, q[0], q[1], q[2], q[3], G, out);
    }
};

template <int KIND, bool KNN>
__global__ void __launch_bounds__(CPD_CELL_WARPS * 32, 2) k_sampler_cpd_cell(CpdCellGrid g, RunArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    CpdWarpLogPCell<KNN> f;
    f.g = g;
    f.s = cpd_cell_stage(g, smem_raw);
    const int64_t c = (int64_t)blockIdx.x * CPD_CELL_WARPS + warp;
    if (c >= a.N) return;
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
    float lp_x = nanf("");  // log p of the current state, carried across transitions (same function, same point)
    for (int64_t t = 0; t < a.n_steps; ++t) {
        dr.begin_step(a.step0 + (uint64_t)t);
        float lp;
        int r = transition_literal<float, KIND>(4, x, u, y, dr, f, &lp, &lp_x);
        rej_total += (unsigned)r;
        rej_window += (unsigned)r;
        if (++since_keep == a.thin) {
            const int64_t k = kept++;
            since_keep = 0;
            if (lane == 0) {
                if (samples) *reinterpret_cast<float4*>(samples + (k * a.ss_step + c * a.ss_chain)) = make_float4(xs[0], xs[1], xs[2], xs[3]);
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

// persistent over points: the lists are staged once per block, then each warp walks its share of the points
template <bool KNN>
__global__ void __launch_bounds__(CPD_CELL_WARPS * 32, 2) k_logp_cpd_cell(CpdCellGrid g, const float* __restrict__ q,
                                                                       float* __restrict__ out, int64_t N) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    CpdWarpLogPCell<KNN> f;
    f.g = g;
    f.s = cpd_cell_stage(g, smem_raw);
    for (int64_t i = (int64_t)blockIdx.x * CPD_CELL_WARPS + warp; i < N; i += (int64_t)gridDim.x * CPD_CELL_WARPS) {
        SVec<const float> v{q + i * 4, 1};
        const float val = f(v);
        if ((threadIdx.x & 31) == 0) out[i] = val;
    }
}


// Baseline samplers on the cutoff lists (random-walk MH, independence, their mixture and spherical HMC: mcmc.py:115-332): the
// loop of k_mh_cpd with the cell-list log-density and gradient above.
// GRAD = false (the gradient-free kinds): the gradient code is left out and two blocks fit an SM.
template <bool KNN, bool GRAD> struct CpdCellMhTarget {
    CpdWarpLogPCell<KNN> f;
    template <class Vec> __device__ __forceinline__ float logp(Vec q) const { return f(q); }
    template <class Vec, class Out> __device__ __forceinline__ void grad(Vec q, Out o) const {
        if constexpr (GRAD) f.grad(q, o);
    }
};

template <bool KNN, bool GRAD>
__global__ void __launch_bounds__(CPD_CELL_WARPS * 32, GRAD ? 1 : 2) k_mh_cpd_cell(CpdCellGrid g, int kind, RunArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    CpdCellMhTarget<KNN, GRAD> f;
    f.f.g = g;
    f.f.s = cpd_cell_stage(g, smem_raw);
    const int64_t c = (int64_t)blockIdx.x * CPD_CELL_WARPS + warp;
    if (c >= a.N) return;
    float xs[4], vs[4], xps[4], vps[4], gs[4];
    Reg4<float> x{xs}, v{vs}, xp{xps}, vp{vps}, gr{gs};
    float* st = static_cast<float*>(a.state) + c * 4;
    float* mom = static_cast<float*>(a.momentum) + c * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) { xs[j] = st[j]; vs[j] = mom[j]; }
    PhiloxSeq<float> dr;
    dr.init(a.seed, a.chain0 + (uint64_t)c);
    MhChain ch{a.stepsize[c], a.adapt_count[c], a.n_walk[c]};
    const MhParams mp{a.alpha, a.n_leap, a.adapt_total};
    unsigned long long acc_total = a.n_reject[c];
    unsigned int acc_window = 0;
    float lp = f.logp(x);
    float* samples = static_cast<float*>(a.samples);
    float* lpo = static_cast<float*>(a.logp);
    float* momenta = static_cast<float*>(a.momenta_out);
    int64_t since_keep = 0, kept = 0;
    for (int64_t t = 0; t < a.n_steps; ++t) {
        dr.begin_step(a.step0 + (uint64_t)t);
        const bool ok = mh_transition<float>(kind, 4, x, v, xp, vp, gr, dr, f, ch, mp, lp);
        acc_total += ok ? 1u : 0u;
        acc_window += ok ? 1u : 0u;
        if (++since_keep == a.thin) {
            const int64_t k = kept++;
            since_keep = 0;
            if (lane == 0) {
                if (samples) *reinterpret_cast<float4*>(samples + (k * a.ss_step + c * a.ss_chain)) = make_float4(xs[0], xs[1], xs[2], xs[3]);
                if (momenta) *reinterpret_cast<float4*>(momenta + (k * a.ss_step + c * a.ss_chain)) = make_float4(vs[0], vs[1], vs[2], vs[3]);
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
        if (a.status) a.status[c] = 0;
    }
}

// batched gradient, persistent over points like k_logp_cpd_cell
template <bool KNN>
__global__ void __launch_bounds__(CPD_CELL_WARPS * 32, 2) k_grad_cpd_cell(CpdCellGrid g, const float* __restrict__ q,
                                                                       float* __restrict__ out, int64_t N) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    CpdWarpLogPCell<KNN> f;
    f.g = g;
    f.s = cpd_cell_stage(g, smem_raw);
    for (int64_t i = (int64_t)blockIdx.x * CPD_CELL_WARPS + warp; i < N; i += (int64_t)gridDim.x * CPD_CELL_WARPS) {
        SVec<const float> v{q + i * 4, 1};
        float gq[4];
        Reg4<float> o{gq};
        f.grad(v, o);
        if ((threadIdx.x & 31) == 0) *reinterpret_cast<float4*>(out + i * 4) = make_float4(gq[0], gq[1], gq[2], gq[3]);
    }
}

}  // namespace geosss
