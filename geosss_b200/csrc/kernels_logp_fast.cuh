// Standalone batched log_prob at the HBM roof (float32): vMF mixture, Bingham(-Fisher), curved vMF.
//
//   geosss/distributions.py:78-86,113-114,156-157,218-221,272-275 evaluated for rows x of an (N, d) array
//   (the only vectorised use upstream: scripts/mixture_vMF.py:104, scripts/curve_vMF.py:119-120).
//
// The operator reads (d+1)*4 bytes per point and does O(Kd) flops, so it is bandwidth-bound.
// Persistent blocks walk over tiles of TILE_ROWS consecutive rows.  A tile is one contiguous
// span of TILE_ROWS*d floats, so it is fetched with ONE bulk asynchronous copy (TMA unit,
// cp.async.bulk global->shared, completion on an mbarrier) into a two-stage ring: tile i+1 is
// in flight while the threads evaluate tile i from shared memory, lane per row.  Target
// parameters (means, A, knots) are staged into shared memory once per block.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "kernels_fast.cuh"
#include "kernels_gc.cuh"

namespace geosss {

constexpr int LP_ROWS = 128;  // rows per tile = threads per block (64 for wide rows, so that several blocks fit an SM)
constexpr int LP_STAGES = 2;
constexpr int LP_MAX_K = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();  // a lost copy must not hang the GPU
    }
}

struct LpParams {
    const float* a0;  // vMF: mu (K,d) | Bingham: A or diag | curved: knots (M,d)
    const float* a1;  // vMF: consts (K) natural log | Bingham: b or null | curved: cos(ang)
    const float* a2;  // curved: sin(ang)
    float kappa;
    int d, K, M, diag, single;
};

// ---- policies: stage parameters, evaluate one row held in shared memory -------------------------
// vMF mixture with compile-time K and dimension DT (the parameter block keeps rows of D = 4 ceil(DT/4) floats):
// means and constants sit in the kernel-parameter constant bank (same block as the sampler kernel), so the
// K*d FMAs per point take them as immediate operands and the only shared-memory traffic is the point itself,
// read with the widest loads the row alignment allows (16 bytes for d % 4 == 0, 8 bytes for even d: with a row
// stride of d floats and lane-per-row access, 8-byte loads of even-d rows are conflict-free per half-warp where
// scalar loads were 2-way conflicted).
template <int K, int DT> struct VmfLpConst {
    using Params = VmfFastParams;
    static constexpr int D = (DT + 3) / 4 * 4;
    static __host__ __device__ size_t smem_floats(const Params&) { return 0; }
    static __host__ __device__ int dim(const Params&) { return DT; }
    __device__ void stage(const Params&, float*) {}
    __device__ __forceinline__ float eval(const Params& P, const float* row) const {
        float x[DT];
        if (DT % 4 == 0) {
#pragma unroll
            for (int j = 0; j + 3 < DT; j += 4) {
                const float4 v = *reinterpret_cast<const float4*>(row + j);
                x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
            }
        } else if (DT % 2 == 0) {
#pragma unroll
            for (int j = 0; j + 1 < DT; j += 2) {
                const float2 v = *reinterpret_cast<const float2*>(row + j);
                x[j] = v.x; x[j + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < DT; ++j) x[j] = row[j];
        }
        float s[K];
#pragma unroll
        for (int k = 0; k < K; ++k) s[k] = P.c[k];
#pragma unroll
        for (int j = 0; j < DT; ++j) {
#pragma unroll
            for (int k = 0; k < K; ++k) s[k] = fmaf(P.mu[k * D + j], x[j], s[k]);
        }
        float m = s[0];
#pragma unroll
        for (int k = 1; k < K; ++k) m = fmaxf(m, s[k]);
        if (!(m > -INFINITY)) return m;  // i0 overflow: every component is -inf, like the reference
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) acc += fast_ex2(s[k] - m);
        return (m + fast_lg2(acc)) * 0.6931471805599453f;
    }
};

// K = 1..8 x d = 2..16 instantiations of k_logp_tiled<VmfLpConst> live in their own translation unit
// (logp_vmf_const.cu).  Returns 0, a negative GEOSSS_E_* code, or a positive cudaError_t.
int launch_logp_vmf_const(int d, int K, const VmfFastParams& P, const float* x, float* out, int64_t N, int sm_count, cudaStream_t st);

// Rows are read with 16-byte shared loads whenever d is a multiple of 4.  With lane-per-row access the row
// stride decides the bank conflicts: a stride of 4 * odd floats is conflict-free for 16-byte loads, so rows
// whose d is a multiple of 8 are landed with a padded stride of d + 4 (one bulk copy per row instead of one
// per tile); other d keep the contiguous tile (stride d: at most 2-way conflicts for scalar loads).
__host__ __device__ inline int lp_row_stride(int d) { return (d & 7) == 0 ? d + 4 : d; }

struct VmfLp {
    using Params = LpParams;
    static __host__ __device__ int dim(const Params& p) { return p.d; }
    const float* muT;  // shared (d, KP): component-minor, times log2(e); KP = K rounded up to 4
    const float* c;    // shared (KP), times log2(e); padding components sit at -inf
    int d, K;
    static __host__ __device__ size_t smem_floats(const LpParams& p) { return (size_t)((p.K + 3) & ~3) * (p.d + 1); }
    __device__ void stage(const LpParams& p, float* sm) {
        d = p.d; K = p.K;
        const int KP = (K + 3) & ~3;
        constexpr float L2E = 1.4426950408889634f;
        for (int i = threadIdx.x; i < KP * d; i += blockDim.x) {
            const int j = i / KP, k = i - j * KP;
            sm[i] = k < K ? p.a0[k * d + j] * L2E : 0.f;
        }
        for (int i = threadIdx.x; i < KP; i += blockDim.x) sm[KP * d + i] = i < K ? p.a1[i] * L2E : -INFINITY;
        muT = sm; c = sm + KP * d;
    }
    __device__ __forceinline__ void add(float (&s)[LP_MAX_K], int j, float xj, int nq) const {
        const float4* col = reinterpret_cast<const float4*>(muT + j * (nq << 2));
#pragma unroll
        for (int q = 0; q < LP_MAX_K / 4; ++q) if (q < nq) {
            const float4 m = col[q];  // one broadcast load feeds four components
            s[4 * q] = fmaf(m.x, xj, s[4 * q]);         s[4 * q + 1] = fmaf(m.y, xj, s[4 * q + 1]);
            s[4 * q + 2] = fmaf(m.z, xj, s[4 * q + 2]); s[4 * q + 3] = fmaf(m.w, xj, s[4 * q + 3]);
        }
    }
    __device__ __forceinline__ float eval(const LpParams&, const float* row) const {
        const int nq = (K + 3) >> 2;
        float s[LP_MAX_K];
#pragma unroll
        for (int k = 0; k < LP_MAX_K; ++k) s[k] = 0.f;
        if ((d & 3) == 0) {
            for (int j = 0; j < d; j += 4) {
                const float4 xv = *reinterpret_cast<const float4*>(row + j);
                add(s, j, xv.x, nq); add(s, j + 1, xv.y, nq); add(s, j + 2, xv.z, nq); add(s, j + 3, xv.w, nq);
            }
        } else {
            for (int j = 0; j < d; ++j) add(s, j, row[j], nq);
        }
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < LP_MAX_K; ++k) if (k < 4 * nq) { s[k] += c[k]; m = fmaxf(m, s[k]); }
        if (!(m > -INFINITY)) return m;  // i0 overflow: every component is -inf, like the reference
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < LP_MAX_K; ++k) if (k < 4 * nq) acc += fast_ex2(s[k] - m);
        return (m + fast_lg2(acc)) * 0.6931471805599453f;
    }
};

struct BinghamLp {
    using Params = LpParams;
    static __host__ __device__ int dim(const Params& p) { return p.d; }
    BinghamGc g;
    static __host__ __device__ size_t smem_floats(const LpParams& p) {
        return (size_t)(p.diag ? p.d : p.d * ((p.d + 3) & ~3)) + (p.a1 ? p.d : 0) + 4;
    }
    __device__ void stage(const LpParams& p, float* sm) {
        GcParams q{};
        q.a0 = p.a0; q.a1 = p.a1; q.d = p.d; q.diag = p.diag;
        g.stage(q, sm);
    }
    __device__ __forceinline__ float eval(const LpParams&, const float* row) const {
        const int d = g.d;
        const bool vec = (d & 3) == 0;
        float q = 0.f;
        if (g.diag) {
            if (vec) {
                for (int j = 0; j < d; j += 4) {
                    const float4 xv = *reinterpret_cast<const float4*>(row + j), a = *reinterpret_cast<const float4*>(g.A + j);
                    q = fmaf(a.x * xv.x, xv.x, q); q = fmaf(a.y * xv.y, xv.y, q);
                    q = fmaf(a.z * xv.z, xv.z, q); q = fmaf(a.w * xv.w, xv.w, q);
                }
            } else {
                for (int j = 0; j < d; ++j) q = fmaf(g.A[j] * row[j], row[j], q);
            }
        } else {
            for (int i = 0; i < g.dp; i += 4) {  // four rows of A x at a time (A symmetric: column block = row block)
                float ax[4] = {0.f, 0.f, 0.f, 0.f};
                auto col = [&](int j, float xj) {
                    const float4 a4 = *reinterpret_cast<const float4*>(g.A + j * g.dp + i);
                    ax[0] = fmaf(a4.x, xj, ax[0]); ax[1] = fmaf(a4.y, xj, ax[1]);
                    ax[2] = fmaf(a4.z, xj, ax[2]); ax[3] = fmaf(a4.w, xj, ax[3]);
                };
                if (vec) {
                    for (int j = 0; j < d; j += 4) {
                        const float4 xv = *reinterpret_cast<const float4*>(row + j);
                        col(j, xv.x); col(j + 1, xv.y); col(j + 2, xv.z); col(j + 3, xv.w);
                    }
                } else {
                    for (int j = 0; j < d; ++j) col(j, row[j]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) if (i + k < d) q = fmaf(row[i + k], ax[k], q);
            }
        }
        if (g.b) for (int j = 0; j < d; ++j) q = fmaf(g.b[j], row[j], q);
        return q;
    }
};

struct CurvedLp {
    using Params = LpParams;
    static __host__ __device__ int dim(const Params& p) { return p.d; }
    CurvedGc g;
    static __host__ __device__ size_t smem_floats(const LpParams& p) { return (size_t)((p.M + 3) & ~3) * p.d + 3 * (size_t)(p.M - 1) + 4; }
    __device__ void stage(const LpParams& p, float* sm) {
        GcParams q{};
        q.a0 = p.a0; q.a1 = p.a1; q.a2 = p.a2; q.d = p.d; q.M = p.M; q.kappa = p.kappa;
        g.stage(q, sm);
    }
    template <int NQ> __device__ __forceinline__ void scores(CurvedGc::Coef& co, const float* row) const {
        if ((g.d & 3) == 0) {
            for (int j = 0; j < g.d; j += 4) {
                const float4 xv = *reinterpret_cast<const float4*>(row + j);
                g.accumulate_x_n<NQ>(co, j, xv.x); g.accumulate_x_n<NQ>(co, j + 1, xv.y);
                g.accumulate_x_n<NQ>(co, j + 2, xv.z); g.accumulate_x_n<NQ>(co, j + 3, xv.w);
            }
        } else {
            for (int j = 0; j < g.d; ++j) g.accumulate_x_n<NQ>(co, j, row[j]);
        }
    }
    __device__ __forceinline__ float eval(const LpParams&, const float* row) const {
        CurvedGc::Coef co;
        g.begin(co);
        switch ((g.M + 3) >> 2) {  // warp-uniform: knot quads
            case 1: scores<1>(co, row); break;
            case 2: scores<2>(co, row); break;
            case 3: scores<3>(co, row); break;
            default: scores<4>(co, row); break;
        }
        return g.kappa * g.best_dot(co, 1.f, 0.f);
    }
};

template <class Policy, int ROWS = LP_ROWS>
__global__ void __launch_bounds__(ROWS) k_logp_tiled(const __grid_constant__ typename Policy::Params P, const float* __restrict__ x,
                                                     float* __restrict__ out, int64_t N) {
    extern __shared__ __align__(128) unsigned char lp_smem_raw[];
    float* lp_smem = reinterpret_cast<float*>(lp_smem_raw);
    __shared__ __align__(8) uint64_t full[LP_STAGES];
    const int d = Policy::dim(P);
    const int ds = lp_row_stride(d);         // row stride in shared memory
    const bool per_row = ds != d;            // padded rows: every thread copies its own row
    const size_t tile_floats = (size_t)ROWS * d;
    float* stage_buf[LP_STAGES];
#pragma unroll
    for (int s = 0; s < LP_STAGES; ++s) stage_buf[s] = lp_smem + s * (size_t)ROWS * ds;
    Policy pol;
    pol.stage(P, lp_smem + LP_STAGES * (size_t)ROWS * ds);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < LP_STAGES; ++s) mbar_init(&full[s], per_row ? ROWS : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t n_tiles = (N + ROWS - 1) / ROWS;
    const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    // a tile can use the bulk engine when it is complete (byte count a multiple of 16) and x is 16-byte aligned
    auto bulk_ok = [&](int64_t tile) { return aligned && (tile + 1) * ROWS <= N; };
    auto issue = [&](int64_t tile, int s) {
        if (!(tile < n_tiles && bulk_ok(tile))) return;
        if (per_row) {  // one arrival + d * 4 bytes per thread; the barrier counts ROWS arrivals
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&full[s], (uint32_t)(d * sizeof(float)));
            bulk_g2s(stage_buf[s] + (size_t)threadIdx.x * ds, x + tile * tile_floats + (size_t)threadIdx.x * d, (uint32_t)(d * sizeof(float)), &full[s]);
        } else if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&full[s], (uint32_t)(tile_floats * sizeof(float)));
            bulk_g2s(stage_buf[s], x + tile * tile_floats, (uint32_t)(tile_floats * sizeof(float)), &full[s]);
        }
    };

    uint32_t phase[LP_STAGES] = {0, 0};
    int64_t tile = blockIdx.x;
    issue(tile, 0);
    for (int it = 0; tile < n_tiles; ++it, tile += gridDim.x) {
        const int s = it & 1;
        issue(tile + gridDim.x, s ^ 1);  // prefetch the next tile of this block
        const int64_t row0 = tile * ROWS;
        if (bulk_ok(tile)) {
            mbar_wait(&full[s], phase[s]);
            phase[s] ^= 1;
        } else {  // ragged last tile / unaligned input: plain coalesced loads
            const int64_t n_here = (N - row0 < ROWS ? N - row0 : ROWS) * d;
            for (int64_t i = threadIdx.x; i < n_here; i += blockDim.x) stage_buf[s][(i / d) * ds + i % d] = x[row0 * d + i];
            __syncthreads();
        }
        const int64_t r = row0 + threadIdx.x;
        if (r < N) out[r] = pol.eval(P, stage_buf[s] + (size_t)threadIdx.x * ds);
        __syncthreads();  // everyone is done with stage s before it is refilled two iterations later
    }
}

}  // namespace geosss
