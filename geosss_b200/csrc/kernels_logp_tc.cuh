// Batched log_prob whose inner contraction S = X W^T runs on the 5th-generation tensor cores
// (tcgen05 + TMEM, 3xTF32), warp-specialised and software-pipelined.
//
//   Bingham(-Fisher)  geosss/distributions.py:78-86,113-114   W = A (d, d)      out = sum_j (S_j + b_j) x_j
//   vMF mixture       distributions.py:156-157,218-221        W = mu (K, d)     out = logsumexp_k(S_k + c_k)
//   curved vMF        distributions.py:272-275,               W = knots (M, d)  out = kappa max over segments of the
//                     spherical_curve.py:10-32,95-102                                  closed-form nearest-point dot
//   (rows x of an (N, d) float32 array; the vectorised use of upstream: scripts/mixture_vMF.py:104,
//   scripts/curve_vMF.py:119-120, scripts/bingham.py)
//
// These are N x d x n_out contractions behind (d+1)*4 bytes per point; on the FP32 pipe the d=100 cases are
// instruction-bound at 0.45-0.67 of the HBM roof (profiles/r1_kernel_table.md).  Here one persistent block per SM
// runs three roles over 128-row tiles:
//   warp 0       producer: cp.async.bulk (TMA) of whole tiles into a ring of R shared-memory stages
//   warp 1       one thread issues 3 ceil(d/8) tcgen05.mma.kind::tf32 (M=128, N=NP, K=8) per tile: Xh Wh + Xh Wl + Xl Wh,
//                A operand (the tile, split in TF32 hi / lo) read from TENSOR MEMORY, B operand (W, split once per
//                block) from shared memory in the canonical K-major layout, fp32 accumulator in tensor memory;
//                tcgen05.commit signals the epilogue
//   warps 2-9    workers, two threads per row: read the row from the landing stage, release the stage, split to
//                hi / lo, tcgen05.st them into the tile's TMEM operand columns (no shared-memory round trip of the
//                operands: the MMA's shared-memory traffic is W only), signal the MMA warp; then the epilogue of the
//                PREVIOUS tile (tcgen05.ld of its accumulator, row-wise reduction, store) while this tile's MMAs run.
// Two TMEM stages (operands + accumulator) and R >= 3 landing stages keep copy, split, MMA and epilogue of
// neighbouring tiles in flight at once; >= 100 KB of loads are outstanding per SM.
// 3xTF32: the dropped lo*lo term is ~2^-22 relative per product, i.e. fp32-like; tests hold 1e-5 max(1, |log p|).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "kernels_logp_fast.cuh"  // smem_u32, mbarrier and bulk-copy helpers
#include "kernels_tc.cuh"         // umma_smem_desc, to_tf32

namespace geosss {

constexpr int TCP_ROWS = 128;                       // UMMA M = TMEM lanes
constexpr int TCP_WORKERS = 8;                      // worker warps (two per TMEM lane quadrant)
constexpr int TCP_THREADS = (TCP_WORKERS + 2) * 32;
constexpr int TCP_MAX_STAGES = 4;
constexpr int TCP_TMEM_COLS = 512;

struct TcpParams {
    const float* W;   // (n_out, d) row-major: A | mu | knots
    const float* v0;  // Bingham: b (d) or null | vMF: constants (K), natural log | curved: cos(ang) (M-1)
    const float* v1;  // curved: sin(ang) (M-1)
    float kappa;
    int d, n_out;
};

__host__ __device__ inline int tcp_row_stride(int d) { return (d & 7) == 0 ? d + 4 : d; }  // as lp_row_stride: conflict-free 16-byte row loads

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 8 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- epilogue policies ---------------------------------------------------------------------------
// NP: accumulator columns (n_out padded; UMMA N, a multiple of 16 for M = 128).  NEEDS_X: the epilogue dots the
// accumulator row with the point itself (kept in registers from the split), both threads of a row take part.
struct TcBinghamEpi {
    static constexpr int NP = 64;
    static constexpr bool NEEDS_X = true;
    static constexpr int AUX = 64;
    static __device__ void stage(const TcpParams& P, float* aux) {
        for (int i = threadIdx.x; i < AUX; i += blockDim.x) aux[i] = (P.v0 && i < P.d) ? P.v0[i] : 0.f;
    }
    static __device__ __forceinline__ float scale() { return 1.f; }
};

struct TcVmfEpi {
    static constexpr int NP = 16;
    static constexpr bool NEEDS_X = false;
    static constexpr int AUX = 16;
    static __device__ void stage(const TcpParams& P, float* aux) {  // constants in log2 units, -inf for the padding
        for (int i = threadIdx.x; i < AUX; i += blockDim.x) aux[i] = i < P.n_out ? P.v0[i] * 1.4426950408889634f : -INFINITY;
    }
    static __device__ __forceinline__ float scale() { return 1.4426950408889634f; }  // W = mu log2(e)
    static __device__ __forceinline__ float finish(const TcpParams&, const float* aux, const float (&s)[16]) {
        float e[16], m = -INFINITY;
#pragma unroll
        for (int k = 0; k < 16; ++k) { e[k] = s[k] + aux[k]; m = fmaxf(m, e[k]); }
        if (!(m > -INFINITY)) return m;  // i0 overflow: every component is -inf, like the reference
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += fast_ex2(e[k] - m);
        return (m + fast_lg2(acc)) * 0.6931471805599453f;
    }
};

struct TcCurvedEpi {
    static constexpr int NP = 16;
    static constexpr bool NEEDS_X = false;
    static constexpr int AUX = 48;  // cos(ang), 1 / (sin(ang) + 1e-10), sin(ang): 15 each
    static __device__ void stage(const TcpParams& P, float* aux) {
        for (int i = threadIdx.x; i < 16; i += blockDim.x) {
            const bool in = i < P.n_out - 1;
            aux[i] = in ? P.v0[i] : 0.f;
            aux[16 + i] = in ? 1.f / (P.v1[i] + 1e-10f) : 0.f;
            aux[32 + i] = in ? P.v1[i] : 0.f;
        }
    }
    static __device__ __forceinline__ float scale() { return 1.f; }
    // max over segments of x . nearest point, first segment wins ties: CurvedGcT::best_dot (kernels_gc.cuh) at theta = 0
    static __device__ __forceinline__ float finish(const TcpParams& P, const float* aux, const float (&g)[16]) {
        float ga = g[0], best = 0.f, best_clipped = -INFINITY;
#pragma unroll
        for (int m = 1; m < 16; ++m) if (m < P.n_out) {
            const float gb = g[m];
            const float p2 = (gb - ga * aux[m - 1]) * aux[16 + m - 1];
            const bool inside = p2 >= 0.f && fmaf(ga, aux[32 + m - 1], -p2 * aux[m - 1]) >= 0.f;
            const float dot = inside ? fast_sqrt(fmaf(ga, ga, p2 * p2)) : (p2 < 0.f ? ga : gb);
            const float dc = fminf(dot, 1.f);
            if (dc > best_clipped) { best_clipped = dc; best = dot; }
            ga = gb;
        }
        return P.kappa * best;
    }
};

// 8 values of a row starting at column k0 (k0 % 8 == 0), zero beyond d; widest loads the row stride allows
__device__ __forceinline__ void tcp_load8(const float* __restrict__ row, int k0, int d, int ds, float (&v)[8]) {
    if (k0 + 8 <= d) {
        if ((ds & 3) == 0) {
            const float4 a = *reinterpret_cast<const float4*>(row + k0), b = *reinterpret_cast<const float4*>(row + k0 + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else if ((ds & 1) == 0) {
#pragma unroll
            for (int q = 0; q < 8; q += 2) { const float2 a = *reinterpret_cast<const float2*>(row + k0 + q); v[q] = a.x; v[q + 1] = a.y; }
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = row[k0 + q];
        }
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = k0 + q < d ? row[k0 + q] : 0.f;
    }
}

template <class Epi>
__global__ void __launch_bounds__(TCP_THREADS, 1) k_logp_tc(TcpParams P, const float* __restrict__ x, float* __restrict__ out,
                                                             int64_t n_tiles, int R) {
    extern __shared__ __align__(128) unsigned char tcp_smem[];
    __shared__ __align__(8) uint64_t raw_full[TCP_MAX_STAGES], raw_empty[TCP_MAX_STAGES], op_full[2], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float part[TCP_ROWS];
    constexpr int NP = Epi::NP;
    const int d = P.d, ksteps = (d + 7) >> 3, KP = ksteps << 3;
    const int ds = tcp_row_stride(d);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t stage_floats = ((size_t)TCP_ROWS * ds + 31) & ~(size_t)31;
    float* raw = reinterpret_cast<float*>(tcp_smem);
    float* Wh = raw + (size_t)R * stage_floats;           // ksteps blocks of NP x 8, canonical K-major / no swizzle
    float* Wl = Wh + (size_t)ksteps * NP * 8;
    float* aux = Wl + (size_t)ksteps * NP * 8;

    // ---- once per block: tensor memory, barriers, the split W operand, epilogue constants
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TCP_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < TCP_MAX_STAGES; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], TCP_WORKERS); }
        for (int i = 0; i < 2; ++i) { mbar_init(&op_full[i], TCP_WORKERS); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], TCP_WORKERS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float sc = Epi::scale();
        constexpr int G = NP / 8;
        for (int i = tid; i < ksteps * NP * 8; i += TCP_THREADS) {  // (kk, chunk, n / 8, n % 8, k % 4) <- W[n][8 kk + 4 chunk + k % 4]
            const int kq = i & 3, r8 = (i >> 2) & 7, g = (i >> 5) % G, chunk = (i / (32 * G)) & 1, kk = i / (64 * G);
            const int n = g * 8 + r8, k = kk * 8 + chunk * 4 + kq;
            const float v = (n < P.n_out && k < d) ? P.W[(size_t)n * d + k] * sc : 0.f;
            const float hi = to_tf32(v);
            Wh[i] = hi;
            Wl[i] = to_tf32(v - hi);
        }
    }
    Epi::stage(P, aux);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // W written by the generic proxy, read by the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const int stage_cols = 2 * KP + NP;  // per TMEM stage: Xh | Xl | accumulator
    const int n_it = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const size_t tile_floats = (size_t)TCP_ROWS * d;

    if (warp == 0) {
        // ===== producer
        for (int it = 0; it < n_it; ++it) {
            const int r = it % R, use = it / R;
            if (use > 0) mbar_wait(&raw_empty[r], (uint32_t)(use - 1) & 1u);
            const float* src = x + ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * (int64_t)tile_floats;
            float* dst = raw + (size_t)r * stage_floats;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (lane == 0) mbar_expect_tx(&raw_full[r], (uint32_t)(tile_floats * sizeof(float)));
            if (ds == d) {
                if (lane == 0) bulk_g2s(dst, src, (uint32_t)(tile_floats * sizeof(float)), &raw_full[r]);
            } else {  // padded rows: one copy per row
                for (int row = lane; row < TCP_ROWS; row += 32)
                    bulk_g2s(dst + (size_t)row * ds, src + (size_t)row * d, (uint32_t)(d * sizeof(float)), &raw_full[r]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===== MMA issuer
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(TCP_ROWS >> 4) << 24);
        for (int it = 0; it < n_it; ++it) {
            const int s = it & 1, use = it >> 1;
            mbar_wait(&op_full[s], (uint32_t)use & 1u);
            if (use > 0) mbar_wait(&acc_empty[s], (uint32_t)(use - 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t xh = tmem + (uint32_t)(s * stage_cols), xl = xh + (uint32_t)KP, acc = xl + (uint32_t)KP;
                bool first = true;
#pragma unroll 1
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t xa = pass == 2 ? xl : xh;
                    const float* Wb = pass == 1 ? Wl : Wh;
                    for (int kk = 0; kk < ksteps; ++kk) {
                        const uint64_t bdesc = umma_smem_desc(Wb + (size_t)kk * NP * 8, NP * 16, 128);
                        const uint32_t accum = first ? 0u : 1u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                            ::"r"(acc), "r"(xa + (uint32_t)(kk << 3)), "l"(bdesc), "r"(idesc), "r"(accum)
                            : "memory");
                        first = false;
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&acc_full[s])) : "memory");
            }
            __syncwarp();
        }
    } else {
        // ===== workers: thread (row, half)
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
        const int khalf = (ksteps + 1) >> 1;
        const int k_lo = half ? khalf : 0, k_hi = half ? ksteps : khalf;
        float xv[Epi::NEEDS_X ? 32 : 1], xp[Epi::NEEDS_X ? 32 : 1];  // this / the previous tile's row values (Bingham: k_hi - k_lo <= 4)

        auto epilogue = [&](int j) {
            const int s = j & 1, use = j >> 1;
            const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)j * gridDim.x) * TCP_ROWS;
            const uint32_t acc = lane_base + (uint32_t)(s * stage_cols + 2 * KP);
            mbar_wait(&acc_full[s], (uint32_t)use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if constexpr (Epi::NEEDS_X) {
                float y[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i) if (k_lo + i < k_hi) tmem_ld8(acc + (uint32_t)((k_lo + i) << 3), y[i]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float q = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) if (k_lo + i < k_hi) {
                    const float* b = aux + ((k_lo + i) << 3);
#pragma unroll
                    for (int c = 0; c < 8; ++c) q = fmaf(y[i][c] + b[c], xp[i * 8 + c], q);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[s]);
                if (half) part[row] = q;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");  // the two warps of this lane quadrant
                if (!half) out[row0 + row] = q + part[row];
                asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");  // part[] may be rewritten
            } else {
                if (!half) {
                    float sc[16];
                    float lo8[8], hi8[8];
                    tmem_ld8(acc, lo8);
                    tmem_ld8(acc + 8u, hi8);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 8; ++c) { sc[c] = lo8[c]; sc[8 + c] = hi8[c]; }
                    out[row0 + row] = Epi::finish(P, aux, sc);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[s]);
            }
        };

        for (int it = 0; it < n_it; ++it) {
            const int s = it & 1, r = it % R;
            mbar_wait(&raw_full[r], (uint32_t)(it / R) & 1u);
            const float* rowp = raw + (size_t)r * stage_floats + (size_t)row * ds;
            const uint32_t xh = lane_base + (uint32_t)(s * stage_cols), xl = xh + (uint32_t)KP;
            // the TMEM operand columns of stage s are free: this thread has seen acc_full[s] of tile it-2 (its MMAs are done)
            if constexpr (Epi::NEEDS_X) {
#pragma unroll
                for (int i = 0; i < 4; ++i) if (k_lo + i < k_hi) {
                    float v[8], h[8], l[8];
                    tcp_load8(rowp, (k_lo + i) << 3, d, ds, v);
#pragma unroll
                    for (int c = 0; c < 8; ++c) { h[c] = to_tf32(v[c]); l[c] = to_tf32(v[c] - h[c]); xv[i * 8 + c] = v[c]; }
                    tmem_st8(xh + (uint32_t)((k_lo + i) << 3), h);
                    tmem_st8(xl + (uint32_t)((k_lo + i) << 3), l);
                }
            } else {
                for (int kk = k_lo; kk < k_hi; ++kk) {
                    float v[8], h[8], l[8];
                    tcp_load8(rowp, kk << 3, d, ds, v);
#pragma unroll
                    for (int c = 0; c < 8; ++c) { h[c] = to_tf32(v[c]); l[c] = to_tf32(v[c] - h[c]); }
                    tmem_st8(xh + (uint32_t)(kk << 3), h);
                    tmem_st8(xl + (uint32_t)(kk << 3), l);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_empty[r]);  // the landing stage may be refilled
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&op_full[s]);
            if (it > 0) epilogue(it - 1);
            if constexpr (Epi::NEEDS_X) {
#pragma unroll
                for (int c = 0; c < 32; ++c) xp[c] = xv[c];
            }
        }
        if (n_it > 0) epilogue(n_it - 1);
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCP_TMEM_COLS) : "memory");
    }
}

// shared memory of one block with R landing stages; R: as many as fit next to the operands (at most TCP_MAX_STAGES)
template <class Epi> inline size_t tcp_smem_bytes(int d, int R) {
    const int ksteps = (d + 7) / 8;
    const size_t stage = ((size_t)TCP_ROWS * tcp_row_stride(d) + 31) & ~(size_t)31;
    return ((size_t)R * stage + 2 * (size_t)ksteps * Epi::NP * 8 + Epi::AUX) * sizeof(float) + 128;
}
template <class Epi> inline int tcp_stages(int d) {
    int R = TCP_MAX_STAGES;
    while (R > 2 && tcp_smem_bytes<Epi>(d, R) > 220 * 1024) --R;
    return tcp_smem_bytes<Epi>(d, R) <= 220 * 1024 ? R : 0;
}
// the two TMEM stages (hi, lo operand columns + accumulator each) must fit the 512 columns
template <class Epi> inline bool tcp_fits(int d) { return 2 * (2 * ((d + 7) / 8 * 8) + Epi::NP) <= TCP_TMEM_COLS && tcp_stages<Epi>(d) >= 2; }

}  // namespace geosss
