// Dense Bingham log_prob on the 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32.
//
//   log p(x) = x^T A x (+ b.x), geosss/distributions.py:78-86,113-114, for rows x of an (N, d) array.
//
// This is the one operator of the path that sits above the FP32 ridge: 2 d^2 flops for (d+1)*4
// bytes (d = 50: 25 flop/B against ~9 flop/B of FP32 FMA peak per HBM byte), SURVEY.md section 8d.
// Y = X A is a genuine GEMM (M = rows, N = K = d), so it goes to the tensor cores; the row-wise
// dot products q_i = Y_i . x_i are the epilogue.
//
// Per 128-row tile:
//   * the tile (one contiguous span of 128*d floats) is fetched by ONE bulk asynchronous copy (TMA); the copy
//     of the next tile is issued as soon as this one has been split, so it is in flight during the MMAs and the
//     epilogue (two blocks per SM interleave the rest); each thread splits its own row into
//     TF32 "hi" (cvt.rna) and "lo" = x - hi parts and writes them in the canonical K-major,
//     no-swizzle UMMA operand layout (8x16-byte core matrices);
//   * one thread issues 3 x ceil(d/8) tcgen05.mma.kind::tf32 (M=128, N=64, K=8) into a 128x64 fp32
//     accumulator in TMEM:  Xh Ah + Xh Al + Xl Ah  (A is split the same way once per block;
//     A symmetric, so its row-major form IS the K-major B operand);
//   * tcgen05.commit -> mbarrier; every thread then pulls its accumulator row out of TMEM with
//     tcgen05.ld and finishes q = Y_row . x_row (+ b.x) in fp32 with x_row = hi + lo read back from the operands.
// Error of 3xTF32 is ~2^-21 relative per product (the dropped lo*lo term), i.e. fp32-like;
// tests hold it to 2e-5 * max(1, |log p|) like the other fp32 kernels.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "kernels_logp_fast.cuh"  // smem_u32, mbarrier helpers

namespace geosss {

constexpr int TC_ROWS = 128;     // UMMA M
constexpr int TC_THREADS = 256;  // two threads per row: thread (row, half) splits K steps [4 half, 4 half + 4) and reduces the same 32 columns
constexpr int TC_N = 64;      // UMMA N: d padded to 64 columns of A
constexpr int TC_MAX_D = 64;

struct TcParams {
    const float* A;  // (d, d) row-major, symmetric
    const float* b;  // (d) or null
    int d;
};

__device__ __forceinline__ uint64_t umma_smem_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start address, leading / stride byte offsets (all >> 4), version 1, no swizzle
    uint64_t desc = 0;
    desc |= (uint64_t)((smem_u32(smem) & 0x3FFFFu) >> 4);
    desc |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    desc |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    desc |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return desc;
}

__device__ __forceinline__ float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// Canonical K-major / no-swizzle operand block for `rows` rows and one 8-wide K step (32 bytes per row):
// element (r, k) lives at  (k / 4) * lbo + (r / 8) * 128 + (r % 8) * 16 + (k % 4) * 4  bytes, lbo = rows * 16.
__device__ __forceinline__ size_t umma_kstep_floats(int rows) { return (size_t)rows * 8; }

static __global__ void __launch_bounds__(TC_THREADS, 2) k_logp_bingham_tc(TcParams P, const float* __restrict__ x, float* __restrict__ out, int64_t N) {
    extern __shared__ __align__(128) unsigned char tc_smem_raw[];
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ __align__(8) uint64_t full;
    __shared__ uint32_t tmem_base_s;
    const int d = P.d;
    const int ksteps = (d + 7) >> 3;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row_t = tid & (TC_ROWS - 1), half = tid >> 7;  // this thread's row of the tile, and which half of the work it takes
    __shared__ float q_part[TC_THREADS];

    const size_t raw_floats = ((size_t)TC_ROWS * d + 31) & ~(size_t)31;
    float* raw = reinterpret_cast<float*>(tc_smem_raw);                  // (128, d) landing buffer of the bulk copy
    float* Xh = raw + raw_floats;                                        // ksteps blocks of 128 x 8
    float* Xl = Xh + ksteps * umma_kstep_floats(TC_ROWS);
    float* Ah = Xl + ksteps * umma_kstep_floats(TC_ROWS);                // ksteps blocks of 64 x 8
    float* Al = Ah + ksteps * umma_kstep_floats(TC_N);
    float* bs = Al + ksteps * umma_kstep_floats(TC_N);                   // (64) Fisher vector or zeros

    // ---- once per block: TMEM, barrier, the split A operand
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(&mma_done, 1);
        mbar_init(&full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < ksteps * TC_N * 8; i += TC_THREADS) {  // (kk, chunk, n / 8, n % 8, k % 4) <- A[n][8 kk + 4 chunk + k % 4]
        const int kq = i & 3, r8 = (i >> 2) & 7, g = (i >> 5) % (TC_N / 8), chunk = (i / (32 * (TC_N / 8))) & 1, kk = i / (32 * (TC_N / 8) * 2);
        const int n = g * 8 + r8, k = kk * 8 + chunk * 4 + kq;
        const float v = (n < d && k < d) ? P.A[(size_t)n * d + k] : 0.f;
        const float hi = to_tf32(v);
        Ah[i] = hi;
        Al[i] = to_tf32(v - hi);
    }
    if (tid < TC_N) bs[tid] = (P.b && tid < d) ? P.b[tid] : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = tmem_base_s;

    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
    const int64_t n_tiles = (N + TC_ROWS - 1) / TC_ROWS;
    uint32_t parity = 0, full_phase = 0;
    const size_t tile_floats = (size_t)TC_ROWS * d;
    const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    auto bulk_ok = [&](int64_t t) { return aligned && t < n_tiles && (t + 1) * TC_ROWS <= N; };
    auto issue = [&](int64_t t) {  // thread 0 only, after a block barrier: nobody reads `raw` any more
        if (bulk_ok(t)) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&full, (uint32_t)(tile_floats * sizeof(float)));
            bulk_g2s(raw, x + t * tile_floats, (uint32_t)(tile_floats * sizeof(float)), &full);
        }
    };

    if (tid == 0) issue(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * TC_ROWS;
        if (bulk_ok(tile)) {
            mbar_wait(&full, full_phase);
            full_phase ^= 1;
        } else {  // ragged last tile / unaligned input
            const int64_t n_here = (N - row0 < TC_ROWS ? N - row0 : TC_ROWS) * d;
            for (int64_t i = tid; i < (int64_t)tile_floats; i += TC_THREADS) raw[i] = i < n_here ? x[row0 * d + i] : 0.f;
            __syncthreads();
        }
        // split this thread's half of its row into the operand layout: one 16-byte store per (k step, chunk).  Thread
        // (row, half) takes k steps [4 half, 4 half + 4) = columns [32 half, 32 half + 32) -- the columns of Y it reduces
        // in the epilogue -- and keeps the values in registers (xv) for that dot product; even d reads the row with
        // 8-byte loads (row stride d floats: conflict-free per half-warp, scalar loads were 2-way conflicted).
        float xv[32];
        {
            const float* row = raw + (size_t)row_t * d;
            const size_t rbase = (size_t)(row_t >> 3) * 32 + (size_t)(row_t & 7) * 4;  // floats
            const bool even = (d & 1) == 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int kk = 4 * half + i;
#pragma unroll
                for (int chunk = 0; chunk < 2; ++chunk) {
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    const int k0 = kk * 8 + chunk * 4;
                    if (kk < ksteps) {
                        if (even) {
                            if (k0 < d) { const float2 a = *reinterpret_cast<const float2*>(row + k0); v[0] = a.x; v[1] = a.y; }
                            if (k0 + 2 < d) { const float2 b2 = *reinterpret_cast<const float2*>(row + k0 + 2); v[2] = b2.x; v[3] = b2.y; }
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) if (k0 + q < d) v[q] = row[k0 + q];
                        }
                        float h[4], l[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) { h[q] = to_tf32(v[q]); l[q] = to_tf32(v[q] - h[q]); }
                        const size_t off = kk * umma_kstep_floats(TC_ROWS) + (size_t)chunk * (TC_ROWS * 4) + rbase;
                        *reinterpret_cast<float4*>(Xh + off) = make_float4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<float4*>(Xl + off) = make_float4(l[0], l[1], l[2], l[3]);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) xv[i * 8 + chunk * 4 + q] = v[q];
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            issue(tile + gridDim.x);  // next tile lands in `raw` while this one is multiplied and reduced
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            bool first = true;
#pragma unroll 1
            for (int pass = 0; pass < 3; ++pass) {
                const float* Xa = pass == 2 ? Xl : Xh;
                const float* Ab = pass == 1 ? Al : Ah;
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint64_t adesc = umma_smem_desc(Xa + kk * umma_kstep_floats(TC_ROWS), TC_ROWS * 16, 128);
                    const uint64_t bdesc = umma_smem_desc(Ab + kk * umma_kstep_floats(TC_N), TC_N * 16, 128);
                    const uint32_t acc = first ? 0u : 1u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem_acc), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
                        : "memory");
                    first = false;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_done)) : "memory");
        }
        mbar_wait(&mma_done, parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // epilogue: thread <-> row <-> TMEM lane; its 32 accumulator columns in two 16-column loads, dotted with the row
        // values kept in registers since the split (`raw` already belongs to the next tile)
        float q = 0.f;
#pragma unroll
        for (int cc = 0; cc < TC_N / 2; cc += 16) {
            const int c0 = half * (TC_N / 2) + cc;  // warps 0-3 reduce columns 0..31, warps 4-7 columns 32..63
            if (c0 < 8 * ksteps) {
                uint32_t r[16];
                const uint32_t taddr = tmem_acc + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j4 = 0; j4 < 16; j4 += 4) {
                    const int n0 = c0 + j4;  // four columns = one (k step, chunk) of the operand layout
                    if (n0 < 8 * ksteps) {
                        const float4 bv = *reinterpret_cast<const float4*>(bs + n0);
                        q = fmaf(__uint_as_float(r[j4]) + bv.x, xv[cc + j4], q);
                        q = fmaf(__uint_as_float(r[j4 + 1]) + bv.y, xv[cc + j4 + 1], q);
                        q = fmaf(__uint_as_float(r[j4 + 2]) + bv.z, xv[cc + j4 + 2], q);
                        q = fmaf(__uint_as_float(r[j4 + 3]) + bv.w, xv[cc + j4 + 3], q);
                    }
                }
            }
        }
        q_part[tid] = q;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // TMEM accumulator and the operand buffers are free for the next tile
        if (half == 0 && row0 + row_t < N) out[row0 + row_t] = q_part[tid] + q_part[tid + TC_ROWS];
    }

    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(64) : "memory");
    }
}

inline size_t tc_smem_bytes(int d) {
    const int ksteps = (d + 7) / 8;
    const size_t raw = ((size_t)TC_ROWS * d + 31) & ~(size_t)31;
    return (raw + 2 * (size_t)ksteps * TC_ROWS * 8 + 2 * (size_t)ksteps * TC_N * 8 + TC_N) * sizeof(float) + 128;
}

}  // namespace geosss
