// Batched log_prob whose inner contraction S = X W^T runs on the 5th-generation tensor cores
// (tcgen05 + TMEM, 3xTF32), warp-specialised and software-pipelined.
//
//   Bingham(-Fisher)  geosss/distributions.py:78-86,113-114   W = eigvecs (+ b) out = sum_j v_j S_j^2 (+ S_d)
//   vMF mixture       distributions.py:156-157,218-221        W = mu (K, d)     out = logsumexp_k(S_k + c_k)
//   curved vMF        distributions.py:272-275,               W = knots (M, d)  out = kappa max over segments of the
//                     spherical_curve.py:10-32,95-102                                  closed-form nearest-point dot
//   (rows x of an (N, d) float32 array; the vectorised use of upstream: scripts/mixture_vMF.py:104,
//   scripts/curve_vMF.py:119-120, scripts/bingham.py)
//
// These are N x d x n_out contractions behind (d+1)*4 bytes per point; on the FP32 pipe the d=100 cases are
// instruction-bound at 0.45-0.67 of the HBM roof (profiles/r1_kernel_table.md).  Dense Bingham is evaluated in the
// eigenbasis of A = U^T diag(v) U (the Bingham object carries it upstream too, distributions.py:67-76; here from a Jacobi
// sweep at target creation): x^T A x = sum_j v_j (u_j . x)^2, so that, like the other two, its epilogue needs the
// accumulator row only (the Fisher vector b rides along as one more row of W).
// One persistent block per SM, 18 warps in four roles over 128-row tiles; tiles alternate between two TMEM stages
// (hi / lo operand columns + accumulator columns each), every stage has its own split and epilogue warps:
//   warp 0        producer: one cp.async.bulk (TMA) per tile into a ring of R shared-memory landing stages
//   warp 1        one thread issues 3 ceil(d/8) tcgen05.mma.kind::tf32 (M=128, N=NP, K=8) per tile: Xh Wh + Xh Wl + Xl Wh,
//                 A operand (the tile, split in TF32 hi / lo) read from TENSOR MEMORY, B operand (W, split once per
//                 block) from shared memory in the canonical K-major layout, fp32 accumulator in tensor memory;
//                 tcgen05.commit signals the stage's epilogue warps (and frees its operand columns)
//   warps 2-9     split warps, four per stage, thread per row: read the row from the landing stage, split to hi / lo,
//                 tcgen05.st into the stage's operand columns (no shared-memory round trip of the operands: the MMA's
//                 shared-memory traffic is W only), release the landing stage, signal the MMA warp
//   warps 10-17   epilogue warps, four per stage, thread per row: tcgen05.ld of the accumulator row, release the
//                 accumulator, row-wise reduction, store
// so that copy, split, MMA and epilogue of neighbouring tiles are all in flight at once, >= 75 KB of loads outstanding
// per SM.  3xTF32: the dropped lo*lo term is ~2^-22 relative per product, i.e. fp32-like; tests hold 1e-5 max(1, |log p|).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "kernels_logp_fast.cuh"  // smem_u32, mbarrier and bulk-copy helpers
#include "kernels_tc.cuh"         // umma_smem_desc, to_tf32

namespace geosss {

constexpr int TCP_ROWS = 128;                       // UMMA M = TMEM lanes
constexpr int TCP_ROLE_WARPS = 4;                   // producer, MMA issuer of the even tiles, of the odd tiles, one idle
constexpr int TCP_SPLIT_GROUPS = 2;                 // split groups of 4 warps, tiles dealt round-robin (3 groups measured no faster: r2_logp_tc_v9.md)
constexpr int TCP_SPLIT_WARPS = 4 * TCP_SPLIT_GROUPS;
constexpr int TCP_THREADS = (TCP_ROLE_WARPS + TCP_SPLIT_WARPS + 8) * 32;  // + 2 x 4 epilogue warps
constexpr int TCP_MAX_STAGES = 4;
constexpr int TCP_TMEM_COLS = 512;
constexpr int TCP_TMEM_STAGES = 4;                  // at most; as many (hi | lo | accumulator) column groups as fit the 512 columns

struct TcpParams {
    const float* W;   // (n_out, d) row-major: eigenvectors of A as rows (+ b as row d) | mu | knots
    const float* v0;  // Bingham: eigenvalues (d) | vMF: constants (K), natural log | curved: cos(ang) (M-1)
    const float* v1;  // curved: sin(ang) (M-1)
    float kappa;
    int d, n_out;
    int np;           // accumulator columns per half: n_out rounded up to 16 (UMMA N is a multiple of 16 at M = 128)
    int spin;         // 1: barrier waits poll with mbarrier.test_wait, 0 (default): mbarrier.try_wait
    float* rows_out;  // TcRotateEpi: (N, d) output rows (may be the input array: a tile is consumed before it is written)
    int sub;          // 128-row MMA groups per tile (0 / 1 .. 4): narrow rows take 256-512-row tiles, one hand-off per tile
};

// Waits of the role warps: mbarrier.try_wait (may suspend the warp for a time slice) or, with GEOSSS_TC_SPIN=1, polling with
// mbarrier.test_wait.  Measured on B200 (r2 sweep): polling is 2-7 % SLOWER on every shape, so try_wait is the default.
__device__ __forceinline__ void tcp_wait(uint64_t* bar, uint32_t parity, int spin) {
    if (!(spin & 1)) { mbar_wait(bar, parity); return; }
    uint32_t done = 0;
    for (uint32_t n = 0; !done; ++n) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (n > (1u << 28)) __trap();  // a lost hand-off must not hang the GPU
    }
}

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

// wider shapes of the same accesses: a tcgen05.st / ld costs the issuing warp its latency whatever the width (phase timers:
// ~100 cycles per x8 store pair in the split loop), so the split stores 32 columns and the epilogue loads 16 per instruction
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31])) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- epilogue policies ---------------------------------------------------------------------------
// NP: accumulator columns (n_out padded; UMMA N, a multiple of 16 for M = 128).
// Every policy consumes the accumulator row in chunks of 16 columns (begin / consume / result), so that the
// epilogue holds 32 accumulator registers at a time whatever NP is.
struct TcBinghamEpi {
    static constexpr int NP_MAX = 64;
    static constexpr int AUX = 64;
    struct State { float q0, q1, q2, q3; };
    static __device__ void stage(const TcpParams& P, float* aux) {  // eigenvalues; the Fisher row (n_out == d + 1) and the padding weigh 0
        for (int i = threadIdx.x; i < AUX; i += blockDim.x) aux[i] = i < P.d ? P.v0[i] : 0.f;
    }
    static __device__ __forceinline__ float scale() { return 1.f; }
    static constexpr bool SCALAR_OUT = true;
    static constexpr bool ROUND_SPLIT = false;
    static constexpr bool ONE_ACC = true;  // wide accumulator: all three products land on the same NP columns
    static __device__ __forceinline__ void begin(State& st, int64_t) { st.q0 = st.q1 = st.q2 = st.q3 = 0.f; }
    // sum_j v_j s_j^2 (+ s_d = b . x)
    static __device__ __forceinline__ void consume(State& st, const TcpParams& P, const float* aux, int c0, const float (&s)[16]) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 v = *reinterpret_cast<const float4*>(aux + c0 + j);
            st.q0 = fmaf(v.x * s[j], s[j], st.q0); st.q1 = fmaf(v.y * s[j + 1], s[j + 1], st.q1);
            st.q2 = fmaf(v.z * s[j + 2], s[j + 2], st.q2); st.q3 = fmaf(v.w * s[j + 3], s[j + 3], st.q3);
        }
        if (P.n_out > P.d && (P.d >> 4) == (c0 >> 4)) {  // Fisher term: accumulator column d (warp-uniform branch)
#pragma unroll
            for (int j = 0; j < 16; ++j) if (c0 + j == P.d) st.q0 += s[j];
        }
    }
    static __device__ __forceinline__ float result(const State& st, const TcpParams&) { return (st.q0 + st.q1) + (st.q2 + st.q3); }
};

struct TcVmfEpi {
    static constexpr int NP_MAX = 16;
    static constexpr int AUX = 16;
    struct State { float v; };
    static __device__ void stage(const TcpParams& P, float* aux) {  // constants in log2 units, -inf for the padding
        for (int i = threadIdx.x; i < AUX; i += blockDim.x) aux[i] = i < P.n_out ? P.v0[i] * 1.4426950408889634f : -INFINITY;
    }
    static __device__ __forceinline__ float scale() { return 1.4426950408889634f; }  // W = mu log2(e)
    static constexpr bool SCALAR_OUT = true;
    static constexpr bool ROUND_SPLIT = false;
    static constexpr bool ONE_ACC = false;
    static __device__ __forceinline__ void begin(State& st, int64_t) { st.v = 0.f; }
    static __device__ __forceinline__ void consume(State& st, const TcpParams& P, const float* aux, int, const float (&s)[16]) {
        if (P.n_out <= 8) { lse<8>(st, aux, s); } else { lse<16>(st, aux, s); }  // uniform branch: half the exponentials for K <= 8
    }
    template <int KN>
    static __device__ __forceinline__ void lse(State& st, const float* aux, const float (&s)[16]) {
        float e[KN], m = -INFINITY;
#pragma unroll
        for (int k = 0; k < KN; ++k) { e[k] = s[k] + aux[k]; m = fmaxf(m, e[k]); }
        if (!(m > -INFINITY)) { st.v = m; return; }  // i0 overflow: every component is -inf, like the reference
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < KN; ++k) acc += fast_ex2(e[k] - m);
        st.v = (m + fast_lg2(acc)) * 0.6931471805599453f;
    }
    static __device__ __forceinline__ float result(const State& st, const TcpParams&) { return st.v; }
};

struct TcCurvedEpi {
    static constexpr int NP_MAX = 16;
    static constexpr int AUX = 48;  // cos(ang), 1 / (sin(ang) + 1e-10), sin(ang): 15 each
    struct State { float v; };
    static __device__ void stage(const TcpParams& P, float* aux) {
        for (int i = threadIdx.x; i < 16; i += blockDim.x) {
            const bool in = i < P.n_out - 1;
            aux[i] = in ? P.v0[i] : 0.f;
            aux[16 + i] = in ? 1.f / (P.v1[i] + 1e-10f) : 0.f;
            aux[32 + i] = in ? P.v1[i] : 0.f;
        }
    }
    static __device__ __forceinline__ float scale() { return 1.f; }
    static constexpr bool SCALAR_OUT = true;
    static constexpr bool ROUND_SPLIT = true;   // the segment maximum amplifies the split error (sqrt of a difference of squares): rounded hi
    static constexpr bool ONE_ACC = false;
    static __device__ __forceinline__ void begin(State& st, int64_t) { st.v = 0.f; }
    // max over segments of x . nearest point, first segment wins ties: CurvedGcT::best_dot (kernels_gc.cuh) at theta = 0
    static __device__ __forceinline__ void consume(State& st, const TcpParams& P, const float* aux, int, const float (&g)[16]) {
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
        st.v = P.kappa * best;
    }
    static __device__ __forceinline__ float result(const State& st, const TcpParams&) { return st.v; }
};

// Rotation of rows, Y = X W^T with all d outputs kept (dense Bingham chains live in the eigenbasis of A: states and retained
// samples are rotated at the C-ABI boundary, capi.cu).  Every thread stores its own row's columns; rows are contiguous, so a
// warp fills a 32 d float span.
struct TcRotateEpi {
    static constexpr int NP_MAX = 64;
    static constexpr int AUX = 4;
    static constexpr bool SCALAR_OUT = false;
    static constexpr bool ROUND_SPLIT = true;
    static constexpr bool ONE_ACC = true;
    struct State { float* o; int64_t row; };
    static __device__ void stage(const TcpParams&, float*) {}
    static __device__ __forceinline__ float scale() { return 1.f; }
    static __device__ __forceinline__ void begin(State& st, int64_t row) { st.o = nullptr; st.row = row; }
    static __device__ __forceinline__ void consume(State& st, const TcpParams& P, const float*, int c0, const float (&s)[16]) {
        float* o = P.rows_out + st.row * P.d + c0;
        if ((P.d & 1) == 0) {  // rows start on 8-byte boundaries
#pragma unroll
            for (int j = 0; j < 16; j += 2) if (c0 + j < P.d) *reinterpret_cast<float2*>(o + j) = make_float2(s[j], s[j + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (c0 + j < P.d) o[j] = s[j];
        }
    }
    static __device__ __forceinline__ float result(const State&, const TcpParams&) { return 0.f; }
};

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// TF32 hi / lo split of the row values (tcp_split_block below).  The tensor core reads an fp32 word as TF32 by dropping its low
// 13 mantissa bits (measured: the 3xTF32 error stays at 1e-6 with this split, r2 mode sweep), so hi is the row value as it is
// and lo = x - trunc(x) (exact in fp32, one AND + one FADD per element).  lo loses its own low bits the same way: an error of
// 2^-10 |lo| < 2^-20 |x| per element, inside the 1e-5 the fp32 log_prob states.

// N (8, 16 or 32) row values from column k0 on, split and stored with one tcgen05.st per half.  ROT (d % 16 == 0, N >= 16): the
// 16-byte pieces of each group of 16 columns are read in the order (p + lane) mod 4 and rotated back in registers.
// LW: floats per shared-memory load the row stride allows (4, 2, 1); CHECK: columns beyond d read as zero (the last block only).
// TRUNC = false (the rotation of rows, whose output is sample coordinates): hi is rounded to TF32 first -- bit pattern + half an
// ulp, low 13 bits cleared -- one more integer instruction per element for half the split error.
template <int N, bool ROT, int LW = 4, bool CHECK = false, bool TRUNC = true>
__device__ __forceinline__ void tcp_split_block(const float* __restrict__ rowp, int k0, int d, int rl, uint32_t xh, uint32_t xl) {
    float v[N], l[N];
    if constexpr (ROT) {
#pragma unroll
        for (int g = 0; g < N; g += 16) {
            const float4* base = reinterpret_cast<const float4*>(rowp + k0 + g);
            float4 t0 = base[rl], t1 = base[(rl + 1) & 3], t2 = base[(rl + 2) & 3], t3 = base[(rl + 3) & 3];
            // t_i = piece (i + rl) mod 4; rotate right by rl: stage 1 by (rl & 1), stage 2 by (rl & 2)
            const bool b0 = rl & 1, b1 = rl & 2;
            float4 u0 = b0 ? t3 : t0, u1 = b0 ? t0 : t1, u2 = b0 ? t1 : t2, u3 = b0 ? t2 : t3;
            float4 p0 = b1 ? u2 : u0, p1 = b1 ? u3 : u1, p2 = b1 ? u0 : u2, p3 = b1 ? u1 : u3;
            v[g + 0] = p0.x; v[g + 1] = p0.y; v[g + 2] = p0.z; v[g + 3] = p0.w; v[g + 4] = p1.x; v[g + 5] = p1.y; v[g + 6] = p1.z; v[g + 7] = p1.w;
            v[g + 8] = p2.x; v[g + 9] = p2.y; v[g + 10] = p2.z; v[g + 11] = p2.w; v[g + 12] = p3.x; v[g + 13] = p3.y; v[g + 14] = p3.z; v[g + 15] = p3.w;
        }
    } else {
        const float* src = rowp + k0;
        if constexpr (CHECK) {
#pragma unroll
            for (int q = 0; q < N; ++q) v[q] = k0 + q < d ? src[q] : 0.f;
        } else if constexpr (LW == 4) {
#pragma unroll
            for (int q = 0; q < N; q += 4) { const float4 a = *reinterpret_cast<const float4*>(src + q); v[q] = a.x; v[q + 1] = a.y; v[q + 2] = a.z; v[q + 3] = a.w; }
        } else if constexpr (LW == 2) {
#pragma unroll
            for (int q = 0; q < N; q += 2) { const float2 a = *reinterpret_cast<const float2*>(src + q); v[q] = a.x; v[q + 1] = a.y; }
        } else {
#pragma unroll
            for (int q = 0; q < N; ++q) v[q] = src[q];
        }
    }
    // hi = the value as it is (the tensor core drops the low 13 bits), lo = x - trunc(x)
    if constexpr (TRUNC) {
#pragma unroll
        for (int c = 0; c < N; ++c) l[c] = v[c] - __uint_as_float(__float_as_uint(v[c]) & 0xFFFFE000u);
    } else {
#pragma unroll
        for (int c = 0; c < N; ++c) {
            const float h = __uint_as_float((__float_as_uint(v[c]) + 0x1000u) & 0xFFFFE000u);
            l[c] = v[c] - h;
            v[c] = h;
        }
    }
    if constexpr (N == 32) { tmem_st32(xh + (uint32_t)k0, v); tmem_st32(xl + (uint32_t)k0, l); }
    else if constexpr (N == 16) { tmem_st16(xh + (uint32_t)k0, v); tmem_st16(xl + (uint32_t)k0, l); }
    else { tmem_st8(xh + (uint32_t)k0, *reinterpret_cast<const float(*)[8]>(v)); tmem_st8(xl + (uint32_t)k0, *reinterpret_cast<const float(*)[8]>(l)); }
}

// one row of a tile: whole blocks of 32 / 16 / 8 columns without bound checks, then the ragged last K step
template <int LW, bool TRUNC>
__device__ __forceinline__ void tcp_split_row(const float* __restrict__ rowp, int d, int KP, uint32_t xh, uint32_t xl) {
    int k0 = 0;
    for (; k0 + 32 <= d; k0 += 32) tcp_split_block<32, false, LW, false, TRUNC>(rowp, k0, d, 0, xh, xl);
    if (k0 + 16 <= d) { tcp_split_block<16, false, LW, false, TRUNC>(rowp, k0, d, 0, xh, xl); k0 += 16; }
    if (k0 + 8 <= d) { tcp_split_block<8, false, LW, false, TRUNC>(rowp, k0, d, 0, xh, xl); k0 += 8; }
    if (k0 < KP) tcp_split_block<8, false, 1, true, TRUNC>(rowp, k0, d, 0, xh, xl);
}

#ifdef TCP_TIMERS
__device__ unsigned long long tcp_timers[32];
#define TCP_T0() long long t_prev = clock64(), t_acc[6] = {0, 0, 0, 0, 0, 0}
#define TCP_TICK(i) { const long long t_now = clock64(); t_acc[i] += t_now - t_prev; t_prev = t_now; }
#define TCP_DUMP(base, n) if (blockIdx.x == 0 && lane == 0) for (int q = 0; q < n; ++q) atomicAdd(&tcp_timers[base + q], (unsigned long long)t_acc[q])
#else
#define TCP_T0()
#define TCP_TICK(i)
#define TCP_DUMP(base, n)
#endif

template <class Epi>
__global__ void __launch_bounds__(TCP_THREADS, 1) k_logp_tc(TcpParams P, const float* __restrict__ x, float* __restrict__ out,
                                                             int64_t n_tiles, int R) {
    extern __shared__ __align__(128) unsigned char tcp_smem[];
    // Every barrier has ONE waiting role instance and one arriving one: the first index is the parity of the WAITER's tile
    // (even tiles: issuer warp 1, split / epilogue group 0; odd tiles: warp 2, group 1).  With an odd ring depth the users of a
    // stage alternate between the two instances, and a parity wait by the instance that runs ahead would take the previous
    // phase for its own (mbarrier waits only know the parity); per-parity barriers keep each one's phases in a single order.
    __shared__ __align__(8) uint64_t raw_full[TCP_SPLIT_GROUPS][TCP_MAX_STAGES], raw_empty[TCP_MAX_STAGES], op_full[2][TCP_TMEM_STAGES],
        op_empty[TCP_SPLIT_GROUPS][TCP_TMEM_STAGES], acc_full[2][TCP_TMEM_STAGES], acc_empty[2][TCP_TMEM_STAGES];
    __shared__ uint32_t tmem_base_s;
    const int NP = P.np, NS = 2 * NP;  // W is stacked: rows 0..NP-1 of a K block hold Wh, rows NP..2NP-1 hold Wl
    const int d = P.d, ksteps = (d + 7) >> 3, KP = ksteps << 3;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int SUBT = P.sub < 1 ? 1 : (P.sub > 4 ? 4 : P.sub), ROWS = TCP_ROWS * SUBT;
    const size_t stage_floats = ((size_t)ROWS * d + 31) & ~(size_t)31;
    float* raw = reinterpret_cast<float*>(tcp_smem);
    float* Ws = raw + (size_t)R * stage_floats;           // ksteps blocks of NS x 8, canonical K-major / no swizzle
    float* aux = Ws + (size_t)ksteps * NS * 8;

    // ---- once per block: tensor memory, barriers, the split W operand, epilogue constants
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TCP_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < TCP_MAX_STAGES; ++i) {
            for (int q = 0; q < TCP_SPLIT_GROUPS; ++q) mbar_init(&raw_full[q][i], 1);
            mbar_init(&raw_empty[i], 4);
        }
        for (int q = 0; q < TCP_SPLIT_GROUPS; ++q)
            for (int i = 0; i < TCP_TMEM_STAGES; ++i) mbar_init(&op_empty[q][i], 1);
        for (int q = 0; q < 2; ++q)
            for (int i = 0; i < TCP_TMEM_STAGES; ++i) {
                mbar_init(&op_full[q][i], 4); mbar_init(&acc_full[q][i], 1); mbar_init(&acc_empty[q][i], 4);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float sc = Epi::scale();
        const int G = NS / 8;
        for (int i = tid; i < ksteps * NS * 8; i += TCP_THREADS) {  // (kk, chunk, n / 8, n % 8, k % 4) <- W[n][8 kk + 4 chunk + k % 4]
            const int kq = i & 3, r8 = (i >> 2) & 7, g = (i >> 5) % G, chunk = (i / (32 * G)) & 1, kk = i / (64 * G);
            const int ns = g * 8 + r8, n = ns < NP ? ns : ns - NP, k = kk * 8 + chunk * 4 + kq;
            const float v = (n < P.n_out && k < d) ? P.W[(size_t)n * d + k] * sc : 0.f;
            const float hi = to_tf32(v);
            Ws[i] = ns < NP ? hi : to_tf32(v - hi);
        }
    }
    Epi::stage(P, aux);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // W written by the generic proxy, read by the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    constexpr bool one_acc = Epi::ONE_ACC;
    // Tensor memory: a ring of SA operand stages (Xh | Xl, 2 KP columns each) followed by a ring of SD accumulator stages
    // (ACCW columns each); tile `it` uses operand stage it mod SA and accumulator stage it mod SD.  Decoupled, so that the
    // split of a later tile overlaps the MMAs of this one even where only two whole (operand + accumulator) groups would fit.
    const int ACCW = one_acc ? NP : NS;
    const int OPW = SUBT * 2 * KP, ACW = SUBT * ACCW;  // columns of one operand / accumulator stage (all sub-tiles)
    int SA = max(2, min(TCP_TMEM_STAGES, (TCP_TMEM_COLS - 2 * ACW) / OPW));
    int SD = max(2, min(TCP_TMEM_STAGES, (TCP_TMEM_COLS - SA * OPW) / ACW));
    {  // experiment knobs: ring depths from GEOSSS_TC_SPIN bits 8-11 (operands) and 12-15 (accumulators), ignored unless they fit
        const int ka = (P.spin >> 8) & 15, kd = (P.spin >> 12) & 15;
        const int sa2 = ka ? ka : SA, sd2 = kd ? kd : SD;
        if (sa2 >= 2 && sd2 >= 2 && sa2 <= TCP_TMEM_STAGES && sd2 <= TCP_TMEM_STAGES && sa2 * OPW + sd2 * ACW <= TCP_TMEM_COLS) { SA = sa2; SD = sd2; }
    }
    const int acc_col0 = SA * OPW;
    const int n_it = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const size_t tile_floats = (size_t)ROWS * d;
    const int quad = warp & 3;  // the TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);

    if (warp == 0) {
        // ===== producer
        TCP_T0();
        for (int it = 0, r = 0, use = 0; it < n_it; ++it, r = r + 1 == R ? 0 : r + 1, use += r == 0) {
            TCP_TICK(1);
            if (use > 0) tcp_wait(&raw_empty[r], (uint32_t)(use - 1) & 1u, P.spin);
            TCP_TICK(0);
            if (lane == 0) {
                const float* src = x + ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * (int64_t)tile_floats;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&raw_full[it % TCP_SPLIT_GROUPS][r], (uint32_t)(tile_floats * sizeof(float)));
                bulk_g2s(raw + (size_t)r * stage_floats, src, (uint32_t)(tile_floats * sizeof(float)), &raw_full[it % TCP_SPLIT_GROUPS][r]);
            }
            __syncwarp();
        }
        TCP_TICK(1);
        TCP_DUMP(0, 2);
    } else if (warp == 1 || warp == 2) {
        // ===== MMA issuers: warp 1 takes the even tiles, warp 2 the odd ones (phase timers: issuing one tile's MMAs costs the
        // warp 50-70 cycles per tcgen05.mma in uniform-datapath bookkeeping, 600-1100 cycles per tile -- with a single issuer that
        // was the period of a narrow tile; the tiles of the two warps use different TMEM stages, so their MMAs interleave freely).  The whole warp walks the loop (descriptor arithmetic stays in uniform registers: a tcgen05.mma
        // costs ~8-32 tensor-pipe cycles at these shapes, so the issue loop itself must not be a chain of vector->uniform moves)
        // and one elected lane issues.  Two passes per tile: Xh [Wh; Wl] (N = 2 NP) and Xl Wh (N = NP) accumulated onto the
        // first NP columns; the epilogue adds the Xh Wl half.
        const uint32_t idesc_a = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NS >> 3) << 17) | ((uint32_t)(TCP_ROWS >> 4) << 24);
        const uint32_t idesc_b = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(TCP_ROWS >> 4) << 24);
        const bool elected = elect_one();
        const uint64_t wdesc = umma_smem_desc(Ws, NS * 16, 128);
        const uint32_t kdesc = (uint32_t)(NS * 8 * 4) >> 4;  // one K block further, in descriptor address units (16 bytes)
        // Two issuers where the operand ring is at least 3 deep; with 2 stages (d >= 88) the interleaved chains only lengthen
        // the stage cycle (measured -3 % at d = 100), so warp 1 issues every tile there and warp 2 idles.
        const bool two = (P.spin & 48) ? (P.spin & 32) != 0 : SA >= 3;
        const int step = two ? 2 : 1;
        uint32_t ph_op = 0, ph_acc = 0;  // parity this warp expects next, one bit per (tile parity, stage)
        TCP_T0();
        // ring positions by counters, not runtime modulo (depths are at least 2)
        for (int it = two ? warp - 1 : (warp == 1 ? 0 : n_it), sa = it % SA, sd = it % SD; it < n_it;
             it += step, sa = sa + step >= SA ? sa + step - SA : sa + step, sd = sd + step >= SD ? sd + step - SD : sd + step) {
            const int par = it & 1;
            TCP_TICK(2);
            tcp_wait(&op_full[par][sa], (ph_op >> (4 * par + sa)) & 1u, P.spin);
            ph_op ^= 1u << (4 * par + sa);
            TCP_TICK(0);
            if (it >= SD) { tcp_wait(&acc_empty[par][sd], (ph_acc >> (4 * par + sd)) & 1u, P.spin); ph_acc ^= 1u << (4 * par + sd); }
            TCP_TICK(1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elected) {
              for (int sb = 0; sb < SUBT; ++sb) {
                const uint32_t xh = tmem + (uint32_t)(sa * OPW + sb * 2 * KP), xl = xh + (uint32_t)KP, acc = tmem + (uint32_t)(acc_col0 + sd * ACW + sb * ACCW);
                if (one_acc) {  // three N = NP passes onto the same columns: Xh Wh, Xh Wl, Xl Wh
#pragma unroll 4
                    for (int kk = 0; kk < ksteps; ++kk)
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                            ::"r"(acc), "r"(xh + (uint32_t)(kk << 3)), "l"(wdesc + (uint64_t)(kk * kdesc)), "r"(idesc_b), "r"(kk > 0 ? 1u : 0u)
                            : "memory");
#pragma unroll 4
                    for (int kk = 0; kk < ksteps; ++kk)
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                            ::"r"(acc), "r"(xh + (uint32_t)(kk << 3)), "l"(wdesc + (uint64_t)(kk * kdesc + NP)), "r"(idesc_b), "r"(1u)
                            : "memory");
                } else {
#pragma unroll 4
                for (int kk = 0; kk < ksteps; ++kk)
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                        ::"r"(acc), "r"(xh + (uint32_t)(kk << 3)), "l"(wdesc + (uint64_t)(kk * kdesc)), "r"(idesc_a), "r"(kk > 0 ? 1u : 0u)
                        : "memory");
                }
#pragma unroll 4
                for (int kk = 0; kk < ksteps; ++kk)
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                        ::"r"(acc), "r"(xl + (uint32_t)(kk << 3)), "l"(wdesc + (uint64_t)(kk * kdesc)), "r"(idesc_b), "r"(1u)
                        : "memory");
              }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&op_empty[(it + SA) % TCP_SPLIT_GROUPS][sa])) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&acc_full[par][sd])) : "memory");
            }
            __syncwarp();
        }
        TCP_TICK(2);
        if (warp == 1) TCP_DUMP(4, 3);
    } else if (warp == 3) {
        // idle role slot (keeps the split / epilogue warps on whole TMEM lane quadrants: quadrant = warp & 3)
    } else if (warp < TCP_ROLE_WARPS + TCP_SPLIT_WARPS) {
        // ===== split warps of stage g: tiles g, g + 2, ...
        const int g = (warp - TCP_ROLE_WARPS) >> 2;
        // Row stride d floats.  d % 16 == 0: the same 16-byte piece of neighbouring rows falls on the same banks (8-way conflict for
        // d % 32 == 0), so a thread reads the four pieces of a pair of K steps in the order (p + lane) mod 4 and rotates them back
        // in registers (two select stages): 2-way conflicts at worst.  Other d: at most 2-way as they are.
        const bool rot = (d & 15) == 0;
        const int rl = lane & 3;
        constexpr bool TR = !Epi::ROUND_SPLIT;
        uint32_t ph_raw = 0, ph_op = 0;
        TCP_T0();
        for (int it = g, sa = g % SA, r = g % R; it < n_it; it += TCP_SPLIT_GROUPS) {
            TCP_TICK(4);
            tcp_wait(&raw_full[g][r], (ph_raw >> r) & 1u, P.spin);
            ph_raw ^= 1u << r;
            TCP_TICK(0);
            if (it >= SA) {  // the MMAs of the tile that used these operand columns last have completed
                tcp_wait(&op_empty[g][sa], (ph_op >> sa) & 1u, P.spin);
                ph_op ^= 1u << sa;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            TCP_TICK(1);
          for (int sb = 0; sb < SUBT; ++sb) {
            const uint32_t xh = lane_base + (uint32_t)(sa * OPW + sb * 2 * KP), xl = xh + (uint32_t)KP;
            const float* rowp = raw + (size_t)r * stage_floats + (size_t)(sb * TCP_ROWS + row) * d;
            int k0 = 0;
            if (rot) {
                for (; k0 + 32 <= KP; k0 += 32) tcp_split_block<32, true, 4, false, TR>(rowp, k0, d, rl, xh, xl);
                if (k0 < KP) tcp_split_block<16, true, 4, false, TR>(rowp, k0, d, rl, xh, xl);
            } else if ((d & 3) == 0) {
                tcp_split_row<4, TR>(rowp, d, KP, xh, xl);
            } else if ((d & 1) == 0) {
                tcp_split_row<2, TR>(rowp, d, KP, xh, xl);
            } else {
                tcp_split_row<1, TR>(rowp, d, KP, xh, xl);
            }
          }
            __syncwarp();
            TCP_TICK(2);
            if (lane == 0) mbar_arrive(&raw_empty[r]);  // the landing stage may be refilled
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            TCP_TICK(3);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&op_full[it & 1][sa]);
            sa += TCP_SPLIT_GROUPS; while (sa >= SA) sa -= SA;  // ring positions by counters, not runtime modulo
            r += TCP_SPLIT_GROUPS; while (r >= R) r -= R;
        }
        TCP_TICK(4);
        if (warp == TCP_ROLE_WARPS) TCP_DUMP(8, 5);
    } else {
        // ===== epilogue warps of stage e
        const int e = (warp - TCP_ROLE_WARPS - TCP_SPLIT_WARPS) >> 2;
        uint32_t ph_acc = 0;
        TCP_T0();
        for (int it = e, sd = e % SD; it < n_it; it += 2, sd = sd + 2 >= SD ? sd + 2 - SD : sd + 2) {
            const int64_t tile_row0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * ROWS;
            TCP_TICK(3);
            tcp_wait(&acc_full[e][sd], (ph_acc >> sd) & 1u, P.spin);
            ph_acc ^= 1u << sd;
            TCP_TICK(0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int sub_i = 0; sub_i < SUBT; ++sub_i) {
            const uint32_t acc = lane_base + (uint32_t)(acc_col0 + sd * ACW + sub_i * ACCW);
            const int64_t row0 = tile_row0 + sub_i * TCP_ROWS;
            typename Epi::State st;
            Epi::begin(st, row0 + row);
            for (int c0 = 0; c0 < NP; c0 += 16) {
                float sc[16], sb[16];
                tmem_ld16(acc + (uint32_t)c0, sc);
                if (!one_acc) tmem_ld16(acc + (uint32_t)(NP + c0), sb);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                TCP_TICK(1);
                if (c0 + 16 >= NP && sub_i == SUBT - 1) {  // the accumulator rows sit in registers now
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[(it + SD) & 1][sd]);
                }
                if (!one_acc) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) sc[q] += sb[q];
                }
                Epi::consume(st, P, aux, c0, sc);
                TCP_TICK(2);
            }
            if (Epi::SCALAR_OUT) out[row0 + row] = Epi::result(st, P);
          }
        }
        TCP_TICK(3);
        if (warp == TCP_ROLE_WARPS + TCP_SPLIT_WARPS) TCP_DUMP(16, 4);
#ifdef TCP_TIMERS
        if (warp == TCP_ROLE_WARPS + TCP_SPLIT_WARPS && blockIdx.x == 0 && lane == 0) atomicAdd(&tcp_timers[24], (unsigned long long)n_it);
#endif
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCP_TMEM_COLS) : "memory");
    }
}

// shared memory of one block with R landing stages; R: as many as fit next to the operands (at most TCP_MAX_STAGES)
inline int tcp_np(int n_out) { return (n_out + 15) / 16 * 16; }
inline size_t tcp_smem_bytes(int d, int np, int aux, int R, int sub = 1) {
    const int ksteps = (d + 7) / 8;
    const size_t stage = ((size_t)TCP_ROWS * sub * d + 31) & ~(size_t)31;
    return ((size_t)R * stage + 2 * (size_t)ksteps * np * 8 + aux) * sizeof(float) + 128;  // landing ring + stacked W + constants
}
inline int tcp_stages(int d, int np, int aux, int sub = 1) {
    int R = TCP_MAX_STAGES;
    while (R > 2 && tcp_smem_bytes(d, np, aux, R, sub) > 220 * 1024) --R;
    return tcp_smem_bytes(d, np, aux, R, sub) <= 220 * 1024 ? R : 0;
}
// Tiles of 256 rows (2 MMA groups per hand-off) where two operand and two accumulator stages of that size fit tensor memory
// and four landing stages fit shared memory: narrow rows, whose 16 KB tiles did not amortise the hand-offs between the role
// warps (measured: vMF d=32 0.67 -> 0.88 of HBM, d=24 0.59 -> 0.77, d=48 0.84 -> 0.89, Bingham d=32 0.62 -> 0.82).  The kernel
// takes up to 4 groups per tile; 3 (d=24: 0.66) and 4 (Bingham d=16: 0.63 against 0.57) were mixed, so 2 is the rule.
inline int tcp_sub(int d, int np, int aux, bool one_acc, int dmax = 64) {
    const int kp = (d + 7) / 8 * 8, accw = one_acc ? np : 2 * np;
    if (d > dmax) return 1;
    for (int sub = 2; sub > 1; --sub)
        if (2 * (sub * 2 * kp) + 2 * (sub * accw) <= TCP_TMEM_COLS && tcp_stages(d, np, aux, sub) >= 4) return sub;
    return 1;
}
// two TMEM stages (hi, lo operand columns + the two accumulator halves each) must fit the 512 columns
template <class Epi> inline bool tcp_fits(int d, int n_out) {
    const int np = tcp_np(n_out);
    return np <= Epi::NP_MAX && 2 * (2 * ((d + 7) / 8 * 8) + 2 * np) <= TCP_TMEM_COLS && tcp_stages(d, np, Epi::AUX) >= 2;
}

}  // namespace geosss
