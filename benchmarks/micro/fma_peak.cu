// Microbenchmark: sustained FP32 FMA and MUFU.EX2 issue rates on this GPU (roofline denominators for
// the issue-bound sampler kernels).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a fma_peak.cu -o fma_peak
#include <cstdio>
#include <cuda_runtime.h>

// packed FP32 pair (sm_100: fma.rn.f32x2 -> FFMA2, two FMAs per lane in one instruction)
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}

template <int MODE>
__global__ void __launch_bounds__(256) k2(float* out, int iters, float a, float b) {
    unsigned long long x0 = pack2(threadIdx.x, 1.f), x1 = pack2(threadIdx.x + 1, 2.f), x2 = pack2(threadIdx.x + 2, 3.f), x3 = pack2(threadIdx.x + 3, 4.f);
    unsigned long long x4 = pack2(threadIdx.x + 4, 1.f), x5 = pack2(threadIdx.x + 5, 2.f), x6 = pack2(threadIdx.x + 6, 3.f), x7 = pack2(threadIdx.x + 7, 4.f);
    const unsigned long long y0 = pack2(a, a), y1 = pack2(b, b), y2 = pack2(a + b, a), y3 = pack2(a - b, b);
    unsigned u0 = threadIdx.x, u1 = u0 * 3, u2 = u0 * 5, u3 = u0 * 7;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {  // 8 independent FFMA2 chains
            x0 = ffma2(x0, y0, y1); x1 = ffma2(x1, y1, y2); x2 = ffma2(x2, y2, y3); x3 = ffma2(x3, y3, y0);
            x4 = ffma2(x4, y0, y2); x5 = ffma2(x5, y1, y3); x6 = ffma2(x6, y2, y0); x7 = ffma2(x7, y3, y1);
        } else {  // 4 FFMA2 + 8 integer ops
            x0 = ffma2(x0, y0, y1); x1 = ffma2(x1, y1, y2); x2 = ffma2(x2, y2, y3); x3 = ffma2(x3, y3, y0);
            u0 = (u0 ^ u1) + i; u1 = (u1 ^ u2) + i; u2 = (u2 ^ u3) + i; u3 = (u3 ^ u0) + i;
        }
    }
    const unsigned long long s = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)s ^ (unsigned)(s >> 32) ^ u0 ^ u1 ^ u2 ^ u3);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    float y0 = a, y1 = b, y2 = a + b, y3 = a - b;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {  // FFMA with three distinct register operands
            x0 = fmaf(x0, y0, y1); x1 = fmaf(x1, y1, y2); x2 = fmaf(x2, y2, y3); x3 = fmaf(x3, y3, y0);
            x4 = fmaf(x4, y0, y2); x5 = fmaf(x5, y1, y3); x6 = fmaf(x6, y2, y0); x7 = fmaf(x7, y3, y1);
        } else if (MODE == 1) {  // FFMA with immediate operands
            x0 = fmaf(x0, 1.0001f, 0.5f); x1 = fmaf(x1, 1.0001f, 0.5f); x2 = fmaf(x2, 1.0001f, 0.5f); x3 = fmaf(x3, 1.0001f, 0.5f);
            x4 = fmaf(x4, 1.0001f, 0.5f); x5 = fmaf(x5, 1.0001f, 0.5f); x6 = fmaf(x6, 1.0001f, 0.5f); x7 = fmaf(x7, 1.0001f, 0.5f);
        } else if (MODE == 2) {  // MUFU.EX2
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x5));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x7));
        } else {  // 4 FFMA + 4 LOP3-like integer ops: mixed FMA/ALU pipes
            unsigned u0 = __float_as_uint(x4), u1 = __float_as_uint(x5), u2 = __float_as_uint(x6), u3 = __float_as_uint(x7);
            x0 = fmaf(x0, y0, y1); x1 = fmaf(x1, y1, y2); x2 = fmaf(x2, y2, y3); x3 = fmaf(x3, y3, y0);
            u0 = (u0 ^ u1) + i; u1 = (u1 ^ u2) + i; u2 = (u2 ^ u3) + i; u3 = (u3 ^ u0) + i;
            x4 = __uint_as_float(u0); x5 = __uint_as_float(u1); x6 = __uint_as_float(u2); x7 = __uint_as_float(u3);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int MODE, bool PACKED = false> double run(const char* name, double ops_per_iter_thread) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    float* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    if (PACKED) k2<MODE><<<blocks, threads>>>(out, 100, 1.0001f, 0.5f); else k<MODE><<<blocks, threads>>>(out, 100, 1.0001f, 0.5f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    if (PACKED) k2<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); else k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = ops_per_iter_thread * iters * (double)blocks * threads;
    const double per_s = ops / (ms * 1e-3);
    printf("{\"test\": \"%s\", \"ms\": %.3f, \"lane_ops_per_s\": %.4g, \"warp_instr_per_clk_per_smsp_at_1965MHz\": %.3f}\n", name, ms, per_s,
           per_s / 32.0 / (sms * 4.0) / 1.965e9);
    cudaFree(out);
    return per_s;
}

int main() {
    run<0>("ffma_3reg", 8);
    run<1>("ffma_imm", 8);
    run<2>("mufu_ex2", 8);
    run<3>("ffma4_plus_alu8", 12);
    // packed pairs: ops counted as warp INSTRUCTIONS per thread and iteration (an FFMA2 is one instruction, two FMAs per lane)
    run<0, true>("ffma2_x8 (instructions)", 8);
    run<1, true>("ffma2_x4_plus_alu8 (instructions)", 12);
    return 0;
}
