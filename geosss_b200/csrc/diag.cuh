// Device diagnostics: partial sums for the multi-chain effective sample size of a scalar test
// function f(x) = x . v (v = e_0 gives the first coordinate the upstream scripts use,
// scripts/ess_mixture_vMF.py:233; v = mode gives scripts/bingham.py:50-52).
//
// Restates the autocovariance step of geosss/utils.py:109-112 (acf) per chain, in the direct
// (non-circular, biased 1/n) form that the multi-chain estimator of Vehtari et al. (2021) uses
// (oracle/ess.py:ess_multichain).  One block per chain: the chain's centred trace is staged in
// shared memory, lags are spread over the threads, and the block adds its chain's
// autocovariances into global double accumulators -- O(L) doubles leave the GPU, which is what
// gets all-reduced over NVLink when chains are sharded over GPUs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace geosss {

constexpr int DIAG_BLOCK = 256;

// samples: draw t of chain c at samples[t * step_stride + c * chain_stride + j] (step-major (n, N, d):
// (N d, d); chain-major (N, n, d): (d, n d)); out: max_lag + 2 doubles, zeroed by the caller:
//   out[l] += acov_c(l) for l < max_lag;  out[max_lag] += mean_c;  out[max_lag+1] += mean_c^2
template <typename T>
__global__ void __launch_bounds__(DIAG_BLOCK) k_ess_partials(const T* __restrict__ samples, int64_t n, int64_t step_stride,
                                                             int64_t chain_stride, int d, const T* __restrict__ v, int max_lag,
                                                             double* __restrict__ out) {
    extern __shared__ float trace[];  // n floats
    __shared__ double red[DIAG_BLOCK / 32];
    __shared__ double mean_s;
    const int64_t c = blockIdx.x;
    double part = 0.0;
    for (int64_t t = threadIdx.x; t < n; t += blockDim.x) {
        const T* row = samples + t * step_stride + c * chain_stride;
        double f = 0.0;
        for (int j = 0; j < d; ++j) f += (double)row[j] * (double)v[j];
        trace[t] = (float)f;
        part += f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < DIAG_BLOCK / 32; ++w) s += red[w];
        mean_s = s / (double)n;
        atomicAdd(out + max_lag, mean_s);
        atomicAdd(out + max_lag + 1, mean_s * mean_s);
    }
    __syncthreads();
    const float mean = (float)mean_s;
    for (int64_t t = threadIdx.x; t < n; t += blockDim.x) trace[t] -= mean;
    __syncthreads();
    for (int l = threadIdx.x; l < max_lag; l += blockDim.x) {
        double acc = 0.0;
        for (int64_t t = 0; t + l < n; ++t) acc += (double)(trace[t] * trace[t + l]);
        atomicAdd(out + l, acc / (double)n);
    }
}

// Streaming form of the same estimator: a run of any length is folded in, chunk by chunk, without ever being retained
// (C3 at 2^20 chains keeps "last state + moments").  The sampler kernels write the scalar test function of a
// chunk of transitions (RunArgs::fout, 4 bytes per kept transition), k_ess_stream_update folds the chunk into per-chain
// accumulators and the chunk buffer is reused:
//   acc[c] = { S = sum f,  P[l] = sum_{t >= l} f_t f_{t-l} (l < L),  head[i] = f_i (i < L),  tail[i] = the last L values }
// (1 + 3 L doubles per chain).  k_ess_stream_finish turns them into the partial sums k_ess_partials produces
// (autocovariance about the chain mean, biased 1/n):  n acov(l) = P[l] - m (2 S - last_l - first_l) + (n - l) m^2  with
// last_l / first_l the sums of the last / first l values.  Products are accumulated in double.
template <typename T>
__global__ void __launch_bounds__(DIAG_BLOCK) k_ess_stream_update(const T* __restrict__ trace, int64_t n_new, int64_t step_stride,
                                                                  int64_t chain_stride, int64_t n_before, int L, double* __restrict__ acc) {
    extern __shared__ float ext[];  // L values carried from earlier chunks, then the n_new values of this one
    __shared__ double red[DIAG_BLOCK / 32];
    const int64_t c = blockIdx.x;
    double* a = acc + c * (1 + 3 * (int64_t)L);
    double *P = a + 1, *head = a + 1 + L, *tail = a + 1 + 2 * (int64_t)L;
    for (int i = threadIdx.x; i < L; i += blockDim.x) ext[i] = (float)tail[i];
    double part = 0.0;
    for (int64_t t = threadIdx.x; t < n_new; t += blockDim.x) {
        const float f = (float)trace[t * step_stride + c * chain_stride];
        ext[L + t] = f;
        part += (double)f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    // lagged products of the new values with everything before them (a partner older than the run does not exist)
    const int64_t valid_before = n_before < L ? n_before : L;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        double s = 0.0;
        int64_t t0 = l - valid_before;  // first new index whose partner exists
        if (t0 < 0) t0 = 0;
        for (int64_t t = t0; t < n_new; ++t) s += (double)ext[L + t] * (double)ext[L + t - l];
        P[l] += s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < DIAG_BLOCK / 32; ++w) s += red[w];
        a[0] += s;
    }
    // the first L values of the run, and the new last L values
    for (int64_t i = threadIdx.x; i < L; i += blockDim.x) {
        if (i >= n_before && i - n_before < n_new) head[i] = (double)ext[L + i - n_before];
        const int64_t src = n_new + i;  // position in ext of the i-th of the last L values (ext holds L + n_new values)
        tail[i] = (double)ext[src];
    }
}

// out: max_lag + 2 doubles as for k_ess_partials, zeroed by the caller; one thread per chain, sums reduced over the warp first
static __global__ void __launch_bounds__(DIAG_BLOCK) k_ess_stream_finish(const double* __restrict__ acc, int64_t n_chains, int64_t n, int L,
                                                                         double* __restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < n_chains;
    const double* a = acc + (live ? c : 0) * (1 + 3 * (int64_t)L);
    const double *P = a + 1, *head = a + 1 + L, *tail = a + 1 + 2 * (int64_t)L;
    const double S = live ? a[0] : 0.0, m = S / (double)n;
    auto warp_add = [&](double v, double* dst) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(dst, v);
    };
    warp_add(live ? m : 0.0, out + L);
    warp_add(live ? m * m : 0.0, out + L + 1);
    double first = 0.0, last = 0.0;
    for (int l = 0; l < L && l < n; ++l) {
        if (l > 0 && live) { first += head[l - 1]; last += tail[L - l]; }  // sums of the first / last l values
        const double cov = (P[l] - m * (2.0 * S - last - first) + (double)(n - l) * m * m) / (double)n;
        warp_add(live ? cov : 0.0, out + l);
    }
}

// Path statistics of the retained draws: mode populations (argmax_k x . mode_k,
// scripts/vMF_diagnostics.py:333-342), hops between labels of successive draws of a chain
// (np.diff(np.sign(x . mode)) != 0, scripts/bingham.py:23-25, when sign_labels; label changes otherwise)
// and the summed geodesic step arccos(clip(x_t . x_{t+1}, -1, 1)) (geosss/sphere.py:64-68 as used in
// scripts/curve_vMF.py:379).  One block per chain; out (zeroed by the caller) holds n_labels counts
// (K, or 3 = {-, 0, +} for sign_labels), then the hop count, then the distance sum.
constexpr int DIAG_MAX_MODES = 64;

template <typename T>
__device__ __forceinline__ int path_label(const T* __restrict__ row, const double* __restrict__ modes, int K, int d, bool sign_labels) {
    if (sign_labels) {
        double s = 0.0;
        for (int j = 0; j < d; ++j) s += (double)row[j] * modes[j];
        return s > 0.0 ? 2 : (s < 0.0 ? 0 : 1);
    }
    int best = 0;
    double bv = 0.0;
    for (int k = 0; k < K; ++k) {
        double s = 0.0;
        for (int j = 0; j < d; ++j) s += (double)row[j] * modes[k * d + j];
        if (k == 0 || s > bv) { bv = s; best = k; }
    }
    return best;
}

template <typename T>
__global__ void __launch_bounds__(DIAG_BLOCK) k_path_stats(const T* __restrict__ samples, int64_t n, int64_t step_stride,
                                                           int64_t chain_stride, int d, const double* __restrict__ modes, int K,
                                                           int sign_labels, double* __restrict__ out) {
    extern __shared__ double modes_s[];  // K * d
    __shared__ unsigned long long counts[DIAG_MAX_MODES];
    __shared__ unsigned long long hops_s;
    __shared__ double red[DIAG_BLOCK / 32];
    const int n_labels = sign_labels ? 3 : K;
    for (int i = threadIdx.x; i < K * d; i += blockDim.x) modes_s[i] = modes[i];
    if (threadIdx.x < DIAG_MAX_MODES) counts[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) hops_s = 0ull;
    __syncthreads();
    const T* chain = samples + (int64_t)blockIdx.x * chain_stride;
    double dist = 0.0;
    unsigned long long hops = 0ull;
    for (int64_t t = threadIdx.x; t < n; t += blockDim.x) {
        const T* row = chain + t * step_stride;
        const int lab = K > 0 ? path_label(row, modes_s, K, d, sign_labels != 0) : 0;
        if (K > 0) atomicAdd(&counts[lab], 1ull);
        if (t + 1 < n) {
            const T* nxt = row + step_stride;
            if (K > 0 && path_label(nxt, modes_s, K, d, sign_labels != 0) != lab) ++hops;
            double c = 0.0;
            for (int j = 0; j < d; ++j) c += (double)row[j] * (double)nxt[j];
            dist += acos(fmin(1.0, fmax(-1.0, c)));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dist += __shfl_xor_sync(0xffffffffu, dist, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dist;
    if (hops) atomicAdd(&hops_s, hops);
    __syncthreads();
    if (threadIdx.x < n_labels && counts[threadIdx.x]) atomicAdd(out + threadIdx.x, (double)counts[threadIdx.x]);
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < DIAG_BLOCK / 32; ++w) s += red[w];
        if (hops_s) atomicAdd(out + n_labels, (double)hops_s);
        atomicAdd(out + n_labels + 1, s);
    }
}

}  // namespace geosss
