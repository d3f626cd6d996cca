// Instantiations and dispatch of the bandwidth-bound vMF-mixture log_prob with the parameters in the constant bank
// (kernels_logp_fast.cuh: k_logp_tiled<VmfLpConst<K, d>>), K = 1..8 components x true dimension d = 2..16.
#include "kernels_logp_fast.cuh"

namespace geosss {

template <int K, int DT> static int launch_kd(const VmfFastParams& P, const float* x, float* out, int64_t N, int sm_count, cudaStream_t st) {
    using Policy = VmfLpConst<K, DT>;
    const size_t smem = (size_t)LP_STAGES * LP_ROWS * lp_row_stride(DT) * sizeof(float);
    // the shared-memory opt-in and the occupancy query cost tens of microseconds: once per kernel and device
    static int per_sm_dev[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int per_sm = (dev >= 0 && dev < 64) ? per_sm_dev[dev] : 0;
    if (per_sm == 0) {
        e = cudaFuncSetAttribute(k_logp_tiled<Policy, LP_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_logp_tiled<Policy, LP_ROWS>, LP_ROWS, smem);
        if (e != cudaSuccess) return (int)e;
        if (per_sm < 1) per_sm = 1;
        if (dev >= 0 && dev < 64) per_sm_dev[dev] = per_sm;
    }
    const int64_t n_tiles = (N + LP_ROWS - 1) / LP_ROWS;
    int64_t grid = (int64_t)sm_count * per_sm;  // persistent: a multiple of the SM count
    if (grid > n_tiles) grid = n_tiles;
    k_logp_tiled<Policy, LP_ROWS><<<(unsigned)grid, LP_ROWS, smem, st>>>(P, x, out, N);
    return (int)cudaGetLastError();
}

template <int K> static int launch_k(int d, const VmfFastParams& P, const float* x, float* out, int64_t N, int sm_count, cudaStream_t st) {
    switch (d) {
        case 2: return launch_kd<K, 2>(P, x, out, N, sm_count, st);
        case 3: return launch_kd<K, 3>(P, x, out, N, sm_count, st);
        case 4: return launch_kd<K, 4>(P, x, out, N, sm_count, st);
        case 5: return launch_kd<K, 5>(P, x, out, N, sm_count, st);
        case 6: return launch_kd<K, 6>(P, x, out, N, sm_count, st);
        case 7: return launch_kd<K, 7>(P, x, out, N, sm_count, st);
        case 8: return launch_kd<K, 8>(P, x, out, N, sm_count, st);
        case 9: return launch_kd<K, 9>(P, x, out, N, sm_count, st);
        case 10: return launch_kd<K, 10>(P, x, out, N, sm_count, st);
        case 11: return launch_kd<K, 11>(P, x, out, N, sm_count, st);
        case 12: return launch_kd<K, 12>(P, x, out, N, sm_count, st);
        case 13: return launch_kd<K, 13>(P, x, out, N, sm_count, st);
        case 14: return launch_kd<K, 14>(P, x, out, N, sm_count, st);
        case 15: return launch_kd<K, 15>(P, x, out, N, sm_count, st);
        case 16: return launch_kd<K, 16>(P, x, out, N, sm_count, st);
        default: return -3;
    }
}

int launch_logp_vmf_const(int d, int K, const VmfFastParams& P, const float* x, float* out, int64_t N, int sm_count, cudaStream_t st) {
    switch (K) {
        case 1: return launch_k<1>(d, P, x, out, N, sm_count, st);
        case 2: return launch_k<2>(d, P, x, out, N, sm_count, st);
        case 3: return launch_k<3>(d, P, x, out, N, sm_count, st);
        case 4: return launch_k<4>(d, P, x, out, N, sm_count, st);
        case 5: return launch_k<5>(d, P, x, out, N, sm_count, st);
        case 6: return launch_k<6>(d, P, x, out, N, sm_count, st);
        case 7: return launch_k<7>(d, P, x, out, N, sm_count, st);
        case 8: return launch_k<8>(d, P, x, out, N, sm_count, st);
        default: return -3;
    }
}

}  // namespace geosss
