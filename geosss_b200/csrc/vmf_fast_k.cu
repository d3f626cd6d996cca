// Instantiations of k_vmf_fast for two values of K (GEOSSS_KPAIR = 0..3 -> K = 2 KPAIR + 1, 2 KPAIR + 2),
// every true dimension 2..16 and both samplers.  See vmf_fast.cu.
#include "kernels_fast.cuh"

#ifndef GEOSSS_KPAIR
#error "compile with -DGEOSSS_KPAIR=0..3"
#endif

namespace geosss {

template <int K, int DT> static int launch_kd(int kind, const VmfFastParams& P, const RunArgs& a, cudaStream_t st) {
    const unsigned grid = (unsigned)((a.N + FAST_BLOCK - 1) / FAST_BLOCK);
    if (kind == SHRINK) k_vmf_fast<K, DT, SHRINK><<<grid, FAST_BLOCK, 0, st>>>(P, a);
    else k_vmf_fast<K, DT, REJECT><<<grid, FAST_BLOCK, 0, st>>>(P, a);
    return (int)cudaGetLastError();
}

template <int K> static int launch_k(int kind, int d, const VmfFastParams& P, const RunArgs& a, cudaStream_t st) {
    switch (d) {
        case 2: return launch_kd<K, 2>(kind, P, a, st);
        case 3: return launch_kd<K, 3>(kind, P, a, st);
        case 4: return launch_kd<K, 4>(kind, P, a, st);
        case 5: return launch_kd<K, 5>(kind, P, a, st);
        case 6: return launch_kd<K, 6>(kind, P, a, st);
        case 7: return launch_kd<K, 7>(kind, P, a, st);
        case 8: return launch_kd<K, 8>(kind, P, a, st);
        case 9: return launch_kd<K, 9>(kind, P, a, st);
        case 10: return launch_kd<K, 10>(kind, P, a, st);
        case 11: return launch_kd<K, 11>(kind, P, a, st);
        case 12: return launch_kd<K, 12>(kind, P, a, st);
        case 13: return launch_kd<K, 13>(kind, P, a, st);
        case 14: return launch_kd<K, 14>(kind, P, a, st);
        case 15: return launch_kd<K, 15>(kind, P, a, st);
        case 16: return launch_kd<K, 16>(kind, P, a, st);
        default: return -3;
    }
}

#define GEOSSS_CAT_(a, b) a##b
#define GEOSSS_CAT(a, b) GEOSSS_CAT_(a, b)

int GEOSSS_CAT(launch_vmf_fast_kpair, GEOSSS_KPAIR)(int K, int kind, int d, const VmfFastParams& P, const RunArgs& a, cudaStream_t st) {
    return K == 2 * GEOSSS_KPAIR + 1 ? launch_k<2 * GEOSSS_KPAIR + 1>(kind, d, P, a, st)
                                     : launch_k<2 * GEOSSS_KPAIR + 2>(kind, d, P, a, st);
}

}  // namespace geosss
