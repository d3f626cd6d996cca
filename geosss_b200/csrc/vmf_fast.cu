// Instantiations and dispatch of the register-resident vMF-mixture sampler kernel (kernels_fast.cuh):
// K = 1..8 components x padded dimension D in {4, 8, 12, 16} x {reject, shrink}.
#include "kernels_fast.cuh"

namespace geosss {

template <int K, int D> static int launch_vmf_fast_kd(int kind, const VmfFastParams& P, const RunArgs& a, cudaStream_t st) {
    const unsigned grid = (unsigned)((a.N + FAST_BLOCK - 1) / FAST_BLOCK);
    if (a.n_steps > 0x7fffffff) return -1;
    if (kind == SHRINK) k_vmf_fast<K, D, SHRINK><<<grid, FAST_BLOCK, 0, st>>>(P, a);
    else k_vmf_fast<K, D, REJECT><<<grid, FAST_BLOCK, 0, st>>>(P, a);
    return (int)cudaGetLastError();
}

template <int K> static int launch_vmf_fast_k(int kind, int Dp, const VmfFastParams& P, const RunArgs& a, cudaStream_t st) {
    switch (Dp) {
        case 4: return launch_vmf_fast_kd<K, 4>(kind, P, a, st);
        case 8: return launch_vmf_fast_kd<K, 8>(kind, P, a, st);
        case 12: return launch_vmf_fast_kd<K, 12>(kind, P, a, st);
        default: return launch_vmf_fast_kd<K, 16>(kind, P, a, st);
    }
}

int launch_fast(int target_kind, int kind, int d, int K, const double* mu, const double* consts, const RunArgs& a, cudaStream_t st) {
    if (fast_supported(target_kind, d, K, 0)) return -3;
    const int Dp = (d + 3) / 4 * 4;
    const VmfFastParams P = make_vmf_fast_params(d, K, mu, consts);
    switch (K) {
        case 1: return launch_vmf_fast_k<1>(kind, Dp, P, a, st);
        case 2: return launch_vmf_fast_k<2>(kind, Dp, P, a, st);
        case 3: return launch_vmf_fast_k<3>(kind, Dp, P, a, st);
        case 4: return launch_vmf_fast_k<4>(kind, Dp, P, a, st);
        case 5: return launch_vmf_fast_k<5>(kind, Dp, P, a, st);
        case 6: return launch_vmf_fast_k<6>(kind, Dp, P, a, st);
        case 7: return launch_vmf_fast_k<7>(kind, Dp, P, a, st);
        default: return launch_vmf_fast_k<8>(kind, Dp, P, a, st);
    }
}

}  // namespace geosss
