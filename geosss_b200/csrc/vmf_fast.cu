// Dispatch of the register-resident vMF-mixture sampler kernel (kernels_fast.cuh): K = 1..8 components x
// true dimension d = 2..16 x {reject, shrink}.  The 240 instantiations are compiled in four translation
// units (vmf_fast_k.cu with GEOSSS_KPAIR = 0..3: K = 2 KPAIR + 1 and 2 KPAIR + 2) so that make -j builds them in parallel.
#include "kernels_fast.cuh"

namespace geosss {

int launch_vmf_fast_kpair0(int K, int kind, int d, const VmfFastParams& P, const RunArgs& a, cudaStream_t st);
int launch_vmf_fast_kpair1(int K, int kind, int d, const VmfFastParams& P, const RunArgs& a, cudaStream_t st);
int launch_vmf_fast_kpair2(int K, int kind, int d, const VmfFastParams& P, const RunArgs& a, cudaStream_t st);
int launch_vmf_fast_kpair3(int K, int kind, int d, const VmfFastParams& P, const RunArgs& a, cudaStream_t st);

int launch_fast(int target_kind, int kind, int d, int K, const double* mu, const double* consts, const RunArgs& a, cudaStream_t st) {
    if (fast_supported(target_kind, d, K, 0)) return -3;
    if (a.n_steps > 0x7fffffff) return -1;
    const VmfFastParams P = make_vmf_fast_params(d, K, mu, consts);
    switch ((K - 1) / 2) {
        case 0: return launch_vmf_fast_kpair0(K, kind, d, P, a, st);
        case 1: return launch_vmf_fast_kpair1(K, kind, d, P, a, st);
        case 2: return launch_vmf_fast_kpair2(K, kind, d, P, a, st);
        default: return launch_vmf_fast_kpair3(K, kind, d, P, a, st);
    }
}

}  // namespace geosss
