// Launchers of the baseline-sampler and gradient kernels (kernels_mh.cuh); a translation unit of its own
// so that it builds in parallel with capi.cu.
#include <cuda_runtime.h>

#include <type_traits>

#include "kernels_exact.cuh"
#include "kernels_mh.cuh"

namespace geosss {

template <typename T, class P> int launch_mh(const P& p, int kind, bool tape, const RunArgs& a, cudaStream_t st) {
    if constexpr (std::is_same<P, CpdP<T>>::value) {
        const size_t smem = (size_t)CPD_WARPS * 3 * p.ns * sizeof(T);
        if (smem > 200 * 1024) return -1;
        const unsigned grid = (unsigned)((a.N + CPD_WARPS - 1) / CPD_WARPS);
        cudaError_t e;
        if (tape) {
            e = cudaFuncSetAttribute(k_mh_cpd<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            k_mh_cpd<T, true><<<grid, CPD_WARPS * 32, smem, st>>>(p, kind, a);
        } else {
            e = cudaFuncSetAttribute(k_mh_cpd<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            k_mh_cpd<T, false><<<grid, CPD_WARPS * 32, smem, st>>>(p, kind, a);
        }
    } else {
        const unsigned grid = (unsigned)((a.N + LANE_BLOCK - 1) / LANE_BLOCK);
        RunArgs b = a;
        const size_t smem = lane_scratch_smem<T>(5, p.d);  // x, v, proposal x / v, gradient in shared memory when they fit
        b.ws_smem = smem ? 1 : 0;
        cudaError_t e = cudaSuccess;
        if (tape) {
            if (smem > 48 * 1024) e = cudaFuncSetAttribute(k_mh_lane<T, true, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            k_mh_lane<T, true><<<grid, LANE_BLOCK, smem, st>>>(p, kind, b);
        } else {
            if (smem > 48 * 1024) e = cudaFuncSetAttribute(k_mh_lane<T, false, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            k_mh_lane<T, false><<<grid, LANE_BLOCK, smem, st>>>(p, kind, b);
        }
    }
    return (int)cudaGetLastError();
}

template <typename T, class P> int launch_grad(const P& p, const T* x, T* out, int64_t N, cudaStream_t st) {
    if constexpr (std::is_same<P, CpdP<T>>::value) {
        const size_t smem = (size_t)CPD_WARPS * 3 * p.ns * sizeof(T);
        if (smem > 200 * 1024) return -1;
        cudaError_t e = cudaFuncSetAttribute(k_grad_cpd<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k_grad_cpd<T><<<(unsigned)((N + CPD_WARPS - 1) / CPD_WARPS), CPD_WARPS * 32, smem, st>>>(p, x, out, N);
    } else {
        k_grad_lane<T><<<(unsigned)((N + LANE_BLOCK - 1) / LANE_BLOCK), LANE_BLOCK, 0, st>>>(p, x, out, N);
    }
    return (int)cudaGetLastError();
}

template <typename T> int launch_exact_vmf(const ExactVmfArgs& a, uint64_t seed, uint64_t offset, T* out, int64_t N, unsigned long long* attempts,
                                           cudaStream_t st) {
    k_exact_vmf<T><<<(unsigned)((N + LANE_BLOCK - 1) / LANE_BLOCK), LANE_BLOCK, 0, st>>>(a, seed, offset, out, N, attempts);
    return (int)cudaGetLastError();
}

template <typename T> int launch_exact_bingham(const ExactBinghamArgs& a, uint64_t seed, uint64_t offset, T* out, double* ws, int64_t N,
                                               unsigned long long* attempts, cudaStream_t st) {
    k_exact_bingham<T><<<(unsigned)((N + LANE_BLOCK - 1) / LANE_BLOCK), LANE_BLOCK, 0, st>>>(a, seed, offset, out, ws, N, attempts);
    return (int)cudaGetLastError();
}

template int launch_exact_vmf<float>(const ExactVmfArgs&, uint64_t, uint64_t, float*, int64_t, unsigned long long*, cudaStream_t);
template int launch_exact_vmf<double>(const ExactVmfArgs&, uint64_t, uint64_t, double*, int64_t, unsigned long long*, cudaStream_t);
template int launch_exact_bingham<float>(const ExactBinghamArgs&, uint64_t, uint64_t, float*, double*, int64_t, unsigned long long*, cudaStream_t);
template int launch_exact_bingham<double>(const ExactBinghamArgs&, uint64_t, uint64_t, double*, double*, int64_t, unsigned long long*, cudaStream_t);

#define GEOSSS_MH_INSTANCES(T)                                                                        \
    template int launch_mh<T, VmfMixP<T>>(const VmfMixP<T>&, int, bool, const RunArgs&, cudaStream_t);   \
    template int launch_mh<T, BinghamP<T>>(const BinghamP<T>&, int, bool, const RunArgs&, cudaStream_t); \
    template int launch_mh<T, CurvedP<T>>(const CurvedP<T>&, int, bool, const RunArgs&, cudaStream_t);   \
    template int launch_mh<T, CpdP<T>>(const CpdP<T>&, int, bool, const RunArgs&, cudaStream_t);         \
    template int launch_grad<T, VmfMixP<T>>(const VmfMixP<T>&, const T*, T*, int64_t, cudaStream_t);     \
    template int launch_grad<T, BinghamP<T>>(const BinghamP<T>&, const T*, T*, int64_t, cudaStream_t);   \
    template int launch_grad<T, CurvedP<T>>(const CurvedP<T>&, const T*, T*, int64_t, cudaStream_t);     \
    template int launch_grad<T, CpdP<T>>(const CpdP<T>&, const T*, T*, int64_t, cudaStream_t);
GEOSSS_MH_INSTANCES(float)
GEOSSS_MH_INSTANCES(double)

}  // namespace geosss
