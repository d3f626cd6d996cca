// Shared definitions for libgeosss_cuda (sm_100a).
//
// Everything marked GEOSSS_HD is plain arithmetic with no CUDA-only constructs, so the
// test-only harness tests/hostsim/ can compile the SAME headers with g++ and check the
// per-chain math against the golden vectors on a machine without a GPU.  The product
// never runs that host build: geosss_b200 loads libgeosss_cuda.so or fails.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GEOSSS_HD __host__ __device__ __forceinline__
#else
#define GEOSSS_HD inline
#endif

namespace geosss {

enum { F32 = 0, F64 = 1 };
enum { REJECT = 0, SHRINK = 1 };
enum { PATH_LITERAL = 0, PATH_FAST = 1 };
enum { TGT_VMF_MIX = 0, TGT_BINGHAM = 1, TGT_CURVED_VMF = 2, TGT_CPD = 3 };

// Proposal cap per transition.  The reference loops for ever (geosss/mcmc.py:394-401 has no
// guard); past the cap the chain keeps its current state, which is always inside the slice.
constexpr int MAX_PROPOSALS = 1 << 17;  // keeps the 16-bit Philox block index from wrapping

// ---- scalar math, one spelling for float and double -------------------------------------
GEOSSS_HD float  m_sqrt(float v)  { return sqrtf(v); }
GEOSSS_HD double m_sqrt(double v) { return sqrt(v); }
GEOSSS_HD float  m_exp(float v)   { return expf(v); }
GEOSSS_HD double m_exp(double v)  { return exp(v); }
GEOSSS_HD float  m_log(float v)   { return logf(v); }
GEOSSS_HD double m_log(double v)  { return log(v); }
GEOSSS_HD float  m_log1p(float v)  { return log1pf(v); }
GEOSSS_HD double m_log1p(double v) { return log1p(v); }
GEOSSS_HD float  m_sin(float v)   { return sinf(v); }
GEOSSS_HD double m_sin(double v)  { return sin(v); }
GEOSSS_HD float  m_cos(float v)   { return cosf(v); }
GEOSSS_HD double m_cos(double v)  { return cos(v); }
GEOSSS_HD float  m_acos(float v)  { return acosf(v); }
GEOSSS_HD double m_acos(double v) { return acos(v); }
GEOSSS_HD float  m_atan2(float a, float b)   { return atan2f(a, b); }
GEOSSS_HD double m_atan2(double a, double b) { return atan2(a, b); }
template <typename T> GEOSSS_HD T m_clip(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }
template <typename T> GEOSSS_HD T m_max(T a, T b) { return a > b ? a : b; }

template <typename T> struct Const;
template <> struct Const<float> {
    static GEOSSS_HD float two_pi() { return 6.283185307179586f; }
    static GEOSSS_HD float log_two_pi() { return 1.8378770664093453f; }
    static GEOSSS_HD float tiny() { return 0.0f; }  // 1e-100 is 0 in fp32
    static GEOSSS_HD float tiny_r() { return 0.0f; }
    static GEOSSS_HD float neg_inf() { return -INFINITY; }
    static GEOSSS_HD float bracket_floor() { return 1e-9f; }  // |theta| below fp32 resolution of a unit vector
};
template <> struct Const<double> {
    static GEOSSS_HD double two_pi() { return 6.283185307179586; }       // 2*np.pi
    static GEOSSS_HD double log_two_pi() { return 1.8378770664093453; }  // np.log(2*np.pi)
    static GEOSSS_HD double tiny() { return 1e-100; }                    // geosss/sphere.py:14
    static GEOSSS_HD double tiny_r() { return 1e-300; }                  // geosss/pointcloud.py:152
    static GEOSSS_HD double neg_inf() { return -INFINITY; }
    static GEOSSS_HD double bracket_floor() { return 1e-18; }
};

#if defined(__CUDACC__)
// MUFU-pipe approximations used by the float32 throughput kernels (relative error ~2^-22)
__device__ __forceinline__ float fast_ex2(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float fast_lg2(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float fast_sqrt(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
#endif

// Strided view of one chain's vector inside a (d, N) column-major workspace: element j of
// chain c lives at base[c + j*N], so the 32 lanes of a warp touch 32 consecutive words.
template <typename T> struct SVec {
    T* p;
    int64_t stride;
    GEOSSS_HD T& operator[](int j) const { return p[(int64_t)j * stride]; }
};

}  // namespace geosss
