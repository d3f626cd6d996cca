// C ABI of libgeosss_cuda (see include/geosss_cuda.h): handles, parameter upload, dispatch.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/geosss_cuda.h"
#include "diag.cuh"
#include "eigh_jacobi.h"
#include "kernels_fast.cuh"
#include "kernels_gc.cuh"
#include "kernels_logp_fast.cuh"
#include "kernels_tc.cuh"
#include "kernels_logp_tc.cuh"
#include "kernels_literal.cuh"
#include "kernels_exact.cuh"
#include "kernels_mh.cuh"
#include "kernels_cpd_cell.cuh"

using namespace geosss;

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return fail(GEOSSS_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)


// ------------------------------------------------------------------ device / pinned scratch kept between calls
// cudaMalloc / cudaFree cost 0.3-1 ms each and cudaFree synchronises the device; a sampler object that lives
// for one sample() call (the reference's usage: construct, sample, drop) would pay that on every call.  Freed
// sampler buffers are therefore parked per device (exact-size reuse, at most POOL_SLOTS blocks / POOL_BYTES) and
// handed to the next sampler of the same shape.  Callers synchronise the device before parking a block.
namespace {
struct DevPool {
    static constexpr int POOL_SLOTS = 16;
    static constexpr size_t POOL_BYTES = size_t(1) << 30;
    struct Slot { void* p; size_t bytes; int dev; };
    std::mutex mu;
    std::vector<Slot> parked;
    size_t parked_bytes = 0;
    cudaError_t get(void** out, size_t bytes) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        {
            std::lock_guard<std::mutex> lock(mu);
            for (size_t i = 0; i < parked.size(); ++i)
                if (parked[i].dev == dev && parked[i].bytes == bytes) {
                    *out = parked[i].p;
                    parked_bytes -= bytes;
                    parked.erase(parked.begin() + (long)i);
                    return cudaSuccess;
                }
        }
        e = cudaMalloc(out, bytes);
        if (e == cudaErrorMemoryAllocation) {  // give parked memory back before giving up
            cudaGetLastError();
            trim();
            e = cudaMalloc(out, bytes);
        }
        return e;
    }
    void put(void* p, size_t bytes) {
        if (!p) return;
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && bytes <= POOL_BYTES / 4) {
            std::lock_guard<std::mutex> lock(mu);
            if ((int)parked.size() < POOL_SLOTS && parked_bytes + bytes <= POOL_BYTES) {
                parked.push_back(Slot{p, bytes, dev});
                parked_bytes += bytes;
                return;
            }
        }
        cudaFree(p);
    }
    void trim() {
        std::lock_guard<std::mutex> lock(mu);
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto& s : parked) { cudaSetDevice(s.dev); cudaFree(s.p); }
        cudaSetDevice(cur);
        parked.clear();
        parked_bytes = 0;
    }
};
DevPool g_pool;

// pinned staging for the small host<->device state transfers (initial states in, rejection counters out)
struct PinnedScratch {
    std::mutex mu;
    void* p = nullptr;
    size_t cap = 0;
    void* ensure(size_t bytes) {
        if (bytes > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr; cap = 0;
            if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); p = nullptr; return nullptr; }
            cap = bytes;
        }
        return p;
    }
};
PinnedScratch g_pinned;
}  // namespace

extern "C" int geosss_release_cached_memory(void) {
    cudaDeviceSynchronize();
    g_pool.trim();
    std::lock_guard<std::mutex> lock(g_pinned.mu);
    if (g_pinned.p) cudaFreeHost(g_pinned.p);
    g_pinned.p = nullptr; g_pinned.cap = 0;
    return 0;
}

// ------------------------------------------------------------------ device arrays in both precisions
struct Arr {
    float* f = nullptr;
    double* d = nullptr;
    size_t n = 0;
    int upload(const double* h, size_t count) {
        n = count;
        if (!count) return 0;
        std::vector<float> hf(count);
        for (size_t i = 0; i < count; ++i) hf[i] = (float)h[i];
        CU(cudaMalloc(&f, count * sizeof(float)));
        CU(cudaMalloc(&d, count * sizeof(double)));
        CU(cudaMemcpy(f, hf.data(), count * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d, h, count * sizeof(double), cudaMemcpyHostToDevice));
        return 0;
    }
    void release() {
        if (f) cudaFree(f);
        if (d) cudaFree(d);
        f = nullptr; d = nullptr;
    }
    template <typename T> const T* get() const;
};
template <> const float* Arr::get<float>() const { return f; }
template <> const double* Arr::get<double>() const { return d; }

struct geosss_target {
    int kind = 0, d = 0;
    // vMF mixture
    int K = 0, single = 0;
    Arr mu, logw, logi0;
    std::vector<double> h_mu, h_const;  // host copies for the fast path's parameter block
    Arr consts;                         // (K) log w - log 2pi - log i0, device
    // Bingham
    Arr A, b, Adiag;
    Arr eig_rows, eig_vals;  // dense A, d <= 64: eigenvectors as rows (+ b as row d when it fits), eigenvalues (kernels_logp_tc.cuh)
    Arr eig_b;               // the Fisher vector in the eigenbasis (W b)
    Arr eig_cols;            // W transposed (rows = columns of W): the B operand of the rotation back, x = W^T y
    std::vector<double> h_eig_rows;  // host copy of the d eigenvector rows (direction of the test function, tests)
    int has_b = 0, diag = 0, eig_n_out = 0;
    // curved vMF
    int M = 0;
    double kappa = 0;
    Arr knots, ang, cosang, sinang;
    // CPD
    int nt = 0, ns = 0, k = 0, outlier = 1;
    double sigma2 = 1, log_const = 0, log_outlier = 0, beta = 1, max_logw = 0;
    double shift[3] = {0, 0, 0};
    int dim = 3;
    Arr tgt, wt, src, logws;
    // cell lists of the float32 cutoff kernels (kernels_cpd_cell.cuh); has_grid = 0: dense pair sum
    int has_grid = 0;
    CpdCellGrid grid{};
    double grid_rc = 0, grid_mean_candidates = 0;
    void *d_cell = nullptr, *d_entry = nullptr, *d_tgt4 = nullptr;

    template <typename T> VmfMixP<T> vmf() const { return {mu.get<T>(), logw.get<T>(), logi0.get<T>(), K, d, single}; }
    template <typename T> BinghamP<T> bingham() const { return {A.get<T>(), has_b ? b.get<T>() : nullptr, d}; }
    template <typename T> CurvedP<T> curved() const {
        return {knots.get<T>(), ang.get<T>(), cosang.get<T>(), sinang.get<T>(), (T)kappa, M, d};
    }
    template <typename T> CpdP<T> cpd() const {
        return {tgt.get<T>(), wt.get<T>(), src.get<T>(), logws.get<T>(), (T)sigma2, (T)log_const,
                (T)log_outlier, (T)beta, nt, ns, k, outlier, {(T)shift[0], (T)shift[1], (T)shift[2]}, dim};
    }
    void release() {
        for (Arr* a : {&mu, &logw, &logi0, &consts, &A, &b, &Adiag, &eig_rows, &eig_vals, &eig_b, &eig_cols, &knots, &ang, &cosang, &sinang, &tgt, &wt, &src, &logws}) a->release();
        for (void** q : {&d_cell, &d_entry, &d_tgt4}) { if (*q) cudaFree(*q); *q = nullptr; }
        has_grid = 0;
    }
};

struct geosss_sampler {
    geosss_target* t = nullptr;
    int kind = 0, dtype = 0, path = 0;
    int64_t N = 0;
    uint64_t seed = 0, chain0 = 0, step = 0;
    void* state = nullptr;  // (N, d) dtype
    void* ws = nullptr;     // literal path scratch (3, d, N)
    unsigned long long* n_reject = nullptr;  // baseline samplers: accepted proposals
    // baseline samplers (kind >= GEOSSS_RWMH)
    double* stepsize = nullptr;
    int64_t* adapt_count = nullptr;
    unsigned long long* n_walk = nullptr;
    void* momentum = nullptr;
    void* momenta_out = nullptr;
    void* fdir = nullptr;   // (d) direction of the scalar test function, sampler dtype (device), see geosss_sampler_trace_out
    void* fout = nullptr;   // where the next run calls write its trace
    int64_t adapt_total = 0;
    double alpha = 0.5;
    int n_leap = 10;
    bool eig = false;                      // dense Bingham on the fast path: the chain lives in the eigenbasis of A
    void* rot_tmp = nullptr;               // one 128-row tile of scratch for the ragged tail of a tensor-core rotation
    void* zbuf = nullptr;                  // great-circle kernels without stored normals: (ceil(d/4), N) float4 scratch (kernels_gc.cuh)
    size_t zbuf_bytes = 0;
    size_t state_bytes = 0, ws_bytes = 0;  // sizes of the pooled blocks (g_pool)
    int device = 0;                        // device the blocks live on (current device at creation)
    bool is_mh() const { return kind >= GEOSSS_RWMH; }
    size_t esize() const { return dtype == GEOSSS_F64 ? 8 : 4; }
};

// Rows of a strided (n_outer, n_inner, d) float array rotated in place: y = W x (to_eig) or x = W^T y; W (d, d) row-major,
// rows = eigenvectors.  A register-tiled FP32 GEMM: a block takes 128 consecutive rows, stages them TRANSPOSED in shared memory
// (Xt[k][row], coalesced global reads), W as Ws[k][out]; thread (row group ri, output group oi = its warp) accumulates a 4 x 4 tile:
// per k one 16-byte load of four rows' x_k, one 16-byte broadcast load of four outputs' weights, 16 FMAs.  Every output is the
// sequential fmaf chain over k = 0..d-1, so a row's result does not depend on where it sits in the array (the continuation
// "first sample of the next call == last sample of this one" stays bit-exact).  First versions (thread per row, 64 row values in
// registers, one shared load per 1-4 FMAs) ran at 1.0-1.7e9 rows/s, load/store-unit bound.
constexpr int ROT_ROWS = 128, ROT_RS = ROT_ROWS + 4;
static __global__ void __launch_bounds__(512) k_rotate_rows(const float* __restrict__ W, int d, int to_eig, float* __restrict__ data,
                                                            int64_t n_outer, int64_t n_inner, int64_t s_outer, int64_t s_inner) {
    extern __shared__ __align__(16) float rot_smem[];
    const int dp = (d + 3) & ~3;
    float* Ws = rot_smem;             // d x dp:      Ws[k * dp + o] multiplies input k into output o
    float* Xt = rot_smem + d * dp;    // dp x ROT_RS: Xt[k * ROT_RS + r], reused for the outputs
    {
        const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5, n_warps_ = blockDim.x >> 5;
        for (int k = warp_; k < d; k += n_warps_)
            for (int o = lane_; o < dp; o += 32) Ws[k * dp + o] = o < d ? (to_eig ? W[o * d + k] : W[k * d + o]) : 0.f;
    }
    __shared__ float* row_base[ROT_ROWS];
    const int64_t n_rows = n_outer * n_inner, row0 = (int64_t)blockIdx.x * ROT_ROWS;
    const int rows_here = (int)(n_rows - row0 < ROT_ROWS ? n_rows - row0 : ROT_ROWS);
    for (int r = threadIdx.x; r < ROT_ROWS; r += blockDim.x) {  // one 64-bit division per row, not per element
        const int64_t idx = row0 + r;
        row_base[r] = r < rows_here ? data + (idx / n_inner) * s_outer + (idx % n_inner) * s_inner : nullptr;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int r = warp; r < ROT_ROWS; r += n_warps) {  // a warp per row: d contiguous floats
        const float* src = row_base[r];
        for (int c = lane; c < d; c += 32) Xt[c * ROT_RS + r] = src ? src[c] : 0.f;
    }
    __syncthreads();
    const int ri = threadIdx.x & 31, oi = threadIdx.x >> 5;  // rows 4 ri .. 4 ri + 3, outputs 4 oi .. 4 oi + 3 (one output group per warp)
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
    for (int k = 0; k < d; ++k) {
        const float4 xv = *reinterpret_cast<const float4*>(Xt + k * ROT_RS + 4 * ri);
        const float4 wv = *reinterpret_cast<const float4*>(Ws + k * dp + 4 * oi);
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, wb[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(wb[b], xa[a], acc[a][b]);
    }
    __syncthreads();  // every warp has read the inputs: the tile now takes the outputs
#pragma unroll
    for (int b = 0; b < 4; ++b)
        *reinterpret_cast<float4*>(Xt + (4 * oi + b) * ROT_RS + 4 * ri) = make_float4(acc[0][b], acc[1][b], acc[2][b], acc[3][b]);
    __syncthreads();
    for (int r = warp; r < rows_here; r += n_warps) {
        float* dst = row_base[r];
        for (int c = lane; c < d; c += 32) dst[c] = Xt[c * ROT_RS + r];
    }
}

// contiguous rows on the tensor cores (defined below, next to the other tcgen05 launches); returns 1 when it does not apply
static int rotate_rows_tc(geosss_sampler* s, bool to_eig, float* data, int64_t rows, cudaStream_t st);

static int rotate_rows(geosss_sampler* s, bool to_eig, void* data, int64_t n_outer, int64_t n_inner, int64_t s_outer, int64_t s_inner,
                       cudaStream_t st) {
    const int64_t rows = n_outer * n_inner;
    if (rows <= 0) return 0;
    const int d = s->t->d, dp = (d + 3) & ~3;
    // one dense (rows, d) matrix: 3xTF32 GEMM on tcgen05 (HBM-bound); strided layouts stay on the FP32 kernel below
    const bool dense_rows = (n_inner == 1 ? s_outer == d : (s_inner == d && (n_outer == 1 || s_outer == n_inner * (int64_t)d)));
    if (dense_rows) {
        const int rc = rotate_rows_tc(s, to_eig, static_cast<float*>(data), rows, st);
        if (rc <= 0) return rc;
    }
    const size_t smem = ((size_t)d * dp + (size_t)dp * ROT_RS) * sizeof(float);
    CU(cudaFuncSetAttribute(k_rotate_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_rotate_rows<<<(unsigned)((rows + ROT_ROWS - 1) / ROT_ROWS), 32 * (dp / 4), smem, st>>>(s->t->eig_rows.f, d, to_eig ? 1 : 0,
                                                                                            static_cast<float*>(data), n_outer, n_inner, s_outer, s_inner);
    CU(cudaGetLastError());
    return 0;
}

static int fill_stepsize(geosss_sampler* s, double value) {
    std::vector<double> h((size_t)s->N, value);
    CU(cudaMemcpy(s->stepsize, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

// ------------------------------------------------------------------ fast-path dispatch
// the cutoff kernels serve the dense sum (k <= 0 or k >= ns) when the lists exist; GEOSSS_CPD_DENSE=1 forces the pair sum
static bool cpd_use_grid(const geosss_target* t) {
    if (!t->has_grid) return false;
    const char* e = std::getenv("GEOSSS_CPD_DENSE");
    return !(e && e[0] == '1');
}

static CpdCellGrid cpd_grid_now(const geosss_target* t) {
    CpdCellGrid g = t->grid;
    for (int a = 0; a < 3; ++a) g.shift[a] = (float)t->shift[a];
    return g;
}

// the normals stay in shared memory for d < 64 and for dense Bingham (kernels_gc.cuh, STORE); measured on
// B200: regenerating them pays at d = 100 (+37 %, twice the resident warps) and costs 6 % at d = 50
// Dense Bingham chains run in the eigenbasis of A (the sampler is equivariant under rotations: great circles, the uniform
// direction and the density are all carried along), i.e. on the DIAGONAL kernel with the eigenvalues; states and retained samples
// are rotated at the boundary (k_rotate_rows).  The reference holds the same decomposition (distributions.py:67-76).
// GEOSSS_NO_EIG=1 keeps the dense matrix-vector kernel.
static bool bingham_eig(const geosss_target* t) {
    return t->kind == TGT_BINGHAM && !t->diag && !t->h_eig_rows.empty() && !std::getenv("GEOSSS_NO_EIG");
}
static bool bingham_diag_like(const geosss_target* t) { return t->diag || bingham_eig(t); }
static bool gc_store_normals(const geosss_target* t) { return t->d < 64 || (t->kind == TGT_BINGHAM && !bingham_diag_like(t)); }

static size_t gc_smem_bytes(const geosss_target* t, const GcParams& p) {
    const size_t pol = t->kind == TGT_BINGHAM ? BinghamGc::smem_floats(p)
                       : t->kind == TGT_VMF_MIX ? VmfGc::smem_floats(p) : CurvedGc::smem_floats(p);
    return ((gc_store_normals(t) ? 2 : 1) * (size_t)t->d * gc_block(gc_store_normals(t)) + pol) * sizeof(float);
}

static GcParams gc_params(const geosss_target* t) {
    GcParams p{};
    p.d = t->d;
    if (t->kind == TGT_BINGHAM) {
        p.diag = bingham_diag_like(t) ? 1 : 0;
        p.a0 = t->diag ? t->Adiag.f : (bingham_eig(t) ? t->eig_vals.f : t->A.f);
        p.a1 = t->has_b ? (bingham_eig(t) ? t->eig_b.f : t->b.f) : nullptr;
    } else if (t->kind == TGT_VMF_MIX) {
        p.M = t->K; p.a0 = t->mu.f; p.a1 = t->consts.f;
    } else {
        p.M = t->M; p.kappa = (float)t->kappa;
        p.a0 = t->knots.f; p.a1 = t->cosang.f; p.a2 = t->sinang.f;
    }
    return p;
}

// 0 when a float32 throughput kernel exists for this target / shape
static int target_fast_supported(const geosss_target* t) {
    if (t->kind == TGT_VMF_MIX) {
        if (!fast_supported(t->kind, t->d, t->K, t->single)) return 0;           // register-resident kernel
        if (t->K > GC_MAX_K) return -1;
        return gc_smem_bytes(t, gc_params(t)) <= 200 * 1024 ? 0 : -1;           // shared-memory kernel
    }
    if (t->kind == TGT_BINGHAM || t->kind == TGT_CURVED_VMF) {
        if (t->kind == TGT_CURVED_VMF && t->M > GC_MAX_KNOTS) return -1;
        return gc_smem_bytes(t, gc_params(t)) <= 200 * 1024 ? 0 : -1;
    }
    if (t->kind == TGT_CPD) {
        // a finite outlier floor, and either the dense sum or the cutoff kernels (which also serve the k-nearest
        // truncation: kernels_cpd_cell.cuh); without the lists the truncation stays on the literal path
        const bool dense = t->k <= 0 || t->k >= t->ns || cpd_use_grid(t);
        return (dense && t->outlier && std::isfinite(t->log_outlier) && std::isfinite(t->log_const) &&
                (size_t)CPD_WARPS * 4 * t->ns * sizeof(float) <= 200 * 1024) ? 0 : -1;
    }
    return -1;
}

template <class Policy, int KIND, bool STORE> static int launch_gc_one(const GcParams& p, size_t smem, const RunArgs& a, cudaStream_t st) {
    constexpr int BLOCK = gc_block(STORE);
    const unsigned grid = (unsigned)((a.N + BLOCK - 1) / BLOCK);
    CU(cudaFuncSetAttribute(k_gc_fast<Policy, KIND, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gc_fast<Policy, KIND, STORE><<<grid, BLOCK, smem, st>>>(p, a);
    CU(cudaGetLastError());
    return 0;
}

template <class Policy> static int launch_gc(const geosss_target* t, int kind_, const RunArgs& a_, cudaStream_t st) {
    // Quorum of the great-circle kernels: their prologue (d normals, 2 M d score FMAs) outweighs the bracket loop, so the
    // prologue should run once per transition with all 32 lanes -- lock-step -- unlike k_vmf_fast (22).  Measured on B200,
    // GeoSSS-shrink, quorum 22 -> 32: Bingham d=50 diagonal +22 %, dense +25 %, vMF mixture d=100 +18 %, curved vMF d=100 +5 %.
    // Scheduling only: chains are bit-identical for every quorum.
    RunArgs a = a_;
    const int kind = kind_;
    if (!std::getenv("GEOSSS_QUORUM")) a.quorum = 32;  // reject too: Bingham d=50 diagonal, 65 536 chains, 2.8 -> 4.1e9 samples/s
    // ... except the curved vMF on wide rows, whose bracket loop (nine segments per proposal) weighs a quarter of a transition:
    // quorum 28 measured +5 % on C3 in round 2 (20 / 24 / 28 / 32: 1.83 / 1.86 / 1.94 / 1.85e9 samples/s)
    if (!std::getenv("GEOSSS_QUORUM") && t->kind == TGT_CURVED_VMF && !gc_store_normals(t) && kind == SHRINK) a.quorum = 28;
    const GcParams p = gc_params(t);
    const size_t smem = gc_smem_bytes(t, p);
    if (a.n_steps > 0x7fffffff) return fail(GEOSSS_E_INVALID, "run: at most 2^31-1 transitions per call");
    if (gc_store_normals(t))
        return kind == SHRINK ? launch_gc_one<Policy, SHRINK, true>(p, smem, a, st) : launch_gc_one<Policy, REJECT, true>(p, smem, a, st);
    return kind == SHRINK ? launch_gc_one<Policy, SHRINK, false>(p, smem, a, st) : launch_gc_one<Policy, REJECT, false>(p, smem, a, st);
}

static int target_launch_fast(const geosss_target* t, int kind, const RunArgs& a, cudaStream_t st) {
    if (t->kind == TGT_VMF_MIX && fast_supported(t->kind, t->d, t->K, t->single)) {
        switch ((t->K + 3) / 4) {  // one instantiation per component-quad count
            case 1: return launch_gc<VmfGcT<1, true>>(t, kind, a, st);
            case 2: return launch_gc<VmfGcT<2, true>>(t, kind, a, st);
            case 3: return launch_gc<VmfGcT<3, true>>(t, kind, a, st);
            default: return launch_gc<VmfGcT<4, true>>(t, kind, a, st);
        }
    }
    if (t->kind == TGT_VMF_MIX) {
        int rc = launch_fast(t->kind, kind, t->d, t->K, t->h_mu.data(), t->h_const.data(), a, st);
        if (rc > 0) return fail(GEOSSS_E_CUDA, "run: %s", cudaGetErrorString((cudaError_t)rc));
        if (rc < 0) return fail(GEOSSS_E_UNSUPPORTED, "run: fast kernel unavailable for this shape");
        return 0;
    }
    if (t->kind == TGT_BINGHAM) return launch_gc<BinghamGc>(t, kind, a, st);
    if (t->kind == TGT_CURVED_VMF) {
        switch ((t->M + 3) / 4) {  // one instantiation per knot-quad count
            case 1: return launch_gc<CurvedGcT<1, true>>(t, kind, a, st);
            case 2: return launch_gc<CurvedGcT<2, true>>(t, kind, a, st);
            case 3: return launch_gc<CurvedGcT<3, true>>(t, kind, a, st);
            default: return launch_gc<CurvedGcT<4, true>>(t, kind, a, st);
        }
    }
    if (t->kind == TGT_CPD && !target_fast_supported(t) && cpd_use_grid(t)) {
        const CpdCellGrid g = cpd_grid_now(t);
        const size_t smem = g.smem_bytes();
        const unsigned grid = (unsigned)((a.N + CPD_CELL_WARPS - 1) / CPD_CELL_WARPS);
        auto kern = kind == SHRINK ? (g.k > 0 ? k_sampler_cpd_cell<SHRINK, true> : k_sampler_cpd_cell<SHRINK, false>)
                                   : (g.k > 0 ? k_sampler_cpd_cell<REJECT, true> : k_sampler_cpd_cell<REJECT, false>);
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, CPD_CELL_WARPS * 32, smem, st>>>(g, a);
        CU(cudaGetLastError());
        return 0;
    }
    if (t->kind == TGT_CPD && !target_fast_supported(t)) {
        const size_t smem = (size_t)CPD_WARPS * 4 * t->ns * sizeof(float);
        const unsigned grid = (unsigned)((a.N + CPD_WARPS - 1) / CPD_WARPS);
        if (kind == SHRINK) {
            CU(cudaFuncSetAttribute(k_sampler_cpd_fast<SHRINK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_sampler_cpd_fast<SHRINK><<<grid, CPD_WARPS * 32, smem, st>>>(t->cpd<float>(), (float)t->max_logw, a);
        } else {
            CU(cudaFuncSetAttribute(k_sampler_cpd_fast<REJECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_sampler_cpd_fast<REJECT><<<grid, CPD_WARPS * 32, smem, st>>>(t->cpd<float>(), (float)t->max_logw, a);
        }
        CU(cudaGetLastError());
        return 0;
    }
    return fail(GEOSSS_E_UNSUPPORTED, "run: no fast kernel for this target");
}

extern "C" int geosss_version(void) { return GEOSSS_VERSION; }
extern "C" const char* geosss_last_error(void) { return g_err; }

// ------------------------------------------------------------------ target creation
extern "C" int geosss_target_create_vmf_mixture(int d, int K, const double* mu, const double* logw,
                                                const double* log_i0, geosss_target_t* out) {
    if (!out || !mu || d < 2 || K < 1) return fail(GEOSSS_E_INVALID, "vmf_mixture: need d >= 2, K >= 1, mu");
    if (!logw && K != 1) return fail(GEOSSS_E_INVALID, "vmf_mixture: logw == NULL means a lone vMF (K == 1)");
    auto* t = new (std::nothrow) geosss_target();
    if (!t) return fail(GEOSSS_E_NOMEM, "out of host memory");
    t->kind = TGT_VMF_MIX; t->d = d; t->K = K; t->single = logw ? 0 : 1;
    std::vector<double> li(K), lw(K, 0.0);
    for (int k = 0; k < K; ++k) {
        if (log_i0) li[k] = log_i0[k];
        else {
            double s = 0;
            for (int j = 0; j < d; ++j) s += mu[(size_t)k * d + j] * mu[(size_t)k * d + j];
            double kap = std::sqrt(s);
            // the reference uses scipy.special.i0, which overflows to inf for kappa >~ 713
            li[k] = kap > 713.0 ? INFINITY : std::log(std::cyl_bessel_i(0.0, kap));
        }
        if (logw) lw[k] = logw[k];
    }
    t->h_mu.assign(mu, mu + (size_t)K * d);
    t->h_const.resize(K);
    for (int k = 0; k < K; ++k) t->h_const[k] = (-1.8378770664093453 - li[k]) + lw[k];
    int rc = t->mu.upload(mu, (size_t)K * d);
    if (!rc) rc = t->logw.upload(lw.data(), K);
    if (!rc) rc = t->logi0.upload(li.data(), K);
    if (!rc) rc = t->consts.upload(t->h_const.data(), K);
    if (rc) { t->release(); delete t; return rc; }
    *out = t;
    return 0;
}

extern "C" int geosss_target_create_bingham(int d, const double* A, const double* b, geosss_target_t* out) {
    if (!out || !A || d < 2) return fail(GEOSSS_E_INVALID, "bingham: need d >= 2 and A");
    bool diag = true;
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            double a = A[(size_t)i * d + j], at = A[(size_t)j * d + i];
            // np.allclose(A, A.T) (geosss/distributions.py:68)
            if (!(std::fabs(a - at) <= 1e-8 + 1e-5 * std::fabs(at))) return fail(GEOSSS_E_INVALID, "bingham: A must be symmetric");
            if (i != j && a != 0.0) diag = false;
        }
    auto* t = new (std::nothrow) geosss_target();
    if (!t) return fail(GEOSSS_E_NOMEM, "out of host memory");
    t->kind = TGT_BINGHAM; t->d = d; t->has_b = b ? 1 : 0; t->diag = diag ? 1 : 0;
    int rc = t->A.upload(A, (size_t)d * d);
    if (!rc && b) rc = t->b.upload(b, d);
    if (!rc && diag) {
        std::vector<double> dg(d);
        for (int i = 0; i < d; ++i) dg[i] = A[(size_t)i * d + i];
        rc = t->Adiag.upload(dg.data(), d);
    }
    if (!rc && !diag && d <= TC_MAX_D) {
        // x^T A x = sum_j w_j (v_j . x)^2: rows of W = eigenvectors (the reference holds the same decomposition, distributions.py:73-76)
        std::vector<double> sym((size_t)d * d), w, V;
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) sym[(size_t)i * d + j] = 0.5 * (A[(size_t)i * d + j] + A[(size_t)j * d + i]);
        eigh_jacobi(d, sym, w, V);
        t->eig_n_out = (!b || d < TcBinghamEpi::NP_MAX) ? d + (b ? 1 : 0) : 0;  // 0: the Fisher row does not fit the tensor-core log_prob
        std::vector<double> rows((size_t)(d + (b ? 1 : 0)) * d);
        for (int j = 0; j < d; ++j)
            for (int k = 0; k < d; ++k) rows[(size_t)j * d + k] = V[(size_t)k * d + j];
        if (b) for (int k = 0; k < d; ++k) rows[(size_t)d * d + k] = b[k];
        t->h_eig_rows.assign(rows.begin(), rows.begin() + (size_t)d * d);
        rc = t->eig_rows.upload(rows.data(), rows.size());
        if (!rc) rc = t->eig_vals.upload(w.data(), d);
        if (!rc) {
            std::vector<double> cols((size_t)d * d);
            for (int j = 0; j < d; ++j)
                for (int k = 0; k < d; ++k) cols[(size_t)k * d + j] = rows[(size_t)j * d + k];
            rc = t->eig_cols.upload(cols.data(), cols.size());
        }
        if (!rc && b) {
            std::vector<double> wb(d, 0.0);
            for (int j = 0; j < d; ++j)
                for (int k = 0; k < d; ++k) wb[j] += rows[(size_t)j * d + k] * b[k];
            rc = t->eig_b.upload(wb.data(), d);
        }
    }
    if (rc) { t->release(); delete t; return rc; }
    *out = t;
    return 0;
}

extern "C" int geosss_target_create_curved_vmf(int d, int n_knots, const double* knots, double kappa, geosss_target_t* out) {
    if (!out || !knots || d < 2 || n_knots < 2) return fail(GEOSSS_E_INVALID, "curved_vmf: need d >= 2 and >= 2 knots");
    auto* t = new (std::nothrow) geosss_target();
    if (!t) return fail(GEOSSS_E_NOMEM, "out of host memory");
    t->kind = TGT_CURVED_VMF; t->d = d; t->M = n_knots; t->kappa = kappa;
    std::vector<double> ang(n_knots - 1), ca(n_knots - 1), sa(n_knots - 1);
    for (int j = 0; j + 1 < n_knots; ++j) {
        double dot = 0;
        for (int i = 0; i < d; ++i) dot += knots[(size_t)j * d + i] * knots[(size_t)(j + 1) * d + i];
        dot = dot < -1 ? -1 : (dot > 1 ? 1 : dot);
        ang[j] = std::acos(dot);  // sphere.distance(a, b), spherical_curve.py:29
        ca[j] = std::cos(ang[j]);
        sa[j] = std::sin(ang[j]);
    }
    int rc = t->knots.upload(knots, (size_t)n_knots * d);
    if (!rc) rc = t->ang.upload(ang.data(), ang.size());
    if (!rc) rc = t->cosang.upload(ca.data(), ca.size());
    if (!rc) rc = t->sinang.upload(sa.data(), sa.size());
    if (rc) { t->release(); delete t; return rc; }
    *out = t;
    return 0;
}


// Cell lists for the cutoff kernels (kernels_cpd_cell.cuh).  Cutoff: a source point farther than rc from a
// target row adds less than exp(CPD_CELL_LOG_EPS) / ns of the outlier floor to the row sum:
//   log_const + log w_k - rc^2 / (2 sigma^2)  <=  log_outlier + CPD_CELL_LOG_EPS - log(ns)   for every k.
// Returns 0 and leaves has_grid = 0 when the lists would not pay (few, large cells) or do not fit shared memory.
static int build_cpd_grid(geosss_target* t, const double* target, const double* wt, const double* source, const double* lws) {
    t->has_grid = 0;
    if (t->dim != 3 || !t->outlier || !std::isfinite(t->log_outlier) || !std::isfinite(t->log_const)) return 0;
    const int ns = t->ns, nt = t->nt;
    const double ln_ratio = t->log_const + t->max_logw + std::log((double)ns) - t->log_outlier - CPD_CELL_LOG_EPS;
    const double rc = std::sqrt(2.0 * t->sigma2 * (ln_ratio > 0.25 ? ln_ratio : 0.25));
    const double h = rc * (1.0 + 1e-4);
    double lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        lo[a] = hi[a] = source[a];
        for (int k = 1; k < ns; ++k) { double v = source[(size_t)k * 3 + a]; lo[a] = v < lo[a] ? v : lo[a]; hi[a] = v > hi[a] ? v : hi[a]; }
        if (!std::isfinite(lo[a]) || !std::isfinite(hi[a])) return 0;
    }
    int n[3];
    double cells = 1;
    for (int a = 0; a < 3; ++a) { n[a] = (int)std::floor((hi[a] - lo[a]) / h) + 3; cells *= n[a]; }
    if (cells > 32768) return 0;
    const int n_cells = n[0] * n[1] * n[2];
    auto cell_of = [&](const double* x, int c[3]) {
        for (int a = 0; a < 3; ++a) {
            c[a] = (int)std::floor((x[a] - (lo[a] - h)) / h);
            c[a] = c[a] < 1 ? 1 : (c[a] > n[a] - 2 ? n[a] - 2 : c[a]);
        }
    };
    std::vector<std::vector<int>> bucket(n_cells);
    for (int k = 0; k < ns; ++k) {
        int c[3];
        cell_of(source + (size_t)k * 3, c);
        bucket[(c[2] * n[1] + c[1]) * n[0] + c[0]].push_back(k);
    }
    const double m0 = t->log_outlier > t->log_const + t->max_logw ? t->log_outlier : t->log_const + t->max_logw;
    const double L2E = 1.4426950408889634;
    std::vector<uint32_t> cell(n_cells);
    std::vector<float4> entry;
    double weighted = 0;
    for (int cz = 0; cz < n[2]; ++cz)
        for (int cy = 0; cy < n[1]; ++cy)
            for (int cx = 0; cx < n[0]; ++cx) {
                const size_t first = entry.size();
                for (int dz = -1; dz <= 1; ++dz)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int x = cx + dx, y = cy + dy, z = cz + dz;
                            if (x < 0 || y < 0 || z < 0 || x >= n[0] || y >= n[1] || z >= n[2]) continue;
                            for (int k : bucket[(z * n[1] + y) * n[0] + x])
                                entry.push_back(make_float4((float)((source[(size_t)k * 3] - (lo[0] - h)) / h),
                                                            (float)((source[(size_t)k * 3 + 1] - (lo[1] - h)) / h),
                                                            (float)((source[(size_t)k * 3 + 2] - (lo[2] - h)) / h),
                                                            (float)((lws[k] + t->log_const - m0) * L2E)));
                        }
                const size_t cnt = entry.size() - first;
                if (cnt >= (1u << (32 - CPD_CELL_COUNT_SHIFT)) || first >= (1u << CPD_CELL_COUNT_SHIFT)) return 0;
                cell[(cz * n[1] + cy) * n[0] + cx] = (uint32_t)first | ((uint32_t)cnt << CPD_CELL_COUNT_SHIFT);
                weighted += (double)cnt * bucket[(cz * n[1] + cy) * n[0] + cx].size();
            }
    const double mean_cand = weighted / ns;  // candidates met by a row sitting on a source point
    CpdCellGrid g{};
    for (int a = 0; a < 3; ++a) { g.org[a] = (float)(lo[a] - h); g.n[a] = n[a]; g.shift[a] = 0.f; }
    g.inv_h = (float)(1.0 / h);
    g.n_cells = n_cells; g.n_entries = (int)entry.size(); g.nt = nt;
    g.q2 = (float)(-0.5 * L2E * h * h / t->sigma2);  // distances are measured in cell units
    g.base = (float)std::exp(t->log_outlier - m0);
    g.m0 = (float)m0;
    g.beta = (float)t->beta;
    g.rc2 = (float)((rc / h) * (rc / h));
    g.k = (t->k <= 0 || t->k >= ns) ? 0 : t->k;
    if (mean_cand > 0.5 * ns || g.smem_bytes() > 200 * 1024 || entry.empty()) return 0;
    std::vector<float4> tgt4(nt);
    for (int l = 0; l < nt; ++l)
        tgt4[l] = make_float4((float)target[(size_t)l * 3], (float)target[(size_t)l * 3 + 1], (float)target[(size_t)l * 3 + 2], (float)wt[l]);
    CU(cudaMalloc(&t->d_cell, cell.size() * sizeof(uint32_t)));
    CU(cudaMalloc(&t->d_entry, entry.size() * sizeof(float4)));
    CU(cudaMalloc(&t->d_tgt4, tgt4.size() * sizeof(float4)));
    CU(cudaMemcpy(t->d_cell, cell.data(), cell.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(t->d_entry, entry.data(), entry.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(t->d_tgt4, tgt4.data(), tgt4.size() * sizeof(float4), cudaMemcpyHostToDevice));
    g.cell = static_cast<const uint32_t*>(t->d_cell);
    g.entry = static_cast<const float4*>(t->d_entry);
    g.tgt4 = static_cast<const float4*>(t->d_tgt4);
    t->grid = g;
    t->grid_rc = rc;
    t->grid_mean_candidates = mean_cand;
    t->has_grid = 1;
    return 0;
}

static int create_cpd(int dim, int n_t, int n_s, const double* target_in, const double* w_t, const double* source,
                      const double* w_s, double sigma, double omega, double beta, int k, int outlier, geosss_target_t* out) {
    if (!out || !target_in || !source || n_t < 1 || n_s < 1) return fail(GEOSSS_E_INVALID, "cpd: need target and source clouds");
    // 2-D targets are stored as (x, y, 0) rows
    std::vector<double> padded;
    const double* target = target_in;
    if (dim == 2) {
        padded.assign((size_t)n_t * 3, 0.0);
        for (int l = 0; l < n_t; ++l) { padded[(size_t)l * 3] = target_in[(size_t)l * 2]; padded[(size_t)l * 3 + 1] = target_in[(size_t)l * 2 + 1]; }
        target = padded.data();
    }
    if (k > 0 && k < n_s && k > CPD_KMAX) return fail(GEOSSS_E_UNSUPPORTED, "cpd: k-nearest truncation supports k <= %d", CPD_KMAX);
    if (!(sigma > 0)) return fail(GEOSSS_E_INVALID, "cpd: sigma must be positive");
    auto* t = new (std::nothrow) geosss_target();
    if (!t) return fail(GEOSSS_E_NOMEM, "out of host memory");
    t->kind = TGT_CPD; t->d = 4; t->nt = n_t; t->ns = n_s; t->k = k; t->outlier = outlier ? 1 : 0;
    t->sigma2 = sigma * sigma; t->beta = beta; t->dim = dim;
    const double D = (double)dim;  // target.dim (registration.py:116, 213)
    const double gauss = D / 2 * std::log(2 * M_PI * sigma * sigma);
    if (outlier) {
        // registration.py:205-215, 237-238
        double vol = 0;
        for (int a = 0; a < dim; ++a) {
            double lo = target[a], hi = target[a];
            for (int l = 1; l < n_t; ++l) { double v = target[(size_t)l * 3 + a]; lo = v < lo ? v : lo; hi = v > hi ? v : hi; }
            vol += std::log(hi - lo);
        }
        t->log_outlier = std::log(omega + 1e-308) - vol;
        t->log_const = std::log(1 - omega) - gauss;
    } else {
        t->log_outlier = -INFINITY;
        t->log_const = -gauss;  // registration.py:116-118
    }
    std::vector<double> wt(n_t, 1.0), lws(n_s, 0.0);
    if (w_t) wt.assign(w_t, w_t + n_t);
    if (w_s) for (int i = 0; i < n_s; ++i) lws[i] = std::log(w_s[i]);
    t->max_logw = lws[0];
    for (int i = 1; i < n_s; ++i) t->max_logw = lws[i] > t->max_logw ? lws[i] : t->max_logw;
    int rc = t->tgt.upload(target, (size_t)n_t * 3);
    if (!rc) rc = t->src.upload(source, (size_t)n_s * 3);
    if (!rc) rc = t->wt.upload(wt.data(), n_t);
    if (!rc) rc = t->logws.upload(lws.data(), n_s);
    if (!rc) rc = build_cpd_grid(t, target, wt.data(), source, lws.data());
    if (rc) { t->release(); delete t; return rc; }
    *out = t;
    return 0;
}

extern "C" int geosss_target_create_cpd(int n_t, int n_s, const double* target, const double* w_t, const double* source,
                                        const double* w_s, double sigma, double omega, double beta, int k, int outlier,
                                        geosss_target_t* out) {
    return create_cpd(3, n_t, n_s, target, w_t, source, w_s, sigma, omega, beta, k, outlier, out);
}

extern "C" int geosss_target_create_cpd_projected(int n_t, int n_s, const double* target_xy, const double* w_t, const double* source,
                                                  const double* w_s, double sigma, double omega, double beta, int k, int outlier,
                                                  geosss_target_t* out) {
    return create_cpd(2, n_t, n_s, target_xy, w_t, source, w_s, sigma, omega, beta, k, outlier, out);
}

extern "C" int geosss_target_cpd_set_translation(geosss_target_t t, const double* translation) {
    if (!t || t->kind != TGT_CPD) return fail(GEOSSS_E_INVALID, "cpd_set_translation: not a registration target");
    for (int a = 0; a < 3; ++a) t->shift[a] = (translation && a < t->dim) ? translation[a] : 0.0;
    return 0;
}

extern "C" int geosss_target_cpd_grid_info(geosss_target_t t, double* cutoff, double* mean_candidates, int* n_cells, int* n_entries) {
    if (!t || t->kind != TGT_CPD) return fail(GEOSSS_E_INVALID, "cpd_grid_info: not a registration target");
    if (cutoff) *cutoff = t->has_grid ? t->grid_rc : 0.0;
    if (mean_candidates) *mean_candidates = t->has_grid ? t->grid_mean_candidates : 0.0;
    if (n_cells) *n_cells = t->has_grid ? t->grid.n_cells : 0;
    if (n_entries) *n_entries = t->has_grid ? t->grid.n_entries : 0;
    return t->has_grid;
}

extern "C" int geosss_target_dim(geosss_target_t t) { return t ? t->d : GEOSSS_E_INVALID; }

// The eigenvector rows W (d, d) the float32 dense-Bingham kernels work in (chain coordinates y = W x); returns 1 when such a
// system is in use for this target, 0 otherwise (rows_out untouched).
extern "C" int geosss_target_bingham_eigenrows(geosss_target_t t, double* rows_out) {
    if (!t || t->kind != TGT_BINGHAM) return fail(GEOSSS_E_INVALID, "bingham_eigenrows: not a Bingham target");
    if (!bingham_eig(t)) return 0;
    if (rows_out) std::memcpy(rows_out, t->h_eig_rows.data(), t->h_eig_rows.size() * sizeof(double));
    return 1;
}

extern "C" int geosss_target_destroy(geosss_target_t t) {
    if (!t) return 0;
    t->release();
    delete t;
    return 0;
}

// ------------------------------------------------------------------ log_prob
template <typename T> static int logp_literal(geosss_target* t, const void* x, void* out, int64_t N, cudaStream_t st) {
    const T* xi = static_cast<const T*>(x);
    T* o = static_cast<T*>(out);
    const unsigned grid = (unsigned)((N + LANE_BLOCK - 1) / LANE_BLOCK);
    switch (t->kind) {
        case TGT_VMF_MIX: k_logp_lane<T><<<grid, LANE_BLOCK, 0, st>>>(t->vmf<T>(), xi, o, N); break;
        case TGT_BINGHAM: k_logp_lane<T><<<grid, LANE_BLOCK, 0, st>>>(t->bingham<T>(), xi, o, N); break;
        case TGT_CURVED_VMF: k_logp_lane<T><<<grid, LANE_BLOCK, 0, st>>>(t->curved<T>(), xi, o, N); break;
        case TGT_CPD: {
            const size_t smem = (size_t)CPD_WARPS * 3 * t->ns * sizeof(T);
            if (smem > 200 * 1024) return fail(GEOSSS_E_UNSUPPORTED, "cpd: source cloud too large for shared memory");
            CU(cudaFuncSetAttribute(k_logp_cpd<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_logp_cpd<T><<<(unsigned)((N + CPD_WARPS - 1) / CPD_WARPS), CPD_WARPS * 32, smem, st>>>(t->cpd<T>(), xi, o, N);
            break;
        }
        default: return fail(GEOSSS_E_INVALID, "unknown target kind");
    }
    CU(cudaGetLastError());
    return 0;
}

static int g_sm_count = 0;

template <class Policy, int ROWS> static int launch_logp_tiled_rows(const typename Policy::Params& p, const float* x, float* out, int64_t N, cudaStream_t st) {
    const size_t smem = ((size_t)LP_STAGES * ROWS * lp_row_stride(Policy::dim(p)) + Policy::smem_floats(p)) * sizeof(float);
    if (smem > 200 * 1024) return fail(GEOSSS_E_UNSUPPORTED, "log_prob: tile does not fit shared memory");
    if (!g_sm_count) {
        int dev = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    CU(cudaFuncSetAttribute(k_logp_tiled<Policy, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_logp_tiled<Policy, ROWS>, ROWS, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t n_tiles = (N + ROWS - 1) / ROWS;
    int64_t grid = (int64_t)g_sm_count * per_sm;  // persistent: a multiple of the SM count
    if (grid > n_tiles) grid = n_tiles;
    k_logp_tiled<Policy, ROWS><<<(unsigned)grid, ROWS, smem, st>>>(p, x, out, N);
    CU(cudaGetLastError());
    return 0;
}

// wide rows (d > 48) use 64-row tiles: a two-stage ring of 128 x d floats would leave one or two blocks per SM
template <class Policy> static int launch_logp_tiled(const typename Policy::Params& p, const float* x, float* out, int64_t N, cudaStream_t st) {
    return Policy::dim(p) > 48 ? launch_logp_tiled_rows<Policy, 64>(p, x, out, N, st) : launch_logp_tiled_rows<Policy, LP_ROWS>(p, x, out, N, st);
}

// small mixtures (d <= 16, K <= 8): parameters in the constant bank (instantiations in logp_vmf_const.cu)
static int logp_vmf_const(geosss_target* t, const float* x, float* o, int64_t N, cudaStream_t st) {
    const VmfFastParams P = make_vmf_fast_params(t->d, t->K, t->h_mu.data(), t->h_const.data());
    if (!g_sm_count) {
        int dev = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const int rc = launch_logp_vmf_const(t->d, t->K, P, x, o, N, g_sm_count, st);
    if (rc > 0) return fail(GEOSSS_E_CUDA, "log_prob: %s", cudaGetErrorString((cudaError_t)rc));
    if (rc < 0) return fail(GEOSSS_E_UNSUPPORTED, "log_prob: no constant-bank kernel for this shape");
    return 0;
}

// Tensor-core log_prob (kernels_logp_tc.cuh): whole 128-row tiles of a 16-byte aligned input; the ragged tail (and an
// unaligned input) goes to the SIMT tiled kernel passed as `simt`.  GEOSSS_NO_TC=1 keeps everything on the SIMT kernels.
template <class Epi, class Simt>
static int logp_tc(const TcpParams& tp, const float* x, float* out, int64_t N, cudaStream_t st, Simt simt) {
    int64_t full = N / TCP_ROWS;
    if (full == 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return simt(x, out, N);
    if (!g_sm_count) {
        int dev = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    TcpParams tq = tp;
    tq.sub = std::getenv("GEOSSS_TC_SUB1") ? 1 : tcp_sub(tp.d, tp.np, Epi::AUX, Epi::ONE_ACC);  // GEOSSS_TC_SUB1=1: 128-row tiles (A/B runs)
    if (const char* e = std::getenv("GEOSSS_TC_SUBCAP")) { const int v = std::atoi(e); if (v >= 1 && v < tq.sub) tq.sub = v; }
    while (tq.sub > 1 && full < 2 * (int64_t)g_sm_count * tq.sub) --tq.sub;  // small batches: keep every SM busy
    full = N / (TCP_ROWS * tq.sub);
    int R = tcp_stages(tp.d, tp.np, Epi::AUX, tq.sub);
    if (const char* e = std::getenv("GEOSSS_TC_STAGES")) { const int v = std::atoi(e); if (v >= 2 && v <= R) R = v; }
    const size_t smem = tcp_smem_bytes(tp.d, tp.np, Epi::AUX, R, tq.sub);
    CU(cudaFuncSetAttribute(k_logp_tc<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = full < g_sm_count ? full : g_sm_count;  // persistent: one block per SM (all 512 TMEM columns)
    k_logp_tc<Epi><<<(unsigned)grid, TCP_THREADS, smem, st>>>(tq, x, out, full, R);
    CU(cudaGetLastError());
    const int64_t done = full * TCP_ROWS * tq.sub;
    return done < N ? simt(x + done * tp.d, out + done, N - done) : 0;
}

static bool tc_enabled() { return !std::getenv("GEOSSS_NO_TC"); }
static int tc_spin() { const char* e = std::getenv("GEOSSS_TC_SPIN"); return e ? std::atoi(e) : 0; }

static int rotate_rows_tc(geosss_sampler* s, bool to_eig, float* data, int64_t rows, cudaStream_t st) {
    const geosss_target* t = s->t;
    const int d = t->d;
    if (!tc_enabled() || d < 8 || !tcp_fits<TcRotateEpi>(d, d) || (reinterpret_cast<uintptr_t>(data) & 15) != 0 || !t->eig_cols.f) return 1;
    if (!g_sm_count) {
        int dev = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    TcpParams tp{to_eig ? t->eig_rows.f : t->eig_cols.f, nullptr, nullptr, 0.f, d, d, tcp_np(d), tc_spin(), data};
    const int R = tcp_stages(d, tp.np, TcRotateEpi::AUX);
    const size_t smem = tcp_smem_bytes(d, tp.np, TcRotateEpi::AUX, R);
    CU(cudaFuncSetAttribute(k_logp_tc<TcRotateEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t full = rows / TCP_ROWS, tail = rows - full * TCP_ROWS;
    if (full > 0) {
        const int64_t grid = full < g_sm_count ? full : g_sm_count;
        k_logp_tc<TcRotateEpi><<<(unsigned)grid, TCP_THREADS, smem, st>>>(tp, data, nullptr, full, R);
        CU(cudaGetLastError());
    }
    if (tail > 0) {  // the ragged tail goes through one zero-padded scratch tile, so that every row sees the same arithmetic
        const size_t tile_bytes = (size_t)TCP_ROWS * d * sizeof(float);
        if (!s->rot_tmp) CU(cudaMalloc(&s->rot_tmp, tile_bytes));
        float* tmp = static_cast<float*>(s->rot_tmp);
        CU(cudaMemsetAsync(tmp, 0, tile_bytes, st));
        CU(cudaMemcpyAsync(tmp, data + full * TCP_ROWS * d, (size_t)tail * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
        tp.rows_out = tmp;
        k_logp_tc<TcRotateEpi><<<1, TCP_THREADS, smem, st>>>(tp, tmp, nullptr, 1, R);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(data + full * TCP_ROWS * d, tmp, (size_t)tail * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

static int logp_fast(geosss_target* t, const void* x, void* out, int64_t N, cudaStream_t st) {
    LpParams p{};
    p.d = t->d;
    const float* xi = static_cast<const float*>(x);
    float* o = static_cast<float*>(out);
    switch (t->kind) {
        case TGT_VMF_MIX: {
            if (!fast_supported(t->kind, t->d, t->K, t->single)) return logp_vmf_const(t, xi, o, N, st);
            p.K = t->K; p.a0 = t->mu.f; p.a1 = t->consts.f; p.single = t->single;
            // chains x components contraction on the tensor cores for wide rows (K padded to 16 accumulator columns); the ragged
            // tail of fewer than 128 rows runs on the SIMT kernel (K <= 8) or the literal one
            auto simt = [&](const float* xs, float* os, int64_t n) {
                return t->K <= LP_MAX_K ? launch_logp_tiled<VmfLp>(p, xs, os, n, st) : logp_literal<float>(t, xs, os, n, st);
            };
            if (tc_enabled() && t->d >= 24 && tcp_fits<TcVmfEpi>(t->d, t->K)) {
                const TcpParams tp{t->mu.f, t->consts.f, nullptr, 0.f, t->d, t->K, tcp_np(t->K), tc_spin(), nullptr};
                return logp_tc<TcVmfEpi>(tp, xi, o, N, st, simt);
            }
            if (t->K > LP_MAX_K) return fail(GEOSSS_E_UNSUPPORTED, "log_prob: fast vMF kernel handles K <= %d", LP_MAX_K);
            return launch_logp_tiled<VmfLp>(p, xi, o, N, st);
        }
        case TGT_BINGHAM:
            // dense A, 8 <= d <= 64: Y = X A on the tensor cores (tcgen05, 3xTF32), kernels_tc.cuh
            if (!t->diag && t->d >= 8 && t->eig_n_out > 0 && tcp_fits<TcBinghamEpi>(t->d, t->eig_n_out) && tc_enabled() && !std::getenv("GEOSSS_TC_V1")) {
                p.diag = 0; p.a0 = t->A.f; p.a1 = t->has_b ? t->b.f : nullptr;
                const TcpParams tp{t->eig_rows.f, t->eig_vals.f, nullptr, 0.f, t->d, t->eig_n_out, tcp_np(t->eig_n_out), tc_spin(), nullptr};
                return logp_tc<TcBinghamEpi>(tp, xi, o, N, st, [&](const float* xs, float* os, int64_t n) { return launch_logp_tiled<BinghamLp>(p, xs, os, n, st); });
            }
            // round-1 kernel (one block-wide phase at a time), kept selectable for A/B runs: GEOSSS_TC_V1=1
            if (!t->diag && t->d >= 8 && t->d <= TC_MAX_D && tc_enabled()) {
                if (!g_sm_count) {
                    int dev = 0;
                    CU(cudaGetDevice(&dev));
                    CU(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
                }
                const size_t smem = tc_smem_bytes(t->d);
                CU(cudaFuncSetAttribute(k_logp_bingham_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                CU(cudaFuncSetAttribute(k_logp_bingham_tc, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                // persistent blocks: as many as fit an SM's 228 KB of shared memory (1 KB reserved per block), at most 4
                int per_sm = (int)((228 * 1024) / (smem + 1024 + 64));
                if (per_sm < 1) per_sm = 1;
                if (per_sm > 4) per_sm = 4;  // 64 TMEM columns per block
                const int64_t n_tiles = (N + TC_ROWS - 1) / TC_ROWS;
                int64_t grid = (int64_t)g_sm_count * per_sm;
                if (grid > n_tiles) grid = n_tiles;
                TcParams tp{t->A.f, t->has_b ? t->b.f : nullptr, t->d};
                k_logp_bingham_tc<<<(unsigned)grid, TC_THREADS, smem, st>>>(tp, xi, o, N);
                CU(cudaGetLastError());
                return 0;
            }
            p.diag = t->diag; p.a0 = t->diag ? t->Adiag.f : t->A.f; p.a1 = t->has_b ? t->b.f : nullptr;
            return launch_logp_tiled<BinghamLp>(p, xi, o, N, st);
        case TGT_CURVED_VMF:
            if (t->M > GC_MAX_KNOTS) return fail(GEOSSS_E_UNSUPPORTED, "log_prob: fast curved kernel handles <= %d knots", GC_MAX_KNOTS);
            p.M = t->M; p.kappa = (float)t->kappa; p.a0 = t->knots.f; p.a1 = t->cosang.f; p.a2 = t->sinang.f;
            if (tc_enabled() && t->d >= 24 && tcp_fits<TcCurvedEpi>(t->d, t->M)) {  // G = X knots^T on the tensor cores
                const TcpParams tp{t->knots.f, t->cosang.f, t->sinang.f, (float)t->kappa, t->d, t->M, tcp_np(t->M), tc_spin(), nullptr};
                return logp_tc<TcCurvedEpi>(tp, xi, o, N, st, [&](const float* xs, float* os, int64_t n) { return launch_logp_tiled<CurvedLp>(p, xs, os, n, st); });
            }
            return launch_logp_tiled<CurvedLp>(p, xi, o, N, st);
        case TGT_CPD: {
            if (target_fast_supported(t)) return fail(GEOSSS_E_UNSUPPORTED, "log_prob: fast CPD kernel needs the dense sum with an outlier term");
            if (cpd_use_grid(t)) {
                const CpdCellGrid g = cpd_grid_now(t);
                const size_t gsmem = g.smem_bytes();
                int dev = 0, sms = 148;
                CU(cudaGetDevice(&dev));
                CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
                const int64_t want = (N + CPD_CELL_WARPS - 1) / CPD_CELL_WARPS;
                const int64_t cap = (int64_t)sms * (gsmem <= 110 * 1024 ? 2 : 1);
                auto kern = g.k > 0 ? k_logp_cpd_cell<true> : k_logp_cpd_cell<false>;
                CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
                kern<<<(unsigned)(want < cap ? want : cap), CPD_CELL_WARPS * 32, gsmem, st>>>(g, xi, o, N);
                CU(cudaGetLastError());
                return 0;
            }
            const size_t smem = (size_t)CPD_WARPS * 4 * t->ns * sizeof(float);
            CU(cudaFuncSetAttribute(k_logp_cpd_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_logp_cpd_fast<<<(unsigned)((N + CPD_WARPS - 1) / CPD_WARPS), CPD_WARPS * 32, smem, st>>>(t->cpd<float>(), (float)t->max_logw, xi, o, N);
            CU(cudaGetLastError());
            return 0;
        }
        default: return fail(GEOSSS_E_UNSUPPORTED, "log_prob: no fast kernel for this target");
    }
}

extern "C" int geosss_log_prob(geosss_target_t t, int dtype, int path, const void* x, void* out, int64_t N, void* stream) {
    if (!t || N < 0) return fail(GEOSSS_E_INVALID, "log_prob: bad arguments");
    if (N == 0) return 0;
    if (!x || !out) return fail(GEOSSS_E_INVALID, "log_prob: null buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (path == GEOSSS_PATH_FAST) {
        if (dtype != GEOSSS_F32) return fail(GEOSSS_E_UNSUPPORTED, "log_prob: the fast path is float32 only");
        return logp_fast(t, x, out, N, st);
    }
    return dtype == GEOSSS_F64 ? logp_literal<double>(t, x, out, N, st) : logp_literal<float>(t, x, out, N, st);
}

// ------------------------------------------------------------------ samplers
extern "C" int geosss_sampler_create(geosss_target_t t, int kind, int dtype, int path, int64_t N, uint64_t seed,
                                     uint64_t chain_offset, geosss_sampler_t* out) {
    if (!t || !out || N < 1) return fail(GEOSSS_E_INVALID, "sampler_create: bad arguments");
    if (kind < GEOSSS_REJECT || kind > GEOSSS_HMC) return fail(GEOSSS_E_INVALID, "sampler_create: unknown sampler kind");
    const bool mh = kind >= GEOSSS_RWMH;
    // the baseline samplers run on the literal kernels, except on a float32 registration target with cutoff lists (k_mh_cpd_cell)
    const bool mh_cell = mh && path == GEOSSS_PATH_FAST && dtype == GEOSSS_F32 && t->kind == TGT_CPD && cpd_use_grid(t);
    if (mh && path == GEOSSS_PATH_FAST && !mh_cell) return fail(GEOSSS_E_UNSUPPORTED, "sampler_create: this baseline sampler runs on the literal path");
    if (dtype != GEOSSS_F32 && dtype != GEOSSS_F64) return fail(GEOSSS_E_INVALID, "sampler_create: dtype must be f32 or f64");
    if (path == GEOSSS_PATH_FAST) {
        if (dtype != GEOSSS_F32) return fail(GEOSSS_E_UNSUPPORTED, "sampler_create: the fast path is float32 only");
        if (target_fast_supported(t)) return fail(GEOSSS_E_UNSUPPORTED, "sampler_create: no fast kernel for this target / shape");
    } else if (path != GEOSSS_PATH_LITERAL) return fail(GEOSSS_E_INVALID, "sampler_create: unknown path");
    auto* s = new (std::nothrow) geosss_sampler();
    if (!s) return fail(GEOSSS_E_NOMEM, "out of host memory");
    s->t = t; s->kind = kind; s->dtype = dtype; s->path = path; s->N = N; s->seed = seed; s->chain0 = chain_offset;
    s->eig = path == GEOSSS_PATH_FAST && bingham_eig(t);
    if (cudaGetDevice(&s->device) != cudaSuccess) { delete s; return fail(GEOSSS_E_CUDA, "sampler_create: no current CUDA device"); }
    const size_t es = s->esize();
    s->state_bytes = (size_t)N * t->d * es;
    cudaError_t e = g_pool.get(&s->state, s->state_bytes);
    if (e == cudaSuccess) e = g_pool.get(reinterpret_cast<void**>(&s->n_reject), (size_t)N * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(s->n_reject, 0, (size_t)N * sizeof(unsigned long long));
    if (e == cudaSuccess && path == GEOSSS_PATH_LITERAL && t->kind != TGT_CPD) {
        s->ws_bytes = (size_t)(mh ? 5 : 3) * N * t->d * es;
        e = g_pool.get(&s->ws, s->ws_bytes);
    }
    if (e == cudaSuccess && mh) e = cudaMalloc(&s->stepsize, (size_t)N * sizeof(double));
    if (e == cudaSuccess && mh) e = cudaMalloc(&s->adapt_count, (size_t)N * sizeof(int64_t));
    if (e == cudaSuccess && mh) e = cudaMemset(s->adapt_count, 0, (size_t)N * sizeof(int64_t));
    if (e == cudaSuccess && mh) e = cudaMalloc(&s->n_walk, (size_t)N * sizeof(unsigned long long));
    if (e == cudaSuccess && mh) e = cudaMemset(s->n_walk, 0, (size_t)N * sizeof(unsigned long long));
    if (e == cudaSuccess && mh) e = cudaMalloc(&s->momentum, (size_t)N * t->d * es);
    if (e == cudaSuccess && mh) e = cudaMemset(s->momentum, 0, (size_t)N * t->d * es);  // mcmc.py:268: zero momenta
    if (e == cudaSuccess && mh && fill_stepsize(s, kind == GEOSSS_HMC ? 1e-3 : 1e-1)) e = cudaErrorUnknown;  // mcmc.py:116,248
    if (e != cudaSuccess) {
        geosss_sampler_destroy(s);
        return fail(e == cudaErrorMemoryAllocation ? GEOSSS_E_NOMEM : GEOSSS_E_CUDA, "sampler_create: %s", cudaGetErrorString(e));
    }
    *out = s;
    return 0;
}

extern "C" int geosss_sampler_destroy(geosss_sampler_t s) {
    if (!s) return 0;
    int cur = s->device;
    cudaGetDevice(&cur);
    if (cur != s->device) cudaSetDevice(s->device);  // blocks are parked under (and freed on) the device they live on
    cudaDeviceSynchronize();  // what cudaFree would do: nothing of this sampler is in flight when its blocks are parked
    g_pool.put(s->state, s->state_bytes);
    g_pool.put(s->ws, s->ws_bytes);
    g_pool.put(s->n_reject, (size_t)s->N * sizeof(unsigned long long));
    if (s->zbuf) cudaFree(s->zbuf);
    if (s->rot_tmp) cudaFree(s->rot_tmp);
    if (s->stepsize) cudaFree(s->stepsize);
    if (s->adapt_count) cudaFree(s->adapt_count);
    if (s->n_walk) cudaFree(s->n_walk);
    if (s->momentum) cudaFree(s->momentum);
    if (s->fdir) cudaFree(s->fdir);
    if (cur != s->device) cudaSetDevice(cur);
    delete s;
    return 0;
}

extern "C" int geosss_sampler_set_state(geosss_sampler_t s, const void* x, void* stream) {
    if (!s || !x) return fail(GEOSSS_E_INVALID, "set_state: bad arguments");
    CU(cudaMemcpyAsync(s->state, x, (size_t)s->N * s->t->d * s->esize(), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    if (s->eig) return rotate_rows(s, true, s->state, s->N, 1, s->t->d, 0, static_cast<cudaStream_t>(stream));
    return 0;
}

extern "C" int geosss_sampler_get_state(geosss_sampler_t s, void* x, uint64_t* n_reject, void* stream) {
    if (!s) return fail(GEOSSS_E_INVALID, "get_state: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (x) CU(cudaMemcpyAsync(x, s->state, (size_t)s->N * s->t->d * s->esize(), cudaMemcpyDeviceToDevice, st));
    if (x && s->eig) { int rc = rotate_rows(s, false, x, s->N, 1, s->t->d, 0, st); if (rc) return rc; }
    if (n_reject) CU(cudaMemcpyAsync(n_reject, s->n_reject, (size_t)s->N * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// Scalar test function f(x) = x . direction, evaluated by the sampler kernels at every kept state (4 or 8 bytes per
// kept transition instead of d): what the ESS diagnostics need from runs too large to retain (scripts/ess_curve_vMF.py:267,
// scripts/bingham.py:50-52).  Sticky like geosss_sampler_mh_momenta_out: applies to the run calls that follow.
extern "C" int geosss_sampler_trace_out(geosss_sampler_t s, const double* direction_host, void* trace_out) {
    if (!s) return fail(GEOSSS_E_INVALID, "trace_out: bad arguments");
    if (!trace_out) { s->fout = nullptr; return 0; }
    if (!direction_host) return fail(GEOSSS_E_INVALID, "trace_out: a direction is needed with an output buffer");
    if (s->t->kind == TGT_CPD) return fail(GEOSSS_E_UNSUPPORTED, "trace_out: registration chains report log_prob (logp_out) instead");
    const size_t d = (size_t)s->t->d;
    if (!s->fdir) CU(cudaMalloc(&s->fdir, d * sizeof(double)));
    if (s->dtype == GEOSSS_F64) {
        CU(cudaMemcpy(s->fdir, direction_host, d * 8, cudaMemcpyHostToDevice));
    } else {
        std::vector<float> f(d);
        for (size_t i = 0; i < d; ++i) f[i] = (float)direction_host[i];
        if (s->eig) {  // x . v = y . (W v) for the chain's eigenbasis coordinates y = W x
            const std::vector<double>& W = s->t->h_eig_rows;
            for (size_t i = 0; i < d; ++i) {
                double acc = 0;
                for (size_t j = 0; j < d; ++j) acc += W[i * d + j] * direction_host[j];
                f[i] = (float)acc;
            }
        }
        CU(cudaMemcpy(s->fdir, f.data(), d * 4, cudaMemcpyHostToDevice));
    }
    s->fout = trace_out;
    return 0;
}

extern "C" int64_t geosss_sampler_steps_done(geosss_sampler_t s) { return s ? (int64_t)s->step : GEOSSS_E_INVALID; }

template <typename T, int KIND, bool TAPE> static int launch_literal(geosss_sampler* s, const RunArgs& a, cudaStream_t st) {
    geosss_target* t = s->t;
    const unsigned grid = (unsigned)((a.N + LANE_BLOCK - 1) / LANE_BLOCK);
    RunArgs b = a;
    const size_t lsm = t->kind == TGT_CPD ? 0 : lane_scratch_smem<T>(3, t->d);  // x / u / y in shared memory when they fit
    b.ws_smem = lsm ? 1 : 0;
    switch (t->kind) {
        case TGT_VMF_MIX:
            if (lsm > 48 * 1024) CU(cudaFuncSetAttribute(k_sampler_lane<T, KIND, TAPE, VmfMixP<T>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
            k_sampler_lane<T, KIND, TAPE><<<grid, LANE_BLOCK, lsm, st>>>(t->vmf<T>(), b);
            break;
        case TGT_BINGHAM:
            if (lsm > 48 * 1024) CU(cudaFuncSetAttribute(k_sampler_lane<T, KIND, TAPE, BinghamP<T>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
            k_sampler_lane<T, KIND, TAPE><<<grid, LANE_BLOCK, lsm, st>>>(t->bingham<T>(), b);
            break;
        case TGT_CURVED_VMF:
            if (lsm > 48 * 1024) CU(cudaFuncSetAttribute(k_sampler_lane<T, KIND, TAPE, CurvedP<T>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
            k_sampler_lane<T, KIND, TAPE><<<grid, LANE_BLOCK, lsm, st>>>(t->curved<T>(), b);
            break;
        case TGT_CPD: {
            const size_t smem = (size_t)CPD_WARPS * 3 * t->ns * sizeof(T);
            if (smem > 200 * 1024) return fail(GEOSSS_E_UNSUPPORTED, "cpd: source cloud too large for shared memory");
            CU(cudaFuncSetAttribute(k_sampler_cpd<T, KIND, TAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_sampler_cpd<T, KIND, TAPE><<<(unsigned)((a.N + CPD_WARPS - 1) / CPD_WARPS), CPD_WARPS * 32, smem, st>>>(t->cpd<T>(), a);
            break;
        }
        default: return fail(GEOSSS_E_INVALID, "unknown target kind");
    }
    CU(cudaGetLastError());
    return 0;
}

template <typename T> static int dispatch_mh(geosss_sampler* s, bool tape, const RunArgs& a, cudaStream_t st) {
    geosss_target* t = s->t;
    int rc;
    switch (t->kind) {
        case TGT_VMF_MIX: rc = launch_mh<T>(t->vmf<T>(), s->kind, tape, a, st); break;
        case TGT_BINGHAM: rc = launch_mh<T>(t->bingham<T>(), s->kind, tape, a, st); break;
        case TGT_CURVED_VMF: rc = launch_mh<T>(t->curved<T>(), s->kind, tape, a, st); break;
        case TGT_CPD: rc = launch_mh<T>(t->cpd<T>(), s->kind, tape, a, st); break;
        default: return fail(GEOSSS_E_INVALID, "unknown target kind");
    }
    if (rc < 0) return fail(GEOSSS_E_UNSUPPORTED, "cpd: source cloud too large for shared memory");
    if (rc > 0) return fail(GEOSSS_E_CUDA, "run: %s", cudaGetErrorString((cudaError_t)rc));
    return 0;
}

template <bool TAPE> static int dispatch_literal(geosss_sampler* s, const RunArgs& a, cudaStream_t st) {
    if (s->is_mh()) return s->dtype == GEOSSS_F64 ? dispatch_mh<double>(s, TAPE, a, st) : dispatch_mh<float>(s, TAPE, a, st);
    if (s->dtype == GEOSSS_F64)
        return s->kind == GEOSSS_SHRINK ? launch_literal<double, SHRINK, TAPE>(s, a, st) : launch_literal<double, REJECT, TAPE>(s, a, st);
    return s->kind == GEOSSS_SHRINK ? launch_literal<float, SHRINK, TAPE>(s, a, st) : launch_literal<float, REJECT, TAPE>(s, a, st);
}

static RunArgs base_args(geosss_sampler* s, int64_t n_steps, int64_t thin) {
    RunArgs a{};
    a.N = s->N; a.n_steps = n_steps; a.thin = thin; a.seed = s->seed; a.chain0 = s->chain0; a.step0 = s->step;
    a.state = s->state; a.ws = s->ws; a.n_reject = s->n_reject;
    a.stepsize = s->stepsize; a.adapt_count = s->adapt_count; a.n_walk = s->n_walk; a.momentum = s->momentum;
    a.momenta_out = s->momenta_out; a.fdir = s->fdir; a.fout = s->fout; a.adapt_total = s->adapt_total; a.alpha = s->alpha; a.n_leap = s->n_leap;
    // quorum of the persistent bracket loop (kernels_fast.cuh); tuned on B200, see DESIGN.md
    a.ss_step = s->N * (int64_t)s->t->d; a.ss_chain = s->t->d; a.sl_step = s->N; a.sl_chain = 1;  // step-major
    a.quorum = s->kind == GEOSSS_SHRINK ? 22 : 20;
    if (const char* q = std::getenv("GEOSSS_QUORUM")) { int v = std::atoi(q); if (v >= 1 && v <= 32) a.quorum = v; }
    return a;
}

// the launch itself: scratch of the great-circle kernels, then the fast or the literal kernels
static int run_launch(geosss_sampler* s, RunArgs& a, cudaStream_t st) {
    // wide rows on the great-circle kernels (d >= 64, normals not kept in shared memory): scratch for the transition's normals
    if (s->path == GEOSSS_PATH_FAST && s->t->kind != TGT_CPD && !s->zbuf && !gc_store_normals(s->t) && !std::getenv("GEOSSS_GC_REGEN") &&
        !(s->t->kind == TGT_VMF_MIX && !fast_supported(s->t->kind, s->t->d, s->t->K, s->t->single))) {
        s->zbuf_bytes = (size_t)((s->t->d + 3) / 4) * (size_t)s->N * 16;
        if (cudaMalloc(&s->zbuf, s->zbuf_bytes) != cudaSuccess) { cudaGetLastError(); s->zbuf = nullptr; }  // fall back to regeneration
    }
    a.zbuf = s->zbuf;
    if (s->path == GEOSSS_PATH_FAST && s->is_mh()) {  // random-walk / independence / mixture MH and HMC on the registration cutoff lists
        const CpdCellGrid g = cpd_grid_now(s->t);
        const size_t smem = g.smem_bytes();
        const unsigned grid = (unsigned)((a.N + CPD_CELL_WARPS - 1) / CPD_CELL_WARPS);
        auto kern = s->kind == GEOSSS_HMC ? (g.k > 0 ? k_mh_cpd_cell<true, true> : k_mh_cpd_cell<false, true>)
                                          : (g.k > 0 ? k_mh_cpd_cell<true, false> : k_mh_cpd_cell<false, false>);
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, CPD_CELL_WARPS * 32, smem, st>>>(g, s->kind, a);
        CU(cudaGetLastError());
        return 0;
    }
    return s->path == GEOSSS_PATH_FAST ? target_launch_fast(s->t, s->kind, a, st) : dispatch_literal<false>(s, a, st);
}

// Sampler.sample on device buffers (mcmc.py:55-77): block_out is one contiguous (N, head_rows + n_steps / thin, d) array; with
// head_rows = 1 row 0 of every chain is its current state -- "the first retained sample is the current state" -- and the kept
// states follow.  Dense-Bingham chains (eigenbasis of A) rotate the whole block on the tensor cores afterwards, the head rows
// included, so that they carry the same arithmetic as the last row of the previous call.
extern "C" int geosss_sampler_run_block(geosss_sampler_t s, int64_t n_steps, int64_t thin, int64_t head_rows, void* block_out, void* stream) {
    if (!s || n_steps < 0 || thin < 1 || head_rows < 0 || head_rows > 1 || !block_out) return fail(GEOSSS_E_INVALID, "run_block: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t n_keep = n_steps / thin, rows = head_rows + n_keep, d = s->t->d;
    const size_t es = s->esize();
    if (rows == 0) return 0;
    if (head_rows)
        CU(cudaMemcpy2DAsync(block_out, (size_t)rows * d * es, s->state, (size_t)d * es, (size_t)d * es, (size_t)s->N, cudaMemcpyDeviceToDevice, st));
    if (n_steps > 0) {
        RunArgs a = base_args(s, n_steps, thin);
        a.ss_step = d; a.ss_chain = rows * d; a.sl_step = 1; a.sl_chain = n_keep;
        a.samples = n_keep > 0 ? static_cast<unsigned char*>(block_out) + (size_t)head_rows * d * es : nullptr;
        const int rc = run_launch(s, a, st);
        if (rc) return rc;
    }
    if (s->eig) {
        const int rc = rotate_rows(s, false, block_out, s->N * rows, 1, d, 0, st);
        if (rc) return rc;
    }
    s->step += (uint64_t)n_steps;
    return 0;
}

extern "C" int geosss_sampler_run(geosss_sampler_t s, int64_t n_steps, int64_t thin, void* samples_out, void* logp_out,
                                  uint32_t* nrej_out, void* stream) {
    return geosss_sampler_run_layout(s, n_steps, thin, GEOSSS_STEP_MAJOR, 0, 0, samples_out, logp_out, nrej_out, stream);
}

extern "C" int geosss_sampler_run_layout(geosss_sampler_t s, int64_t n_steps, int64_t thin, int layout, int64_t chain_stride,
                                         int64_t scalar_chain_stride, void* samples_out, void* logp_out, uint32_t* nrej_out,
                                         void* stream) {
    if (!s || n_steps < 0 || thin < 1) return fail(GEOSSS_E_INVALID, "run: need n_steps >= 0 and thin >= 1");
    if (layout != GEOSSS_STEP_MAJOR && layout != GEOSSS_CHAIN_MAJOR) return fail(GEOSSS_E_INVALID, "run: unknown output layout");
    if (n_steps == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RunArgs a = base_args(s, n_steps, thin);
    if (layout == GEOSSS_CHAIN_MAJOR) {
        const int64_t n_keep = n_steps / thin;
        a.ss_step = s->t->d; a.ss_chain = chain_stride > 0 ? chain_stride : n_keep * (int64_t)s->t->d;
        a.sl_step = 1; a.sl_chain = scalar_chain_stride > 0 ? scalar_chain_stride : n_keep;
        if (a.ss_chain < n_keep * (int64_t)s->t->d || a.sl_chain < n_keep) return fail(GEOSSS_E_INVALID, "run: chain stride smaller than a chain's output");
    }
    a.samples = samples_out; a.logp = logp_out; a.nrej = nrej_out;
    int rc = run_launch(s, a, st);
    if (rc) return rc;
    if (s->eig && samples_out) {  // retained samples back into the caller's coordinates, in place
        // rows in memory order: the smaller stride is the inner index, so that a block's 128 rows are neighbours
        rc = a.ss_step <= a.ss_chain ? rotate_rows(s, false, samples_out, s->N, n_steps / thin, a.ss_chain, a.ss_step, st)
                                     : rotate_rows(s, false, samples_out, n_steps / thin, s->N, a.ss_step, a.ss_chain, st);
        if (rc) return rc;
    }
    s->step += (uint64_t)n_steps;
    return 0;
}

extern "C" int geosss_sampler_run_tape(geosss_sampler_t s, int64_t n_steps, const double* tape, const int64_t* tape_offsets,
                                       void* samples_out, uint32_t* nrej_out, int32_t* status, void* stream) {
    if (!s || n_steps < 0 || !tape || !tape_offsets) return fail(GEOSSS_E_INVALID, "run_tape: bad arguments");
    if (s->path != GEOSSS_PATH_LITERAL) return fail(GEOSSS_E_UNSUPPORTED, "run_tape: tape mode runs on the literal path");
    if (n_steps == 0) return 0;
    RunArgs a = base_args(s, n_steps, 1);
    a.samples = samples_out; a.nrej = nrej_out; a.tape = tape; a.tape_off = tape_offsets; a.status = status;
    int rc = dispatch_literal<true>(s, a, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
    s->step += (uint64_t)n_steps;
    return 0;
}

// ------------------------------------------------------------------ host-buffer entry points
extern "C" int geosss_sampler_set_state_host(geosss_sampler_t s, const double* x_host) {
    if (!s || !x_host) return fail(GEOSSS_E_INVALID, "set_state_host: bad arguments");
    const size_t n = (size_t)s->N * s->t->d;
    if (s->dtype == GEOSSS_F64) {
        CU(cudaMemcpy(s->state, x_host, n * 8, cudaMemcpyHostToDevice));
    } else {
        std::lock_guard<std::mutex> lock(g_pinned.mu);
        float* f = static_cast<float*>(g_pinned.ensure(n * 4));
        if (!f) return fail(GEOSSS_E_NOMEM, "set_state_host: no pinned staging memory");
        for (size_t i = 0; i < n; ++i) f[i] = (float)x_host[i];
        CU(cudaMemcpy(s->state, f, n * 4, cudaMemcpyHostToDevice));
    }
    if (s->eig) {
        int rc = rotate_rows(s, true, s->state, s->N, 1, s->t->d, 0, nullptr);
        if (rc) return rc;
        CU(cudaStreamSynchronize(nullptr));
    }
    return 0;
}

extern "C" int geosss_sampler_get_state_host(geosss_sampler_t s, double* x_host, uint64_t* n_reject_host) {
    if (!s) return fail(GEOSSS_E_INVALID, "get_state_host: bad arguments");
    const size_t n = (size_t)s->N * s->t->d;
    if (x_host) {
        if (s->dtype == GEOSSS_F64) {
            CU(cudaMemcpy(x_host, s->state, n * 8, cudaMemcpyDeviceToHost));
        } else {
            std::vector<float> f(n);
            if (s->eig) {  // back to the caller's coordinates through a scratch copy
                void* tmp = nullptr;
                CU(cudaMalloc(&tmp, n * 4));
                cudaError_t e = cudaMemcpy(tmp, s->state, n * 4, cudaMemcpyDeviceToDevice);
                int rc = e == cudaSuccess ? rotate_rows(s, false, tmp, s->N, 1, s->t->d, 0, nullptr) : 0;
                if (e == cudaSuccess && !rc) e = cudaMemcpy(f.data(), tmp, n * 4, cudaMemcpyDeviceToHost);
                cudaFree(tmp);
                if (rc) return rc;
                if (e != cudaSuccess) return fail(GEOSSS_E_CUDA, "get_state_host: %s", cudaGetErrorString(e));
            } else {
                CU(cudaMemcpy(f.data(), s->state, n * 4, cudaMemcpyDeviceToHost));
            }
            for (size_t i = 0; i < n; ++i) x_host[i] = (double)f[i];
        }
    }
    if (n_reject_host) CU(cudaMemcpy(n_reject_host, s->n_reject, (size_t)s->N * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return 0;
}

namespace {
// Streams, events and the two device staging buffers of the host-buffer pipeline.  They are kept
// per device between calls (allocation and release of 2 x 64 MB cost more than the kernels they
// serve); a call takes the pipe of the current device under a lock and grows it when needed.
struct HostPipe {
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t filled[2] = {nullptr, nullptr}, drained[2] = {nullptr, nullptr};
    void* buf[2] = {nullptr, nullptr};
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (!compute) {
            CU(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking));
            CU(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) {
                CU(cudaEventCreateWithFlags(&filled[i], cudaEventDisableTiming));
                CU(cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming));
            }
        }
        if (bytes > cap) {
            for (int i = 0; i < 2; ++i) {
                if (buf[i]) CU(cudaFree(buf[i]));
                buf[i] = nullptr;
            }
            cap = 0;
            for (int i = 0; i < 2; ++i) CU(cudaMalloc(&buf[i], bytes));
            cap = bytes;
        }
        return 0;
    }
};
constexpr int MAX_DEVICES = 64;
HostPipe g_pipes[MAX_DEVICES];
std::mutex g_pipe_mutex[MAX_DEVICES];  // one pipeline per device; calls on different devices do not serialise
}  // namespace

// n_skip transitions without output, then n_keep transitions whose states stream to the host:
// chunk i is sampled into buf[i & 1] on one stream while chunk i-1 drains over PCIe on another.
extern "C" int geosss_sampler_run_host(geosss_sampler_t s, int64_t n_skip, int64_t n_keep, void* samples_host) {
    if (!s || n_skip < 0 || n_keep < 0 || (n_keep > 0 && !samples_host)) return fail(GEOSSS_E_INVALID, "run_host: bad arguments");
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) return fail(GEOSSS_E_UNSUPPORTED, "run_host: device index out of range");
    std::lock_guard<std::mutex> lock(g_pipe_mutex[dev]);
    HostPipe& h = g_pipes[dev];
    CU(cudaDeviceSynchronize());  // earlier work on other streams (set_state, run) must have landed
    const size_t row = (size_t)s->N * s->t->d * s->esize();  // one retained state of every chain
    int64_t chunk = (int64_t)((size_t)(64u << 20) / row);
    if (chunk < 1) chunk = 1;
    if (n_keep > 0 && chunk > n_keep) chunk = n_keep;
    int rc = h.ensure(n_keep > 0 ? (size_t)chunk * row : 0);
    if (rc) return rc;
    if (n_skip > 0) {
        rc = geosss_sampler_run(s, n_skip, 1, nullptr, nullptr, nullptr, h.compute);
        if (rc) return rc;
    }
    if (n_keep > 0) {
        unsigned char* out = static_cast<unsigned char*>(samples_host);
        int64_t done = 0;
        for (int i = 0; done < n_keep; ++i) {
            const int b = i & 1;
            const int64_t n = n_keep - done < chunk ? n_keep - done : chunk;
            if (i >= 2) CU(cudaStreamWaitEvent(h.compute, h.drained[b], 0));
            rc = geosss_sampler_run(s, n, 1, h.buf[b], nullptr, nullptr, h.compute);
            if (rc) return rc;
            CU(cudaEventRecord(h.filled[b], h.compute));
            CU(cudaStreamWaitEvent(h.copy, h.filled[b], 0));
            CU(cudaMemcpyAsync(out + (size_t)done * row, h.buf[b], (size_t)n * row, cudaMemcpyDeviceToHost, h.copy));
            CU(cudaEventRecord(h.drained[b], h.copy));
            done += n;
        }
    }
    CU(cudaStreamSynchronize(h.compute));
    CU(cudaStreamSynchronize(h.copy));
    return 0;
}

extern "C" int geosss_sample_host(geosss_target_t t, int kind, int dtype, int path, int64_t N, const double* x0_host,
                                  uint64_t seed, uint64_t chain_offset, int64_t n_samples, int64_t burnin,
                                  void* samples_host, uint64_t* n_reject_host) {
    if (!t || !x0_host || !samples_host || N < 1) return fail(GEOSSS_E_INVALID, "sample_host: bad arguments");
    if (n_samples < 1) return fail(GEOSSS_E_INVALID, "sample_host: n_samples must be positive");  // mcmc.py:62
    if (burnin < 0) return fail(GEOSSS_E_INVALID, "sample_host: burnin must be >= 0");            // mcmc.py:18
    geosss_sampler_t s = nullptr;
    int rc = geosss_sampler_create(t, kind, dtype, path, N, seed, chain_offset, &s);
    if (rc) return rc;
    rc = geosss_sampler_set_state_host(s, x0_host);
    if (!rc && s->is_mh()) s->adapt_total = burnin;  // MetropolisHastings.sample: reset(burnin), mcmc.py:171
    // the first retained state is the initial state itself when there is no burn-in (mcmc.py:67-77)
    unsigned char* out = static_cast<unsigned char*>(samples_host);
    int64_t n_skip = burnin - 1, n_keep = n_samples;
    if (!rc && burnin == 0) {
        const size_t n = (size_t)N * t->d;
        if (dtype == GEOSSS_F64) std::memcpy(out, x0_host, n * 8);
        else { float* f = reinterpret_cast<float*>(out); for (size_t i = 0; i < n; ++i) f[i] = (float)x0_host[i]; }
        out += n * s->esize();
        n_skip = 0;
        n_keep = n_samples - 1;
    }
    if (!rc) rc = geosss_sampler_run_host(s, n_skip, n_keep, out);
    if (!rc && n_reject_host) rc = geosss_sampler_get_state_host(s, nullptr, n_reject_host);
    geosss_sampler_destroy(s);
    return rc;
}

// ------------------------------------------------------------------ baseline samplers: parameters
extern "C" int geosss_sampler_mh_configure(geosss_sampler_t s, double stepsize, double mixing_probability, int n_leapfrog) {
    if (!s || !s->is_mh()) return fail(GEOSSS_E_INVALID, "mh_configure: not a baseline sampler");
    if (stepsize > 0) { int rc = fill_stepsize(s, stepsize); if (rc) return rc; }
    else if (stepsize != 0) return fail(GEOSSS_E_INVALID, "mh_configure: stepsize must be positive (0 keeps the current values)");  // mcmc.py:98
    if (mixing_probability >= 0) {
        if (mixing_probability > 1) return fail(GEOSSS_E_INVALID, "mh_configure: mixing probability must lie in [0, 1]");
        s->alpha = mixing_probability;
    }
    if (n_leapfrog > 0) s->n_leap = n_leapfrog;
    return 0;
}

extern "C" int geosss_sampler_mh_reset(geosss_sampler_t s, int64_t adapt_steps) {
    if (!s || !s->is_mh() || adapt_steps < 0) return fail(GEOSSS_E_INVALID, "mh_reset: need a baseline sampler and adapt_steps >= 0");
    CU(cudaMemset(s->adapt_count, 0, (size_t)s->N * sizeof(int64_t)));
    s->adapt_total = adapt_steps;
    return 0;
}

extern "C" int geosss_sampler_mh_momenta_out(geosss_sampler_t s, void* momenta_out) {
    if (!s || s->kind != GEOSSS_HMC) return fail(GEOSSS_E_INVALID, "mh_momenta_out: not an HMC sampler");
    s->momenta_out = momenta_out;
    return 0;
}

extern "C" int geosss_sampler_mh_get_host(geosss_sampler_t s, double* stepsize_host, uint64_t* n_walk_host, double* momentum_host) {
    if (!s || !s->is_mh()) return fail(GEOSSS_E_INVALID, "mh_get_host: not a baseline sampler");
    CU(cudaDeviceSynchronize());
    if (stepsize_host) CU(cudaMemcpy(stepsize_host, s->stepsize, (size_t)s->N * sizeof(double), cudaMemcpyDeviceToHost));
    if (n_walk_host) CU(cudaMemcpy(n_walk_host, s->n_walk, (size_t)s->N * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (momentum_host) {
        const size_t n = (size_t)s->N * s->t->d;
        if (s->dtype == GEOSSS_F64) {
            CU(cudaMemcpy(momentum_host, s->momentum, n * 8, cudaMemcpyDeviceToHost));
        } else {
            std::vector<float> f(n);
            CU(cudaMemcpy(f.data(), s->momentum, n * 4, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n; ++i) momentum_host[i] = (double)f[i];
        }
    }
    return 0;
}

// ------------------------------------------------------------------ gradient
template <typename T> static int gradient_any(geosss_target* t, const void* x, void* out, int64_t N, cudaStream_t st) {
    const T* xi = static_cast<const T*>(x);
    T* o = static_cast<T*>(out);
    int rc;
    if (sizeof(T) == 4 && t->kind == TGT_CPD && cpd_use_grid(t) && !std::getenv("GEOSSS_GRAD_DENSE")) {  // float32: cutoff lists
        const CpdCellGrid g = cpd_grid_now(t);
        const size_t gsmem = g.smem_bytes();
        int dev = 0, sms = 148;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int64_t want = (N + CPD_CELL_WARPS - 1) / CPD_CELL_WARPS;
        const int64_t cap = (int64_t)sms * (gsmem <= 110 * 1024 ? 2 : 1);
        auto kern = g.k > 0 ? k_grad_cpd_cell<true> : k_grad_cpd_cell<false>;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        kern<<<(unsigned)(want < cap ? want : cap), CPD_CELL_WARPS * 32, gsmem, st>>>(g, reinterpret_cast<const float*>(xi), reinterpret_cast<float*>(o), N);
        CU(cudaGetLastError());
        return 0;
    }
    switch (t->kind) {
        case TGT_VMF_MIX: rc = launch_grad<T>(t->vmf<T>(), xi, o, N, st); break;
        case TGT_BINGHAM: rc = launch_grad<T>(t->bingham<T>(), xi, o, N, st); break;
        case TGT_CURVED_VMF: rc = launch_grad<T>(t->curved<T>(), xi, o, N, st); break;
        case TGT_CPD: rc = launch_grad<T>(t->cpd<T>(), xi, o, N, st); break;
        default: return fail(GEOSSS_E_INVALID, "unknown target kind");
    }
    if (rc < 0) return fail(GEOSSS_E_UNSUPPORTED, "cpd: source cloud too large for shared memory");
    if (rc > 0) return fail(GEOSSS_E_CUDA, "gradient: %s", cudaGetErrorString((cudaError_t)rc));
    return 0;
}

extern "C" int geosss_gradient(geosss_target_t t, int dtype, const void* x, void* out, int64_t N, void* stream) {
    if (!t || N < 0) return fail(GEOSSS_E_INVALID, "gradient: bad arguments");
    if (N == 0) return 0;
    if (!x || !out) return fail(GEOSSS_E_INVALID, "gradient: null buffer");
    if (dtype != GEOSSS_F32 && dtype != GEOSSS_F64) return fail(GEOSSS_E_INVALID, "gradient: dtype must be f32 or f64");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dtype == GEOSSS_F64 ? gradient_any<double>(t, x, out, N, st) : gradient_any<float>(t, x, out, N, st);
}

// ------------------------------------------------------------------ exact i.i.d. samplers (geosss/rand.py)
namespace {
struct DevBuf {  // small scratch freed on scope exit (after the stream has been synchronised)
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};
}  // namespace

extern "C" int geosss_exact_vmf(int d, const double* mu_host, uint64_t seed, uint64_t offset, int dtype, void* out, int64_t N,
                                uint64_t* attempts_host, void* stream) {
    if (d < 2 || !mu_host || N < 0 || (N > 0 && !out)) return fail(GEOSSS_E_INVALID, "exact_vmf: need d >= 2, mu and an output buffer");
    if (dtype != GEOSSS_F32 && dtype != GEOSSS_F64) return fail(GEOSSS_E_INVALID, "exact_vmf: dtype must be f32 or f64");
    if (attempts_host) *attempts_host = 0;
    if (N == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double kap = 0;
    for (int j = 0; j < d; ++j) kap += mu_host[j] * mu_host[j];
    kap = std::sqrt(kap);
    const double p = d - 1;
    ExactVmfArgs a{};
    a.d = d; a.kappa = kap; a.max_attempts = 1 << 14;
    a.b0 = (-2 * kap + std::sqrt(4 * kap * kap + p * p)) / p;  // rand.py:49-51
    a.x0 = (1 - a.b0) / (1 + a.b0);
    a.c = kap * a.x0 + p * std::log(1 - a.x0 * a.x0);
    DevBuf mu, cnt;
    CU(cudaMalloc(&mu.p, (size_t)d * sizeof(double)));
    CU(cudaMalloc(&cnt.p, sizeof(unsigned long long)));
    CU(cudaMemcpyAsync(mu.p, mu_host, (size_t)d * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), st));
    a.mu = static_cast<const double*>(mu.p);
    auto* c = static_cast<unsigned long long*>(cnt.p);
    const int rc = dtype == GEOSSS_F64 ? launch_exact_vmf<double>(a, seed, offset, static_cast<double*>(out), N, c, st)
                                       : launch_exact_vmf<float>(a, seed, offset, static_cast<float*>(out), N, c, st);
    if (rc) return fail(GEOSSS_E_CUDA, "exact_vmf: %s", cudaGetErrorString((cudaError_t)rc));
    unsigned long long tried = 0;
    CU(cudaMemcpyAsync(&tried, cnt.p, sizeof tried, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (attempts_host) *attempts_host = tried;
    return 0;
}

extern "C" int geosss_exact_bingham(int d, const double* v_host, const double* U_host, double b, double log_m, uint64_t seed,
                                    uint64_t offset, int dtype, void* out, int64_t N, int max_attempts, uint64_t* attempts_host,
                                    void* stream) {
    if (d < 2 || !v_host || N < 0 || (N > 0 && !out) || !(b > 0) || max_attempts < 1)
        return fail(GEOSSS_E_INVALID, "exact_bingham: need d >= 2, eigenvalues, b > 0 and an output buffer");
    if (dtype != GEOSSS_F32 && dtype != GEOSSS_F64) return fail(GEOSSS_E_INVALID, "exact_bingham: dtype must be f32 or f64");
    if (attempts_host) *attempts_host = 0;
    if (N == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DevBuf v, U, ws, cnt;
    CU(cudaMalloc(&v.p, (size_t)d * sizeof(double)));
    CU(cudaMalloc(&cnt.p, sizeof(unsigned long long)));
    CU(cudaMemcpyAsync(v.p, v_host, (size_t)d * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), st));
    if (U_host) {
        CU(cudaMalloc(&U.p, (size_t)d * d * sizeof(double)));
        CU(cudaMalloc(&ws.p, (size_t)N * d * sizeof(double)));
        CU(cudaMemcpyAsync(U.p, U_host, (size_t)d * d * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    ExactBinghamArgs a{static_cast<const double*>(v.p), static_cast<const double*>(U.p), b, log_m, d, max_attempts};
    auto* c = static_cast<unsigned long long*>(cnt.p);
    const int rc = dtype == GEOSSS_F64
                       ? launch_exact_bingham<double>(a, seed, offset, static_cast<double*>(out), static_cast<double*>(ws.p), N, c, st)
                       : launch_exact_bingham<float>(a, seed, offset, static_cast<float*>(out), static_cast<double*>(ws.p), N, c, st);
    if (rc) return fail(GEOSSS_E_CUDA, "exact_bingham: %s", cudaGetErrorString((cudaError_t)rc));
    unsigned long long tried = 0;
    CU(cudaMemcpyAsync(&tried, cnt.p, sizeof tried, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (attempts_host) *attempts_host = tried;
    return 0;
}

// ------------------------------------------------------------------ diagnostics
extern "C" int geosss_diag_ess_partials(const void* samples, int dtype, int layout, int64_t n_draws, int64_t n_chains, int d,
                                        const void* direction, int max_lag, double* out, void* stream) {
    if (layout != GEOSSS_STEP_MAJOR && layout != GEOSSS_CHAIN_MAJOR) return fail(GEOSSS_E_INVALID, "diag_ess_partials: unknown layout");
    const int64_t step_stride = layout == GEOSSS_STEP_MAJOR ? n_chains * (int64_t)d : d;
    const int64_t chain_stride = layout == GEOSSS_STEP_MAJOR ? d : n_draws * (int64_t)d;
    if (!samples || !direction || !out || n_draws < 2 || n_chains < 1 || d < 1 || max_lag < 1 || max_lag > n_draws)
        return fail(GEOSSS_E_INVALID, "diag_ess_partials: bad arguments");
    const size_t smem = (size_t)n_draws * sizeof(float);
    if (smem > 200 * 1024) return fail(GEOSSS_E_UNSUPPORTED, "diag_ess_partials: at most 51200 draws per chain per call");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU(cudaMemsetAsync(out, 0, (size_t)(max_lag + 2) * sizeof(double), st));
    if (dtype == GEOSSS_F64) {
        CU(cudaFuncSetAttribute(k_ess_partials<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_ess_partials<double><<<(unsigned)n_chains, DIAG_BLOCK, smem, st>>>(static_cast<const double*>(samples), n_draws, step_stride, chain_stride, d,
                                                                             static_cast<const double*>(direction), max_lag, out);
    } else {
        CU(cudaFuncSetAttribute(k_ess_partials<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_ess_partials<float><<<(unsigned)n_chains, DIAG_BLOCK, smem, st>>>(static_cast<const float*>(samples), n_draws, step_stride, chain_stride, d,
                                                                            static_cast<const float*>(direction), max_lag, out);
    }
    CU(cudaGetLastError());
    return 0;
}

extern "C" int geosss_diag_stream_update(const void* trace, int dtype, int layout, int64_t n_new, int64_t n_chains, int64_t n_before,
                                         int max_lag, double* acc, void* stream) {
    if (layout != GEOSSS_STEP_MAJOR && layout != GEOSSS_CHAIN_MAJOR) return fail(GEOSSS_E_INVALID, "diag_stream_update: unknown layout");
    if (!trace || !acc || n_new < 1 || n_chains < 1 || n_before < 0 || max_lag < 1) return fail(GEOSSS_E_INVALID, "diag_stream_update: bad arguments");
    const size_t smem = (size_t)(max_lag + n_new) * sizeof(float);
    if (smem > 200 * 1024) return fail(GEOSSS_E_UNSUPPORTED, "diag_stream_update: max_lag + chunk length must stay below 51200");
    const int64_t step_stride = layout == GEOSSS_STEP_MAJOR ? n_chains : 1, chain_stride = layout == GEOSSS_STEP_MAJOR ? 1 : n_new;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == GEOSSS_F64) {
        CU(cudaFuncSetAttribute(k_ess_stream_update<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_ess_stream_update<double><<<(unsigned)n_chains, DIAG_BLOCK, smem, st>>>(static_cast<const double*>(trace), n_new, step_stride, chain_stride,
                                                                                  n_before, max_lag, acc);
    } else {
        CU(cudaFuncSetAttribute(k_ess_stream_update<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_ess_stream_update<float><<<(unsigned)n_chains, DIAG_BLOCK, smem, st>>>(static_cast<const float*>(trace), n_new, step_stride, chain_stride,
                                                                                 n_before, max_lag, acc);
    }
    CU(cudaGetLastError());
    return 0;
}

extern "C" int geosss_diag_stream_finish(const double* acc, int64_t n_chains, int64_t n_total, int max_lag, double* out, void* stream) {
    if (!acc || !out || n_chains < 1 || n_total < 2 || max_lag < 1) return fail(GEOSSS_E_INVALID, "diag_stream_finish: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU(cudaMemsetAsync(out, 0, (size_t)(max_lag + 2) * sizeof(double), st));
    k_ess_stream_finish<<<(unsigned)((n_chains + DIAG_BLOCK - 1) / DIAG_BLOCK), DIAG_BLOCK, 0, st>>>(acc, n_chains, n_total, max_lag, out);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int geosss_diag_path_stats(const void* samples, int dtype, int layout, int64_t n_draws, int64_t n_chains, int d,
                                      const double* modes, int n_modes, int sign_labels, double* out, void* stream) {
    if (layout != GEOSSS_STEP_MAJOR && layout != GEOSSS_CHAIN_MAJOR) return fail(GEOSSS_E_INVALID, "diag_path_stats: unknown layout");
    if (!samples || !out || n_draws < 1 || n_chains < 1 || d < 1 || n_modes < 0 || (n_modes > 0 && !modes))
        return fail(GEOSSS_E_INVALID, "diag_path_stats: bad arguments");
    if (n_modes > DIAG_MAX_MODES || (sign_labels && n_modes != 1))
        return fail(GEOSSS_E_UNSUPPORTED, "diag_path_stats: at most 64 modes; sign labels take exactly one direction");
    const int64_t step_stride = layout == GEOSSS_STEP_MAJOR ? n_chains * (int64_t)d : d;
    const int64_t chain_stride = layout == GEOSSS_STEP_MAJOR ? d : n_draws * (int64_t)d;
    const int n_labels = sign_labels ? 3 : n_modes;
    const size_t smem = (size_t)n_modes * d * sizeof(double);
    if (smem > 200 * 1024) return fail(GEOSSS_E_UNSUPPORTED, "diag_path_stats: modes do not fit in shared memory");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU(cudaMemsetAsync(out, 0, (size_t)(n_labels + 2) * sizeof(double), st));
    if (dtype == GEOSSS_F64) {
        CU(cudaFuncSetAttribute(k_path_stats<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_path_stats<double><<<(unsigned)n_chains, DIAG_BLOCK, smem, st>>>(static_cast<const double*>(samples), n_draws, step_stride, chain_stride,
                                                                           d, modes, n_modes, sign_labels, out);
    } else {
        CU(cudaFuncSetAttribute(k_path_stats<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_path_stats<float><<<(unsigned)n_chains, DIAG_BLOCK, smem, st>>>(static_cast<const float*>(samples), n_draws, step_stride, chain_stride,
                                                                          d, modes, n_modes, sign_labels, out);
    }
    CU(cudaGetLastError());
    return 0;
}

#ifdef TCP_TIMERS
extern "C" int geosss_debug_tc_timers(unsigned long long* out) {  // experiment builds only: read and clear the phase timers
    unsigned long long z[32] = {0};
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, geosss::tcp_timers, sizeof(z));
    cudaMemcpyToSymbol(geosss::tcp_timers, z, sizeof(z));
    return 0;
}
#endif
