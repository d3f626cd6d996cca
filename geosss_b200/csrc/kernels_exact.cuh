// Exact i.i.d. samplers the reference uses as ground truth (geosss/rand.py), one lane per draw:
//
//   k_exact_vmf      Wood (1994) rejection sampler for vMF(mu)              rand.py:31-66
//   k_exact_bingham  Kent-Ganeiber-Mardia rejection from an angular central  rand.py:132-193
//                    Gaussian envelope for Bingham(A), any dimension
//
// Draw i reads the Philox stream of "chain" offset + i; attempt a of the rejection loop is "step" a, so a
// draw does not depend on the batch it is generated in.  Differences from upstream, by construction: the
// global numpy state is replaced by (seed, index); Beta(p/2, p/2) is formed from 2p normals
// (chi-square ratio); the tangent part is placed around the mode by a Householder reflection instead
// of an SVD rotation (the tangent direction is uniform, so the law is the same).
#pragma once
#include "kernels_literal.cuh"
#include "transition_mh.cuh"

namespace geosss {

struct ExactVmfArgs {
    const double* mu;  // (d) device
    double kappa, b0, x0, c;
    int d;
    int max_attempts;
};

template <typename T>
__global__ void __launch_bounds__(LANE_BLOCK) k_exact_vmf(ExactVmfArgs a, uint64_t seed, uint64_t offset, T* __restrict__ out, int64_t N,
                                                          unsigned long long* __restrict__ attempts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int d = a.d, p = a.d - 1;
    PhiloxSeq<double> dr;
    dr.init(seed, offset + (uint64_t)i);
    double W = 0.0;
    int tries = 0;
    if (a.kappa > 1e-8) {
        for (; tries < a.max_attempts;) {
            dr.begin_step((uint64_t)tries++);
            double s1 = 0.0, s2 = 0.0;
            for (int j = 0; j < p; ++j) { const double z = dr.normal(); s1 += z * z; }
            for (int j = 0; j < p; ++j) { const double z = dr.normal(); s2 += z * z; }
            const double Z = s1 / (s1 + s2);  // Beta(p/2, p/2)
            const double U = log(dr.uniform());
            W = (1.0 - (1.0 + a.b0) * Z) / (1.0 - (1.0 - a.b0) * Z);
            if (a.kappa * W + p * log(1.0 - a.x0 * W) - a.c >= U) break;
        }
    }
    // uniform direction: the tangent part (p normals), or the whole point when kappa = 0
    T* o = out + i * d;
    dr.begin_step((uint64_t)a.max_attempts);
    const int m = a.kappa > 1e-8 ? p : d;
    double nn = 0.0;
    for (int j = 0; j < m; ++j) { const double z = dr.normal(); o[j] = (T)z; nn += z * z; }
    // (re-reading o[j] keeps the lane free of a d-sized local array; T rounding of the raw normals is harmless)
    if (a.kappa > 1e-8) {
        const double r = sqrt(fmax(0.0, 1.0 - W * W)) / sqrt(nn);
        // x = [r v, W] around e_d, then reflect e_d onto u = mu / kappa:  x - 2 h (h.x)/(h.h), h = e_d - u
        double hx = 0.0, hh = 0.0;
        for (int j = 0; j < d; ++j) {
            const double h = (j == p ? 1.0 : 0.0) - a.mu[j] / a.kappa;
            const double xj = j == p ? W : r * (double)o[j];
            hx += h * xj; hh += h * h;
        }
        const double f = hh > 1e-30 ? 2.0 * hx / hh : 0.0;
        for (int j = 0; j < d; ++j) {
            const double h = (j == p ? 1.0 : 0.0) - a.mu[j] / a.kappa;
            const double xj = j == p ? W : r * (double)o[j];
            o[j] = (T)(xj - f * h);
        }
    } else {
        const double r = 1.0 / sqrt(nn);
        for (int j = 0; j < d; ++j) o[j] = (T)(r * (double)o[j]);
    }
    if (attempts) atomicAdd(attempts, (unsigned long long)(tries > 0 ? tries : 1));
}

struct ExactBinghamArgs {
    const double* v;  // (d) eigenvalues of -A shifted to start at 0, device
    const double* U;  // (d, d) eigenvectors in columns, row-major, or null (A given by its diagonal)
    double b, log_m;
    int d;
    int max_attempts;
};

// ws: (N, d) doubles of scratch when U != null (the accepted point in the eigenbasis), else unused
template <typename T>
__global__ void __launch_bounds__(LANE_BLOCK) k_exact_bingham(ExactBinghamArgs a, uint64_t seed, uint64_t offset, T* __restrict__ out, double* __restrict__ ws,
                                                              int64_t N, unsigned long long* __restrict__ attempts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int d = a.d;
    PhiloxSeq<double> dr;
    dr.init(seed, offset + (uint64_t)i);
    int tries = 0;
    double nn = 1.0;
    for (; tries < a.max_attempts;) {
        dr.begin_step((uint64_t)tries++);
        // ACG proposal z_j / sqrt(1 + 2 v_j / b), normalised; the test only needs sums
        nn = 0.0;
        double uv = 0.0;
        for (int j = 0; j < d; ++j) {
            const double z = dr.normal();
            const double y = z / sqrt(1.0 + 2.0 * a.v[j] / a.b);
            nn += y * y; uv += y * y * a.v[j];
        }
        const double u = uv / nn;
        const double log_acc = -u + 0.5 * d * log(1.0 + 2.0 * u / a.b) - a.log_m;
        if (log(dr.uniform()) < log_acc) break;
    }
    // regenerate the accepted proposal from its counter (no d-sized local array)
    dr.begin_step((uint64_t)(tries - 1));
    const double r = 1.0 / sqrt(nn);
    if (a.U) {
        double* y = ws + i * d;
        for (int j = 0; j < d; ++j) y[j] = r * dr.normal() / sqrt(1.0 + 2.0 * a.v[j] / a.b);
        for (int k = 0; k < d; ++k) {  // x = U y  (samples @ U.T, rand.py:186)
            double s = 0.0;
            for (int j = 0; j < d; ++j) s += a.U[(int64_t)k * d + j] * y[j];
            out[i * d + k] = (T)s;
        }
    } else {
        for (int j = 0; j < d; ++j) out[i * d + j] = (T)(r * dr.normal() / sqrt(1.0 + 2.0 * a.v[j] / a.b));
    }
    if (attempts) atomicAdd(attempts, (unsigned long long)tries);
}

// defined in mh.cu; return a cudaError_t value
template <typename T> int launch_exact_vmf(const ExactVmfArgs& a, uint64_t seed, uint64_t offset, T* out, int64_t N, unsigned long long* attempts,
                                           cudaStream_t st);
template <typename T> int launch_exact_bingham(const ExactBinghamArgs& a, uint64_t seed, uint64_t offset, T* out, double* ws, int64_t N,
                                               unsigned long long* attempts, cudaStream_t st);

}  // namespace geosss
