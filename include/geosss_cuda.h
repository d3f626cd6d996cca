/*
 * libgeosss_cuda -- C ABI of the B200 (sm_100a) engine for geosss's geodesic slice-sampling
 * hot path.  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * The reference (microscopic-image-analysis/geosss v0.3.5) is pure Python and has no FFI; the
 * entry points below are what a binding for the path would need, one per reference interface:
 *
 *   geosss_target_create_vmf_mixture   VonMisesFisher.__init__ / MixtureModel.__init__
 *                                      (geosss/distributions.py:126-127, :211-216)
 *   geosss_target_create_bingham       Bingham.__init__ / BinghamFisher.__init__ (:67-76, :108-111)
 *   geosss_target_create_curved_vmf    CurvedVonMisesFisher.__init__ + SlerpCurve.__init__
 *                                      (distributions.py:263-265, spherical_curve.py:79-85)
 *   geosss_target_create_cpd           CoherentPointDrift.__init__ / GaussianMixtureModel.__init__
 *                                      (geosss/registration.py:69-99, :172-203)
 *   geosss_log_prob                    Distribution.log_prob(x) for x of shape (N, d)
 *                                      (distributions.py:78-86, :156-157, :218-221, :272-275,
 *                                       registration.py:47-53)
 *   geosss_sampler_create              Sampler.__init__ (geosss/mcmc.py:28-45, :340-355)
 *   geosss_sampler_set_state/get_state Sampler.state, .n_reject (mcmc.py:44, :355)
 *   geosss_sampler_run                 repeated __next__ inside Sampler.sample (mcmc.py:55-77,
 *                                      :357-374 reject, :382-401 shrink), own Philox stream
 *   geosss_sampler_run_tape            the same transitions fed with recorded draws (parity mode)
 *   geosss_sample_host                 Sampler(pdf, x0, seed).sample(n, burnin) with HOST buffers
 *
 * Conventions: every function returns 0 on success or a negative GEOSSS_E_* code, with a
 * message available from geosss_last_error() (thread-local).  Unless the name ends in _host,
 * data pointers are DEVICE pointers owned by the caller; work is enqueued on the given
 * cudaStream_t (passed as void*; NULL = default stream) and is asynchronous -- no hidden
 * synchronisation or allocation in log_prob / run.  Handles are opaque and belong to the
 * device that was current when they were created; a handle must not be used from two threads
 * at once.  Row-major everywhere; "dtype" selects float32 (fast) or float64 (parity).
 */
#ifndef GEOSSS_CUDA_H
#define GEOSSS_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GEOSSS_VERSION 100 /* 0.1.0 */

enum { GEOSSS_F32 = 0, GEOSSS_F64 = 1 };
enum {
    GEOSSS_REJECT = 0,       /* RejectionSphericalSliceSampler  (geosss/mcmc.py:335-374) */
    GEOSSS_SHRINK = 1,       /* ShrinkageSphericalSliceSampler  (:377-401) */
    GEOSSS_RWMH = 2,         /* MetropolisHastings              (:115-173) */
    GEOSSS_INDEPENDENCE = 3, /* IndependenceSampler             (:176-179) */
    GEOSSS_MIXTURE = 4,      /* MixtureRWMHIndependenceSampler  (:182-234) */
    GEOSSS_HMC = 5           /* SphericalHMC                    (:237-332) */
};
/* LITERAL: reference operation order (any target, either dtype).
 * FAST:    float32 throughput kernels that use the great-circle algebra. */
enum { GEOSSS_PATH_LITERAL = 0, GEOSSS_PATH_FAST = 1 };
/* layout of retained outputs: STEP_MAJOR (n_keep, N, d) -- contiguous per step, what host streaming wants;
 * CHAIN_MAJOR (N, n_keep, d) -- the (chain, draw, dim) array the reference's scripts build
 * (scripts/protein_reg3d3d.py:346-349); every lane appends to its own contiguous stream, so sectors are
 * completed by the chain that opened them. */
enum { GEOSSS_STEP_MAJOR = 0, GEOSSS_CHAIN_MAJOR = 1 };

enum {
    GEOSSS_OK = 0,
    GEOSSS_E_INVALID = -1,     /* bad argument (the reference would assert / raise ValueError) */
    GEOSSS_E_CUDA = -2,        /* a CUDA runtime call failed */
    GEOSSS_E_UNSUPPORTED = -3, /* combination not available (e.g. FAST path in float64) */
    GEOSSS_E_NOMEM = -4
};

typedef struct geosss_target* geosss_target_t;
typedef struct geosss_sampler* geosss_sampler_t;

int geosss_version(void);
const char* geosss_last_error(void);

/* ---- targets (host parameter arrays, float64; copied to the device in both precisions) ---- */

/* mu: (K, d).  logw: (K) log of NORMALISED weights, or NULL for a lone VonMisesFisher (no
 * logsumexp).  log_i0: (K) log(i0(|mu_k|)) as the reference computes it (scipy), may hold +inf;
 * NULL lets the library compute it with std::cyl_bessel_i. */
int geosss_target_create_vmf_mixture(int d, int K, const double* mu, const double* logw,
                                     const double* log_i0, geosss_target_t* out);
/* A: (d, d) symmetric.  b: (d) Fisher term or NULL. */
int geosss_target_create_bingham(int d, const double* A, const double* b, geosss_target_t* out);
/* knots: (n_knots, d), consecutive knots are joined by slerps. */
int geosss_target_create_curved_vmf(int d, int n_knots, const double* knots, double kappa,
                                    geosss_target_t* out);
/* target: (n_t, 3), w_t: (n_t); source: (n_s, 3), w_s: (n_s).  k <= 0 or k >= n_s sums every
 * source point, else the k nearest (k <= 64).  outlier = 0 selects the GaussianMixtureModel
 * parent (no outlier column, omega ignored). */
int geosss_target_create_cpd(int n_t, int n_s, const double* target, const double* w_t,
                             const double* source, const double* w_s, double sigma, double omega,
                             double beta, int k, int outlier, geosss_target_t* out);
/* 3D -> 2D registration (source = RotationProjection, geosss/pointcloud.py:273-293): target_xy: (n_t, 2);
 * the rotated source is projected onto its first two coordinates; constants use target.dim = 2. */
int geosss_target_create_cpd_projected(int n_t, int n_s, const double* target_xy, const double* w_t,
                                       const double* source, const double* w_s, double sigma, double omega,
                                       double beta, int k, int outlier, geosss_target_t* out);
/* Fixed translation added to the rotated (projected) source in later log_prob / gradient / sampler calls:
 * (3) or (2) host doubles, NULL = none (the `translation` argument of Registration.log_prob / gradient,
 * registration.py:47-59; pointcloud.py:260-264). */
int geosss_target_cpd_set_translation(geosss_target_t t, const double* translation);
/* Cell lists of the float32 cutoff kernels (dense sum with an outlier floor, 3-D clouds): returns 1 and fills
 * the cutoff radius, the mean number of candidate source points a target row meets (instead of n_s), and
 * the list sizes; 0 when the target uses the dense pair sum (no floor, 2-D target, lists would not pay or do
 * not fit shared memory).  Replaces the per-call KD-tree of geosss/registration.py:101-109.
 * GEOSSS_CPD_DENSE=1 in the environment forces the dense pair sum. */
int geosss_target_cpd_grid_info(geosss_target_t t, double* cutoff, double* mean_candidates, int* n_cells, int* n_entries);
int geosss_target_dim(geosss_target_t t); /* ambient dimension d of the state */
/* Dense Bingham targets (d <= 64) hold the eigenvector rows W of A (reference: Bingham.U, geosss/distributions.py:67-76): the
 * float32 throughput kernels evaluate x^T A x and run the chains in the coordinates y = W x.  Returns 1 and fills rows_out
 * ((d, d) row-major, may be NULL) when that system is in use, 0 when not (diagonal A, d > 64, GEOSSS_NO_EIG). */
int geosss_target_bingham_eigenrows(geosss_target_t t, double* rows_out);
int geosss_target_destroy(geosss_target_t t);

/* x: (N, d) device, out: (N) device, both of `dtype`. */
int geosss_log_prob(geosss_target_t t, int dtype, int path, const void* x, void* out, int64_t N,
                    void* stream);

/* ---- samplers: N independent chains ------------------------------------------------------ */

/* chain_offset: global id of this sampler's chain 0; the Philox counter uses the global id, so
 * results do not depend on how chains are split over samplers / GPUs. */
int geosss_sampler_create(geosss_target_t t, int kind, int dtype, int path, int64_t N,
                          uint64_t seed, uint64_t chain_offset, geosss_sampler_t* out);
int geosss_sampler_set_state(geosss_sampler_t s, const void* x /* (N, d) dtype */, void* stream);
/* x: (N, d) or NULL; n_reject: (N) running totals or NULL. */
int geosss_sampler_get_state(geosss_sampler_t s, void* x, uint64_t* n_reject, void* stream);
/* Runs n_steps transitions of every chain.  Every thin-th state is kept:
 * samples_out (n_steps / thin, N, d), logp_out (n_steps / thin, N), nrej_out (n_steps / thin, N)
 * rejections accumulated since the previous kept state; any of them may be NULL. */
int geosss_sampler_run(geosss_sampler_t s, int64_t n_steps, int64_t thin, void* samples_out,
                       void* logp_out, uint32_t* nrej_out, void* stream);
/* The same with the output layout chosen: CHAIN_MAJOR writes samples_out (N, n_keep, d), logp_out and
 * nrej_out (N, n_keep).  chain_stride / scalar_chain_stride (in elements, 0 = dense) let a chain's rows
 * land inside a larger (N, n_total, d) array, e.g. behind an initial-state slot. */
int geosss_sampler_run_layout(geosss_sampler_t s, int64_t n_steps, int64_t thin, int layout,
                              int64_t chain_stride, int64_t scalar_chain_stride, void* samples_out,
                              void* logp_out, uint32_t* nrej_out, void* stream);
/* Scalar test function f(x) = x . direction (direction: d float64 on the host), written for every kept state of the
 * run calls that follow into trace_out (device, sampler dtype, same scalar layout as logp_out: (n_keep, N), or
 * (N, n_keep) with CHAIN_MAJOR).  trace_out == NULL switches it off.  Replaces retaining (n, N, d) samples for the
 * ESS of a coordinate / of x . mode (reference: scripts/ess_mixture_vMF.py:233, scripts/ess_curve_vMF.py:267,
 * scripts/bingham.py:50-52, where the full chains are kept and sliced afterwards). */
int geosss_sampler_trace_out(geosss_sampler_t s, const double* direction_host, void* trace_out);
/* Sampler.sample on device buffers (geosss/mcmc.py:55-77): block_out is ONE contiguous (N, head_rows + n_steps / thin, d) array
 * of the sampler's dtype; with head_rows = 1 row 0 of every chain holds its state before the run ("the first retained sample
 * is the current state", mcmc.py:67-68), the kept states follow.  head_rows is 0 or 1. */
int geosss_sampler_run_block(geosss_sampler_t s, int64_t n_steps, int64_t thin, int64_t head_rows, void* block_out, void* stream);
/* Parity mode: draws come from a flat float64 tape per chain, in the order the reference consumes
 * them (d normals, level uniform, [bracket uniform,] one uniform per proposal; next transition).
 * tape_offsets: (N + 1) int64.  status: (N) 0 ok / 1 tape exhausted early, or NULL. */
int geosss_sampler_run_tape(geosss_sampler_t s, int64_t n_steps, const double* tape,
                            const int64_t* tape_offsets, void* samples_out, uint32_t* nrej_out,
                            int32_t* status, void* stream);
/* transitions done so far (the Philox step counter) */
int64_t geosss_sampler_steps_done(geosss_sampler_t s);
/* Parks the sampler's device blocks for the next sampler of the same shape on this device (the reference's
 * usage constructs a sampler per sample() call: geosss/utils.py:198-232); synchronises the device first. */
int geosss_sampler_destroy(geosss_sampler_t s);
/* Returns every parked device block and the pinned staging buffer to the driver. */
int geosss_release_cached_memory(void);

/* ---- HOST-buffer entry points (synchronous; copies are part of the call) ------------------- */

/* x_host: (N, d) float64 on the host, converted to the sampler's dtype. */
int geosss_sampler_set_state_host(geosss_sampler_t s, const double* x_host);
/* x_host: (N, d) float64 or NULL; n_reject_host: (N) or NULL. */
int geosss_sampler_get_state_host(geosss_sampler_t s, double* x_host, uint64_t* n_reject_host);
/* n_skip transitions without output, then n_keep transitions whose states are written to
 * samples_host (n_keep, N, d) of the sampler's dtype (pinned memory recommended).  Sampling of
 * chunk i overlaps the device->host copy of chunk i-1 (two streams, double buffering). */
int geosss_sampler_run_host(geosss_sampler_t s, int64_t n_skip, int64_t n_keep, void* samples_host);

/* Sampler(pdf, x0, seed).sample(n_samples, burnin) for N chains: x0_host (N, d) float64 on the
 * host; samples_host (n_samples, N, d) of `dtype` on the host (pinned memory recommended).
 * Reference semantics per chain (mcmc.py:55-77): the first retained state is the initial state,
 * n_samples + burnin - 1 transitions run, the first `burnin` states are dropped.  The device
 * work is chunked and the device->host copies are overlapped with sampling on two streams.
 * n_reject_host: (N) or NULL. */
int geosss_sample_host(geosss_target_t t, int kind, int dtype, int path, int64_t N,
                       const double* x0_host, uint64_t seed, uint64_t chain_offset,
                       int64_t n_samples, int64_t burnin, void* samples_host,
                       uint64_t* n_reject_host);

/* ---- baseline samplers of the reference's comparison ("next" row f3) ------------------------ */
/* Kinds GEOSSS_RWMH .. GEOSSS_HMC are created with geosss_sampler_create (literal path; also the fast
 * path for a float32 registration target with cutoff lists: k_mh_cpd_cell) and driven by
 * the same run / run_layout / run_tape / run_host / sample_host entry points.  For them the counter
 * reported as n_reject counts ACCEPTED proposals (MetropolisHastings.n_accept, mcmc.py:133,158) and
 * the per-transition output of run() is 1 if the transition accepted.  Tape order per transition:
 * [mixture: choice uniform,] [random walk: the Gamma(d/2) variate,] d normals, acceptance uniform. */

/* stepsize > 0 sets every chain's step size (0 keeps them; defaults 1e-1, HMC 1e-3: mcmc.py:116,248);
 * mixing_probability in [0,1] (< 0 keeps; mcmc.py:194); n_leapfrog > 0 (<= 0 keeps; default 10). */
int geosss_sampler_mh_configure(geosss_sampler_t s, double stepsize, double mixing_probability, int n_leapfrog);
/* AdaptiveStepsize.reset(burnin) (mcmc.py:101-104): the next adapt_steps adapting transitions of every
 * chain multiply its step size by 1.02 (accepted) or 0.98 (rejected). */
int geosss_sampler_mh_reset(geosss_sampler_t s, int64_t adapt_steps);
/* HMC: device buffer (same dtype / layout as the samples of the following run calls) that receives the
 * momenta of the retained states (SphericalHMC.sample(return_momenta=True), mcmc.py:320-332); NULL stops. */
int geosss_sampler_mh_momenta_out(geosss_sampler_t s, void* momenta_out);
/* Host copies (any may be NULL): stepsize (N), random-walk proposals of the mixture sampler (N,
 * rwmh_counter of mcmc.py:199,217), current momenta (N, d). */
int geosss_sampler_mh_get_host(geosss_sampler_t s, double* stepsize_host, uint64_t* n_walk_host, double* momentum_host);

/* Distribution.gradient(x) for x of shape (N, d) (distributions.py:88-89, :159-160, :223-227, :277-278;
 * registration.py:55-59 with pointcloud.py:135-203).  x, out: (N, d) device, `dtype`.  A float32 registration target with
 * cutoff lists is evaluated on the lists (k_grad_cpd_cell; GEOSSS_GRAD_DENSE=1 keeps the dense sum). */
int geosss_gradient(geosss_target_t t, int dtype, const void* x, void* out, int64_t N, void* stream);

/* ---- exact i.i.d. samplers used as ground truth ("next" row f4; geosss/rand.py) ----------------- */
/* Both are synchronous on `stream` (they return the number of proposals tried).  Draw i uses the Philox
 * stream (seed, offset + i), so results do not depend on how a request is split into calls or GPUs. */

/* sample_vMF (rand.py:31-66, Wood 1994): N draws from vMF(mu), mu: (d) host doubles; out: (N, d) device. */
int geosss_exact_vmf(int d, const double* mu_host, uint64_t seed, uint64_t offset, int dtype, void* out, int64_t N,
                     uint64_t* attempts_host, void* stream);
/* sample_bingham (rand.py:132-193, Kent, Ganeiber & Mardia): rejection from the angular central Gaussian
 * envelope.  v_host: (d) eigenvalues of -A shifted so the smallest is 0; U_host: (d, d) row-major matrix with
 * the eigenvectors in its columns, or NULL when A was given by its diagonal; b: root of sum 1/(b + 2 v_j) = 1;
 * log_m: log of the envelope bound.  max_attempts per draw (upstream: n_iter). */
int geosss_exact_bingham(int d, const double* v_host, const double* U_host, double b, double log_m, uint64_t seed,
                         uint64_t offset, int dtype, void* out, int64_t N, int max_attempts, uint64_t* attempts_host,
                         void* stream);

/* ---- diagnostics ("next" row of the scope table: ESS half of the metric) --------------------- */

/* Partial sums for the multi-chain ESS of f(x) = x . direction over samples in STEP_MAJOR
 * (n_draws, n_chains, d) or CHAIN_MAJOR (n_chains, n_draws, d) layout (device, `dtype`); direction: (d)
 * device, `dtype`.  out: (max_lag + 2) doubles on the
 * device: out[l] = sum over chains of the biased autocovariance at lag l (geosss/utils.py:91-112
 * per chain), out[max_lag] = sum of chain means, out[max_lag + 1] = sum of squared chain means.
 * The sums are additive over chain shards, so ranks all-reduce `out` and finish on the host. */
int geosss_diag_ess_partials(const void* samples, int dtype, int layout, int64_t n_draws, int64_t n_chains, int d,
                             const void* direction, int max_lag, double* out, void* stream);

/* Streaming form of geosss_diag_ess_partials for runs that are never retained: a scalar trace (geosss_sampler_trace_out)
 * is produced chunk by chunk and folded into per-chain accumulators acc ((n_chains, 1 + 3 max_lag) device doubles, zeroed
 * by the caller before the first chunk); trace: (n_new, n_chains) or, CHAIN_MAJOR, (n_chains, n_new) of `dtype`; n_before:
 * draws already folded in.  _finish writes the same max_lag + 2 partial sums as geosss_diag_ess_partials for the n_total
 * draws seen.  Replaces keeping whole chains for `arviz.ess` / `acf` (scripts/ess_curve_vMF.py:97-130, utils.py:91-112). */
int geosss_diag_stream_update(const void* trace, int dtype, int layout, int64_t n_new, int64_t n_chains, int64_t n_before,
                              int max_lag, double* acc, void* stream);
int geosss_diag_stream_finish(const double* acc, int64_t n_chains, int64_t n_total, int max_lag, double* out, void* stream);

/* Path statistics of retained draws, same layouts as above.  modes: (n_modes, d) device doubles (NULL
 * when n_modes = 0).  out (device doubles, additive over chain shards):
 *   out[0 .. L-1]  draws per label; label = argmax_k x . mode_k, L = n_modes (mode populations of
 *                  scripts/vMF_diagnostics.py:333-342), or with sign_labels (n_modes = 1) the sign of
 *                  x . mode as {-, 0, +}, L = 3 (scripts/bingham.py:23-25);
 *   out[L]         successive draws of one chain whose labels differ ("hops", bingham.py:25);
 *   out[L + 1]     sum over successive draws of arccos(clip(x_t . x_{t+1}, -1, 1)), the geodesic
 *                  step of geosss/sphere.py:64-68 (scripts/curve_vMF.py:379). */
int geosss_diag_path_stats(const void* samples, int dtype, int layout, int64_t n_draws, int64_t n_chains, int d,
                           const double* modes, int n_modes, int sign_labels, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOSSS_CUDA_H */
