"""
Generate tests/golden/*.npz by running the UPSTREAM geosss package
(oracle tooling; runs only in the build container where /root/reference exists).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Nothing here is imported at test/bench time.  Each fixture stores the inputs
(parameters, points, initial state, the flat draw tape recorded from the
reference's own PCG64 run) and the reference's outputs (log-densities,
trajectories, cumulative rejection counts), all float64.
"""

import os
import sys

import numpy as np

REF = os.environ.get("GEOSSS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import geosss as gs  # noqa: E402  (the upstream package)
from geosss.distributions import CurvedVonMisesFisher  # noqa: E402
from geosss.pointcloud import PointCloud  # noqa: E402
from geosss.registration import CoherentPointDrift, GaussianMixtureModel  # noqa: E402
from geosss.spherical_curve import SlerpCurve, brownian_curve  # noqa: E402

from oracle.tape import TapeRNG  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)

README_MUS = np.array([[0.87, -0.37, 0.33], [-0.20, -0.89, -0.40], [0.19, 0.22, -0.96]])
README_X0 = np.array([-0.86, 0.19, -0.47])


def points(d, n, seed, scale_some=True):
    """Unit points plus a few deliberately un-normalised rows."""
    x = gs.sample_sphere(d - 1, n, seed=seed)
    if scale_some:
        x[::7] *= 0.998
        x[3::11] *= 1.07
    return x


def vmf_mixture(mus, weights=None):
    return gs.MixtureModel([gs.VonMisesFisher(m) for m in mus], weights)


def record_run(cls, pdf, x0, seed, n_states):
    """Run the upstream sampler with a recording shim; return tape + trajectory."""
    smp = cls(pdf, np.array(x0, dtype=float), seed)
    rec = TapeRNG(source=smp.rng)
    smp.rng = rec
    traj = [np.array(smp.state)]
    cum_rej = [0]
    while len(traj) < n_states:
        traj.append(np.copy(next(smp)))
        cum_rej.append(smp.n_reject)
    return rec.recorded(), np.array(traj), np.array(cum_rej, dtype=np.int64)


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def golden_logprob():
    out = {}
    # C1 README mixture, C2 d=10 K=5 at three kappas, a weighted mixture, a lone vMF
    out["c1_mus"] = 80.0 * README_MUS
    out["c1_x"] = points(3, 96, 11)
    out["c1_logp"] = vmf_mixture(out["c1_mus"]).log_prob(out["c1_x"])
    modes = gs.sample_sphere(9, 5, seed=1234)
    out["c2_modes"] = modes
    out["c2_x"] = points(10, 256, 12)
    for kappa in (50, 100, 500):
        out[f"c2_logp_k{kappa}"] = vmf_mixture(kappa * modes).log_prob(out["c2_x"])
    out["wmix_mus"] = np.array([30.0, 5.0, 120.0, 60.0])[:, None] * gs.sample_sphere(5, 4, seed=5)
    out["wmix_w"] = np.array([0.5, 2.0, 1.0, 0.25])
    out["wmix_x"] = points(6, 64, 13)
    out["wmix_logp"] = vmf_mixture(out["wmix_mus"], out["wmix_w"]).log_prob(out["wmix_x"])
    out["vmf_mu"] = 42.0 * gs.sample_sphere(7, seed=6)
    out["vmf_x"] = points(8, 64, 14)
    out["vmf_logp"] = gs.VonMisesFisher(out["vmf_mu"]).log_prob(out["vmf_x"])
    # overflow quirk: kappa > ~713 -> i0 = inf -> log_prob = -inf
    with np.errstate(over="ignore"):
        out["vmf_big_logp"] = vmf_mixture(800.0 * modes).log_prob(out["c2_x"][:8])
    save("logprob_vmf", **out)

    out = {}
    # C4 diagonal, its dense variant, a small dense one, Bingham-Fisher
    for tag, d, vmax, eig, seed in (("c4diag", 50, 300, True, 6982), ("c4dense", 50, 300, False, 6982),
                                   ("d10", 10, 30, False, 6982)):
        pdf = gs.random_bingham(d, vmax, 0, eig, seed)
        x = points(d, 128, 20 + d)
        out[f"{tag}_A"], out[f"{tag}_x"], out[f"{tag}_logp"] = pdf.A, x, pdf.log_prob(x)
        out[f"{tag}_mode"] = pdf.mode
    A = gs.random_bingham(6, 20, 0, False, 3).A
    b = 4.0 * gs.sample_sphere(5, seed=4)
    x = points(6, 64, 27)
    out["bf_A"], out["bf_b"], out["bf_x"] = A, b, x
    out["bf_logp"] = gs.BinghamFisher(A, b).log_prob(x)
    save("logprob_bingham", **out)

    out = {}
    for tag, d in (("c3", 100), ("d10", 10), ("d3", 3)):
        knots = brownian_curve(10, d, 0.5, seed=4562)
        pdf = CurvedVonMisesFisher(SlerpCurve(knots), 500.0)
        x = np.vstack([points(d, 96, 30 + d), knots[:3], gs.sphere.radial_projection(knots[4] + knots[5])])
        out[f"{tag}_knots"], out[f"{tag}_x"], out[f"{tag}_logp"] = knots, x, pdf.log_prob(x)
    save("logprob_curved", **out)

    out = {}
    data = np.load(os.path.join(REF, "data", "protein_registration.npz"))
    tgt, src = data["target"], data["source"]
    out["c5_target"], out["c5_source"] = tgt, src
    out["c5_rotation"], out["c5_sigma"], out["c5_omega"] = data["rotation"], float(data["sigma"]), float(
        data["prob_outlier"])
    cpd = CoherentPointDrift(PointCloud(tgt, np.ones(len(tgt))), PointCloud(src, np.ones(len(src)) / len(src)),
                             sigma=1.0, k=20, omega=0.4)
    q = points(4, 48, 41)
    from scipy.spatial.transform import Rotation
    q = np.vstack([q, Rotation.from_matrix(data["rotation"]).as_quat(), [0, 0, 0, 1.0]])
    out["c5_q"] = q
    out["c5_logp"] = np.array([cpd.log_prob(row) for row in q])
    # cube case in the style of tests/test_cpd.py:16-141 (k=8, sigma=0.5, omega in {0,.2,.4}), seeded
    gen = np.random.default_rng(77)
    cube = np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)], dtype=float)
    cube = np.vstack([cube, gen.uniform(0, 1, (22, 3))])
    Rtrue = Rotation.from_euler("zyx", [30, -20, 55], degrees=True).as_matrix()
    tcl = cube @ Rtrue.T + 0.02 * gen.standard_normal(cube.shape)
    w_s = gen.uniform(0.5, 1.5, len(cube))
    w_t = gen.uniform(0.5, 1.5, len(cube))
    out["cube_target"], out["cube_source"], out["cube_ws"], out["cube_wt"] = tcl, cube, w_s, w_t
    qc = points(4, 32, 43)
    out["cube_q"] = qc
    for om in (0.0, 0.2, 0.4):
        m = CoherentPointDrift(PointCloud(tcl, w_t), PointCloud(cube, w_s), sigma=0.5, k=8, omega=om, beta=1.5)
        out[f"cube_logp_om{int(om * 10)}"] = np.array([m.log_prob(row) for row in qc])
    g = GaussianMixtureModel(PointCloud(tcl, w_t), PointCloud(cube, w_s), sigma=0.5, k=8)
    out["cube_logp_gmm"] = np.array([g.log_prob(row) for row in qc])
    save("logprob_cpd", **out)
    return tgt, src


def golden_trajectories(tgt, src):
    shrink, reject = gs.ShrinkageSphericalSliceSampler, gs.RejectionSphericalSliceSampler

    def pack(out, tag, cls, pdf, x0, seed, n_states):
        tape, traj, rej = record_run(cls, pdf, x0, seed, n_states)
        out[f"{tag}_x0"], out[f"{tag}_tape"], out[f"{tag}_traj"], out[f"{tag}_cumrej"] = x0, tape, traj, rej

    # C1: the README demo itself (1000 samples + 100 burn-in, seed 3521)
    out = {"mus": 80.0 * README_MUS}
    pdf = vmf_mixture(80.0 * README_MUS)
    pack(out, "shrink", shrink, pdf, README_X0, 3521, 1100)
    pack(out, "reject", reject, pdf, README_X0, 3521, 1100)
    # what the documented call returns (burn-in chopped)
    out["readme_shrink_samples"] = shrink(pdf, README_X0.copy(), 3521).sample(1000, 100)
    out["readme_reject_samples"] = reject(pdf, README_X0.copy(), 3521).sample(1000, 100)
    save("traj_c1", **out)

    # C2: d=10 K=5, kappa 100 and 500, distinct init/sampler seeds
    modes = gs.sample_sphere(9, 5, seed=1234)
    out = {"modes": modes}
    for kappa in (100, 500):
        pdf = vmf_mixture(kappa * modes)
        for s in (0, 1):
            x0 = gs.sample_sphere(9, seed=1000 + s)
            pack(out, f"shrink_k{kappa}_s{s}", shrink, pdf, x0, 2000 + s, 1501)
        pack(out, f"reject_k{kappa}_s0", reject, pdf, gs.sample_sphere(9, seed=1000), 2000, 201)
    save("traj_c2", **out)

    # C3: curved vMF d=100 kappa=500
    knots = brownian_curve(10, 100, 0.5, seed=4562)
    pdf = CurvedVonMisesFisher(SlerpCurve(knots), 500.0)
    out = {"knots": knots, "kappa": 500.0}
    pack(out, "shrink", shrink, pdf, gs.sample_sphere(99, seed=1002), 2002, 201)
    pack(out, "reject", reject, pdf, gs.sample_sphere(99, seed=1003), 2003, 41)
    save("traj_c3", **out)

    # C4: Bingham d=50 (diagonal as in scripts/bingham.py:123-134, start at the mode) + dense variant
    out = {}
    for tag, eig in (("diag", True), ("dense", False)):
        pdf = gs.random_bingham(50, 300, 0, eig, 6982)
        out[f"{tag}_A"] = pdf.A
        pack(out, f"{tag}_shrink", shrink, pdf, pdf.mode.copy(), 2004, 301)
        pack(out, f"{tag}_reject", reject, pdf, pdf.mode.copy(), 2005, 61)
    save("traj_c4", **out)

    # C5: CPD registration on S^3
    cpd = CoherentPointDrift(PointCloud(tgt, np.ones(len(tgt))), PointCloud(src, np.ones(len(src)) / len(src)),
                             sigma=1.0, k=20, omega=0.4)
    out = {}
    pack(out, "shrink", shrink, cpd, gs.sample_sphere(3, seed=1006), 2006, 41)
    pack(out, "reject", reject, cpd, gs.sample_sphere(3, seed=1007), 2007, 9)
    save("traj_c5", **out)


def golden_misc():
    out = {}
    out["sphere_d9_n5_seed1234"] = gs.sample_sphere(9, 5, seed=1234)
    out["sphere_d2_seed7"] = gs.sample_sphere(2, seed=7)
    gen = np.random.default_rng(99)
    z, v = gen.standard_normal((16, 100)), gen.standard_normal((16, 100))
    out["proj_z"], out["proj_v"] = z, v
    out["proj_out"] = np.array([gs.spherical_projection(a, b) for a, b in zip(z, v)])
    out["radial_out"] = gs.radial_projection(z)
    out["dist_out"] = gs.distance(gs.radial_projection(z), gs.radial_projection(v))
    out["bingham_c4_A"] = gs.random_bingham(50, 300, 0, True, 6982).A
    out["bingham_d10_dense_A"] = gs.random_bingham(10, 30, 0, False, 6982).A
    out["brownian_c3"] = brownian_curve(10, 100, 0.5, seed=4562)
    # library ESS helpers on an AR(1) trace
    ar = np.zeros(4000)
    e = gen.standard_normal(4000)
    for i in range(1, 4000):
        ar[i] = 0.8 * ar[i - 1] + e[i]
    out["ar1"], out["ar1_acf"] = ar, gs.acf_fft(ar)
    out["ar1_iat"], out["ar1_neff"] = gs.IAT(ar), gs.n_eff(ar)
    save("misc", **out)




def golden_pathstats():
    """Reference runs + the scripts' diagnostic one-liners applied to them: mode populations / KL
    (scripts/vMF_diagnostics.py:331-342), hopping frequency (scripts/bingham.py:23-25), geodesic steps
    (scripts/curve_vMF.py:379)."""
    shrink = gs.ShrinkageSphericalSliceSampler
    modes = gs.sample_sphere(9, 5, seed=1234)
    pdf = vmf_mixture(100.0 * modes)
    xs = np.array([shrink(pdf, gs.sample_sphere(9, seed=50 + c), 60 + c).sample(1000) for c in range(3)])
    counts, kls, dist = np.zeros(5), [], []
    for x in xs:
        m = np.argmax(x @ modes.T, axis=1)
        i, c = np.unique(m, return_counts=True)
        counts[i] += c
        dist.append(gs.distance(x[:-1], x[1:]))
    x = xs.reshape(-1, 10)  # pooled draws, as the scripts pool their runs
    m = np.argmax(x @ modes.T, axis=1)
    i, c = np.unique(m, return_counts=True)
    p = np.full(len(modes), 1e-100)
    p[i] = c
    p[i] /= p.sum()
    kl = p @ np.log(p / pdf.weights)
    bing = gs.random_bingham(10, 30, 0, True, 6982)
    bx = np.array([shrink(bing, bing.mode.copy(), 70 + c).sample(1000) for c in range(3)])
    hop = np.array([np.mean(np.diff(np.sign(b @ bing.mode)) != 0.0) for b in bx])
    save("pathstats", vmf_modes=modes, vmf_weights=pdf.weights, vmf_samples=xs, vmf_counts=counts, vmf_kl=kl,
         vmf_dist=np.array(dist), bingham_A=bing.A, bingham_mode=bing.mode, bingham_samples=bx, bingham_hop=hop,
         bingham_dist=np.array([gs.distance(b[:-1], b[1:]) for b in bx]))


def record_baseline(make, n_states, adapt):
    """Run an upstream baseline sampler (mcmc.py:115-332) with the recording shim.  ``adapt`` is the
    burn-in length handed to ``reset`` (what ``sample(n, burnin)`` does, mcmc.py:171)."""
    smp = make()
    rec = TapeRNG(source=smp.rng)
    smp.rng = rec
    smp.reset(adapt)
    traj, steps, acc = [np.array(smp.state)], [smp.stepsize], [0]
    while len(traj) < n_states:
        traj.append(np.copy(next(smp)))
        steps.append(smp.stepsize)
        acc.append(smp.n_accept)
    return {"tape": rec.recorded(), "traj": np.array(traj), "stepsize": np.array(steps),
            "cumacc": np.array(acc, dtype=np.int64), "adapt": adapt}


def golden_baselines(tgt, src):
    """Gradients of every target and draw-taped trajectories of RWMH / independence / mixture / HMC."""
    from geosss.mcmc import IndependenceSampler, MetropolisHastings, MixtureRWMHIndependenceSampler, SphericalHMC
    out = {}
    modes = gs.sample_sphere(9, 5, seed=1234)
    c1 = vmf_mixture(80.0 * README_MUS)
    c2 = vmf_mixture(100.0 * modes, np.array([1.0, 2.0, 3.0, 4.0, 5.0]))
    single = gs.VonMisesFisher(40.0 * modes[0])
    dense = gs.random_bingham(10, 30, 0, False, 6982)
    fisher = gs.BinghamFisher(dense.A, 3.0 * modes[1])
    knots = brownian_curve(10, 100, 0.5, seed=4562)
    curved = CurvedVonMisesFisher(SlerpCurve(knots), 500.0)
    n = len(src)
    cpd = CoherentPointDrift(PointCloud(tgt, np.ones(len(tgt))), PointCloud(src, np.ones(n) / n), sigma=1.0, k=20, omega=0.4)
    gmm = GaussianMixtureModel(PointCloud(tgt, np.ones(len(tgt))), PointCloud(src, np.ones(n) / n), sigma=1.0, k=20)
    full = CoherentPointDrift(PointCloud(tgt, np.ones(len(tgt))), PointCloud(src, np.ones(n) / n), sigma=2.0, k=n, omega=0.1, beta=0.5)
    out.update(c1_mus=80.0 * README_MUS, c2_mus=100.0 * modes, c2_weights=c2.weights, single_mu=single.mu, dense_A=dense.A,
               fisher_b=fisher.b, knots=knots, kappa=500.0)

    # gradients (distributions.py:88-89,159-160,223-227,277-278; registration.py:55-59)
    x3, x10, x100 = points(3, 24, 31), points(10, 24, 32), points(100, 12, 33)
    q = np.vstack([points(4, 10, 34), matrix2quat_ref(tgt, src)])
    out.update(x3=x3, x10=x10, x100=x100, q=q)
    out["grad_c1"] = np.array([c1.gradient(x) for x in x3])
    out["grad_c2"] = np.array([c2.gradient(x) for x in x10])
    out["grad_single"] = np.array([single.gradient(x) for x in x10])
    out["grad_dense"] = np.array([dense.gradient(x) for x in x10])
    out["grad_fisher"] = np.array([fisher.gradient(x) for x in x10])
    out["grad_curved"] = np.array([curved.gradient(x) for x in x100])
    out["grad_cpd"] = np.array([cpd.gradient(row) for row in q])
    out["grad_gmm"] = np.array([gmm.gradient(row) for row in q])
    out["grad_full"] = np.array([full.gradient(row) for row in q])

    def pack(tag, rec):
        for k, v in rec.items():
            out[f"{tag}_{k}"] = v

    x0_3, x0_10 = README_X0, gs.sample_sphere(9, seed=1000)
    pack("rwmh_c1", record_baseline(lambda: MetropolisHastings(c1, x0_3.copy(), 11, stepsize=1e-1), 401, 100))
    pack("rwmh_c2", record_baseline(lambda: MetropolisHastings(c2, x0_10.copy(), 12, stepsize=1e-1), 401, 150))
    pack("rwmh_dense", record_baseline(lambda: MetropolisHastings(dense, dense.mode.copy(), 13, stepsize=0.3), 301, 0))
    pack("rwmh_curved", record_baseline(lambda: MetropolisHastings(curved, gs.sample_sphere(99, seed=1002), 14), 201, 80))
    pack("rwmh_cpd", record_baseline(lambda: MetropolisHastings(cpd, gs.sample_sphere(3, seed=1006), 15), 81, 30))
    pack("indep_c1", record_baseline(lambda: IndependenceSampler(c1, x0_3.copy(), 16), 301, 50))
    pack("mix_c2", record_baseline(lambda: MixtureRWMHIndependenceSampler(c2, x0_10.copy(), 17, stepsize=1e-1,
                                                                          mixing_probability=0.5), 401, 150))
    pack("mix_c1", record_baseline(lambda: MixtureRWMHIndependenceSampler(c1, x0_3.copy(), 18, stepsize=0.2,
                                                                          mixing_probability=0.8), 301, 100))
    pack("hmc_c1", record_baseline(lambda: SphericalHMC(c1, x0_3.copy(), 19, stepsize=1e-1), 201, 80))
    pack("hmc_c2", record_baseline(lambda: SphericalHMC(c2, x0_10.copy(), 20, stepsize=1e-2, n_steps=7), 201, 80))
    pack("hmc_dense", record_baseline(lambda: SphericalHMC(dense, dense.mode.copy(), 21, stepsize=5e-2), 151, 50))
    pack("hmc_fisher", record_baseline(lambda: SphericalHMC(fisher, dense.mode.copy(), 22, stepsize=5e-2), 101, 0))
    pack("hmc_curved", record_baseline(lambda: SphericalHMC(curved, gs.sample_sphere(99, seed=1002), 23, stepsize=1e-3), 61, 20))
    pack("hmc_cpd", record_baseline(lambda: SphericalHMC(cpd, gs.sample_sphere(3, seed=1006), 24, stepsize=1e-3), 41, 15))
    # what the documented calls return
    out["rwmh_c1_sample"] = MetropolisHastings(c1, x0_3.copy(), 25).sample(200, 0.2)
    pos, mom = SphericalHMC(c1, x0_3.copy(), 26, stepsize=1e-1).sample(100, 20, return_momenta=True)
    out["hmc_c1_positions"], out["hmc_c1_momenta"] = pos, mom
    save("baselines", **out)


def matrix2quat_ref(tgt, src):
    """Quaternion of the true rotation of the registration data set (data file, 'rotation')."""
    from geosss.pointcloud import matrix2quat
    return matrix2quat(np.load(os.path.join(REF, "data", "protein_registration.npz"))["rotation"])


def golden_exact():
    """Moments of the reference's exact samplers (geosss/rand.py: Wood vMF, Kent Bingham) under a fixed
    global numpy seed: the statistical pin for the device versions."""
    from geosss.rand import sample_bingham, sample_bingham_2d, sample_bingham_3d, sample_vMF
    np.random.seed(20240607)
    out = {}
    modes = gs.sample_sphere(9, 5, seed=1234)
    for tag, mu in (("d10", 40.0 * modes[0]), ("d3", 80.0 * README_MUS[0]), ("d2", np.array([3.0, -4.0]))):
        x = sample_vMF(gs.VonMisesFisher(mu), 20000)
        out[f"vmf_{tag}_mu"], out[f"vmf_{tag}_mean"], out[f"vmf_{tag}_n"] = mu, x.mean(0), len(x)
        out[f"vmf_{tag}_cos_quantiles"] = np.quantile(x @ mu / np.linalg.norm(mu), [0.05, 0.25, 0.5, 0.75, 0.95])
    for tag, pdf in (("d10", gs.random_bingham(10, 30, 0, False, 6982)), ("d50", gs.random_bingham(50, 300, 0, True, 6982)),
                     ("d3", gs.random_bingham(3, 12, 0, False, 11)), ("d2", gs.random_bingham(2, 5, 0, False, 12))):
        x, eff = sample_bingham(pdf.A, 200000, return_efficiency=True)
        out[f"bing_{tag}_A"], out[f"bing_{tag}_second"], out[f"bing_{tag}_eff"], out[f"bing_{tag}_n"] = pdf.A, x.T @ x / len(x), eff, len(x)
    x2 = sample_bingham_2d(gs.random_bingham(2, 5, 0, False, 12), 200000)
    x3 = sample_bingham_3d(gs.random_bingham(3, 12, 0, False, 11), 20000)
    out["bing_d2_second_2d"], out["bing_d3_second_gibbs"] = x2.T @ x2 / len(x2), x3.T @ x3 / len(x3)
    save("exact", **out)


def golden_projection():
    """3D-3D with a translation and 3D-2D (RotationProjection) registration on the toy cube of the
    reference's tests/test_cpd.py: log_prob and gradient at several rotations, omega in {0, 0.2, 0.4}."""
    from geosss.pointcloud import RotationMatrix, RotationProjection, matrix2quat
    gen = np.random.default_rng(77)
    cube = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=float)
    R_true = RotationMatrix().rotation3d(euler_angles=np.array([0.2, 0.3, 0.1]))
    q = np.vstack([matrix2quat(R_true), points(4, 9, 35, scale_some=False)])
    out = {"cube": cube, "R_true": R_true, "q": q}
    for tag, src, t_true, D in (("p3", PointCloud(cube), np.array([0.5, -0.3, 0.2]), 3), ("p2", RotationProjection(cube), np.array([0.5, -0.3]), 2)):
        tgt = src.transform_positions(R_true, t_true) + gen.normal(0, 0.05, (8, D))
        tgt = np.vstack([tgt, gen.uniform(-3, 3, (4, D))])
        out[f"{tag}_target"], out[f"{tag}_t"] = tgt, t_true
        for om in (0, 2, 4):
            cpd = CoherentPointDrift(PointCloud(tgt), src, sigma=0.5, k=8, beta=1.0, omega=om / 10)
            out[f"{tag}_logp_om{om}"] = np.array([cpd.log_prob(row, t_true) for row in q])
            out[f"{tag}_grad_om{om}"] = np.array([cpd.gradient(row, t_true) for row in q])
            out[f"{tag}_logp0_om{om}"] = np.array([cpd.log_prob(row) for row in q])
        out[f"{tag}_logp_matrix"] = CoherentPointDrift(PointCloud(tgt), src, sigma=0.5, k=8, omega=0.2).log_prob(R_true, t_true)
        gmm = GaussianMixtureModel(PointCloud(tgt), src, sigma=0.5, k=8)
        out[f"{tag}_logp_gmm"] = np.array([gmm.log_prob(row, t_true) for row in q])
        out[f"{tag}_grad_gmm"] = np.array([gmm.gradient(row, t_true) for row in q])
    save("projection", **out)


def golden_tessellation(tgt, src):
    """600-cell rotation grid (one subdivision) and the reference CPD log-probabilities on it: the
    'systematic search' of scripts/protein_reg3d3d.py:629-648 at a size that runs in seconds."""
    from geosss.sphere_tesselation import tessellate_rotations
    quats = tessellate_rotations(1)
    cpd = CoherentPointDrift(PointCloud(tgt, np.ones(len(tgt))), PointCloud(src, np.ones(len(src)) / len(src)),
                             sigma=1.0, k=20, omega=0.4)
    logp = np.array([cpd.log_prob(q) for q in quats])
    save("tessellation", quats_n1=quats, logp_n1=logp, n0_count=len(tessellate_rotations(0)),
         n2_count=len(tessellate_rotations(2)))


def golden_api_surface():
    """Names and call signatures of the reference's public interface for the path (drop-in check)."""
    import inspect
    import json
    from geosss import mcmc, distributions, registration, pointcloud, spherical_curve, sphere, utils

    def params(fn):
        return [p for p in inspect.signature(fn).parameters if p != "self"]

    items = {
        "ShrinkageSphericalSliceSampler": mcmc.ShrinkageSphericalSliceSampler,
        "RejectionSphericalSliceSampler": mcmc.RejectionSphericalSliceSampler,
        "MetropolisHastings": mcmc.MetropolisHastings, "IndependenceSampler": mcmc.IndependenceSampler,
        "MixtureRWMHIndependenceSampler": mcmc.MixtureRWMHIndependenceSampler, "SphericalHMC": mcmc.SphericalHMC,
        "VonMisesFisher": distributions.VonMisesFisher, "MixtureModel": distributions.MixtureModel,
        "Bingham": distributions.Bingham, "BinghamFisher": distributions.BinghamFisher,
        "CurvedVonMisesFisher": distributions.CurvedVonMisesFisher,
        "CoherentPointDrift": registration.CoherentPointDrift, "GaussianMixtureModel": registration.GaussianMixtureModel,
        "RotationProjection": pointcloud.RotationProjection, "RotationMatrix": pointcloud.RotationMatrix,
        "PointCloud": pointcloud.PointCloud, "SlerpCurve": spherical_curve.SlerpCurve, "SamplerLauncher": utils.SamplerLauncher,
    }
    surface = {"classes": {}, "functions": {}}
    for name, cls in items.items():
        entry = {"init": params(cls.__init__)}
        for meth in ("sample", "log_prob", "gradient", "run", "run_rwmh", "run_hmc", "reset", "find_nearest", "transform_positions",
                     "create", "rotation2d", "rotation3d"):
            if hasattr(cls, meth):
                fn = getattr(cls, meth)
                entry[meth] = params(getattr(fn, "__wrapped__", fn))
        surface["classes"][name] = entry
    for name, fn in {"determine_burnin": mcmc.determine_burnin, "random_bingham": distributions.random_bingham,
                     "brownian_curve": spherical_curve.brownian_curve, "sample_sphere": sphere.sample_sphere,
                     "radial_projection": sphere.radial_projection, "orthogonal_projection": sphere.orthogonal_projection,
                     "spherical_projection": sphere.spherical_projection, "sample_subsphere": sphere.sample_subsphere,
                     "distance": sphere.distance, "quat2matrix": pointcloud.quat2matrix,
                     "cartesian2polar": sphere.cartesian2polar, "polar2cartesian": sphere.polar2cartesian,
                     "spherical2cartesian": sphere.spherical2cartesian, "cartesian2spherical": sphere.cartesian2spherical,
                     "givens": sphere.givens, "acf": utils.acf, "acf_fft": utils.acf_fft, "IAT": utils.IAT, "n_eff": utils.n_eff,
                     "counter": utils.counter, "take_time": utils.take_time,
                     "jacobian_rotation_matrix": pointcloud.jacobian_rotation_matrix,
                     "sample_vMF": __import__("geosss.rand", fromlist=["x"]).sample_vMF,
                     "sample_bingham": __import__("geosss.rand", fromlist=["x"]).sample_bingham,
                     "sample_bingham_2d": __import__("geosss.rand", fromlist=["x"]).sample_bingham_2d,
                     "sample_bingham_3d": __import__("geosss.rand", fromlist=["x"]).sample_bingham_3d,
                     "tessellate_rotations": __import__("geosss.sphere_tesselation", fromlist=["x"]).tessellate_rotations}.items():
        surface["functions"][name] = params(fn)
    surface["attributes"] = {"Bingham": ["A", "v", "U", "d", "mode", "max_log_prob"], "MixtureModel": ["pdfs", "weights"],
                             "VonMisesFisher": ["mu", "d", "kappa", "mode", "max_log_prob"], "CurvedVonMisesFisher": ["curve", "kappa", "d"],
                             "sampler": ["state", "rng", "target", "n_reject"],
                             "MetropolisHastings": ["state", "rng", "target", "n_accept", "stepsize"],
                             "MixtureRWMHIndependenceSampler": ["alpha", "rwmh_counter", "indep_counter"], "SphericalHMC": ["n_steps"]}
    with open(os.path.join(OUT, "api_surface.json"), "w") as f:
        json.dump(surface, f, indent=1, sort_keys=True)
    print("api_surface.json written")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "api":
        golden_api_surface()
    elif len(sys.argv) > 1 and sys.argv[1] == "pathstats":
        golden_pathstats()
    elif len(sys.argv) > 1 and sys.argv[1] == "baselines":
        d = np.load(os.path.join(REF, "data", "protein_registration.npz"))
        golden_baselines(d["target"], d["source"])
    elif len(sys.argv) > 1 and sys.argv[1] == "projection":
        golden_projection()
    elif len(sys.argv) > 1 and sys.argv[1] == "exact":
        golden_exact()
    elif len(sys.argv) > 1 and sys.argv[1] == "tessellation":
        d = np.load(os.path.join(REF, "data", "protein_registration.npz"))
        golden_tessellation(d["target"], d["source"])
    else:
        with np.errstate(over="ignore", divide="ignore"):
            t, s = golden_logprob()
            golden_trajectories(t, s)
            golden_misc()
            golden_pathstats()
            golden_exact()
            golden_projection()
            golden_baselines(t, s)
            golden_tessellation(t, s)
        golden_api_surface()
