"""
GPU parity of the baseline samplers (random-walk MH, independence, mixture, spherical HMC) and of
``gradient`` through the C ABI, against the reference's own runs (tests/golden/baselines.npz), the
oracle fed with the same Philox draws, and distributional properties.

Tolerances: float64 gradients <= 1e-11 relative, float32 <= 2e-4; float64 tape trajectories <= 1e-8 abs
with identical accept decisions and step sizes.
"""

import numpy as np
import pytest
import torch

import geosss_b200 as gs
from oracle import baselines as ob
from oracle import philox
from test_oracle_baselines import GRADS, RUNS, oracle_targets

pytestmark = pytest.mark.gpu
F32, F64 = np.float32, np.float64

CLASSES = {ob.RWMH: gs.MetropolisHastings, ob.INDEPENDENCE: gs.IndependenceSampler,
           ob.MIXTURE: gs.MixtureRWMHIndependenceSampler, ob.HMC: gs.SphericalHMC}


def device_targets(g, golden):
    c = golden("logprob_cpd")
    n = len(c["c5_source"])
    tgt, src = gs.PointCloud(c["c5_target"]), gs.PointCloud(c["c5_source"], np.ones(n) / n)
    return {
        "c1": gs.MixtureModel([gs.VonMisesFisher(m) for m in g["c1_mus"]]),
        "c2": gs.MixtureModel([gs.VonMisesFisher(m) for m in g["c2_mus"]], g["c2_weights"].copy()),
        "single": gs.VonMisesFisher(g["single_mu"]),
        "dense": gs.Bingham(g["dense_A"]),
        "fisher": gs.BinghamFisher(g["dense_A"], g["fisher_b"]),
        "curved": gs.CurvedVonMisesFisher(gs.SlerpCurve(g["knots"]), float(g["kappa"])),
        "cpd": gs.CoherentPointDrift(tgt, src, sigma=1.0, k=20, omega=0.4),
        "gmm": gs.GaussianMixtureModel(tgt, src, sigma=1.0, k=20),
        "full": gs.CoherentPointDrift(tgt, src, sigma=2.0, k=n, omega=0.1, beta=0.5),
    }


@pytest.mark.parametrize("dtype,tol", [(F64, 1e-11), (F32, 2e-4)])
@pytest.mark.parametrize("name,pts", GRADS)
def test_gradient(golden, name, pts, dtype, tol):
    g = golden("baselines")
    pdf = device_targets(g, golden)[name]
    want = g[f"grad_{name}"]
    got = pdf.gradient(g[pts].astype(dtype))
    scale = np.maximum(1.0, np.max(np.abs(want), axis=1, keepdims=True))
    assert got.dtype == dtype and np.max(np.abs(got - want) / scale) <= tol
    one = pdf.gradient(g[pts][3].astype(dtype))  # a single point, as the reference calls it
    assert one.shape == want[3].shape and np.max(np.abs(one - want[3]) / scale[3]) <= tol
    dev = pdf.gradient(torch.as_tensor(g[pts].astype(dtype)).cuda())
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), got)


@pytest.mark.parametrize("tag,kind,target,kw", RUNS)
def test_tape_trajectory(golden, tag, kind, target, kw):
    """The reference's own draws replayed on the device: same accept decisions, states, step sizes."""
    g = golden("baselines")
    traj = g[f"{tag}_traj"]
    d = traj.shape[1] // 2 if kind == ob.HMC else traj.shape[1]
    n = len(traj) - 1
    smp = CLASSES[kind](device_targets(g, golden)[target], traj[0, :d].copy(), 0, dtype="float64", **kw)
    smp.reset(int(g[f"{tag}_adapt"]))
    x, acc, status = smp.run_tape(g[f"{tag}_tape"], n)
    assert status[0] == 0
    assert np.array_equal(np.cumsum(acc[0]), g[f"{tag}_cumacc"][1:])
    assert np.max(np.abs(x[0] - traj[1:, :d])) <= 1e-8
    assert smp.n_accept == g[f"{tag}_cumacc"][-1]
    assert abs(smp.stepsize - g[f"{tag}_stepsize"][-1]) <= 1e-12 * g[f"{tag}_stepsize"][-1]
    if kind == ob.HMC:
        v = smp.state[d:]
        assert np.max(np.abs(v - traj[-1, d:])) <= 1e-6 * max(1.0, np.max(np.abs(traj[-1, d:])))


@pytest.mark.parametrize("kind,target,kw", [(ob.RWMH, "c2", dict(stepsize=0.1)), (ob.INDEPENDENCE, "c1", {}),
                                            (ob.MIXTURE, "c2", dict(stepsize=0.1, mixing_probability=0.3)),
                                            (ob.HMC, "c2", dict(stepsize=0.02, n_steps=5)), (ob.HMC, "dense", dict(stepsize=0.05)),
                                            (ob.RWMH, "cpd", {}), (ob.HMC, "cpd", dict(stepsize=1e-3, n_steps=4))])
def test_free_running_vs_oracle(golden, kind, target, kw):
    """Philox-driven kernels vs the oracle algorithm fed by oracle/philox.py:PhiloxSeqRNG, several chains,
    a chain offset and a continued run."""
    g = golden("baselines")
    tgt = oracle_targets(g, golden)[target]
    N, n, seed, off = 3, 40 if target == "cpd" else 90, 0xFEED1234ABCD, 11
    x0 = gs.sample_sphere(tgt.d - 1, N, seed=8)
    smp = CLASSES[kind](device_targets(g, golden)[target], x0.copy(), seed, dtype="float64", chain_offset=off, **kw)
    got = smp.sample(n, burnin=n // 3, return_all_samples=True)
    more = smp.sample(11)  # continues; first row is the current state
    for c in range(N):
        ref = ob.Baseline(kind, tgt, x0[c], philox.PhiloxSeqRNG(seed, off + c, F64), **kw)
        want = ref.sample(n, burnin=n // 3, return_all_samples=True)
        assert np.max(np.abs(got[c] - want)) <= 1e-8, c
        ref.reset(0)
        cont = np.array([np.copy(next(ref)) for _ in range(10)])
        assert np.max(np.abs(more[c, 1:] - cont)) <= 1e-8 and np.array_equal(more[c, 0], got[c, -1])
        assert abs(smp.stepsize[c] - ref.stepsize) <= 1e-12 * ref.stepsize
        assert smp.n_accept_per_chain[c] == ref.n_accept


def test_sample_semantics_and_counters(golden):
    g = golden("baselines")
    pdf = device_targets(g, golden)["c1"]
    x0 = np.array([-0.86, 0.19, -0.47])
    type(pdf).log_prob.reset_counters()
    smp = gs.MetropolisHastings(pdf, x0, 5)
    out = smp.sample(200, 0.2)
    assert out.shape == (200, 3) and out.dtype == F32 and smp.path == "literal"
    assert np.allclose(np.linalg.norm(out, axis=1), 1, atol=1e-5)
    assert pdf.log_prob.num_calls == 2 * 239 and 0 < smp.n_accept <= 239
    assert smp.stepsize != 0.1  # adapted during the 40 burn-in transitions
    with pytest.raises(AttributeError):
        smp.n_reject
    many = gs.MixtureRWMHIndependenceSampler(pdf, gs.sample_sphere(2, 64, seed=1), 6, mixing_probability=0.25)
    xs = many.sample(50)
    assert xs.shape == (64, 50, 3) and many.rwmh_counter + many.indep_counter == 49 * 64
    assert 0.15 < many.rwmh_counter / (49 * 64) < 0.35
    hmc = gs.SphericalHMC(pdf, x0, 7, stepsize=0.05, n_steps=6, dtype="float64")
    type(pdf).gradient.reset_counters()
    pos, mom = hmc.sample(30, 10, return_momenta=True)
    assert pos.shape == (30, 3) and mom.shape == (30, 3)
    assert np.max(np.abs(np.sum(pos * mom, axis=1))) < 1e-9  # momenta are tangent
    assert hmc.state.shape == (6,) and np.array_equal(hmc.state[:3], pos[-1]) and np.array_equal(hmc.state[3:], mom[-1])
    assert pdf.gradient.num_calls == 7 * 39
    assert next(hmc).shape == (6,)
    # reference-generated outputs have the same shape / burn-in handling
    assert out.shape == g["rwmh_c1_sample"].shape and pos.shape[1] == g["hmc_c1_positions"].shape[1]
    dev = gs.SphericalHMC(pdf, torch.as_tensor(gs.sample_sphere(2, 8, seed=2)).cuda(), 8, stepsize=0.05).sample(20, 5)
    assert dev.is_cuda and tuple(dev.shape) == (8, 20, 3)
    with pytest.raises(RuntimeError):
        gs.MetropolisHastings(pdf, x0, 5, path="fast").sample(3)


def test_launcher_runs_all_four_methods(golden):
    g = golden("baselines")
    pdf = device_targets(g, golden)["c2"]
    launcher = gs.SamplerLauncher(pdf, gs.sample_sphere(9, seed=3), 100, burnin=0.2, seed=9)
    for method in ("sss-reject", "sss-shrink", "rwmh", "hmc"):
        assert launcher.run(method).shape == (100, 10)
    assert launcher.rwmh.n_accept > 0 and launcher.hmc.stepsize != 1e-1
    with pytest.raises(ValueError):
        launcher.run("gibbs")


def test_stationary_distribution_agrees_with_slice_sampler(golden):
    """All samplers target the same density: mean cosine to the mode of a single vMF (d=10, kappa=40)."""
    g = golden("baselines")
    pdf = device_targets(g, golden)["single"]
    mode = g["single_mu"] / np.linalg.norm(g["single_mu"])
    x0 = np.tile(mode, (4096, 1))

    def mean_cos(smp, n, burnin):
        x = smp.sample(n, burnin)
        return float(np.mean(x[:, n // 2:] @ mode))

    want = mean_cos(gs.ShrinkageSphericalSliceSampler(pdf, x0, 1), 200, 50)
    assert abs(mean_cos(gs.MetropolisHastings(pdf, x0, 2, stepsize=0.3), 600, 200) - want) < 4e-3
    assert abs(mean_cos(gs.SphericalHMC(pdf, x0, 3, stepsize=0.05, n_steps=10), 300, 100) - want) < 4e-3
    assert abs(mean_cos(gs.MixtureRWMHIndependenceSampler(pdf, x0, 4, stepsize=0.3, mixing_probability=0.9), 600, 200) - want) < 4e-3


def test_shard_invariance(golden):
    g = golden("baselines")
    pdf = device_targets(g, golden)["c2"]
    x0 = gs.sample_sphere(9, 10, seed=4)
    whole = gs.SphericalHMC(pdf, x0, 21, stepsize=0.02, dtype="float64").sample(30, 10)
    lo = gs.SphericalHMC(pdf, x0[:4], 21, stepsize=0.02, dtype="float64").sample(30, 10)
    hi = gs.SphericalHMC(pdf, x0[4:], 21, stepsize=0.02, dtype="float64", chain_offset=4).sample(30, 10)
    assert np.array_equal(whole, np.concatenate([lo, hi]))


@pytest.mark.parametrize("cls,kw,tol", [(gs.MetropolisHastings, dict(stepsize=0.05), 1e-5),
                                        (gs.MixtureRWMHIndependenceSampler, dict(stepsize=0.05, mixing_probability=0.7), 1e-5),
                                        (gs.SphericalHMC, dict(stepsize=1e-3, n_steps=4), 1e-4)])
def test_baselines_on_the_registration_cutoff_lists(golden, cls, kw, tol):
    """float32 random-walk MH, its mixture with the independence proposal and spherical HMC on the registration target run on the
    cutoff-list kernel (k_mh_cpd_cell, path 'fast'): same Philox draws, same proposals, a log-density and a gradient that differ
    from the dense float32 sums only by the terms below float32 resolution -- so the chains follow the literal float32 kernel
    step for step (an accept / reject decision within 1e-4 of the threshold may flip; none does at this size), and the adapted
    step sizes and the acceptance counts agree."""
    g = golden("logprob_cpd")
    n = len(g["c5_source"])
    pdf = gs.CoherentPointDrift(gs.PointCloud(g["c5_target"]), gs.PointCloud(g["c5_source"], np.ones(n) / n), sigma=1.0, k=20, omega=0.4)
    N, steps = 64, 60
    x0 = gs.sample_sphere(3, N, seed=31)
    fast = cls(pdf, x0.copy(), 77, dtype="float32", **kw)
    a = fast.sample(steps, burnin=20, return_all_samples=True)
    lit = cls(pdf, x0.copy(), 77, dtype="float32", path="literal", **kw)
    b = lit.sample(steps, burnin=20, return_all_samples=True)
    assert fast.path == "fast" and lit.path == "literal"
    same = np.max(np.abs(a - b), axis=(1, 2)) <= tol
    assert same.mean() >= 0.95, same.mean()
    assert np.allclose(fast.stepsize[same], lit.stepsize[same], rtol=1e-6)
    assert np.array_equal(np.asarray(fast.n_accept_per_chain)[same], np.asarray(lit.n_accept_per_chain)[same])
    assert 0.05 < fast.n_accept / (N * (steps + 20 - 1)) <= 1.0  # HMC at this step size accepts nearly always
    # continuation and shard invariance on the fast path
    more = fast.sample(5)
    assert np.array_equal(more[:, 0], a[:, -1])
    part = cls(pdf, x0[16:32].copy(), 77, dtype="float32", chain_offset=16, **kw)
    assert np.array_equal(part.sample(steps, burnin=20, return_all_samples=True), a[16:32]) and part.path == "fast"


def test_registration_gradient_on_the_cutoff_lists(golden):
    """float32 `gradient` of the registration target runs on the cutoff lists (k_grad_cpd_cell): against the float64 oracle
    gradient at random rotations and near the optimum, k-nearest and dense sums; relative to the largest component."""
    from oracle import targets as T
    g = golden("logprob_cpd")
    n = len(g["c5_source"])
    q = np.concatenate([gs.sample_sphere(3, 64, seed=3), g["c5_q"][:16]])
    for k in (20, None):
        kw = dict(sigma=1.0, omega=0.4)
        pdf = gs.CoherentPointDrift(gs.PointCloud(g["c5_target"]), gs.PointCloud(g["c5_source"], np.ones(n) / n), k=k if k else n, **kw)
        ora = T.CPD(g["c5_target"], g["c5_source"], k=k if k else n, source_weights=np.ones(n) / n, **kw)
        got = pdf.gradient(q.astype(np.float32))
        want = np.array([ora.gradient(qi) for qi in q])
        scale = np.max(np.abs(want), axis=1, keepdims=True)
        assert np.max(np.abs(got - want) / scale) <= 2e-4, k
