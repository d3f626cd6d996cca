"""
CPU checks of the DEVICE math of the baseline samplers and gradients (csrc/transition_mh.cuh,
targets.cuh: grad) through the test-only host build, against the reference's own runs
(tests/golden/baselines.npz) and against the oracle fed with the same Philox draws.

Tolerances: float64 gradients <= 1e-11 relative; float64 tape trajectories <= 1e-8 abs with identical
accept decisions and step sizes; float32 gradients <= 2e-4 relative.
"""

import numpy as np
import pytest

import hostsim_util as hs
from oracle import baselines as ob
from oracle import philox
from test_oracle_baselines import GRADS, RUNS, oracle_targets

F32, F64 = np.float32, np.float64


@pytest.fixture(scope="module")
def lib():
    return hs.load()


def host_targets(g, golden):
    c = golden("logprob_cpd")
    n = len(c["c5_source"])
    ws = np.ones(n) / n
    return {
        "c1": hs.vmf_mixture(g["c1_mus"]),
        "c2": hs.vmf_mixture(g["c2_mus"], g["c2_weights"]),
        "single": hs.vmf_mixture(g["single_mu"], single=True),
        "dense": hs.bingham(g["dense_A"]),
        "fisher": hs.bingham(g["dense_A"], g["fisher_b"]),
        "curved": hs.curved(g["knots"], float(g["kappa"])),
        "cpd": hs.cpd(c["c5_target"], c["c5_source"], 1.0, 20, 0.4, ws=ws),
        "gmm": hs.cpd(c["c5_target"], c["c5_source"], 1.0, 20, 0.0, ws=ws, outlier=False),
        "full": hs.cpd(c["c5_target"], c["c5_source"], 2.0, n, 0.1, beta=0.5, ws=ws),
    }


@pytest.mark.parametrize("dtype,tol", [(F64, 1e-11), (F32, 2e-4)])
@pytest.mark.parametrize("name,pts", GRADS)
def test_gradient_device_math(lib, golden, name, pts, dtype, tol):
    g = golden("baselines")
    got = hs.grad(lib, host_targets(g, golden)[name], g[pts], dtype)
    want = g[f"grad_{name}"]
    scale = np.maximum(1.0, np.max(np.abs(want), axis=1, keepdims=True))
    assert np.max(np.abs(got - want) / scale) <= tol


@pytest.mark.parametrize("tag,kind,target,kw", RUNS)
def test_baseline_tape_trajectory_device_math(lib, golden, tag, kind, target, kw):
    g = golden("baselines")
    traj = g[f"{tag}_traj"]
    d = traj.shape[1] // 2 if kind == ob.HMC else traj.shape[1]
    n = len(traj) - 1
    step = kw.get("stepsize", 1e-3 if kind == ob.HMC else 1e-1)
    rc, x, v, steps, acc, n_walk = hs.run_mh(lib, host_targets(g, golden)[target], kind, traj[0, :d], n, F64, step,
                                             tape=g[f"{tag}_tape"], alpha=kw.get("mixing_probability", 0.5),
                                             n_leap=kw.get("n_steps", 10), adapt=int(g[f"{tag}_adapt"]))
    assert rc == 0
    assert np.array_equal(np.cumsum(acc), g[f"{tag}_cumacc"][1:])
    assert np.max(np.abs(x - traj[1:, :d])) <= 1e-8
    assert np.allclose(steps, g[f"{tag}_stepsize"][1:], rtol=1e-12, atol=0)
    if kind == ob.HMC:
        assert np.max(np.abs(v - traj[1:, d:])) <= 1e-6 * max(1.0, np.max(np.abs(traj[1:, d:])))


@pytest.mark.parametrize("kind,target,kw", [(ob.RWMH, "c2", dict(stepsize=0.1)), (ob.INDEPENDENCE, "c1", {}),
                                            (ob.MIXTURE, "c2", dict(stepsize=0.1, mixing_probability=0.3)),
                                            (ob.HMC, "c2", dict(stepsize=0.02, n_steps=5)), (ob.HMC, "dense", dict(stepsize=0.05)),
                                            (ob.RWMH, "curved", {})])
def test_baseline_philox_stream_vs_oracle(lib, golden, kind, target, kw):
    """Free-running device math and the oracle algorithm fed by oracle/philox.py:PhiloxSeqRNG agree."""
    g = golden("baselines")
    tgt = oracle_targets(g, golden)[target]
    x0 = np.random.default_rng(5).standard_normal(tgt.d)
    x0 /= np.linalg.norm(x0)
    n, seed, chain, step0, adapt = 120, 0x1234ABCD5678, 7, 3, 40
    smp = ob.Baseline(kind, tgt, x0, philox.PhiloxSeqRNG(seed, chain, F64, first_step=step0), **kw)
    smp.reset(adapt)
    want = np.array([np.copy(next(smp)) for _ in range(n)])
    step = kw.get("stepsize", 1e-3 if kind == ob.HMC else 1e-1)
    rc, x, v, steps, acc, n_walk = hs.run_mh(lib, host_targets(g, golden)[target], kind, x0, n, F64, step, seed=seed, chain=chain,
                                             step0=step0, alpha=kw.get("mixing_probability", 0.5), n_leap=kw.get("n_steps", 10),
                                             adapt=adapt)
    assert rc == 0 and np.max(np.abs(x - want)) <= 1e-8
    assert int(acc.sum()) == smp.n_accept and abs(steps[-1] - smp.stepsize) <= 1e-12 * smp.stepsize
    if kind == ob.MIXTURE:
        assert n_walk == smp.rwmh_counter
