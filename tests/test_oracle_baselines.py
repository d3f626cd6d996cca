"""Oracle restatement of the baseline samplers and target gradients vs the reference's own runs
(tests/golden/baselines.npz, made by oracle/make_golden.py from geosss/mcmc.py:78-332)."""

import numpy as np
import pytest

from oracle import baselines as ob
from oracle import targets as ot
from oracle.tape import TapeRNG


def oracle_targets(g, golden):
    c = golden("logprob_cpd")
    n = len(c["c5_source"])
    return {
        "c1": ot.VMFMixture(g["c1_mus"]),
        "c2": ot.VMFMixture(g["c2_mus"], g["c2_weights"]),
        "single": ot.SingleVMF(g["single_mu"]),
        "dense": ot.Bingham(g["dense_A"]),
        "fisher": ot.Bingham(g["dense_A"], g["fisher_b"]),
        "curved": ot.CurvedVMF(g["knots"], float(g["kappa"])),
        "cpd": ot.CPD(c["c5_target"], c["c5_source"], 1.0, 20, 1.0, 0.4, source_weights=np.ones(n) / n),
        "gmm": ot.GMM(c["c5_target"], c["c5_source"], 1.0, 20, source_weights=np.ones(n) / n),
        "full": ot.CPD(c["c5_target"], c["c5_source"], 2.0, n, 0.5, 0.1, source_weights=np.ones(n) / n),
    }


GRADS = [("c1", "x3"), ("c2", "x10"), ("single", "x10"), ("dense", "x10"), ("fisher", "x10"), ("curved", "x100"),
         ("cpd", "q"), ("gmm", "q"), ("full", "q")]


@pytest.mark.parametrize("name,pts", GRADS)
def test_gradient_oracle(golden, name, pts):
    g = golden("baselines")
    t = oracle_targets(g, golden)[name]
    got = np.array([t.gradient(x) for x in g[pts]])
    want = g[f"grad_{name}"]
    assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) <= 1e-11


RUNS = [("rwmh_c1", ob.RWMH, "c1", dict(stepsize=1e-1)), ("rwmh_c2", ob.RWMH, "c2", dict(stepsize=1e-1)),
        ("rwmh_dense", ob.RWMH, "dense", dict(stepsize=0.3)), ("rwmh_curved", ob.RWMH, "curved", {}),
        ("rwmh_cpd", ob.RWMH, "cpd", {}), ("indep_c1", ob.INDEPENDENCE, "c1", {}),
        ("mix_c2", ob.MIXTURE, "c2", dict(stepsize=1e-1, mixing_probability=0.5)),
        ("mix_c1", ob.MIXTURE, "c1", dict(stepsize=0.2, mixing_probability=0.8)),
        ("hmc_c1", ob.HMC, "c1", dict(stepsize=1e-1)), ("hmc_c2", ob.HMC, "c2", dict(stepsize=1e-2, n_steps=7)),
        ("hmc_dense", ob.HMC, "dense", dict(stepsize=5e-2)), ("hmc_fisher", ob.HMC, "fisher", dict(stepsize=5e-2)),
        ("hmc_curved", ob.HMC, "curved", dict(stepsize=1e-3)), ("hmc_cpd", ob.HMC, "cpd", dict(stepsize=1e-3))]


@pytest.mark.parametrize("tag,kind,target,kw", RUNS)
def test_baseline_trajectory_oracle(golden, tag, kind, target, kw):
    g = golden("baselines")
    traj = g[f"{tag}_traj"]
    d = traj.shape[1] // 2 if kind == ob.HMC else traj.shape[1]
    smp = ob.Baseline(kind, oracle_targets(g, golden)[target], traj[0, :d], TapeRNG(tape=g[f"{tag}_tape"]), **kw)
    smp.reset(int(g[f"{tag}_adapt"]))
    for t in range(1, len(traj)):
        x = next(smp)
        assert np.max(np.abs(x - traj[t, :d])) <= 1e-9, (tag, t)
        if kind == ob.HMC:
            assert np.max(np.abs(smp.v - traj[t, d:])) <= 1e-7 * max(1.0, np.max(np.abs(traj[t, d:]))), (tag, t)
        assert smp.n_accept == g[f"{tag}_cumacc"][t]
        assert abs(smp.stepsize - g[f"{tag}_stepsize"][t]) <= 1e-15 * g[f"{tag}_stepsize"][t] * (t + 1)
    assert smp.rng.cursor == len(g[f"{tag}_tape"])
