"""
3D-2D registration (RotationProjection source) and per-call translations, the cases of the reference's
own tests/test_cpd.py: oracle and device math on the CPU, the C ABI on the GPU, all against values the
reference produced (tests/golden/projection.npz, oracle/make_golden.py:golden_projection).
"""

import numpy as np
import pytest

import geosss_b200 as gs
import hostsim_util as hs
from oracle import targets as ot

F32, F64 = np.float32, np.float64
CASES = [(tag, om) for tag in ("p3", "p2") for om in (0, 2, 4)]


def rel(got, want):
    return np.max(np.abs(np.asarray(got, float) - want) / np.maximum(1.0, np.abs(want)))


@pytest.mark.parametrize("tag,om", CASES)
def test_oracle(golden, tag, om):
    g = golden("projection")
    cpd = ot.CPD(g[f"{tag}_target"], g["cube"], 0.5, 8, 1.0, om / 10, translation=g[f"{tag}_t"])
    assert rel(cpd.log_prob(g["q"]), g[f"{tag}_logp_om{om}"]) <= 1e-12
    assert rel(np.array([cpd.gradient(q) for q in g["q"]]), g[f"{tag}_grad_om{om}"]) <= 1e-10
    plain = ot.CPD(g[f"{tag}_target"], g["cube"], 0.5, 8, 1.0, om / 10)
    assert rel(plain.log_prob(g["q"]), g[f"{tag}_logp0_om{om}"]) <= 1e-12
    gmm = ot.GMM(g[f"{tag}_target"], g["cube"], 0.5, 8, translation=g[f"{tag}_t"])
    assert rel(gmm.log_prob(g["q"]), g[f"{tag}_logp_gmm"]) <= 1e-12
    assert rel(np.array([gmm.gradient(q) for q in g["q"]]), g[f"{tag}_grad_gmm"]) <= 1e-10


@pytest.mark.parametrize("dtype,tol_lp,tol_g", [(F64, 1e-12, 1e-10), (F32, 2e-5, 5e-4)])
@pytest.mark.parametrize("tag,om", CASES)
def test_device_math_on_host(golden, tag, om, dtype, tol_lp, tol_g):
    g = golden("projection")
    lib = hs.load()
    t = hs.cpd(g[f"{tag}_target"], g["cube"], 0.5, 8, om / 10, translation=g[f"{tag}_t"])
    assert rel(hs.logp(lib, t, g["q"], dtype), g[f"{tag}_logp_om{om}"]) <= tol_lp
    want = g[f"{tag}_grad_om{om}"]
    assert np.max(np.abs(hs.grad(lib, t, g["q"], dtype) - want) / np.maximum(1.0, np.max(np.abs(want), axis=1, keepdims=True))) <= tol_g


def clouds(g, tag):
    src = gs.PointCloud(g["cube"]) if tag == "p3" else gs.RotationProjection(g["cube"])
    return gs.PointCloud(g[f"{tag}_target"]), src


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol_lp,tol_g", [(F64, 1e-12, 1e-10), (F32, 2e-5, 5e-4)])
@pytest.mark.parametrize("tag,om", CASES)
def test_gpu(golden, tag, om, dtype, tol_lp, tol_g):
    g = golden("projection")
    tgt, src = clouds(g, tag)
    cpd = gs.CoherentPointDrift(tgt, src, sigma=0.5, k=8, beta=1.0, omega=om / 10)
    q, t = g["q"].astype(dtype), g[f"{tag}_t"]
    assert rel(cpd.log_prob(q, t), g[f"{tag}_logp_om{om}"]) <= tol_lp
    assert rel(cpd.log_prob(q), g[f"{tag}_logp0_om{om}"]) <= tol_lp  # the per-call translation does not stick
    want = g[f"{tag}_grad_om{om}"]
    assert np.max(np.abs(cpd.gradient(q, t) - want) / np.maximum(1.0, np.max(np.abs(want), axis=1, keepdims=True))) <= tol_g
    if om == 2:
        assert abs(cpd.log_prob(g["R_true"], t) - g[f"{tag}_logp_matrix"]) <= 1e-9  # rotation-matrix input
        gmm = gs.GaussianMixtureModel(tgt, src, sigma=0.5, k=8)
        assert rel(gmm.log_prob(q, t), g[f"{tag}_logp_gmm"]) <= tol_lp


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["p3", "p2"])
def test_gpu_gradient_is_the_derivative_and_samplers_see_the_translation(golden, tag):
    """The check upstream's tests/test_cpd.py prints: analytical vs numerical gradient at the true pose."""
    from scipy.optimize import approx_fprime
    g = golden("projection")
    tgt, src = clouds(g, tag)
    cpd = gs.CoherentPointDrift(tgt, src, sigma=0.5, k=8, omega=0.2)
    cpd.translation = g[f"{tag}_t"]
    q = g["q"][0]
    num = approx_fprime(q, lambda v: float(cpd.log_prob(v)), 1e-6)
    assert np.max(np.abs(cpd.gradient(q) - num)) < 1e-3
    # sampling with the persistent translation reaches poses as likely as the true one (the cube has 24
    # equivalent rotations, so the states themselves are not compared)
    x0 = gs.sample_sphere(3, 256, seed=3)
    x = gs.ShrinkageSphericalSliceSampler(cpd, x0, 5, dtype="float64").sample(40, 39)[:, -1]
    lp = cpd.log_prob(x)
    assert np.median(lp) > float(cpd.log_prob(q)) - 25
    cpd.translation = None
    assert np.median(cpd.log_prob(x)) < np.median(lp) - 0.5  # without the shift the same poses score worse
    with pytest.raises(ValueError):
        gs.CoherentPointDrift(tgt, gs.PointCloud(g["cube"]) if tag == "p2" else gs.RotationProjection(g["cube"]))
