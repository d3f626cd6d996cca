"""GPU exact samplers (geosss/rand.py on the device) against the moments of the reference's own samplers
(tests/golden/exact.npz), the oracle restatement, and the slice sampler they serve as ground truth for."""

import numpy as np
import pytest
from scipy import stats

import geosss_b200 as gs
from oracle import exact

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["d10", "d3", "d2"])
def test_wood_vmf(golden, tag):
    g = golden("exact")
    mu = g[f"vmf_{tag}_mu"]
    pdf = gs.VonMisesFisher(mu)
    x = gs.sample_vMF(pdf, 200000, seed=3)
    assert x.shape == (200000, len(mu)) and x.dtype == np.float64
    assert np.max(np.abs(np.linalg.norm(x, axis=1) - 1)) < 1e-12
    assert np.max(np.abs(x.mean(0) - g[f"vmf_{tag}_mean"])) < 5 * np.sqrt(1.0 / int(g[f"vmf_{tag}_n"]) + 5e-6)
    u = mu / np.linalg.norm(mu)
    assert np.max(np.abs(np.quantile(x @ u, [0.05, 0.25, 0.5, 0.75, 0.95]) - g[f"vmf_{tag}_cos_quantiles"])) < 0.02
    # two-sample Kolmogorov-Smirnov against the oracle's Wood sampler, along the mode and across it
    ref = exact.wood_vmf(mu, 50000, seed=4)
    assert stats.ks_2samp(x[:50000] @ u, ref @ u).pvalue > 1e-3
    t = np.zeros(len(mu))
    t[0] = 1.0
    t -= u * (t @ u)
    assert stats.ks_2samp(x[:50000] @ t, ref @ t).pvalue > 1e-3
    # reproducible, independent of batching, float32 output, single draw shape
    assert np.array_equal(gs.sample_vMF(pdf, 1000, seed=3), x[:1000])
    assert gs.sample_vMF(pdf, seed=3).shape == (len(mu),)
    assert gs.sample_vMF(pdf, 16, seed=3, dtype="float32").dtype == np.float32
    assert gs.sample_vMF(gs.VonMisesFisher(np.zeros(4)), 50000, seed=1).mean(0) == pytest.approx(np.zeros(4), abs=0.02)


@pytest.mark.parametrize("tag", ["d10", "d50", "d3", "d2"])
def test_kent_bingham(golden, tag):
    g = golden("exact")
    A = g[f"bing_{tag}_A"]
    x, eff = gs.sample_bingham(A, 200000, return_efficiency=True, seed=5)
    assert x.shape == (200000, len(A)) and np.max(np.abs(np.linalg.norm(x, axis=1) - 1)) < 1e-12
    assert np.max(np.abs(x.T @ x / len(x) - g[f"bing_{tag}_second"])) < 5 * np.sqrt(1e-5)
    assert abs(eff - g[f"bing_{tag}_eff"]) < 0.06
    if tag == "d50":  # diagonal A handed over as a vector (rand.py:150-151)
        xd = gs.sample_bingham(np.diag(A).copy(), 100000, seed=6)
        assert np.max(np.abs(xd.T @ xd / len(xd) - g[f"bing_{tag}_second"])) < 5 * np.sqrt(2e-5)
    assert np.array_equal(gs.sample_bingham(A, 500, seed=5), x[:500])


def test_low_dimensional_wrappers_and_ground_truth_use(golden):
    g = golden("exact")
    x2 = gs.sample_bingham_2d(gs.Bingham(g["bing_d2_A"]), 100000, seed=7)
    assert np.max(np.abs(x2.T @ x2 / len(x2) - g["bing_d2_second_2d"])) < 0.02
    x3 = gs.sample_bingham_3d(gs.Bingham(g["bing_d3_A"]), 100000, seed=8)
    assert np.max(np.abs(x3.T @ x3 / len(x3) - g["bing_d3_second_gibbs"])) < 0.03  # the Gibbs reference is itself noisy
    with pytest.raises(ValueError):
        gs.sample_bingham_2d(gs.Bingham(g["bing_d3_A"]), 10)
    # the use the scripts make of it (scripts/bingham.py:17-18): the slice sampler reproduces the exact moments
    pdf = gs.Bingham(g["bing_d10_A"])
    chains = gs.ShrinkageSphericalSliceSampler(pdf, np.tile(pdf.mode, (8192, 1)), 9).sample(60, 40)
    xs = chains[:, 30:].reshape(-1, 10).astype(float)
    assert np.max(np.abs(xs.T @ xs / len(xs) - g["bing_d10_second"])) < 0.01
