"""
Pins the oracle (oracle/*.py, float64 numpy restatement) against vectors produced
by the upstream geosss package itself (tests/golden/*.npz, oracle/make_golden.py).
CPU only.
"""

import numpy as np
import pytest

from oracle import ess, philox, samplers, sphere, targets
from oracle.tape import TapeRNG


def close(a, b, tol=1e-13):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = np.maximum(1.0, np.abs(b))
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    return np.all(both_inf | (np.abs(a - b) <= tol * scale))


# ---------------------------------------------------------------- log_prob
def test_vmf_mixture_logprob(golden):
    g = golden("logprob_vmf")
    assert close(targets.VMFMixture(g["c1_mus"]).log_prob(g["c1_x"]), g["c1_logp"])
    for kappa in (50, 100, 500):
        got = targets.VMFMixture(kappa * g["c2_modes"]).log_prob(g["c2_x"])
        assert close(got, g[f"c2_logp_k{kappa}"])
    assert close(targets.VMFMixture(g["wmix_mus"], g["wmix_w"]).log_prob(g["wmix_x"]), g["wmix_logp"])
    assert close(targets.SingleVMF(g["vmf_mu"]).log_prob(g["vmf_x"]), g["vmf_logp"])
    # single point form
    assert close(targets.VMFMixture(g["c1_mus"]).log_prob(g["c1_x"][5]), g["c1_logp"][5])


def test_vmf_overflow_quirk(golden):
    g = golden("logprob_vmf")
    with np.errstate(over="ignore"):
        got = targets.VMFMixture(800.0 * g["c2_modes"]).log_prob(g["c2_x"][:8])
    assert np.all(np.isneginf(g["vmf_big_logp"])) and np.all(np.isneginf(got))


def test_bingham_logprob(golden):
    g = golden("logprob_bingham")
    for tag in ("c4diag", "c4dense", "d10"):
        pdf = targets.Bingham(g[f"{tag}_A"])
        assert close(pdf.log_prob(g[f"{tag}_x"]), g[f"{tag}_logp"])
        assert close(np.abs(pdf.mode @ g[f"{tag}_mode"]), 1.0, 1e-10)
    assert close(targets.Bingham(g["bf_A"], g["bf_b"]).log_prob(g["bf_x"]), g["bf_logp"])
    with pytest.raises(AssertionError):
        targets.Bingham(g["d10_A"]).log_prob(np.ones(9))


def test_curved_vmf_logprob(golden):
    g = golden("logprob_curved")
    for tag in ("c3", "d10", "d3"):
        pdf = targets.CurvedVMF(g[f"{tag}_knots"], 500.0)
        assert close(pdf.log_prob(g[f"{tag}_x"]), g[f"{tag}_logp"])


def test_cpd_logprob(golden):
    g = golden("logprob_cpd")
    n = len(g["c5_source"])
    cpd = targets.CPD(g["c5_target"], g["c5_source"], sigma=1.0, k=20, omega=0.4,
                      source_weights=np.ones(n) / n)
    assert close(cpd.log_prob(g["c5_q"]), g["c5_logp"], 1e-13)
    # dense sum == k=20 truncation at sigma=1: <= 5e-12 abs
    dense = targets.CPD(g["c5_target"], g["c5_source"], sigma=1.0, k=0, omega=0.4,
                        source_weights=np.ones(n) / n)
    assert np.max(np.abs(dense.log_prob(g["c5_q"]) - g["c5_logp"])) < 5e-12
    # closed-form quaternion -> R equals scipy's
    for q in g["c5_q"][:8]:
        assert np.allclose(targets.quat_to_matrix_closed_form(q), targets.quat_to_matrix(q), atol=1e-15)
    for om in (0, 2, 4):
        m = targets.CPD(g["cube_target"], g["cube_source"], sigma=0.5, k=8, omega=om / 10, beta=1.5,
                        target_weights=g["cube_wt"], source_weights=g["cube_ws"])
        assert close(m.log_prob(g["cube_q"]), g[f"cube_logp_om{om}"])
    gm = targets.GMM(g["cube_target"], g["cube_source"], sigma=0.5, k=8,
                     target_weights=g["cube_wt"], source_weights=g["cube_ws"])
    assert close(gm.log_prob(g["cube_q"]), g["cube_logp_gmm"])


# ---------------------------------------------------------------- geometry / fixtures
def test_sphere_helpers(golden):
    g = golden("misc")
    assert np.array_equal(sphere.sample_sphere(9, 5, seed=1234), g["sphere_d9_n5_seed1234"])
    assert np.array_equal(sphere.sample_sphere(2, seed=7), g["sphere_d2_seed7"])
    got = np.array([sphere.great_subsphere_point(a, b) for a, b in zip(g["proj_z"], g["proj_v"])])
    assert np.array_equal(got, g["proj_out"])
    assert np.array_equal(sphere.unit(g["proj_z"]), g["radial_out"])
    assert np.array_equal(sphere.distance(sphere.unit(g["proj_z"]), sphere.unit(g["proj_v"])), g["dist_out"])
    # tests/test_sphere.py:44-51 property: projected direction is unit and orthogonal to the pole
    pole = sphere.unit(g["proj_v"])
    assert np.allclose(np.sum(got * pole, axis=1), 0, atol=1e-14)
    assert np.allclose(np.linalg.norm(got, axis=1), 1, atol=1e-14)


def test_fixture_generators(golden):
    g = golden("misc")
    assert np.array_equal(targets.random_bingham_matrix(50, 300, 0, True, 6982), g["bingham_c4_A"])
    assert np.array_equal(targets.random_bingham_matrix(10, 30, 0, False, 6982), g["bingham_d10_dense_A"])
    assert np.array_equal(targets.brownian_knots(10, 100, 0.5, seed=4562), g["brownian_c3"])


def test_library_ess_helpers(golden):
    g = golden("misc")
    assert np.allclose(ess.acf_fft(g["ar1"]), g["ar1_acf"], atol=1e-15)
    assert np.isclose(ess.iat(g["ar1"]), g["ar1_iat"], rtol=1e-14)
    assert np.isclose(ess.n_eff(g["ar1"]), g["ar1_neff"], rtol=1e-14)
    # the multi-chain estimator agrees with the AR(1) closed form (1-r)/(1+r) within MC error
    gen = np.random.default_rng(3)
    x = np.zeros((8, 4000))
    e = gen.standard_normal(x.shape)
    for i in range(1, x.shape[1]):
        x[:, i] = 0.8 * x[:, i - 1] + e[:, i]
    rel = ess.ess_multichain(x, 400) / x.size
    assert abs(rel - (1 - 0.8) / (1 + 0.8)) < 0.03


# ---------------------------------------------------------------- trajectories under the reference's own draws
def replay(kind, log_prob, x0, tape, n_states):
    smp = samplers.SliceSampler(kind, log_prob, np.array(x0, float), TapeRNG(tape=tape))
    traj = smp.sample(n_states, 0)
    assert smp.rng.cursor == len(tape)  # consumed exactly what the reference drew
    return traj, np.concatenate([[0], np.cumsum(smp.rejects_per_step)])


def check_traj(g, tag, kind, log_prob, tol=1e-14):
    traj, cum = replay(kind, log_prob, g[f"{tag}_x0"], g[f"{tag}_tape"], len(g[f"{tag}_traj"]))
    assert np.array_equal(cum, g[f"{tag}_cumrej"])
    assert np.max(np.abs(traj - g[f"{tag}_traj"])) <= tol


def test_traj_c1_readme(golden):
    g = golden("traj_c1")
    pdf = targets.VMFMixture(g["mus"])
    check_traj(g, "shrink", samplers.SHRINK, pdf.log_prob)
    check_traj(g, "reject", samplers.REJECT, pdf.log_prob)
    # documented call: .sample(1000, 100) == states 100..1099 of the 1100-state run
    assert np.array_equal(g["readme_shrink_samples"], g["shrink_traj"][100:])
    assert np.array_equal(g["readme_reject_samples"], g["reject_traj"][100:])


def test_traj_c2(golden):
    g = golden("traj_c2")
    for kappa in (100, 500):
        pdf = targets.VMFMixture(kappa * g["modes"])
        check_traj(g, f"shrink_k{kappa}_s0", samplers.SHRINK, pdf.log_prob)
        check_traj(g, f"shrink_k{kappa}_s1", samplers.SHRINK, pdf.log_prob)
        check_traj(g, f"reject_k{kappa}_s0", samplers.REJECT, pdf.log_prob)


def test_traj_c3_c4_c5(golden):
    g = golden("traj_c3")
    pdf = targets.CurvedVMF(g["knots"], float(g["kappa"]))
    check_traj(g, "shrink", samplers.SHRINK, pdf.log_prob)
    check_traj(g, "reject", samplers.REJECT, pdf.log_prob)
    g = golden("traj_c4")
    for tag in ("diag", "dense"):
        pdf = targets.Bingham(g[f"{tag}_A"])
        check_traj(g, f"{tag}_shrink", samplers.SHRINK, pdf.log_prob)
        check_traj(g, f"{tag}_reject", samplers.REJECT, pdf.log_prob)
    g, c = golden("traj_c5"), golden("logprob_cpd")
    n = len(c["c5_source"])
    cpd = targets.CPD(c["c5_target"], c["c5_source"], sigma=1.0, k=20, omega=0.4, source_weights=np.ones(n) / n)
    check_traj(g, "shrink", samplers.SHRINK, cpd.log_prob)
    check_traj(g, "reject", samplers.REJECT, cpd.log_prob)


def test_burnin_rule():
    assert samplers.burnin_count(1000, 100) == 100
    assert samplers.burnin_count(1000, 0.2) == 200
    assert samplers.burnin_count(1000, 0) == 0
    with pytest.raises(AssertionError):
        samplers.burnin_count(10, 1.5)
    with pytest.raises(AssertionError):
        samplers.burnin_count(10, -1)


# ---------------------------------------------------------------- Philox known answers (Random123 kat_vectors)
def test_philox_known_answers():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        assert tuple(philox.philox4x32_10(ctr, key)) == want


def test_philox_chain_rng_layout():
    # d = 10: aux words ride in the last normal block, proposals start at block 3
    rng = philox.PhiloxChainRNG(seed=12345, chain=7, kind=samplers.SHRINK, d=10, dtype=np.float32)
    z = rng.standard_normal(10)
    assert z.shape == (10,) and np.all(np.isfinite(z))
    u = rng.random()
    t0 = rng.uniform(0, 1)
    p = [rng.uniform(0, 1) for _ in range(7)]
    last = philox.philox4x32_10((7, 0, 0, 2), (12345, 0))
    assert u == philox.u01(last[2], np.float32) and t0 == philox.u01(last[3], np.float32)
    blk = philox.philox4x32_10((7, 0, 0, 3), (12345, 0))
    assert p[:4] == [philox.u01(w, np.float32) for w in blk]
    blk2 = philox.philox4x32_10((7, 0, 0, 4), (12345, 0))
    assert p[4:7] == [philox.u01(w, np.float32) for w in blk2[:3]]
    # reject sampler skips the bracket word
    rj = philox.PhiloxChainRNG(seed=12345, chain=7, kind=samplers.REJECT, d=10, dtype=np.float32)
    rj.standard_normal(10)
    rj.random()
    assert rj.uniform(0, 1) == p[0]
    # d = 4: dedicated aux block 1, proposals from block 2; second transition uses step 1
    r4 = philox.PhiloxChainRNG(seed=1, chain=0, kind=samplers.SHRINK, d=4, dtype=np.float64)
    r4.standard_normal(4)
    r4.standard_normal(4)
    aux = philox.philox4x32_10((0, 0, 1, 1), (1, 0))
    assert r4.random() == philox.u01(aux[0], np.float64) and r4.uniform(0, 1) == philox.u01(aux[1], np.float64)
    assert r4.uniform(0, 1) == philox.u01(philox.philox4x32_10((0, 0, 1, 2), (1, 0))[0], np.float64)


def test_pathstats_oracle_matches_reference_expressions(golden):
    from oracle import pathstats
    g = golden("pathstats")
    xs = g["vmf_samples"]
    assert np.array_equal(sum(pathstats.mode_populations(x, g["vmf_modes"]) for x in xs), g["vmf_counts"])
    assert abs(pathstats.mode_kl(xs.reshape(-1, xs.shape[-1]), g["vmf_modes"], g["vmf_weights"]) - g["vmf_kl"]) < 1e-14
    for x, want in zip(xs, g["vmf_dist"]):
        assert np.array_equal(pathstats.geodesic_steps(x), want)
    for b, hop in zip(g["bingham_samples"], g["bingham_hop"]):
        assert pathstats.hops(b, g["bingham_mode"]) / (len(b) - 1) == hop


def test_exact_samplers_oracle_matches_reference_moments(golden):
    """oracle/exact.py (Wood vMF, Kent Bingham) vs moments of the reference's own exact samplers
    (geosss/rand.py run under a fixed numpy seed).  5-sigma bounds: both sides are Monte-Carlo estimates."""
    from oracle import exact
    g = golden("exact")
    for tag in ("d10", "d3", "d2"):
        mu = g[f"vmf_{tag}_mu"]
        x = exact.wood_vmf(mu, 100000, seed=1)
        se = np.sqrt(1.0 / int(g[f"vmf_{tag}_n"]) + 1e-5)
        assert np.max(np.abs(x.mean(0) - g[f"vmf_{tag}_mean"])) < 5 * se
    for tag in ("d10", "d3", "d2"):
        x, eff = exact.kent_bingham(g[f"bing_{tag}_A"], 100000, seed=2, return_efficiency=True)
        assert np.max(np.abs(x.T @ x / len(x) - g[f"bing_{tag}_second"])) < 5 * np.sqrt(1.5e-5)
        assert abs(eff - g[f"bing_{tag}_eff"]) < 0.06


def test_bulk_ess_restatement_known_answers():
    """oracle/ess.py:bulk_ess (arviz-style rank-normalised split ESS, restated from Vehtari et al. 2021): i.i.d. draws
    give ESS ~ S, AR(1) chains ESS / S ~ (1 - phi) / (1 + phi), average ranks equal scipy's, and the truncation rules
    written for the device path (geosss_b200/diagnostics.py:_tau_bulk, from all-reduced partial sums) agree with the
    literal arviz-order loop of the oracle on random traces, including runs that stop for lack of lags."""
    from scipy.stats import rankdata
    from geosss_b200.diagnostics import _tau_bulk
    gen = np.random.default_rng(0)
    assert abs(ess.bulk_ess(gen.standard_normal((8, 4000)), relative=True) - 1.0) < 0.05
    for phi in (0.5, 0.9):
        e = gen.standard_normal((16, 20000))
        y = np.zeros_like(e)
        for t in range(1, e.shape[1]):
            y[:, t] = phi * y[:, t - 1] + e[:, t]
        assert abs(ess.bulk_ess(y, relative=True) / ((1 - phi) / (1 + phi)) - 1.0) < 0.08
    v = gen.integers(0, 50, 1000).astype(float)
    assert np.array_equal(ess.average_ranks(v), rankdata(v))
    assert ess.split_chains(np.arange(14.0).reshape(2, 7)).tolist() == [[0, 1, 2], [7, 8, 9], [4, 5, 6], [11, 12, 13]]
    for trial in range(200):
        m, n, phi = int(gen.integers(1, 6)), int(gen.integers(8, 200)), gen.uniform(-0.9, 0.99)
        e = gen.standard_normal((m, n))
        y = np.zeros_like(e)
        for t in range(1, n):
            y[:, t] = phi * y[:, t - 1] + e[:, t]
        acov = ess.autocov_fft_chains(y)
        assert np.allclose(acov[:, :min(n, 30)], ess.autocov_multichain(y, min(n, 30)), atol=1e-12)
        mean_var = acov[:, 0].mean() * n / (n - 1.0)
        var_plus = mean_var * (n - 1.0) / n + (y.mean(1).var(ddof=1) if m > 1 else 0.0)
        tau_o, _ = ess.geyer_tau(acov.mean(0), mean_var, var_plus, n, m * n)
        cm = y.mean(1)
        tau_d, exhausted = _tau_bulk(acov.sum(0), cm.sum(), (cm * cm).sum(), n, m)
        assert not exhausted and abs(tau_o - tau_d) <= 1e-12 * tau_o
        L = max(4, n // 3)
        tau_l, exhausted = _tau_bulk(acov.sum(0)[:L], cm.sum(), (cm * cm).sum(), n, m)
        assert exhausted or abs(tau_l - tau_o) <= 1e-9 * tau_o
