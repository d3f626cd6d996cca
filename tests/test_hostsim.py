"""
CPU checks of the DEVICE math: csrc/{philox,targets,transition}.cuh are compiled for the host
by tests/hostsim (test-only, serial loops) and compared with the golden vectors recorded from
the upstream package.  The same comparisons run on the GPU through the C ABI in
tests/test_gpu_parity.py; this file catches arithmetic mistakes before GPU time is spent.

Stated tolerances:
  float64  log_prob <= 1e-12 * max(1, |logp|);  CPD <= 5e-12 abs;  tape trajectories <= 1e-12
           abs with identical per-step rejection counts
  float32  log_prob <= 1e-5 * max(1, |logp|);   CPD <= 1e-3 abs;   tape trajectories <= 1e-4
           abs up to the first accept/reject flip (flip step reported)
"""

import numpy as np
import pytest

import hostsim_util as hs
from oracle import philox, samplers, targets

F32, F64 = np.float32, np.float64


@pytest.fixture(scope="module")
def lib():
    return hs.load()


def rel_err(got, want):
    return np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want)))


def test_philox_matches_known_answers(lib):
    import ctypes as C
    for ctr, key in [((0, 0, 0, 0), (0, 0)), ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2),
                     ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)), ((7, 0, 12, 3), (99, 1))]:
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        lib.hs_philox(c, k, o)
        assert list(o) == philox.philox4x32_10(ctr, key)


@pytest.mark.parametrize("dtype,tol", [(F64, 1e-12), (F32, 1e-5)])
def test_logprob_vmf_bingham_curved(lib, golden, dtype, tol):
    g = golden("logprob_vmf")
    assert rel_err(hs.logp(lib, hs.vmf_mixture(g["c1_mus"]), g["c1_x"], dtype), g["c1_logp"]) <= tol
    for kappa in (50, 100, 500):
        got = hs.logp(lib, hs.vmf_mixture(kappa * g["c2_modes"]), g["c2_x"], dtype)
        assert rel_err(got, g[f"c2_logp_k{kappa}"]) <= tol
    assert rel_err(hs.logp(lib, hs.vmf_mixture(g["wmix_mus"], g["wmix_w"]), g["wmix_x"], dtype), g["wmix_logp"]) <= tol
    assert rel_err(hs.logp(lib, hs.vmf_mixture(g["vmf_mu"], single=True), g["vmf_x"], dtype), g["vmf_logp"]) <= tol
    assert np.all(np.isneginf(hs.logp(lib, hs.vmf_mixture(800.0 * g["c2_modes"]), g["c2_x"][:8], dtype)))
    g = golden("logprob_bingham")
    for tag in ("c4diag", "c4dense", "d10"):
        assert rel_err(hs.logp(lib, hs.bingham(g[f"{tag}_A"]), g[f"{tag}_x"], dtype), g[f"{tag}_logp"]) <= tol
    assert rel_err(hs.logp(lib, hs.bingham(g["bf_A"], g["bf_b"]), g["bf_x"], dtype), g["bf_logp"]) <= tol
    g = golden("logprob_curved")
    for tag in ("c3", "d10", "d3"):
        got = hs.logp(lib, hs.curved(g[f"{tag}_knots"], 500.0), g[f"{tag}_x"], dtype)
        assert rel_err(got, g[f"{tag}_logp"]) <= (tol if dtype == F64 else 2e-5)


@pytest.mark.parametrize("dtype,tol", [(F64, 5e-12), (F32, 1e-3)])
def test_logprob_cpd(lib, golden, dtype, tol):
    g = golden("logprob_cpd")
    n = len(g["c5_source"])
    for k in (20, 0):  # exact k-nearest truncation, and the dense sum (equal at sigma = 1)
        t = hs.cpd(g["c5_target"], g["c5_source"], 1.0, k, 0.4, ws=np.ones(n) / n)
        assert np.max(np.abs(hs.logp(lib, t, g["c5_q"], dtype) - g["c5_logp"])) <= tol
    for om in (0, 2, 4):
        t = hs.cpd(g["cube_target"], g["cube_source"], 0.5, 8, om / 10, beta=1.5, wt=g["cube_wt"], ws=g["cube_ws"])
        assert rel_err(hs.logp(lib, t, g["cube_q"], dtype), g[f"cube_logp_om{om}"]) <= (1e-12 if dtype == F64 else 1e-5)
    t = hs.cpd(g["cube_target"], g["cube_source"], 0.5, 8, 0.0, wt=g["cube_wt"], ws=g["cube_ws"], outlier=False)
    assert rel_err(hs.logp(lib, t, g["cube_q"], dtype), g["cube_logp_gmm"]) <= (1e-12 if dtype == F64 else 1e-5)


def tape_case(lib, tgt, kind, g, tag, dtype):
    want, cum = g[f"{tag}_traj"], g[f"{tag}_cumrej"]
    rc, got, nrej = hs.run(lib, tgt, kind, g[f"{tag}_x0"], len(want) - 1, dtype, tape=g[f"{tag}_tape"])
    rej_want = np.diff(cum)
    if dtype == F64:
        assert rc == 0
        assert np.array_equal(nrej, rej_want), tag
        assert np.max(np.abs(got - want[1:])) <= 1e-12, tag
        return None
    flips = np.nonzero(nrej != rej_want)[0]
    upto = flips[0] if len(flips) else len(rej_want)
    assert np.max(np.abs(got[:upto] - want[1:1 + upto])) <= 1e-4, tag
    return int(upto), len(rej_want)


def cases(golden):
    g = golden("traj_c1")
    t = hs.vmf_mixture(g["mus"])
    yield t, samplers.SHRINK, g, "shrink"
    yield t, samplers.REJECT, g, "reject"
    g = golden("traj_c2")
    for kappa in (100, 500):
        t = hs.vmf_mixture(kappa * g["modes"])
        yield t, samplers.SHRINK, g, f"shrink_k{kappa}_s0"
        yield t, samplers.SHRINK, g, f"shrink_k{kappa}_s1"
        yield t, samplers.REJECT, g, f"reject_k{kappa}_s0"
    g = golden("traj_c3")
    t = hs.curved(g["knots"], float(g["kappa"]))
    yield t, samplers.SHRINK, g, "shrink"
    yield t, samplers.REJECT, g, "reject"
    g = golden("traj_c4")
    for tag in ("diag", "dense"):
        t = hs.bingham(g[f"{tag}_A"])
        yield t, samplers.SHRINK, g, f"{tag}_shrink"
        yield t, samplers.REJECT, g, f"{tag}_reject"
    g, c = golden("traj_c5"), golden("logprob_cpd")
    n = len(c["c5_source"])
    t = hs.cpd(c["c5_target"], c["c5_source"], 1.0, 20, 0.4, ws=np.ones(n) / n)
    yield t, samplers.SHRINK, g, "shrink"
    yield t, samplers.REJECT, g, "reject"


def test_tape_trajectories_float64(lib, golden):
    for t, kind, g, tag in cases(golden):
        tape_case(lib, t, kind, g, tag, F64)


def test_tape_trajectories_float32(lib, golden):
    total, clean = 0, 0
    for t, kind, g, tag in cases(golden):
        upto, n = tape_case(lib, t, kind, g, tag, F32)
        total += n
        clean += upto
    # flips are rare (~1e-5..1e-6 per proposal); most transitions precede any flip
    assert clean >= 0.8 * total, (clean, total)


@pytest.mark.parametrize("kind", [samplers.SHRINK, samplers.REJECT])
def test_philox_free_running_matches_oracle_fed_same_draws(lib, golden, kind):
    """Device Philox layout + literal transition == reference algorithm fed by oracle/philox.py."""
    g = golden("traj_c2")
    mus = 100.0 * g["modes"]
    pdf = targets.VMFMixture(mus)
    x0 = g["shrink_k100_s0_x0"]
    n = 120 if kind == samplers.SHRINK else 40
    for chain in (0, 5, 2**33 + 1):
        rng = philox.PhiloxChainRNG(seed=0xDEADBEEF12345, chain=chain, kind=kind, d=10, dtype=F64, first_step=3)
        want = samplers.SliceSampler(kind, pdf.log_prob, x0.copy(), rng)
        traj = want.sample(n + 1)
        rc, got, nrej = hs.run(lib, hs.vmf_mixture(mus), kind, x0, n, F64, seed=0xDEADBEEF12345, chain=chain, step0=3)
        assert rc == 0
        assert np.array_equal(nrej, want.rejects_per_step)
        assert np.max(np.abs(got - traj[1:])) <= 1e-11


def test_philox_step_counter_crosses_32_bits(lib, golden):
    """The transition index is 48 bits wide in the counter: a chain that has already done 2^32 - 3 transitions
    keeps following the oracle across the 32-bit boundary (checkpoint / resume of very long runs)."""
    g = golden("traj_c1")
    pdf = targets.VMFMixture(g["mus"])
    x0 = np.array([0.6, -0.64, 0.48])
    first = 2**32 - 3
    rng = philox.PhiloxChainRNG(seed=7, chain=11, kind=samplers.SHRINK, d=3, dtype=F64, first_step=first)
    want = samplers.SliceSampler(samplers.SHRINK, pdf.log_prob, x0.copy(), rng).sample(9)
    rc, got, _ = hs.run(lib, hs.vmf_mixture(g["mus"]), samplers.SHRINK, x0, 8, F64, seed=7, chain=11, step0=first)
    assert rc == 0 and np.max(np.abs(got - want[1:])) <= 1e-11


def test_jacobi_eigensystem_for_dense_bingham():
    """csrc/eigh_jacobi.h (cyclic Jacobi, what the dense-Bingham tensor-core path evaluates x^T A x in) against numpy:
    reconstruction, orthonormality and eigenvalues for random symmetric, diagonal, rank-deficient and C4-like matrices."""
    import ctypes as C
    lib = hs.load()
    gen = np.random.default_rng(0)
    mats = [np.diag(np.arange(5.0)), np.zeros((3, 3)), np.ones((4, 4))]
    for n in (2, 7, 50, 64):
        G = gen.standard_normal((n, n))
        mats += [G + G.T, G.T @ G, (G[:, :3] @ G[:, :3].T)]
    from oracle import targets
    mats.append(targets.random_bingham_matrix(50, 300, 0, False, 6982))
    for A in mats:
        n = len(A)
        A = np.ascontiguousarray(A, dtype=float)
        w, V = np.empty(n), np.empty((n, n))
        lib.hs_eigh(n, A.ctypes.data_as(hs._pd), w.ctypes.data_as(hs._pd), V.ctypes.data_as(hs._pd))
        scale = max(1.0, np.abs(A).max())
        assert np.max(np.abs((V * w) @ V.T - A)) <= 1e-12 * scale * n
        assert np.max(np.abs(V.T @ V - np.eye(n))) <= 1e-13 * n
        assert np.allclose(np.sort(w), np.linalg.eigvalsh(A), atol=1e-11 * scale * n)
