"""
GPU parity tests: the CUDA path, called through the Python mirror -> ctypes -> C ABI of
libgeosss_cuda.so, against the golden vectors recorded from the upstream package and against
the float64 oracle fed the same draws.  Nothing here reads /root/reference.

Stated tolerances:
  float64  log_prob <= 1e-12 * max(1, |logp|);  CPD <= 5e-12 abs;  tape trajectories <= 1e-12
           abs with identical per-step rejection counts
  float32  log_prob <= 1e-5 * max(1, |logp|);  CPD <= 1e-3 abs;  trajectories <= 1e-4 abs STRICTLY
           before the first accept/reject flip of a chain (a flip = the per-transition rejection count
           differs from the float64 reference algorithm fed the same draws); the flip count per 1000
           transitions is printed, appended to gpurun_out/flips.jsonl and bounded per test
  own-RNG marginals: Kolmogorov-Smirnov against exact draws at alpha = 1e-3
"""

import json
import os

import numpy as np
import pytest
import torch

import geosss_b200 as gs
from geosss_b200 import _lib
from oracle import exact, philox, samplers, targets

pytestmark = pytest.mark.gpu

F32, F64 = np.float32, np.float64


def rel_err(got, want):
    return np.max(np.abs(np.asarray(got, float) - want) / np.maximum(1.0, np.abs(want)))


def mixture(mus, weights=None):
    return gs.MixtureModel([gs.VonMisesFisher(m) for m in mus], weights)


def oracle_of(pdf):
    """The float64 numpy restatement (oracle/targets.py) of a device target."""
    if isinstance(pdf, gs.MixtureModel):
        return targets.VMFMixture(np.stack([p.mu for p in pdf.pdfs]), pdf.weights)
    if isinstance(pdf, gs.VonMisesFisher):
        return targets.SingleVMF(pdf.mu)
    if isinstance(pdf, gs.BinghamFisher):
        return targets.Bingham(pdf.A, pdf.b)
    if isinstance(pdf, gs.Bingham):
        return targets.Bingham(pdf.A)
    if isinstance(pdf, gs.CurvedVonMisesFisher):
        return targets.CurvedVMF(pdf.curve.knots, pdf.kappa)
    raise TypeError(type(pdf))


def run_with_rejections(smp, n):
    """n transitions through the C ABI: states (N, n, d) and per-transition rejection counts (N, n)."""
    smp._sampler()
    buf = torch.empty((smp.N, n, smp.d), dtype=smp._tdtype, device=smp.device)
    nrej = torch.zeros((smp.N, n), dtype=torch.int32, device=smp.device)
    smp._run_device(n, buf, nrej=nrej, chain_major=True)
    torch.cuda.synchronize()
    return buf.cpu().numpy(), nrej.cpu().numpy()


class FlipCount:
    """Accept/reject flips of float32 chains against the float64 reference algorithm fed the same draws.
    A chain is compared up to its first flip (first transition whose rejection count differs); before it the
    states must agree to `tol`, strictly.  rate() = flipped chains per 1000 compared transitions."""

    def __init__(self, name, tol=1e-4):
        self.name, self.tol, self.flips, self.compared, self.max_err = name, tol, 0, 0, 0.0

    def add(self, got, nrej, want, rej_want):
        n = len(rej_want)
        diff = np.nonzero(np.asarray(nrej[:n]) != np.asarray(rej_want))[0]
        upto = int(diff[0]) if len(diff) else n
        if upto:
            err = float(np.max(np.abs(got[:upto] - want[:upto])))
            self.max_err = max(self.max_err, err)
            assert err <= self.tol, (self.name, "deviation before any flip", err, upto)
        self.flips += 1 if len(diff) else 0
        self.compared += upto + (1 if len(diff) else 0)

    def rate(self):
        return 1000.0 * self.flips / max(self.compared, 1)

    def check(self, max_per_1000):
        rec = {"test": self.name, "flips": self.flips, "transitions_compared": self.compared,
               "flips_per_1000_transitions": round(self.rate(), 4), "max_abs_err_before_flip": self.max_err, "bound": max_per_1000}
        print("FLIPS", json.dumps(rec))
        try:
            os.makedirs("gpurun_out", exist_ok=True)
            with open(os.path.join("gpurun_out", "flips.jsonl"), "a") as f:
                f.write(json.dumps(rec) + "\n")
        except OSError:
            pass
        # the bound is a rate; a run shorter than 1000 / bound transitions may still hold one flip
        assert self.flips <= max(1, int(max_per_1000 * self.compared / 1000.0)), rec


def c5_target(golden, k=20):
    c = golden("logprob_cpd")
    n = len(c["c5_source"])
    return gs.CoherentPointDrift(gs.PointCloud(c["c5_target"], np.ones(len(c["c5_target"]))),
                                 gs.PointCloud(c["c5_source"], np.ones(n) / n), sigma=1.0, k=k, omega=0.4)


# ------------------------------------------------------------------ log_prob
@pytest.mark.parametrize("dtype,tol", [(F64, 1e-12), (F32, 1e-5)])
def test_logprob_vmf(golden, dtype, tol):
    g = golden("logprob_vmf")
    assert rel_err(mixture(g["c1_mus"]).log_prob(g["c1_x"].astype(dtype)), g["c1_logp"]) <= tol
    for kappa in (50, 100, 500):
        got = mixture(kappa * g["c2_modes"]).log_prob(g["c2_x"].astype(dtype))
        assert got.dtype == dtype and got.shape == (256,)
        assert rel_err(got, g[f"c2_logp_k{kappa}"]) <= tol
    assert rel_err(mixture(g["wmix_mus"], g["wmix_w"]).log_prob(g["wmix_x"].astype(dtype)), g["wmix_logp"]) <= tol
    assert rel_err(gs.VonMisesFisher(g["vmf_mu"]).log_prob(g["vmf_x"].astype(dtype)), g["vmf_logp"]) <= tol
    # single point -> scalar; torch tensor in -> torch tensor out
    one = mixture(g["c1_mus"]).log_prob(g["c1_x"][5].astype(dtype))
    assert np.ndim(one) == 0 and abs(one - g["c1_logp"][5]) <= tol * max(1, abs(g["c1_logp"][5]))
    tt = mixture(g["c1_mus"]).log_prob(torch.as_tensor(g["c1_x"].astype(dtype)).cuda())
    assert isinstance(tt, torch.Tensor) and tt.is_cuda and rel_err(tt.cpu().numpy(), g["c1_logp"]) <= tol
    # i0 overflow quirk: kappa > ~713 gives -inf like the reference
    assert np.all(np.isneginf(mixture(800.0 * g["c2_modes"]).log_prob(g["c2_x"][:8].astype(dtype))))


@pytest.mark.parametrize("dtype,tol", [(F64, 1e-12), (F32, 1e-5)])
def test_logprob_bingham_curved(golden, dtype, tol):
    g = golden("logprob_bingham")
    for tag in ("c4diag", "c4dense", "d10"):
        pdf = gs.Bingham(g[f"{tag}_A"])
        assert rel_err(pdf.log_prob(g[f"{tag}_x"].astype(dtype)), g[f"{tag}_logp"]) <= tol
    assert rel_err(gs.BinghamFisher(g["bf_A"], g["bf_b"]).log_prob(g["bf_x"].astype(dtype)), g["bf_logp"]) <= tol
    with pytest.raises(AssertionError):
        gs.Bingham(g["d10_A"]).log_prob(np.ones(9))
    g = golden("logprob_curved")
    for tag in ("c3", "d10", "d3"):
        pdf = gs.CurvedVonMisesFisher(gs.SlerpCurve(g[f"{tag}_knots"]), 500.0)
        err = rel_err(pdf.log_prob(g[f"{tag}_x"].astype(dtype)), g[f"{tag}_logp"])
        print(f"curved log_prob {tag} {np.dtype(dtype).name}: max rel err {err:.3g}")
        assert err <= tol, (tag, err)


@pytest.mark.parametrize("dtype,tol", [(F64, 5e-12), (F32, 1e-3)])
def test_logprob_cpd(golden, dtype, tol):
    g = golden("logprob_cpd")
    for k in (20, 0):
        assert np.max(np.abs(c5_target(golden, k).log_prob(g["c5_q"].astype(dtype)) - g["c5_logp"])) <= tol
    # rotation matrix input (registration.py:47-53)
    want = g["c5_logp"][-2]  # the quaternion of the true rotation
    assert abs(c5_target(golden).log_prob(g["c5_rotation"]) - want) <= 1e-9
    rtol = 1e-12 if dtype == F64 else 1e-5
    for om in (0, 2, 4):
        m = gs.CoherentPointDrift(gs.PointCloud(g["cube_target"], g["cube_wt"]), gs.PointCloud(g["cube_source"], g["cube_ws"]),
                                  sigma=0.5, k=8, omega=om / 10, beta=1.5)
        assert rel_err(m.log_prob(g["cube_q"].astype(dtype)), g[f"cube_logp_om{om}"]) <= rtol
    gm = gs.GaussianMixtureModel(gs.PointCloud(g["cube_target"], g["cube_wt"]), gs.PointCloud(g["cube_source"], g["cube_ws"]),
                                 sigma=0.5, k=8)
    assert rel_err(gm.log_prob(g["cube_q"].astype(dtype)), g["cube_logp_gmm"]) <= rtol


def test_cpd_fast_path(golden):
    """float32 throughput CPD kernels (dense sum): log_prob within 1e-3 abs of the reference's k=20 values,
    and the sampler follows the float64 reference algorithm under shared Philox draws."""
    g = golden("logprob_cpd")
    n = len(g["c5_source"])
    pdf = gs.CoherentPointDrift(gs.PointCloud(g["c5_target"], np.ones(len(g["c5_target"]))),
                                gs.PointCloud(g["c5_source"], np.ones(n) / n), sigma=1.0, k=20, omega=0.4, dense=True)
    got = pdf.log_prob(g["c5_q"].astype(F32))
    assert np.max(np.abs(got - g["c5_logp"])) <= 1e-3
    ora = targets.CPD(g["c5_target"], g["c5_source"], sigma=1.0, k=20, omega=0.4, source_weights=np.ones(n) / n)
    x0 = gs.sample_sphere(3, 6, seed=17)
    fc = FlipCount("cpd_fast_dense_c5", tol=2e-4)
    for cls, kind, steps in ((gs.ShrinkageSphericalSliceSampler, samplers.SHRINK, 12), (gs.RejectionSphericalSliceSampler, samplers.REJECT, 4)):
        smp = cls(pdf, x0, 4242, dtype="float32", path="fast")
        got, nrej = run_with_rejections(smp, steps)
        assert smp.path == "fast"
        for c in range(6):
            want, rej = oracle_chain(ora, kind, x0[c], 4242, c, steps, 4, F32)
            fc.add(got[c], nrej[c], want, rej)
    fc.check(1.0)  # measured: 0 flips in 96 transitions (profiles/r2_flips.jsonl)


def test_cpd_cutoff_cells_equal_dense_sum(golden, monkeypatch):
    """Cell-list kernels (target rows rotated into the source frame, candidates from the 27-cell neighbourhood)
    against the dense 214 x 214 pair sum of the same float32 path and against the float64 literal kernel:
    random rotations, rotations near the optimum, a translation that moves rows out of the grid, a NaN."""
    g = golden("logprob_cpd")
    n = len(g["c5_source"])
    mk = lambda **kw: gs.CoherentPointDrift(gs.PointCloud(g["c5_target"]), gs.PointCloud(g["c5_source"], np.ones(n) / n),
                                            sigma=1.0, k=20, omega=0.4, **kw)
    fast, exact = mk(dense=True), mk()
    info = fast.cutoff_grid_info()
    assert info is not None and 6.0 < info["cutoff"] < 8.0 and info["mean_candidates"] < 0.25 * n, info
    assert exact.cutoff_grid_info() is not None  # k = 20 keeps its meaning on the cutoff kernels (next test)
    gen = np.random.default_rng(5)
    q = gs.sample_sphere(3, 777, seed=3)
    best = g["c5_q"][int(np.argmax(g["c5_logp"]))]
    near = best + 0.05 * gen.standard_normal((200, 4))
    q = np.vstack([q, near / np.linalg.norm(near, axis=1, keepdims=True), best[None]])
    for shift in (None, np.array([3.0, -2.0, 1.5]), np.array([500.0, 0.0, 0.0])):
        want = exact.log_prob(q, shift)
        cells = fast.log_prob(q.astype(F32), shift)
        monkeypatch.setenv("GEOSSS_CPD_DENSE", "1")
        dense = fast.log_prob(q.astype(F32), shift)
        monkeypatch.delenv("GEOSSS_CPD_DENSE")
        assert np.max(np.abs(cells - want)) <= 1e-3, np.max(np.abs(cells - want))
        assert np.max(np.abs(cells - dense)) <= 5e-4, np.max(np.abs(cells - dense))
    assert np.isnan(fast.log_prob(np.array([[np.nan, 0, 0, 1]], dtype=F32)))[0]
    # sampler on the cutoff kernels against the float64 reference algorithm (KD-tree k = 20) fed the same draws
    x0 = gs.sample_sphere(3, 16, seed=9)
    ora = targets.CPD(g["c5_target"], g["c5_source"], sigma=1.0, k=20, omega=0.4, source_weights=np.ones(n) / n)
    smp = gs.ShrinkageSphericalSliceSampler(fast, x0, 77, dtype="float32", path="fast")
    got, nrej = run_with_rejections(smp, 6)
    fc = FlipCount("cpd_cutoff_cells_c5_shrink", tol=2e-4)
    for c in range(16):
        want_c, rej_c = oracle_chain(ora, samplers.SHRINK, x0[c], 77, c, 6, 4, F32)
        fc.add(got[c], nrej[c], want_c, rej_c)
    fc.check(1.0)


@pytest.mark.parametrize("k", [1, 3, 8, 20])
def test_cpd_cutoff_cells_keep_k_nearest_truncation(golden, k):
    """registration.py:107-108 sums the k nearest rotated source points.  The cutoff kernels reproduce that:
    rows with more than k points inside the cutoff rank their candidates exactly (small k here forces that
    branch; at k = 20 it never runs on these clouds).  Checked against the float64 literal kernel, which
    selects the k nearest of all pairs, and -- for k = 20 -- against the reference's own values."""
    g = golden("logprob_cpd")
    n = len(g["c5_source"])
    pdf = gs.CoherentPointDrift(gs.PointCloud(g["c5_target"]), gs.PointCloud(g["c5_source"], np.ones(n) / n),
                                sigma=1.0, k=k, omega=0.4)
    gen = np.random.default_rng(k)
    best = g["c5_q"][int(np.argmax(g["c5_logp"]))]
    near = best + 0.03 * gen.standard_normal((150, 4))
    q = np.vstack([gs.sample_sphere(3, 300, seed=k), near / np.linalg.norm(near, axis=1, keepdims=True), g["c5_q"]])
    ora = targets.CPD(g["c5_target"], g["c5_source"], sigma=1.0, k=k, omega=0.4, source_weights=np.ones(n) / n)
    want = ora.log_prob(q)               # float64 numpy restatement: KD-tree query of the k nearest (registration.py:101-109)
    assert np.max(np.abs(pdf.log_prob(q) - want)) <= 5e-12 * 1e3  # float64 literal kernel: exact selection over all pairs
    got = pdf.log_prob(q.astype(F32))    # float32: cutoff kernels
    assert np.max(np.abs(got - want)) <= 1e-3, np.max(np.abs(got - want))
    if k == 20:
        assert np.max(np.abs(got[-len(g["c5_q"]):] - g["c5_logp"])) <= 1e-3
    smp = gs.ShrinkageSphericalSliceSampler(pdf, gs.sample_sphere(3, 32, seed=1), 5, dtype="float32")
    out = smp.sample(4)
    assert smp.path == "fast" and np.all(np.isfinite(out))


def test_logprob_empty_and_ragged(golden):
    g = golden("logprob_vmf")
    pdf = mixture(g["c1_mus"])
    assert pdf.log_prob(np.zeros((0, 3))).shape == (0,)
    for n in (1, 31, 33, 129):
        assert rel_err(pdf.log_prob(g["c1_x"][:n] if n <= 96 else np.tile(g["c1_x"], (2, 1))[:n]),
                       (g["c1_logp"][:n] if n <= 96 else np.tile(g["c1_logp"], 2)[:n])) <= 1e-12
    with pytest.raises(ValueError):
        pdf.log_prob(np.zeros(4))


def test_fast_logprob_tiles_ragged_unaligned_large(golden):
    """Tiled float32 log_prob: complete tiles go through the bulk-copy ring, the ragged last tile and
    16-byte-unaligned inputs through plain loads; all must agree with the float64 literal kernel."""
    g = golden("logprob_vmf")
    pdf = mixture(100.0 * g["c2_modes"])
    gen = np.random.default_rng(0)
    for n in (127, 128, 129, 1000, 128 * 300 + 5):
        x = gs.radial_projection(gen.standard_normal((n + 1, 10)))
        want = oracle_of(pdf).log_prob(x)  # float64 numpy restatement
        xt = torch.as_tensor(x, dtype=torch.float32).cuda()
        got = pdf.log_prob(xt).cpu().numpy()
        assert rel_err(got, want) <= 1e-5
        got_unaligned = pdf.log_prob(xt[1:]).cpu().numpy()  # row offset 40 B: not 16-byte aligned
        assert rel_err(got_unaligned, want[1:]) <= 1e-5
    for pdf2, d in ((gs.random_bingham(50, 300, 0, True, 6982), 50), (gs.random_bingham(7, 30, 0, False, 1), 7),
                    (gs.CurvedVonMisesFisher(gs.SlerpCurve(gs.brownian_curve(10, 100, 0.5, seed=4562)), 500.0), 100)):
        x = gs.radial_projection(gen.standard_normal((128 * 9 + 17, d)))
        err = rel_err(pdf2.log_prob(x.astype(F32)), oracle_of(pdf2).log_prob(x))
        print(f"tiled log_prob {type(pdf2).__name__} d={d}: max rel err {err:.3g}")
        assert err <= 1e-5, (type(pdf2).__name__, d, err)


@pytest.mark.parametrize("d", [8, 12, 16, 20, 32, 33, 48, 64, 100, 128])
def test_fast_logprob_row_layouts(d):
    """Every shared-memory row layout of the tiled float32 log_prob: d a multiple of 8 (padded rows, one bulk
    copy per row), a multiple of 4 (contiguous tile, 16-byte loads) and odd sizes (scalar loads); aligned,
    unaligned and ragged inputs; all four targets against the float64 literal kernels."""
    gen = np.random.default_rng(d)
    K = 3 if d <= 16 else 7
    pdfs = [gs.MixtureModel([gs.VonMisesFisher(30.0 * m) for m in gs.sample_sphere(d - 1, K, seed=2)], np.arange(1.0, K + 1)),
            gs.random_bingham(d, 40.0, 0, True, 3), gs.BinghamFisher(gs.random_bingham(d, 40.0, 0, True, 3).A, gen.standard_normal(d)),
            gs.CurvedVonMisesFisher(gs.SlerpCurve(gs.brownian_curve(9, d, 0.5, seed=4)), 200.0)]
    if d <= 64:
        pdfs.append(gs.random_bingham(d, 40.0, 0, False, 5))  # dense (tensor-core kernel for 8 <= d <= 64)
    x = gs.radial_projection(gen.standard_normal((128 * 5 + 77, d)))
    xt = torch.as_tensor(x, dtype=torch.float32).cuda()
    for pdf in pdfs:
        want = oracle_of(pdf).log_prob(x)  # float64 numpy restatement
        err = rel_err(pdf.log_prob(xt).cpu().numpy(), want)
        print(f"row layout d={d} {type(pdf).__name__}: max rel err {err:.3g}")
        assert err <= 1e-5, (type(pdf).__name__, err)
        assert rel_err(pdf.log_prob(xt[1:]).cpu().numpy(), want[1:]) <= 1e-5  # rows start 4 d bytes later: unaligned for odd d
        assert rel_err(pdf.log_prob(xt[:128]).cpu().numpy(), want[:128]) <= 1e-5


def test_dense_bingham_logprob_on_tensor_cores():
    """Dense Bingham(-Fisher) log_prob, 8 <= d <= 64: tcgen05 3xTF32 kernel vs the float64 literal kernel,
    ragged tile counts, unaligned input, and the SIMT fallback (GEOSSS_NO_TC) as a second opinion."""
    import os
    gen = np.random.default_rng(0)
    for d in (8, 17, 32, 50, 64):
        pdf = gs.random_bingham(d, 300.0, 0.0, False, 6982)
        for n in (1, 127, 128, 129, 128 * 37 + 5):
            x = gs.radial_projection(gen.standard_normal((n + 1, d)))
            want = oracle_of(pdf).log_prob(x)  # float64 numpy restatement
            xt = torch.as_tensor(x, dtype=torch.float32).cuda()
            assert rel_err(pdf.log_prob(xt).cpu().numpy(), want) <= 1e-5
            assert rel_err(pdf.log_prob(xt[1:]).cpu().numpy(), want[1:]) <= 1e-5  # 4d-byte offset: plain-load path when not 16-byte aligned
    pf = gs.BinghamFisher(gs.random_bingham(20, 20.0, 0.0, False, 3).A, 4.0 * gs.sample_sphere(19, seed=4))
    x = gs.radial_projection(gen.standard_normal((1000, 20)))
    tc = pf.log_prob(x.astype(F32))
    want = oracle_of(pf).log_prob(x)
    assert rel_err(tc, want) <= 1e-5
    os.environ["GEOSSS_NO_TC"] = "1"
    try:
        simt = pf.log_prob(x.astype(F32))
    finally:
        del os.environ["GEOSSS_NO_TC"]
    assert rel_err(simt, want) <= 1e-5 and not np.array_equal(simt, tc)


@pytest.mark.parametrize("d", [24, 32, 48])
def test_tensor_core_logprob_on_256_row_tiles(d, monkeypatch):
    """Large batches of narrow rows take 256-row tiles on the tcgen05 kernel (two MMA groups per hand-off; batches below
    2 x 148 x 256 points keep 128-row tiles, so the small-N tests above never reach this path): vMF mixture, curved vMF and dense
    Bingham against the float64 oracle on a strided subset and the ragged tail, and bit-equal to the 128-row-tile kernel
    (GEOSSS_TC_SUB1) -- a row's 3xTF32 arithmetic does not depend on the tile it sits in."""
    n = 2 * 148 * 256 * 2 + 256 + 77
    dev = torch.device("cuda", 0)
    x = torch.nn.functional.normalize(torch.randn(n, d, device=dev, generator=torch.Generator(dev).manual_seed(d)), dim=1)
    pick = np.r_[0:300, n // 2:n // 2 + 300, n - 400:n]
    pdfs = [gs.MixtureModel([gs.VonMisesFisher(60.0 * m) for m in gs.sample_sphere(d - 1, 6, seed=2)]),
            gs.CurvedVonMisesFisher(gs.SlerpCurve(gs.brownian_curve(9, d, 0.5, seed=4)), 200.0),
            gs.random_bingham(d, 100.0, 0, False, 5)]
    for pdf in pdfs:
        got = pdf.log_prob(x)
        want = oracle_of(pdf).log_prob(x[pick].double().cpu().numpy())
        assert rel_err(got[pick].cpu().numpy(), want) <= 1e-5, type(pdf).__name__
        monkeypatch.setenv("GEOSSS_TC_SUB1", "1")
        small_tiles = pdf.log_prob(x)
        monkeypatch.delenv("GEOSSS_TC_SUB1")
        assert torch.equal(got, small_tiles), type(pdf).__name__


# ------------------------------------------------------------------ trajectories under the reference's draws
def tape_cases(golden):
    g = golden("traj_c1")
    pdf = mixture(g["mus"])
    yield pdf, gs.ShrinkageSphericalSliceSampler, g, "shrink"
    yield pdf, gs.RejectionSphericalSliceSampler, g, "reject"
    g = golden("traj_c2")
    for kappa in (100, 500):
        pdf = mixture(kappa * g["modes"])
        yield pdf, gs.ShrinkageSphericalSliceSampler, g, f"shrink_k{kappa}_s0"
        yield pdf, gs.ShrinkageSphericalSliceSampler, g, f"shrink_k{kappa}_s1"
        yield pdf, gs.RejectionSphericalSliceSampler, g, f"reject_k{kappa}_s0"
    g = golden("traj_c3")
    pdf = gs.CurvedVonMisesFisher(gs.SlerpCurve(g["knots"]), float(g["kappa"]))
    yield pdf, gs.ShrinkageSphericalSliceSampler, g, "shrink"
    yield pdf, gs.RejectionSphericalSliceSampler, g, "reject"
    g = golden("traj_c4")
    for tag in ("diag", "dense"):
        pdf = gs.Bingham(g[f"{tag}_A"])
        yield pdf, gs.ShrinkageSphericalSliceSampler, g, f"{tag}_shrink"
        yield pdf, gs.RejectionSphericalSliceSampler, g, f"{tag}_reject"
    g = golden("traj_c5")
    pdf = c5_target(golden)
    yield pdf, gs.ShrinkageSphericalSliceSampler, g, "shrink"
    yield pdf, gs.RejectionSphericalSliceSampler, g, "reject"


def test_tape_trajectories_float64(golden):
    for pdf, cls, g, tag in tape_cases(golden):
        want, cum = g[f"{tag}_traj"], g[f"{tag}_cumrej"]
        smp = cls(pdf, g[f"{tag}_x0"], 0, dtype="float64", path="literal")
        got, nrej, status = smp.run_tape(g[f"{tag}_tape"], len(want) - 1)
        assert status[0] == 0, tag
        assert np.array_equal(nrej[0], np.diff(cum)), tag
        assert np.max(np.abs(got[0] - want[1:])) <= 1e-12, tag
        assert smp.n_reject == cum[-1]


def test_tape_trajectories_float32(golden):
    """Literal float32 kernels fed the reference's own recorded draws, C1-C5, both samplers (1100 + 8 x 1500 + ... transitions)."""
    fc = FlipCount("tape_literal_float32_c1_to_c5")
    for pdf, cls, g, tag in tape_cases(golden):
        want, rej_want = g[f"{tag}_traj"], np.diff(g[f"{tag}_cumrej"])
        got, nrej, _ = cls(pdf, g[f"{tag}_x0"], 0, dtype="float32", path="literal").run_tape(g[f"{tag}_tape"], len(want) - 1)
        fc.add(got[0], nrej[0], want[1:], rej_want)
    fc.check(1.0)  # measured on B200: 1 flip in 8994 transitions (profiles/r2_flips.jsonl)


def test_tape_many_chains_ragged_block(golden):
    """The same tape on 130 chains (not a multiple of the block size) gives 130 identical chains."""
    g = golden("traj_c2")
    pdf = mixture(100 * g["modes"])
    tag = "shrink_k100_s0"
    n = 200
    N = 130
    x0 = np.tile(g[f"{tag}_x0"], (N, 1))
    got, nrej, status = gs.ShrinkageSphericalSliceSampler(pdf, x0, 0, dtype="float64", path="literal").run_tape(
        [g[f"{tag}_tape"]] * N, n)
    assert np.all(status == 0)
    assert np.max(np.abs(got - g[f"{tag}_traj"][None, 1:n + 1])) <= 1e-12
    assert np.all(nrej == np.diff(g[f"{tag}_cumrej"])[None, :n])
    # a tape that is too short is reported, not read past its end
    _, _, status = gs.ShrinkageSphericalSliceSampler(pdf, x0[:2], 0, dtype="float64", path="literal").run_tape(
        [g[f"{tag}_tape"][:100]] * 2, 50)
    assert np.all(status == 1)


# ------------------------------------------------------------------ own Philox stream
def oracle_chain(pdf_oracle, kind, x0, seed, chain, n, d, dtype, first_step=0, normals_to_caller=None):
    """The float64 reference algorithm fed the device's Philox draws.  ``normals_to_caller`` (d, d): the device used the
    normal draws in rotated coordinates y = W x (dense Bingham in its eigenbasis); the reference then gets z = W.T @ draws."""
    rng = philox.PhiloxChainRNG(seed=seed, chain=chain, kind=kind, d=d, dtype=dtype, first_step=first_step)
    if normals_to_caller is not None:
        draw = rng.standard_normal
        rng.standard_normal = lambda m: normals_to_caller @ np.asarray(draw(m), dtype=float)
    smp = samplers.SliceSampler(kind, pdf_oracle.log_prob, np.array(x0, float), rng)
    return smp.sample(n + 1)[1:], np.array(smp.rejects_per_step)


@pytest.mark.parametrize("cls,kind", [(gs.ShrinkageSphericalSliceSampler, samplers.SHRINK),
                                      (gs.RejectionSphericalSliceSampler, samplers.REJECT)])
def test_free_running_literal_float64_matches_oracle(golden, cls, kind):
    """Literal float64 kernel + device Philox == reference algorithm fed oracle/philox.py's draws."""
    g = golden("traj_c2")
    mus = 100.0 * g["modes"]
    x0 = gs.sample_sphere(9, 6, seed=5)
    n = 60 if kind == samplers.SHRINK else 25
    seed, off = 0xABCDEF0123456789, 2**32 - 3  # chain ids straddle the 32-bit boundary
    smp = cls(mixture(mus), x0, seed, dtype="float64", path="literal", chain_offset=off)
    got = smp.sample(n + 1)
    assert got.shape == (6, n + 1, 10) and np.array_equal(got[:, 0], x0)
    for c in range(6):
        want, rej = oracle_chain(targets.VMFMixture(mus), kind, x0[c], seed, off + c, n, 10, F64)
        assert np.max(np.abs(got[c, 1:] - want)) <= 1e-11
    assert smp.n_reject == sum(oracle_chain(targets.VMFMixture(mus), kind, x0[c], seed, off + c, n, 10, F64)[1].sum() for c in range(6))


@pytest.mark.parametrize("cls,kind", [(gs.ShrinkageSphericalSliceSampler, samplers.SHRINK),
                                      (gs.RejectionSphericalSliceSampler, samplers.REJECT)])
@pytest.mark.parametrize("cfg", ["c1", "c2_k100", "c2_k500"])
def test_fast_path_follows_reference_under_shared_draws(golden, cls, kind, cfg):
    """fp32 throughput kernel vs the float64 reference algorithm fed the same Philox draws: <= 1e-4 strictly
    before a chain's first accept/reject flip; flips per 1000 transitions reported and bounded."""
    if cfg == "c1":
        mus, d = golden("traj_c1")["mus"], 3
    else:
        mus, d = float(cfg.split("k")[1]) * golden("traj_c2")["modes"], 10
    N, n = (8, 200) if kind == samplers.SHRINK else (8, 190)  # >= 1500 transitions per configuration
    x0 = gs.sample_sphere(d - 1, N, seed=11)
    seed = 20240607
    smp = cls(mixture(mus), x0, seed, dtype="float32", path="fast")
    got, nrej = run_with_rejections(smp, n)
    assert smp.path == "fast" and got.dtype == F32
    fc = FlipCount(f"k_vmf_fast_{cfg}_{'shrink' if kind == samplers.SHRINK else 'reject'}")
    for c in range(N):
        want, rej = oracle_chain(targets.VMFMixture(mus), kind, x0[c], seed, c, n, d, F32)
        fc.add(got[c], nrej[c], want, rej)
    fc.check(1.0)


@pytest.mark.parametrize("d", list(range(2, 17)))
def test_fast_kernel_every_dimension_follows_reference(d):
    """k_vmf_fast is instantiated per true dimension (2..16): each one against the float64 reference algorithm
    fed the same Philox draws (the draw layout -- aux words in the last normal block or in their own -- depends
    on d mod 4), both samplers, K = 3."""
    mus = 25.0 * gs.sample_sphere(d - 1, 3, seed=100 + d)
    x0 = gs.sample_sphere(d - 1, 6, seed=d)
    fc = FlipCount(f"k_vmf_fast_d{d}")
    for cls, kind, n in ((gs.ShrinkageSphericalSliceSampler, samplers.SHRINK, 16), (gs.RejectionSphericalSliceSampler, samplers.REJECT, 8)):
        smp = cls(mixture(mus), x0, 977 + d, dtype="float32", path="fast")
        got, nrej = run_with_rejections(smp, n)
        assert smp.path == "fast"
        for c in range(6):
            want, rej = oracle_chain(targets.VMFMixture(mus), kind, x0[c], 977 + d, c, n, d, F32)
            fc.add(got[c], nrej[c], want, rej)
    fc.check(1.0)  # 144 transitions per dimension; measured: 0 flips in all 15 dimensions


def test_parked_device_blocks_are_reused_cleanly():
    """Short-lived samplers (the reference constructs one per sample() call) reuse parked device blocks: a
    second sampler of the same shape must start from its own state and zeroed counters, results must not
    depend on what the previous owner left behind, and the cache can be released at any time."""
    from geosss_b200 import _lib
    pdf = mixture(30.0 * gs.sample_sphere(4, 3, seed=2))
    x0 = gs.sample_sphere(4, 257, seed=3)
    first = gs.ShrinkageSphericalSliceSampler(pdf, x0, 5)
    a = first.sample(20)
    rej_a = first.n_reject
    del first
    other = gs.ShrinkageSphericalSliceSampler(pdf, gs.sample_sphere(4, 257, seed=4), 6)  # same shape: takes the parked blocks
    other.sample(7)
    del other
    again = gs.ShrinkageSphericalSliceSampler(pdf, x0, 5)
    b = again.sample(20)
    assert np.array_equal(a, b) and again.n_reject == rej_a
    assert _lib.lib().geosss_release_cached_memory() == 0
    c = gs.ShrinkageSphericalSliceSampler(pdf, x0, 5).sample(20)
    assert np.array_equal(a, c)


def gc_case(golden, cfg):
    """(device pdf, oracle pdf, d) for the great-circle fast kernels."""
    if cfg in ("c4diag", "c4dense", "d10"):
        A = golden("logprob_bingham")[f"{cfg}_A"]
        return gs.Bingham(A), targets.Bingham(A), len(A)
    if cfg == "bf":
        g = golden("logprob_bingham")
        return gs.BinghamFisher(g["bf_A"], g["bf_b"]), targets.Bingham(g["bf_A"], g["bf_b"]), len(g["bf_A"])
    knots = golden("logprob_curved")[f"{cfg}_knots"]
    return gs.CurvedVonMisesFisher(gs.SlerpCurve(knots), 500.0), targets.CurvedVMF(knots, 500.0), knots.shape[1]


@pytest.mark.parametrize("cls,kind", [(gs.ShrinkageSphericalSliceSampler, samplers.SHRINK),
                                      (gs.RejectionSphericalSliceSampler, samplers.REJECT)])
@pytest.mark.parametrize("cfg", ["c4diag", "c4dense", "d10", "bf", "c3", "d3"])
def test_great_circle_fast_kernels_follow_reference_under_shared_draws(golden, cls, kind, cfg):
    """Bingham(-Fisher) and curved-vMF throughput kernels vs the float64 reference algorithm fed the
    same Philox draws: <= 1e-4 until the first accept/reject flip."""
    pdf, ora, d = gc_case(golden, cfg)
    N, n = 12, (60 if kind == samplers.SHRINK else 20)
    if cfg in ("c3",):
        N, n = 6, (40 if kind == samplers.SHRINK else 12)  # the float64 oracle of the d=100 curve costs ~3 ms per evaluation
    x0 = gs.sample_sphere(d - 1, N, seed=13)
    if cfg.startswith("c4") or cfg == "d10":
        x0[: N // 2] = gs.radial_projection(pdf.mode + 0.3 * x0[: N // 2])  # some chains start near the mode
    seed = 777
    smp = cls(pdf, x0, seed, dtype="float32", path="fast")
    got, nrej = run_with_rejections(smp, n)
    assert smp.path == "fast" and got.dtype == F32 and got.shape == (N, n, d)
    # dense Bingham chains live in the eigenbasis of A on the device: the same draws are normals in those coordinates
    W = pdf.device_eigenrows() if isinstance(pdf, gs.Bingham) else None
    assert (W is not None) == (cfg in ("c4dense", "d10", "bf"))  # the dense matrices of the fixtures
    fc = FlipCount(f"k_gc_fast_{cfg}_{'shrink' if kind == samplers.SHRINK else 'reject'}")
    for c in range(N):
        want, rej = oracle_chain(ora, kind, x0[c], seed, c, n, d, F32, normals_to_caller=None if W is None else W.T)
        fc.add(got[c], nrej[c], want, rej)
    fc.check(1.0)


def test_dense_bingham_chains_live_in_the_eigenbasis(golden, monkeypatch):
    """Dense A on the float32 path: the chain runs as y = W x on the diagonal kernel and states / samples are rotated at the
    C-ABI boundary (tensor cores for dense blocks, one scratch tile for the ragged tail).  W is an eigen-system of A; the
    first retained sample is the current state; a second call continues bit-exactly from the last sample of the first; shards
    reproduce the full run bit for bit (the rotation of a row does not depend on where the row sits); the dense matrix-vector
    kernel (GEOSSS_NO_EIG) has the same law."""
    A = golden("logprob_bingham")["c4dense_A"]
    pdf = gs.Bingham(A)
    W = pdf.device_eigenrows()
    assert W is not None and np.max(np.abs(W @ W.T - np.eye(50))) < 1e-12
    assert np.max(np.abs(W @ A @ W.T - np.diag(np.diag(W @ A @ W.T)))) < 1e-9 * np.abs(A).max()
    N = 300  # two full 128-row tiles and a ragged tail in every rotation
    x0 = torch.as_tensor(gs.sample_sphere(49, N, seed=8), dtype=torch.float32).cuda()
    smp = gs.ShrinkageSphericalSliceSampler(pdf, x0, 5)
    first = smp.sample(7)
    assert smp.path == "fast" and tuple(first.shape) == (N, 7, 50)
    assert torch.max(torch.abs(first[:, 0] - x0)) < 2e-6                      # W^T (W x0) in 3xTF32 / float32
    assert torch.max(torch.abs(first.double().norm(dim=-1) - 1.0)) < 1e-5
    second = smp.sample(5)
    assert torch.equal(second[:, 0], first[:, -1])                             # continuation, bit for bit
    assert np.array_equal(smp.state.astype(np.float32), second[:, -1].cpu().numpy())
    full = torch.cat([first, second[:, 1:]], dim=1)
    part = gs.ShrinkageSphericalSliceSampler(pdf, x0[100:200], 5, chain_offset=100).sample(11)
    assert torch.equal(part[:, 1:], full[100:200, 1:])                         # shard invariance through the rotations
    lp = pdf.log_prob(full[:, -1])
    assert rel_err(lp.cpu().numpy(), oracle_of(pdf).log_prob(full[:, -1].double().cpu().numpy())) <= 1e-5
    # host-buffer entry point (step-major staging buffers are rotated too)
    host = gs.ShrinkageSphericalSliceSampler(pdf, x0.cpu().numpy(), 5).sample(11)
    assert np.max(np.abs(host - full.cpu().numpy())) < 1e-6
    # same law on the dense matrix-vector kernel: mean log-density after burn-in
    many = torch.as_tensor(gs.sample_sphere(49, 4096, seed=9), dtype=torch.float32).cuda()
    a = pdf.log_prob(gs.ShrinkageSphericalSliceSampler(pdf, many, 6).sample(1, 150)[:, 0]).double().mean().item()
    monkeypatch.setenv("GEOSSS_NO_EIG", "1")
    b = pdf.log_prob(gs.ShrinkageSphericalSliceSampler(gs.Bingham(A), many, 7).sample(1, 150)[:, 0]).double().mean().item()
    monkeypatch.delenv("GEOSSS_NO_EIG")
    assert abs(a - b) < 1.0, (a, b)  # log p ~ 280 +- 5 per chain: means of 4096 chains agree to a fraction of a unit


@pytest.mark.parametrize("cls,kind", [(gs.ShrinkageSphericalSliceSampler, samplers.SHRINK),
                                      (gs.RejectionSphericalSliceSampler, samplers.REJECT)])
@pytest.mark.parametrize("d,K,kappa", [(40, 12, 60.0), (17, 3, 200.0), (100, 16, 30.0)])
def test_high_dimensional_vmf_mixture_fast_kernel(cls, kind, d, K, kappa):
    """vMF mixtures beyond the register kernel's d <= 16 / K <= 8 run on the shared-memory kernel."""
    mus = kappa * gs.sample_sphere(d - 1, K, seed=d)
    w = np.linspace(1.0, 2.0, K)
    pdf, ora = mixture(mus, w), targets.VMFMixture(mus, w)
    N, n = 10, (20 if kind == samplers.SHRINK else 6)
    x0 = gs.sample_sphere(d - 1, N, seed=3)
    x0[: N // 2] = gs.radial_projection(mus[0] / kappa + 0.2 * x0[: N // 2])
    smp = cls(pdf, x0, 31337, dtype="float32", path="fast")
    got, nrej = run_with_rejections(smp, n)
    assert smp.path == "fast"
    fc = FlipCount(f"k_gc_fast_vmf_d{d}_K{K}_{'shrink' if kind == samplers.SHRINK else 'reject'}")
    for c in range(N):
        want, rej = oracle_chain(ora, kind, x0[c], 31337, c, n, d, F32)
        fc.add(got[c], nrej[c], want, rej)
    fc.check(1.0)
    # literal fp32 agrees in distribution of rejections (same algorithm): compare mean rejections loosely
    lit = cls(pdf, x0, 31337, dtype="float32", path="literal")
    lit.sample(n + 1)
    assert abs(lit.n_reject - int(nrej.sum())) <= 0.2 * max(lit.n_reject, 10)


def test_shard_invariance_and_continuation(golden):
    """Chains depend on (seed, global chain id, step) only: splitting chains over samplers or
    steps over calls does not change them (what makes 1/2/4/8-GPU runs identical)."""
    g = golden("traj_c2")
    for pdf, d in ((gc_case(golden, "c4diag")[0], 50), (gc_case(golden, "d3")[0], 3)):
        x0 = gs.sample_sphere(d - 1, 100, seed=3)
        whole = gs.ShrinkageSphericalSliceSampler(pdf, x0, 5, path="fast").sample(21)
        b = gs.ShrinkageSphericalSliceSampler(pdf, x0[37:], 5, path="fast", chain_offset=37).sample(21)
        assert np.array_equal(whole[37:], b)
        s = gs.ShrinkageSphericalSliceSampler(pdf, x0, 5, path="fast")
        first = s.sample(9).copy()
        second = s.sample(13)
        assert np.array_equal(np.concatenate([first, second[:, 1:]], axis=1), whole)
    pdf = mixture(100.0 * g["modes"])
    x0 = gs.sample_sphere(9, 200, seed=3)
    for path in ("fast", "literal"):
        whole = gs.ShrinkageSphericalSliceSampler(pdf, x0, 99, path=path).sample(30)
        a = gs.ShrinkageSphericalSliceSampler(pdf, x0[:77], 99, path=path).sample(30)
        b = gs.ShrinkageSphericalSliceSampler(pdf, x0[77:], 99, path=path, chain_offset=77).sample(30)
        assert np.array_equal(whole, np.concatenate([a, b]))
        s = gs.ShrinkageSphericalSliceSampler(pdf, x0, 99, path=path)
        first = s.sample(12).copy()
        second = s.sample(19)  # starts from the current state: 18 new transitions
        assert np.array_equal(second[:, 0], first[:, -1])
        assert np.array_equal(np.concatenate([first, second[:, 1:]], axis=1), whole)


def test_sample_semantics(golden):
    g = golden("traj_c1")
    pdf = mixture(g["mus"])
    x0 = np.array([-0.86, 0.19, -0.47])
    s = gs.ShrinkageSphericalSliceSampler(pdf, x0, 3521, dtype="float64")
    out = s.sample(1000, 100)
    assert out.shape == (1000, 3) and out.dtype == F64
    assert s.rng.steps_done == 1099  # n + burnin - 1 transitions (mcmc.py:67-68)
    allout = gs.ShrinkageSphericalSliceSampler(pdf, x0, 3521, dtype="float64").sample(1000, 100, return_all_samples=True)
    assert allout.shape == (1100, 3) and np.array_equal(allout[0], x0) and np.array_equal(allout[100:], out)
    frac = gs.ShrinkageSphericalSliceSampler(pdf, x0, 3521, dtype="float64").sample(1000, 0.1)
    assert np.array_equal(frac, out)
    # norms relax towards 1 but the initial state is not normalised (README.md:52)
    assert abs(np.linalg.norm(allout[0]) - 1) > 1e-3 and np.all(np.abs(np.linalg.norm(out, axis=1) - 1) < 0.05)
    # iterator protocol + counters
    it = gs.RejectionSphericalSliceSampler(pdf, x0, 1, dtype="float64")
    gs.MixtureModel.log_prob.reset_counters()
    y = next(it)
    assert y.shape == (3,) and np.array_equal(y, it.state)
    rejected = it.n_reject  # reading it folds the device counters into log_prob.num_calls
    assert gs.MixtureModel.log_prob.num_calls == 2 + rejected
    with pytest.raises(AssertionError):
        it.sample(0)
    # torch in -> torch out, on the device
    xt = torch.as_tensor(gs.sample_sphere(2, 64, seed=1)).cuda()
    tt = gs.ShrinkageSphericalSliceSampler(pdf, xt, 7).sample(50, 10)
    assert isinstance(tt, torch.Tensor) and tt.is_cuda and tuple(tt.shape) == (64, 50, 3)


def test_host_entry_point_equals_device_run(golden):
    import ctypes as C
    g = golden("traj_c2")
    pdf = mixture(100.0 * g["modes"])
    N, n, burn = 300, 40, 7
    x0 = np.ascontiguousarray(gs.sample_sphere(9, N, seed=8))
    want = gs.ShrinkageSphericalSliceSampler(pdf, x0, 5).sample(n, burn)
    dev = torch.device("cuda", torch.cuda.current_device())
    out = torch.empty((n, N, 10), dtype=torch.float32, pin_memory=True)
    nrej = np.zeros(N, dtype=np.uint64)
    _lib.check(_lib.lib().geosss_sample_host(pdf._handle(dev), _lib.SHRINK, _lib.F32, _lib.PATH_FAST, N, _lib.dptr(x0), 5, 0,
                                             n, burn, out.data_ptr(), nrej.ctypes.data_as(C.c_void_p)))
    assert np.array_equal(out.numpy().transpose(1, 0, 2), want)
    assert nrej.sum() > 0
    out0 = torch.empty((n, N, 10), dtype=torch.float32, pin_memory=True)
    _lib.check(_lib.lib().geosss_sample_host(pdf._handle(dev), _lib.SHRINK, _lib.F32, _lib.PATH_FAST, N, _lib.dptr(x0), 5, 0,
                                             n, 0, out0.data_ptr(), None))
    assert np.array_equal(out0.numpy()[0], x0.astype(F32))


# ------------------------------------------------------------------ statistical agreement under own RNG
def ks_2samp_p(a, b):
    from scipy.stats import ks_2samp
    return ks_2samp(a, b).pvalue


@pytest.mark.parametrize("cls", [gs.ShrinkageSphericalSliceSampler, gs.RejectionSphericalSliceSampler])
@pytest.mark.parametrize("path", ["fast", "literal"])
def test_marginals_match_exact_vmf_mixture(cls, path):
    """Final states of many independent chains are i.i.d.; compare with exact Wood draws."""
    from scipy.special import i0, iv
    mus = np.array([[6.0, 0, 0, 0, 0], [0, -4.0, 3.0, 0, 0], [0, 0, 0, 5.0, -2.0]])
    w = np.array([0.5, 0.3, 0.2])
    pdf = mixture(mus, w)
    # The reference's component "normaliser" is 2 pi i0(kappa) whatever d is (distributions.py:157), so
    # with unequal kappas the mixture's true component masses are w_k Z_d(kappa_k) / (2 pi i0(kappa_k)).
    kap = np.linalg.norm(mus, axis=1)
    w_eff = w * ((2 * np.pi) ** 2.5 * iv(1.5, kap) / kap**1.5) / (2 * np.pi * i0(kap))
    N = 8192
    x0 = gs.sample_sphere(4, N, seed=21)
    burn = 150 if cls is gs.ShrinkageSphericalSliceSampler else 40
    got = cls(pdf, x0, 77, path=path).sample(1, burn)[:, 0].astype(float)
    ref = exact.wood_vmf_mixture(mus, w_eff, N, seed=5)
    assert np.allclose(np.linalg.norm(got, axis=1), 1, atol=1e-4)
    for j in range(5):
        assert ks_2samp_p(got[:, j], ref[:, j]) > 1e-3, j
    for m in mus:
        assert ks_2samp_p(got @ m, ref @ m) > 1e-3


def test_cpd_cutoff_sampler_has_the_law_of_the_literal_float64_sampler(golden):
    """Own-RNG check of the registration sampler: the float32 cutoff kernel and the float64 literal kernel (exact
    k-nearest over all pairs) are the same Markov kernel, so after the same number of transitions from the same
    initial states their chains have the same law -- compared through log p and the rotation angle of the final
    states (two-sample KS, alpha = 1e-3), with independent seeds."""
    g = golden("logprob_cpd")
    n = len(g["c5_source"])
    pdf = gs.CoherentPointDrift(gs.PointCloud(g["c5_target"]), gs.PointCloud(g["c5_source"], np.ones(n) / n),
                                sigma=1.0, k=20, omega=0.4)
    N, steps = 1024, 25
    x0 = gs.sample_sphere(3, N, seed=31)
    fast = gs.ShrinkageSphericalSliceSampler(pdf, x0, 101, dtype="float32")
    a = fast.sample(1, steps)[:, 0].astype(float)
    lit = gs.ShrinkageSphericalSliceSampler(pdf, x0, 202, dtype="float64")
    b = lit.sample(1, steps)[:, 0]
    assert fast.path == "fast" and lit.path == "literal"
    la, lb = pdf.log_prob(a), pdf.log_prob(b)
    assert ks_2samp_p(la, lb) > 1e-3, (np.mean(la), np.mean(lb))
    assert ks_2samp_p(np.abs(a[:, 3]), np.abs(b[:, 3])) > 1e-3
    # both climbed: the initial states sit near log p = -2475
    assert np.mean(la) > np.mean(pdf.log_prob(x0)) + 20 and np.mean(lb) > np.mean(pdf.log_prob(x0)) + 20
    assert abs(fast.n_reject - lit.n_reject) < 0.1 * lit.n_reject


# ------------------------------------------------------------------ diagnostics
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_device_ess_matches_oracle(dtype):
    from geosss_b200 import diagnostics
    from oracle import ess as E
    gen = np.random.default_rng(2)
    n, N, d = 600, 50, 4
    x = np.zeros((n, N, d))
    e = gen.standard_normal(x.shape)
    for t in range(1, n):
        x[t] = 0.75 * x[t - 1] + e[t]
    xt = torch.as_tensor(x, dtype=dtype).cuda()
    v = np.array([0.5, -1.0, 0.25, 2.0])
    for direction, trace in ((None, x[:, :, 0].T), (v, (x @ v).T)):
        got, rel = diagnostics.ess(xt, direction, 200, chain_major=False)
        want = E.ess_multichain(trace, 200)
        got_cm, _ = diagnostics.ess(xt.permute(1, 0, 2).contiguous(), direction, 200)  # (chains, draws, d), as sample() returns
        assert abs(got_cm - got) <= 1e-9 * got
        assert abs(got - want) <= (1e-4 if dtype == torch.float64 else 2e-3) * want  # the trace is staged in fp32 shared memory
        assert abs(rel - want / (n * N)) < 1e-3
    # relative ESS of an AR(1) with r = 0.75 is (1 - r) / (1 + r)
    assert abs(diagnostics.ess(xt, None, 200, chain_major=False)[1] - 0.25 / 1.75) < 0.03


def test_device_bulk_ess_matches_independent_restatement():
    """Rank-normalised split bulk ESS (what the upstream scripts get from arviz.ess): device implementation
    (ranks by a device sort, autocovariance partial sums in csrc/diag.cuh, truncation rules on the host) against
    the independently written numpy restatement oracle/ess.py:bulk_ess; known answers for AR(1) traces."""
    from geosss_b200 import diagnostics
    from oracle import ess as E
    gen = np.random.default_rng(4)
    for phi, N, n in ((0.0, 16, 400), (0.75, 24, 1001), (0.97, 8, 3000), (-0.5, 5, 64)):
        e = gen.standard_normal((N, n))
        x = np.zeros_like(e)
        for t in range(1, n):
            x[:, t] = phi * x[:, t - 1] + e[:, t]
        x = x.astype(np.float32)  # the device ranks float32 values: give both sides the same ties
        x[:, ::7] = np.round(x[:, ::7], 1)  # ... and some exact ties
        got = diagnostics.bulk_ess(torch.as_tensor(x).cuda())
        want = E.bulk_ess(x.astype(float))
        assert abs(got - want) <= 2e-4 * want, (phi, got, want)
        rel = diagnostics.bulk_ess(torch.as_tensor(x).cuda(), relative=True)
        assert abs(rel - want / (N * (2 * (n // 2)))) <= 2e-4 * rel
    # AR(1): ESS / S -> (1 - phi) / (1 + phi)
    e = gen.standard_normal((64, 4000))
    x = np.zeros_like(e)
    for t in range(1, 4000):
        x[:, t] = 0.75 * x[:, t - 1] + e[:, t]
    assert abs(diagnostics.bulk_ess(torch.as_tensor(x, dtype=torch.float32).cuda(), relative=True) - 0.25 / 1.75) < 0.01


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_streaming_ess_equals_retained_trace_estimate(golden, dtype):
    """Chunks of a scalar trace folded into per-chain accumulators on the device (never retained) give the estimator of the
    whole trace: against the float64 oracle on AR(1) traces cut into ragged chunks (shorter than, equal to and longer than
    max_lag), and for a real run -- streaming_ess(sampler) against ess() of the same chain's retained trace."""
    from geosss_b200 import diagnostics
    from oracle import ess as E
    gen = np.random.default_rng(7)
    N, n, L = 37, 1500, 64
    x = np.zeros((N, n))
    e = gen.standard_normal((N, n))
    for t in range(1, n):
        x[:, t] = 0.8 * x[:, t - 1] + e[:, t]
    x += 3.0  # a mean far from zero: the raw lagged sums must cancel it
    x = x.astype(np.float32).astype(float)  # the accumulators stage the trace in float32 (as k_ess_partials does), products in double
    want = E.ess_multichain(x, L)
    for cuts in ([n], [10, 40, 64, 65, 300, 7, 1014], [1] * 5 + [n - 5]):
        acc = diagnostics.StreamingESS(N, L)
        pos = 0
        for m in cuts:
            acc.update(torch.as_tensor(x[:, pos:pos + m], dtype=dtype).cuda())
            pos += m
        assert pos == n and acc.n == n
        got, rel = acc.result()
        assert abs(got - want) <= 1e-9 * want, (cuts[:3], got, want)
        assert abs(rel - want / (N * n)) <= 1e-12
    # a real run: the same seed gives the same chain with and without retaining the trace
    g = golden("traj_c2")
    pdf = mixture(100.0 * g["modes"])
    x0 = torch.as_tensor(gs.sample_sphere(9, 300, seed=4)).cuda()
    v = gs.sample_sphere(9, seed=5)
    kw = dict(dtype="float32" if dtype == torch.float32 else "float64")
    full = gs.ShrinkageSphericalSliceSampler(pdf, x0, 12, **kw).sample_trace(700, v, burnin=50)
    want2, _ = diagnostics.ess(full.unsqueeze(-1), np.ones(1), 100)
    got2, _ = diagnostics.streaming_ess(gs.ShrinkageSphericalSliceSampler(pdf, x0, 12, **kw), 700, v, burnin=50, chunk=128, max_lag=100)
    assert abs(got2 - want2) <= 1e-6 * want2, (got2, want2)


def test_sample_trace_equals_projection_of_retained_samples(golden):
    """The scalar test-function trace written by the sampler kernels (4 bytes per kept transition) equals
    x . direction of the fully retained chain, for the register-resident, the shared-memory and the literal kernels."""
    g = golden("traj_c2")
    cases = [(mixture(100.0 * g["modes"]), "float32", 2e-6), (gs.random_bingham(50, 300, 0, True, 6982), "float32", 2e-6),
             (gs.random_bingham(12, 30, 0, False, 3), "float32", 2e-6),
             (gs.CurvedVonMisesFisher(gs.SlerpCurve(gs.brownian_curve(10, 100, 0.5, seed=4562)), 500.0), "float32", 2e-6),
             (mixture(100.0 * g["modes"]), "float64", 1e-14)]
    for pdf, dtype, tol in cases:
        d = pdf.d
        x0 = gs.sample_sphere(d - 1, 70, seed=d)
        v = gs.sample_sphere(d - 1, seed=1)
        for thin in (1, 3):
            a = gs.ShrinkageSphericalSliceSampler(pdf, torch.as_tensor(x0).cuda(), 5, dtype=dtype)
            full = a.sample(41 * thin + 10, 10, return_all_samples=True)  # (N, 41 thin + 20, d): initial state first
            b = gs.ShrinkageSphericalSliceSampler(pdf, torch.as_tensor(x0).cuda(), 5, dtype=dtype)
            tr = b.sample_trace(41, v, burnin=19, thin=thin)
            assert tuple(tr.shape) == (70, 41) and tr.is_cuda
            kept = full[:, 19 + thin::thin][:, :41]
            want = kept.double().cpu().numpy() @ v
            assert np.max(np.abs(tr.double().cpu().numpy() - want)) <= tol * 10, (type(pdf).__name__, dtype, thin)
            assert a.n_reject >= b.n_reject > 0
    # default direction: the first coordinate (scripts/ess_mixture_vMF.py:233)
    one = gs.ShrinkageSphericalSliceSampler(cases[0][0], x0[0], 9).sample_trace(30)
    ref = gs.ShrinkageSphericalSliceSampler(cases[0][0], torch.as_tensor(x0[0]).cuda(), 9).sample(31)
    assert tuple(one.shape) == (30,) and torch.equal(one, ref[1:, 0])


def test_launcher_and_bingham_relative_ess(golden):
    """Published sampler quality (scripts/Bingham.ipynb:254, arviz bulk ESS of x . mode over 10 runs of
    1e5 draws started at the mode, scripts/bingham.py:28-70,113-134): GeoSSS-reject 99.73 %, GeoSSS-shrink 15.2 %
    on the d=10 Bingham target.  Here: 64 chains x 20,000 draws after 2,000 of burn-in, the same estimator."""
    from geosss_b200 import diagnostics
    pdf = gs.random_bingham(10, 30, 0, True, 6982)
    la = gs.SamplerLauncher(pdf, torch.as_tensor(np.tile(pdf.mode, (16, 1))).cuda(), 300, burnin=0.2, seed=3)
    assert tuple(la.run("sss-shrink").shape) == (16, 300, 10)
    x0 = np.tile(pdf.mode, (64, 1))
    out = {}
    for name, cls in (("sss-reject", gs.RejectionSphericalSliceSampler), ("sss-shrink", gs.ShrinkageSphericalSliceSampler)):
        tr = cls(pdf, torch.as_tensor(x0).cuda(), 3).sample_trace(20000, pdf.mode, burnin=2000)
        out[name] = diagnostics.bulk_ess(tr, relative=True)
    print("relative bulk ESS of x . mode, Bingham d=10:", out)
    assert out["sss-reject"] >= 0.9
    assert abs(out["sss-shrink"] - 0.152) <= 0.04


@pytest.mark.parametrize("d,vmax,want", [(10, 30.0, {"sss-reject": 1.58, "sss-shrink": 0.54}),
                                         (50, 300.0, {"sss-reject": 1.57, "sss-shrink": 0.20})])
def test_published_mean_geodesic_step_on_bingham(d, vmax, want):
    """BASELINE.md / scripts/Bingham.ipynb:276-279, 515-518: mean geodesic distance between successive samples
    on random_bingham(d, vmax, 0, eigensystem=True, seed=6982), chains started at the mode."""
    pdf = gs.random_bingham(d, vmax, 0.0, True, 6982)
    x0 = torch.as_tensor(np.tile(pdf.mode, (128, 1))).cuda()
    for method, target in want.items():
        s = gs.SamplerLauncher(pdf, x0, 3000, burnin=0.1, seed=11).run(method)  # (N, n, d) on the device
        step = torch.arccos(torch.clamp((s[:, 1:] * s[:, :-1]).sum(-1), -1, 1)).mean().item()
        assert abs(step - target) <= 0.04, (method, step, target)


# ------------------------------------------------------------------ C-ABI details: thinning, logp / rejection outputs, edge shapes
def _run_raw(pdf, cls, x0, seed, n_steps, thin, path):
    """geosss_sampler_run through the Python wrapper's private hook, with every optional output."""
    smp = cls(pdf, torch.as_tensor(x0).cuda(), seed, path=path)
    n_keep = n_steps // thin
    N, d = x0.shape
    samples = torch.empty((n_keep, N, d), dtype=torch.float32, device="cuda")
    logp = torch.empty((n_keep, N), dtype=torch.float32, device="cuda")
    nrej = torch.zeros((n_keep, N), dtype=torch.int32, device="cuda")
    smp._run_device(n_steps, samples, thin=thin, logp=logp, nrej=nrej)
    torch.cuda.synchronize()
    return smp, samples.cpu().numpy(), logp.cpu().numpy(), nrej.cpu().numpy()


@pytest.mark.parametrize("cfg", ["vmf", "c4diag", "d3", "cpd"])
@pytest.mark.parametrize("cls", [gs.ShrinkageSphericalSliceSampler, gs.RejectionSphericalSliceSampler])
def test_thinning_logp_and_rejection_outputs(golden, cfg, cls):
    if cfg == "vmf":
        pdf, d = mixture(100.0 * golden("traj_c2")["modes"]), 10
    elif cfg == "cpd":
        g = golden("logprob_cpd")
        n = len(g["c5_source"])
        pdf, d = gs.CoherentPointDrift(gs.PointCloud(g["c5_target"]), gs.PointCloud(g["c5_source"], np.ones(n) / n),
                                       sigma=1.0, k=20, omega=0.4, dense=True), 4
    else:
        pdf, _, d = gc_case(golden, cfg)
    N = 70 if cfg != "cpd" else 9
    steps = 24 if cfg != "cpd" else 6
    x0 = gs.sample_sphere(d - 1, N, seed=4)
    smp1, s1, lp1, r1 = _run_raw(pdf, cls, x0, 9, steps, 1, "fast")
    smp3, s3, lp3, r3 = _run_raw(pdf, cls, x0, 9, steps, 3, "fast")
    assert smp1.path == "fast"
    # thinning keeps every third state of the same chain and sums the rejections in between
    assert np.array_equal(s3, s1[2::3]) and np.array_equal(lp3, lp1[2::3])
    assert np.array_equal(r3, r1.reshape(steps // 3, 3, N).sum(1))
    assert int(r1.sum()) == smp1.n_reject
    # logp_out is the log-density of the kept state
    want = np.asarray(pdf.log_prob(s1.reshape(-1, d).astype(np.float64))).reshape(steps, N)
    tol = 2e-3 if cfg == "cpd" else 3e-4
    assert np.max(np.abs(lp1 - want) / np.maximum(1.0, np.abs(want))) <= tol


def test_edge_shapes_fast_path():
    """One chain, K = 1, d = 2..16 and a chain count that is not a multiple of the block size."""
    gen = np.random.default_rng(3)
    for d in (2, 3, 4, 5, 8, 13, 16):
        mus = 12.0 * gs.sample_sphere(d - 1, 2, seed=d)
        for pdf in (mixture(mus), gs.VonMisesFisher(mus[0])):
            for N in (1, 33):
                x0 = gs.sample_sphere(d - 1, N, seed=1)
                out = gs.ShrinkageSphericalSliceSampler(pdf, x0 if N > 1 else x0[0], 5, path="fast").sample(30, 5)
                assert out.shape == ((N, 30, d) if N > 1 else (30, d))
                assert np.all(np.isfinite(out)) and np.allclose(np.linalg.norm(out, axis=-1), 1, atol=1e-4)
    # n_steps = 0 is a no-op; a second call continues
    pdf = mixture(12.0 * gs.sample_sphere(4, 3, seed=2))
    smp = gs.RejectionSphericalSliceSampler(pdf, torch.as_tensor(gs.sample_sphere(4, 5, seed=3)).cuda(), 1)
    smp._run_device(0, None)
    assert smp.rng.steps_done == 0 and smp.n_reject == 0


# ------------------------------------------------------------------ BASELINE.json full sizes: size-independent properties
@pytest.mark.parametrize("cfg,N,steps", [("c2", 65536, 40), ("c3", 1 << 20, 3), ("c4", 1 << 20, 4), ("c5", 1 << 18, 2)])
def test_full_size_configs_properties(golden, cfg, N, steps):
    """At the chain counts BASELINE.json names: every retained state is finite and on the sphere, the
    first 96 chains are bit-identical to a 96-chain run (chains depend on (seed, chain id, step) only),
    logp_out equals log_prob of the retained state, rejection totals add up."""
    if cfg == "c2":
        pdf, d = mixture(100.0 * golden("traj_c2")["modes"]), 10
    elif cfg == "c3":
        pdf, d = gs.CurvedVonMisesFisher(gs.SlerpCurve(gs.brownian_curve(10, 100, 0.5, seed=4562)), 500.0), 100
    elif cfg == "c4":
        pdf, d = gs.random_bingham(50, 300, 0, True, 6982), 50
    else:
        g = golden("logprob_cpd")
        n = len(g["c5_source"])
        pdf, d = gs.CoherentPointDrift(gs.PointCloud(g["c5_target"]), gs.PointCloud(g["c5_source"], np.ones(n) / n),
                                       sigma=1.0, k=20, omega=0.4, dense=True), 4
    gen = torch.Generator(device="cuda").manual_seed(5)
    x0 = torch.nn.functional.normalize(torch.randn(N, d, device="cuda", generator=gen, dtype=torch.float64), dim=1)
    smp = gs.ShrinkageSphericalSliceSampler(pdf, x0, 2024)
    out = torch.empty((N, steps, d), dtype=torch.float32, device="cuda")
    logp = torch.empty((N, steps), dtype=torch.float32, device="cuda")
    nrej = torch.zeros((N, steps), dtype=torch.int32, device="cuda")
    smp._run_device(steps, out, logp=logp, nrej=nrej, chain_major=True)
    torch.cuda.synchronize()
    assert smp.path == "fast"
    assert bool(torch.isfinite(out).all())
    assert float((out.norm(dim=-1) - 1).abs().max()) < 2e-4
    assert int(nrej.sum()) == smp.n_reject and smp.n_reject > 0
    small = gs.ShrinkageSphericalSliceSampler(pdf, x0[:96], 2024)
    out_s = torch.empty((96, steps, d), dtype=torch.float32, device="cuda")
    small._run_device(steps, out_s, chain_major=True)
    assert torch.equal(out[:96], out_s)
    idx = torch.arange(0, N, max(1, N // 512), device="cuda")
    want = pdf.log_prob(out[idx, -1].double())
    tol = 2e-3 if cfg == "c5" else 3e-4
    assert float(((logp[idx, -1].double() - want).abs() / want.abs().clamp(min=1)).max()) <= tol


def test_long_run_stays_on_the_sphere_and_stationary():
    """2 * 10^5 transitions per chain in fp32: states stay unit and finite, the rejection rate of the second
    half equals that of the first half (no drift), every chain sits at a mode-like density."""
    modes = gs.sample_sphere(9, 5, seed=1234)
    pdf = mixture(100.0 * modes)
    x0 = torch.as_tensor(gs.sample_sphere(9, 2048, seed=2)).cuda()
    smp = gs.ShrinkageSphericalSliceSampler(pdf, x0, 99)
    smp._run_device(100_000, None)
    r1 = smp.n_reject
    smp._run_device(100_000, None)
    r2 = smp.n_reject - r1
    x = smp.state
    assert np.all(np.isfinite(x)) and np.max(np.abs(np.linalg.norm(x, axis=1) - 1)) < 1e-5
    assert abs(r2 / r1 - 1) < 0.01
    assert np.mean(np.max(x @ modes.T, axis=1)) > 0.95
    assert smp.rng.steps_done == 200_000


def test_systematic_rotation_search(golden):
    """scripts/protein_reg3d3d.py:629-648: log_prob over the 600-cell grid in one batched call; same values
    and the same best rotation as the reference's Python loop."""
    g = golden("tessellation")
    pdf64 = c5_target(golden)  # k = 20, exact truncation
    q_best, lp_best, quats, logp = gs.systematic_search(pdf64, 1)
    assert rel_err(logp, g["logp_n1"]) <= 1e-11, rel_err(logp, g["logp_n1"])
    best_ref = int(np.argmax(g["logp_n1"]))
    assert abs(abs(q_best @ g["quats_n1"][best_ref]) - 1) < 1e-12 and abs(lp_best - g["logp_n1"][best_ref]) <= 1e-8
    c = golden("logprob_cpd")
    n = len(c["c5_source"])
    fast = gs.CoherentPointDrift(gs.PointCloud(c["c5_target"]), gs.PointCloud(c["c5_source"], np.ones(n) / n),
                                 sigma=1.0, k=20, omega=0.4, dense=True)
    _, _, _, logp32 = gs.systematic_search(fast, quaternions=quats.astype(F32))
    assert rel_err(logp32, g["logp_n1"]) <= 2e-5, rel_err(logp32, g["logp_n1"])
    assert int(np.argmax(logp32)) == best_ref


@pytest.mark.parametrize("dtype", [F64, F32])
@pytest.mark.parametrize("chain_major", [True, False])
def test_device_path_statistics(golden, dtype, chain_major):
    """Mode populations / KL, hopping frequency and geodesic steps on the device vs the reference's
    script expressions evaluated on reference runs (tests/golden/pathstats.npz)."""
    from geosss_b200 import diagnostics as dg
    g = golden("pathstats")

    def dev(x):
        t = torch.as_tensor(x.astype(dtype)).cuda()
        return t if chain_major else t.transpose(0, 1).contiguous()

    xs, tol = dev(g["vmf_samples"]), (1e-12 if dtype == F64 else 2e-4)
    st = dg.path_stats(xs, g["vmf_modes"], chain_major=chain_major)
    assert np.array_equal(st["counts"], g["vmf_counts"]) and st["pairs"] == 3 * 999
    assert abs(st["distance_sum"] - g["vmf_dist"].sum()) <= tol * g["vmf_dist"].sum()
    pdf = mixture(100.0 * g["vmf_modes"])
    assert abs(dg.mode_kl(xs, pdf, chain_major=chain_major) - g["vmf_kl"]) < 1e-12
    assert abs(dg.mean_geodesic_step(xs, chain_major=chain_major) - g["vmf_dist"].mean()) <= tol
    bx = dev(g["bingham_samples"])
    bing = gs.Bingham(g["bingham_A"])
    assert np.allclose(np.abs(bing.mode @ g["bingham_mode"]), 1, atol=1e-12)
    assert abs(dg.hopping_frequency(bx, bing, chain_major=chain_major) - g["bingham_hop"].mean()) < 1e-12
    assert abs(dg.mean_geodesic_step(bx, chain_major=chain_major) - g["bingham_dist"].mean()) <= tol
    with pytest.raises(RuntimeError):
        dg.path_stats(xs, np.zeros((65, 10)), chain_major=chain_major)
    with pytest.raises(ValueError):
        dg.path_stats(g["vmf_samples"], g["vmf_modes"])
