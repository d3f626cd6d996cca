"""The paper's four-way comparison on the GPU: sss-reject, sss-shrink, RWMH and spherical HMC on one target, every sampler
through the public API with the scalar test function written by the kernel (no states retained), throughput from CUDA events
and ESS/s from the rank-normalised split bulk ESS of the trace (scripts/mixture_vMF.py:406-411, scripts/vMF_diagnostics.py:451-492).

    python benchmarks/bench_baselines.py [--chains 65536] [--draws 400] [--kappa 100]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geosss_b200 as gs  # noqa: E402
from geosss_b200 import diagnostics  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=65536)
    ap.add_argument("--draws", type=int, default=400)
    ap.add_argument("--burnin", type=int, default=100)
    ap.add_argument("--kappa", type=float, default=100.0)
    ap.add_argument("--d", type=int, default=10)
    ap.add_argument("--K", type=int, default=5)
    ap.add_argument("--ess-chains", type=int, default=2048)
    ap.add_argument("--target", default="vmf", choices=["vmf", "cpd"], help="cpd: C5 registration (214 x 214 points, sigma=1, omega=0.4, k=20)")
    a = ap.parse_args()
    if a.target == "cpd":
        return registration(a)
    dev = torch.device("cuda", 0)
    modes = gs.sample_sphere(a.d - 1, a.K, seed=1234)
    pdf = gs.MixtureModel([gs.VonMisesFisher(a.kappa * m) for m in modes])
    x0 = torch.as_tensor(gs.sample_sphere(a.d - 1, a.chains, seed=4), dtype=torch.float32, device=dev)
    makers = {
        "sss-reject": lambda: gs.RejectionSphericalSliceSampler(pdf, x0, 7),
        "sss-shrink": lambda: gs.ShrinkageSphericalSliceSampler(pdf, x0, 7),
        "rwmh": lambda: gs.MetropolisHastings(pdf, x0, 7, stepsize=1e-1),
        "hmc": lambda: gs.SphericalHMC(pdf, x0, 7, stepsize=1e-1),
    }
    for name, make in makers.items():
        smp = make()
        smp.sample_trace(8, burnin=a.burnin)  # burn-in = the adaptation window of the two baselines
        out = torch.empty((a.chains, a.draws), dtype=torch.float32, device=dev)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        smp.sample_trace(a.draws, out=out)
        t1.record()
        torch.cuda.synchronize()
        dt = t0.elapsed_time(t1) * 1e-3
        sub = out[: a.ess_chains].contiguous()
        ess = diagnostics.bulk_ess(sub)
        row = {"sampler": name, "path": smp.path, "chains": a.chains, "draws": a.draws, "d": a.d, "K": a.K, "kappa": a.kappa,
               "seconds": dt, "samples_per_s": a.chains * a.draws / dt,
               "relative_bulk_ess": ess / sub.numel(), "ess_per_s": ess / sub.numel() * a.chains * a.draws / dt,
               "test_function": "first coordinate", "ess_chains": a.ess_chains}
        if name in ("rwmh", "hmc"):
            row["acceptance"] = float(smp.n_accept) / (a.chains * (a.draws + a.burnin + 8))
            row["stepsize_mean"] = float(np.mean(smp.stepsize))
        else:
            row["mean_rejections"] = float(smp.n_reject) / (a.chains * (a.draws + a.burnin + 8))
        print(json.dumps(row), flush=True)


def registration(a):
    """C5: slice samplers, random-walk MH and HMC (10 leapfrog steps) on the cutoff-list kernels (path 'fast'), RWMH and HMC on the
    dense literal float32 kernels beside them (fewer chains: they are 50-90x slower)."""
    dev = torch.device("cuda", 0)
    g = np.load(os.path.join(ROOT, "tests", "golden", "logprob_cpd.npz"))
    n = len(g["c5_source"])
    pdf = gs.CoherentPointDrift(gs.PointCloud(g["c5_target"]), gs.PointCloud(g["c5_source"], np.ones(n) / n), sigma=1.0, k=20, omega=0.4)
    cases = [
        ("sss-reject", a.chains, lambda x0: gs.RejectionSphericalSliceSampler(pdf, x0, 7)),
        ("sss-shrink", a.chains, lambda x0: gs.ShrinkageSphericalSliceSampler(pdf, x0, 7)),
        ("rwmh", a.chains, lambda x0: gs.MetropolisHastings(pdf, x0, 7, stepsize=0.05)),
        ("rwmh (literal float32, dense sum)", a.chains // 16, lambda x0: gs.MetropolisHastings(pdf, x0, 7, stepsize=0.05, path="literal")),
        ("hmc", a.chains, lambda x0: gs.SphericalHMC(pdf, x0, 7, stepsize=1e-3, n_steps=10)),
        ("hmc (literal float32, dense sum)", a.chains // 16, lambda x0: gs.SphericalHMC(pdf, x0, 7, stepsize=1e-3, n_steps=10, path="literal")),
    ]
    for name, N, make in cases:
        x0 = torch.as_tensor(gs.sample_sphere(3, N, seed=4), dtype=torch.float32, device=dev)
        smp = make(x0)
        smp.sample(2, burnin=4)
        out = torch.empty((N, a.draws + 1, 4), dtype=torch.float32, device=dev)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        got = smp.sample(a.draws + 1, out=out)
        t1.record()
        torch.cuda.synchronize()
        dt = t0.elapsed_time(t1) * 1e-3
        lp = pdf.log_prob(got[:, -1])
        print(json.dumps({"sampler": name, "path": smp.path, "chains": N, "draws": a.draws, "target": "C5 registration 214x214 k=20",
                          "seconds": dt, "samples_per_s": N * a.draws / dt, "mean_log_prob_last": float(lp.mean())}), flush=True)


if __name__ == "__main__":
    main()
