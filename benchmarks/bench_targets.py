"""
Secondary measurements (not the driver's bench line): samples/s of the other BASELINE configs and
standalone batched log_prob bandwidth, one JSON line per case.

    python benchmarks/bench_targets.py [--quick]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geosss_b200 as gs  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e-3 / reps


def c5_target(dense=True):
    g = np.load(os.path.join(ROOT, "tests", "golden", "logprob_cpd.npz"))
    n = len(g["c5_source"])
    return gs.CoherentPointDrift(gs.PointCloud(g["c5_target"], np.ones(len(g["c5_target"]))),
                                 gs.PointCloud(g["c5_source"], np.ones(n) / n), sigma=1.0, k=20, omega=0.4, dense=dense)


def cases(quick):
    modes = gs.sample_sphere(9, 5, seed=1234)
    yield "C2 vmf-mix d=10 K=5 kappa=100", gs.MixtureModel([gs.VonMisesFisher(100.0 * m) for m in modes]), 65536, 200, (1 << 22 if quick else 1 << 24)
    yield "C3 curved-vmf d=100 kappa=500", gs.CurvedVonMisesFisher(gs.SlerpCurve(gs.brownian_curve(10, 100, 0.5, seed=4562)), 500.0), \
        (1 << 16 if quick else 1 << 20), 20, (1 << 20 if quick else 1 << 22)
    # log_prob sizes: launches of >= 150 us, so that the host side of pdf.log_prob (tens of us per call) stays hidden
    yield "C4 bingham d=50 diag", gs.random_bingham(50, 300, 0, True, 6982), (1 << 16 if quick else 1 << 20), 20, (1 << 20 if quick else 1 << 22)
    yield "C4 bingham d=50 dense", gs.random_bingham(50, 300, 0, False, 6982), (1 << 16 if quick else 1 << 20), 20, (1 << 20 if quick else 1 << 22)
    yield "C5 cpd 214x214 dense", c5_target(), (1 << 12 if quick else 1 << 15), 5, (1 << 14 if quick else 1 << 17)
    yield "C5 cpd 214x214 k=20 (scripts/protein_reg3d3d.py as written)", c5_target(False), (1 << 12 if quick else 1 << 15), 5, (1 << 14 if quick else 1 << 17)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--baselines", action="store_true", help="also time the RWMH / HMC baselines (literal fp32 kernels)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    for name, pdf, N, T, n_lp in cases(a.quick):
        if a.only and a.only not in name:
            continue
        d = pdf.d
        x0 = torch.as_tensor(gs.sample_sphere(d - 1, N, seed=1)).to(dev)
        for cls, tag in ((gs.ShrinkageSphericalSliceSampler, "shrink"), (gs.RejectionSphericalSliceSampler, "reject")):
            if tag == "reject" and "C2" not in name and not a.only:
                continue
            smp = cls(pdf, x0, 11)
            smp._run_device(10, None)
            torch.cuda.synchronize()
            r0 = int(smp.n_reject_per_chain.sum())
            dt = timed(lambda: smp._run_device(T, None), 2)
            rej = (int(smp.n_reject_per_chain.sum()) - r0) / (3 * T * N)
            print(json.dumps({"case": name, "sampler": tag, "path": smp.path, "chains": N, "transitions": T,
                              "samples_per_s": N * T / dt, "ms": dt * 1e3, "rejections_per_transition": rej}), flush=True)
        if a.baselines:
            Nb = min(N, 1 << 16)
            for cls, tag, kw, Tb in ((gs.MetropolisHastings, "rwmh", dict(stepsize=0.1), T),
                                     (gs.SphericalHMC, "hmc", dict(stepsize=0.02 if d > 4 else 1e-3, n_steps=10), max(T // 10, 2))):
                smp = cls(pdf, x0[:Nb], 11, **kw)
                smp.reset(10)
                smp._run_device(10, None)
                torch.cuda.synchronize()
                a0 = int(smp.n_accept_per_chain.sum())
                dt = timed(lambda: smp._run_device(Tb, None), 2)
                acc = (int(smp.n_accept_per_chain.sum()) - a0) / (3 * Tb * Nb)
                print(json.dumps({"case": name, "sampler": tag, "path": smp.path, "chains": Nb, "transitions": Tb,
                                  "samples_per_s": Nb * Tb / dt, "ms": dt * 1e3, "acceptance": acc}), flush=True)
            xg = torch.as_tensor(gs.sample_sphere(d - 1, min(n_lp, 1 << 20), seed=3), dtype=torch.float32).to(dev)
            dt = timed(lambda: pdf.gradient(xg), 3)
            print(json.dumps({"case": name, "op": "gradient fp32", "points": len(xg), "points_per_s": len(xg) / dt, "ms": dt * 1e3}), flush=True)
        # standalone batched log_prob (fp32): algorithmic bytes = (d + 1) * 4 per point
        x = torch.nn.functional.normalize(torch.randn(n_lp, d, device=dev, generator=torch.Generator(dev).manual_seed(2)), dim=1)
        dt = timed(lambda: pdf.log_prob(x), 5)
        print(json.dumps({"case": name, "op": "log_prob fp32", "points": n_lp, "points_per_s": n_lp / dt,
                          "GBps_algorithmic": n_lp * (d + 1) * 4 / dt / 1e9, "ms": dt * 1e3}), flush=True)


if __name__ == "__main__":
    main()
