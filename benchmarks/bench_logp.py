"""
Standalone batched log_prob bandwidth (float32), tensor-core kernels against the SIMT tiled kernels:

    python benchmarks/bench_logp.py [--points 4194304]

One JSON line per (target, variant): algorithmic GB/s = points * (d + 1) * 4 / launch time (CUDA events, inputs larger
than L2), fraction of the measured HBM copy bandwidth (MEASURED_PEAKS.json), max relative error against the float64
oracle on the first 512 points.  Variants are switched by the library's environment knobs, read at every launch.
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
from oracle import targets  # noqa: E402  (checker only)


def oracle_of(pdf):
    if isinstance(pdf, gs.MixtureModel):
        return targets.VMFMixture(np.stack([p.mu for p in pdf.pdfs]), pdf.weights)
    if isinstance(pdf, gs.BinghamFisher):
        return targets.Bingham(pdf.A, pdf.b)
    if isinstance(pdf, gs.Bingham):
        return targets.Bingham(pdf.A)
    return targets.CurvedVMF(pdf.curve.knots, pdf.kappa)


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e-3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1 << 22)
    ap.add_argument("--only", default="")
    ap.add_argument("--sweep", action="store_true", help="tensor-core kernel tuning knobs (landing stages, TMEM rings, stacked W) instead of the standard variants")
    ap.add_argument("--modes", default="", help="comma-separated landing-ring depths (GEOSSS_TC_STAGES) to compare")
    a = ap.parse_args()
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    dev = torch.device("cuda", 0)
    mix = lambda d, K, kappa: gs.MixtureModel([gs.VonMisesFisher(kappa * m) for m in gs.sample_sphere(d - 1, K, seed=1234)])
    cases = [
        ("bingham dense d=50", gs.random_bingham(50, 300, 0, False, 6982), [("tcgen05 pipelined", {}), ("tcgen05 round-1 kernel", {"GEOSSS_TC_V1": "1"}), ("fp32 SIMT", {"GEOSSS_NO_TC": "1"})]),
        ("bingham dense d=64", gs.random_bingham(64, 300, 0, False, 6982), [("tcgen05 pipelined", {}), ("tcgen05 round-1 kernel", {"GEOSSS_TC_V1": "1"})]),
        ("bingham dense d=32", gs.random_bingham(32, 300, 0, False, 6982), [("tcgen05 pipelined", {}), ("tcgen05 round-1 kernel", {"GEOSSS_TC_V1": "1"})]),
        ("vmf mixture d=100 K=8", mix(100, 8, 100.0), [("tcgen05 pipelined", {}), ("fp32 SIMT", {"GEOSSS_NO_TC": "1"})]),
        ("vmf mixture d=32 K=8", mix(32, 8, 100.0), [("tcgen05 pipelined", {}), ("fp32 SIMT", {"GEOSSS_NO_TC": "1"})]),
        ("vmf mixture d=24 K=8", mix(24, 8, 100.0), [("tcgen05 pipelined", {}), ("fp32 SIMT", {"GEOSSS_NO_TC": "1"})]),
        ("vmf mixture d=40 K=8", mix(40, 8, 100.0), [("tcgen05 pipelined", {}), ("fp32 SIMT", {"GEOSSS_NO_TC": "1"})]),
        ("vmf mixture d=48 K=8", mix(48, 8, 100.0), [("tcgen05 pipelined", {}), ("fp32 SIMT", {"GEOSSS_NO_TC": "1"})]),
        ("bingham dense d=40", gs.random_bingham(40, 300, 0, False, 6982), [("tcgen05 pipelined", {})]),
        ("bingham dense d=16", gs.random_bingham(16, 300, 0, False, 6982), [("tcgen05 pipelined", {})]),
        ("vmf mixture d=64 K=12", mix(64, 12, 100.0), [("tcgen05 pipelined", {}), ("fp32 SIMT", {"GEOSSS_NO_TC": "1"})]),
        ("curved vmf d=100 10 knots (C3)", gs.CurvedVonMisesFisher(gs.SlerpCurve(gs.brownian_curve(10, 100, 0.5, seed=4562)), 500.0),
         [("tcgen05 pipelined", {}), ("tcgen05, 2 landing stages", {"GEOSSS_TC_STAGES": "2"}), ("fp32 SIMT", {"GEOSSS_NO_TC": "1"})]),
        ("bingham diagonal d=50 (C4)", gs.random_bingham(50, 300, 0, True, 6982), [("fp32 SIMT", {})]),
    ]
    if a.sweep:  # tuning knobs of the tensor-core kernel
        knobs = [("default", {}), ("2 landing stages", {"GEOSSS_TC_STAGES": "2"}), ("3 landing stages", {"GEOSSS_TC_STAGES": "3"}),
                 ("polling waits", {"GEOSSS_TC_SPIN": "1"})]
        cases = [(n, p, knobs) for n, p, v in cases if "SIMT" not in v[0][0]]
    if a.modes:
        knobs = [(f"stages {m}", {"GEOSSS_TC_STAGES": m}) for m in a.modes.split(",")]
        cases = [(n, p, knobs) for n, p, v in cases if "SIMT" not in v[0][0]]
    for name, pdf, variants in cases:
        if a.only and a.only not in name:
            continue
        d = pdf.d
        n = a.points
        x = torch.nn.functional.normalize(torch.randn(n, d, device=dev, generator=torch.Generator(dev).manual_seed(2)), dim=1)
        want = oracle_of(pdf).log_prob(x[:512].double().cpu().numpy())
        want_tail = oracle_of(pdf).log_prob(x[-200:].double().cpu().numpy())
        for tag, env in variants:
            for k, v in env.items():
                os.environ[k] = v
            try:
                got = pdf.log_prob(x)
                err = float(np.max(np.abs(got[:512].double().cpu().numpy() - want) / np.maximum(1.0, np.abs(want))))
                err = max(err, float(np.max(np.abs(got[-200:].double().cpu().numpy() - want_tail) / np.maximum(1.0, np.abs(want_tail)))))
                dt = timed(lambda: pdf.log_prob(x), 10)
                gbs = n * (d + 1) * 4 / dt / 1e9
                print(json.dumps({"case": name, "variant": tag, "points": n, "ms": dt * 1e3, "GBps_algorithmic": gbs, "frac_of_measured_hbm": gbs / peak,
                                  "points_per_s": n / dt, "max_rel_err_vs_oracle": err}), flush=True)
            except Exception as exc:
                print(json.dumps({"case": name, "variant": tag, "error": f"{type(exc).__name__}: {exc}"}), flush=True)
            finally:
                for k in env:
                    del os.environ[k]
        del x


if __name__ == "__main__":
    main()
