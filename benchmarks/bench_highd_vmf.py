"""GeoSSS-shrink on vMF mixtures beyond the register-resident kernel (d > 16 or K > 8): python benchmarks/bench_highd_vmf.py"""
import sys, os, json, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geosss_b200 as gs
from bench_targets import timed
for d, K, kappa in ((32, 8, 60.0), (100, 5, 100.0), (64, 12, 80.0)):
    pdf = gs.MixtureModel([gs.VonMisesFisher(kappa * m) for m in gs.sample_sphere(d - 1, K, seed=5)])
    N, T = 1 << 18, 20
    smp = gs.ShrinkageSphericalSliceSampler(pdf, torch.as_tensor(gs.sample_sphere(d - 1, N, seed=1)).cuda(), 11)
    smp._run_device(10, None)
    dt = timed(lambda: smp._run_device(T, None), 2)
    print(json.dumps({"case": f"vmf-mix d={d} K={K}", "path": smp.path, "samples_per_s": N * T / dt}), flush=True)
