"""One large batched log_prob of the C5 registration target (for ncu): python benchmarks/cpd_logp_probe.py [n_points]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geosss_b200 as gs
from bench_targets import c5_target, timed

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
pdf = c5_target()
print(pdf.cutoff_grid_info())
x = torch.as_tensor(gs.sample_sphere(3, n, seed=2), dtype=torch.float32).cuda()
dt = timed(lambda: pdf.log_prob(x), 5)
print({"points": n, "points_per_s": n / dt, "ms": dt * 1e3})
