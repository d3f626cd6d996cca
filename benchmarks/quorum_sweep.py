"""Quorum sweep of the persistent bracket loop on the bench workload (C2, 65536 chains x 500 transitions):
python benchmarks/quorum_sweep.py [shrink|reject]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geosss_b200 as gs

kind = sys.argv[1] if len(sys.argv) > 1 else "shrink"
N, T, d = 65536, 500 if kind == "shrink" else 200, 10
modes = gs.sample_sphere(d - 1, 5, seed=1234)
pdf = gs.MixtureModel([gs.VonMisesFisher(100.0 * m) for m in modes])
cls = gs.ShrinkageSphericalSliceSampler if kind == "shrink" else gs.RejectionSphericalSliceSampler
x0 = torch.as_tensor(gs.sample_sphere(d - 1, N, seed=5)).cuda()
buf = torch.empty((N, T, d), dtype=torch.float32, device="cuda")
for q in ([14, 16, 18, 20, 22, 24, 26, 28, 32] if len(sys.argv) < 3 else [int(v) for v in sys.argv[2:]]):
    os.environ["GEOSSS_QUORUM"] = str(q)
    smp = cls(pdf, x0, 48385)
    smp._run_device(200, None)
    for _ in range(2):
        smp._run_device(T, buf, chain_major=True)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        smp._run_device(T, buf, chain_major=True)
    e.record()
    torch.cuda.synchronize()
    print(json.dumps({"sampler": kind, "quorum": q, "samples_per_s": 5 * N * T / (s.elapsed_time(e) * 1e-3)}), flush=True)
