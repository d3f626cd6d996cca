"""Dense-Bingham GeoSSS-shrink with every state retained (what sample() returns): sampler in the eigenbasis + rotation of the samples.
    python benchmarks/bench_dense_retained.py"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geosss_b200 as gs

dev = torch.device("cuda", 0)
N, T = 1 << 18, 40
for tag, pdf in (("dense", gs.random_bingham(50, 300, 0, False, 6982)), ("diag", gs.random_bingham(50, 300, 0, True, 6982))):
    for env in ({}, {"GEOSSS_NO_EIG": "1"}):
        if tag == "diag" and env:
            continue
        os.environ.update(env)
        x0 = torch.as_tensor(gs.sample_sphere(49, N, seed=1)).to(dev)
        smp = gs.ShrinkageSphericalSliceSampler(pdf, x0, 3)
        buf = torch.empty((N, T, 50), dtype=torch.float32, device=dev)
        smp._run_device(50, None)
        smp._run_device(T, buf, chain_major=True)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            smp._run_device(T, buf, chain_major=True)
        e.record()
        torch.cuda.synchronize()
        dt = s.elapsed_time(e) * 1e-3 / 3
        for k in env:
            del os.environ[k]
        print(json.dumps({"case": f"Bingham d=50 {tag}, every state retained", "env": env, "chains": N, "transitions": T, "samples_per_s": N * T / dt, "ms": dt * 1e3}), flush=True)
