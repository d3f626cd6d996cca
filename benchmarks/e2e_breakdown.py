"""Where the end-to-end (host buffer) call spends its time. python benchmarks/e2e_breakdown.py"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geosss_b200 as gs
from geosss_b200 import _lib
import ctypes as C

N, T, d = 65536, 500, 10
modes = gs.sample_sphere(9, 5, seed=1234)
pdf = gs.MixtureModel([gs.VonMisesFisher(100.0 * m) for m in modes])
x0 = np.ascontiguousarray(gs.sample_sphere(9, N, seed=1))
out = torch.empty((T + 1, N, d), dtype=torch.float32, pin_memory=True)
dev = torch.device("cuda", 0)
lib = _lib.lib()

def tick():
    torch.cuda.synchronize()
    return time.perf_counter()

def api_call(seed):
    s = gs.ShrinkageSphericalSliceSampler(pdf, x0, seed)
    s.sample(T + 1, 0, out=out)

api_call(1)
for rep in range(4):
    t0 = tick()
    api_call(100 + rep)
    t1 = tick()
    print(f"whole API call (construct, sample, drop): {1e3*(t1-t0):.2f} ms = {N*T/(t1-t0):.3e} samples/s")

for rep in range(3):
    t0 = tick()
    s = gs.ShrinkageSphericalSliceSampler(pdf, x0, 5 + rep)
    t1 = tick()
    h = s._sampler()
    t2 = tick()
    _lib.check(lib.geosss_sampler_run_host(h, 0, T, out[1:].data_ptr()))
    t3 = tick()
    s._sync_counters()
    t4 = tick()
    del s
    t4b = tick()
    # raw D2H of the same bytes
    devbuf = torch.empty((T, N, d), dtype=torch.float32, device=dev)
    t5 = tick()
    out[1:].copy_(devbuf)
    t6 = tick()
    print(f"construct {1e3*(t1-t0):.2f} ms | device sampler + H2D {1e3*(t2-t1):.2f} | run_host {1e3*(t3-t2):.2f} | counters {1e3*(t4-t3):.2f} | drop {1e3*(t4b-t4):.2f} | "
          f"raw pinned D2H of the same {out[1:].numel()*4/1e9:.2f} GB: {1e3*(t6-t5):.2f} ms = {out[1:].numel()*4/(t6-t5)/1e9:.1f} GB/s")
