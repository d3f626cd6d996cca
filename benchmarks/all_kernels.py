"""
Launch every kernel of libgeosss_cuda once at a moderate size (for the ncu kernel table in profiles/).

    python benchmarks/all_kernels.py
    ncu --metrics <list in profiles/kernel_table.py> --csv --log-file gpurun_out/kernels.csv python benchmarks/all_kernels.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geosss_b200 as gs  # noqa: E402
from geosss_b200 import diagnostics as dg  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    N = 1 << 16
    modes = gs.sample_sphere(9, 5, seed=1234)
    c2 = gs.MixtureModel([gs.VonMisesFisher(100.0 * m) for m in modes])
    hi = gs.MixtureModel([gs.VonMisesFisher(60.0 * m) for m in gs.sample_sphere(31, 8, seed=5)])
    curved = gs.CurvedVonMisesFisher(gs.SlerpCurve(gs.brownian_curve(10, 100, 0.5, seed=4562)), 500.0)
    diag, dense = gs.random_bingham(50, 300, 0, True, 6982), gs.random_bingham(50, 300, 0, False, 6982)
    g = np.load(os.path.join(ROOT, "tests", "golden", "logprob_cpd.npz"))
    n = len(g["c5_source"])
    clouds = gs.PointCloud(g["c5_target"]), gs.PointCloud(g["c5_source"], np.ones(n) / n)
    cpd_dense = gs.CoherentPointDrift(*clouds, sigma=1.0, k=20, omega=0.4, dense=True)
    cpd_knn = gs.CoherentPointDrift(*clouds, sigma=1.0, k=20, omega=0.4)

    def x0(pdf, count):
        return torch.as_tensor(gs.sample_sphere(pdf.d - 1, count, seed=1)).to(dev)

    # throughput samplers (fp32 fast path)
    for pdf, T, count in ((c2, 100, N), (hi, 10, N), (curved, 5, N), (diag, 10, N), (dense, 10, N), (cpd_dense, 2, 2048)):
        for cls in (gs.ShrinkageSphericalSliceSampler, gs.RejectionSphericalSliceSampler):
            if cls is gs.RejectionSphericalSliceSampler and pdf is not c2:
                continue
            cls(pdf, x0(pdf, count), 11)._run_device(T, None)
    # literal samplers (fp64 lane / warp per chain) and baselines (fp32)
    gs.ShrinkageSphericalSliceSampler(c2, x0(c2, 1 << 14), 12, dtype="float64")._run_device(5, None)
    gs.ShrinkageSphericalSliceSampler(cpd_knn, x0(cpd_knn, 1024), 12, dtype="float32", path="literal")._run_device(1, None)
    gs.MetropolisHastings(c2, x0(c2, N), 13)._run_device(50, None)
    gs.SphericalHMC(c2, x0(c2, N), 14, stepsize=0.02)._run_device(5, None)
    gs.SphericalHMC(cpd_knn, x0(cpd_knn, 1024), 15)._run_device(1, None)
    gs.MetropolisHastings(cpd_knn, x0(cpd_knn, 16384), 16, stepsize=0.05)._run_device(10, None)  # cutoff lists: k_mh_cpd_cell
    # log_prob / gradient
    for pdf, count in ((c2, 1 << 22), (hi, 1 << 20), (curved, 1 << 19), (diag, 1 << 20), (dense, 1 << 20), (cpd_dense, 1 << 13)):
        x = x0(pdf, count).float()
        pdf.log_prob(x)
    for pdf, count in ((c2, 1 << 18), (dense, 1 << 16), (cpd_knn, 1 << 11)):
        x = x0(pdf, count)
        pdf.log_prob(x)   # float64: literal kernels
        pdf.gradient(x.float())
    # tensor-core log_prob at the shapes it is tuned on, retained dense-Bingham samples (rotation kernel)
    big = gs.MixtureModel([gs.VonMisesFisher(100.0 * m) for m in gs.sample_sphere(99, 8, seed=5)])
    for pdf, count in ((big, 1 << 20), (curved, 1 << 20), (dense, 1 << 21)):
        pdf.log_prob(x0(pdf, count).float())
    gs.ShrinkageSphericalSliceSampler(dense, x0(dense, 1 << 14), 17).sample(8)
    # diagnostics and exact samplers
    chains = gs.ShrinkageSphericalSliceSampler(c2, x0(c2, 4096), 16).sample(500)
    dg.ess(chains)
    dg.bulk_ess(chains[:, :, 0])
    dg.streaming_ess(gs.ShrinkageSphericalSliceSampler(c2, x0(c2, 4096), 18), 400, chunk=100, max_lag=50)
    dg.mode_kl(chains, c2)
    gs.sample_vMF(c2.pdfs[0], 1 << 18, seed=1, as_torch=True)
    gs.sample_bingham(dense.A, 1 << 18, seed=2, as_torch=True)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
