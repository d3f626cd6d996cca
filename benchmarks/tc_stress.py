"""Stress of k_logp_tc: every row width the tensor-core path accepts, ragged batch sizes on both sides of the 256-row-tile
threshold, repeated launches; each result checked against the SIMT kernels (GEOSSS_NO_TC).  A hang shows as the timeout of the
caller; run under `timeout`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geosss_b200 as gs  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    gen = np.random.default_rng(0)
    worst = 0.0
    launches = 0
    for d in list(range(8, 113, 4)) + [25, 33, 50, 63, 99]:
        pdfs = []
        if d >= 24:
            pdfs.append(gs.MixtureModel([gs.VonMisesFisher(40.0 * m) for m in gs.sample_sphere(d - 1, int(gen.integers(2, 17)), seed=d)]))
            pdfs.append(gs.CurvedVonMisesFisher(gs.SlerpCurve(gs.brownian_curve(int(gen.integers(3, 17)), d, 0.5, seed=d)), 100.0))
        if d <= 64:
            pdfs.append(gs.random_bingham(d, 50.0, 0, False, d))
        for pdf in pdfs:
            for rep in range(6):
                n = int(gen.integers(1, 400_000)) if rep % 2 else int(gen.integers(1, 3000))
                x = torch.nn.functional.normalize(torch.randn(n, d, device=dev), dim=1)
                a = pdf.log_prob(x)
                b = pdf.log_prob(x)
                assert torch.equal(a, b), (d, type(pdf).__name__, n, "not reproducible")
                os.environ["GEOSSS_NO_TC"] = "1"
                ref = pdf.log_prob(x)
                del os.environ["GEOSSS_NO_TC"]
                err = float(((a - ref).abs() / ref.abs().clamp(min=1.0)).max())
                worst = max(worst, err)
                assert err <= 2e-5, (d, type(pdf).__name__, n, err)
                launches += 2
    torch.cuda.synchronize()
    print(f"ok: {launches} tensor-core launches, worst difference to the SIMT kernels {worst:.3g}")


if __name__ == "__main__":
    main()
