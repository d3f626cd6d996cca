"""Multi-GPU parity on real devices: chains sharded over the ranks of one box, one process per GPU,
NCCL only for the final gather and the diagnostic reductions.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/check_multi_gpu.py

Checks (rank 0 prints one JSON line, exit code 1 on failure):
  * the chains of G ranks, gathered over NCCL, equal a single-GPU run of all chains BIT FOR BIT (the Philox counter
    uses the global chain id), for the register-resident vMF kernel, a great-circle kernel and the cutoff CPD kernel;
  * ESS and path statistics from all-reduced partial sums equal the single-GPU values.
"""
import json, os, sys
import numpy as np, torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geosss_b200 as gs
from geosss_b200 import diagnostics, distributed

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

g = np.load(os.path.join(ROOT, "tests", "golden", "logprob_cpd.npz"))
n_src = len(g["c5_source"])
cases = {
    "vmf d=10 K=5 (k_vmf_fast)": (gs.MixtureModel([gs.VonMisesFisher(100.0 * m) for m in gs.sample_sphere(9, 5, seed=1234)]), 4099, 60),
    "bingham d=50 (k_gc_fast)": (gs.random_bingham(50, 300, 0, True, 6982), 2051, 12),
    "bingham d=50 dense (eigenbasis chains, tcgen05 rotation of the samples)": (gs.random_bingham(50, 300, 0, False, 6982), 2051, 12),
    "curved vmf d=100 (k_gc_fast, normals parked in global scratch)": (gs.CurvedVonMisesFisher(gs.SlerpCurve(gs.brownian_curve(10, 100, 0.5, seed=4562)), 500.0), 1031, 10),
    "cpd k=20 (k_sampler_cpd_cell)": (gs.CoherentPointDrift(gs.PointCloud(g["c5_target"]), gs.PointCloud(g["c5_source"], np.ones(n_src) / n_src),
                                                            sigma=1.0, k=20, omega=0.4), 301, 4),
}
report, ok = {}, True
for name, (pdf, N, n) in cases.items():
    x0 = gs.sample_sphere(pdf.d - 1, N, seed=7)
    smp = distributed.make_sampler(gs.ShrinkageSphericalSliceSampler, pdf, torch.as_tensor(x0).to(dev), 2024)
    mine = smp.sample(n)                                   # (count_r, n, d) on this GPU
    every = distributed.gather_chains(mine, N)             # NCCL all_gather
    ess_all = diagnostics.ess(mine, None, n // 2)          # partial sums all-reduced over NCCL
    bulk_all = None if n < 8 else diagnostics.bulk_ess(mine[:, :, 0].contiguous())  # ranks over all chains of all processes (all_gather), partial sums all-reduced
    rej_all = distributed.reduce_sum(torch.tensor([float(smp.n_reject)], dtype=torch.float64, device=dev))
    if rank == 0:
        dist_state = dist.group.WORLD
        single = gs.ShrinkageSphericalSliceSampler(pdf, torch.as_tensor(x0).to(dev), 2024)
        want = single.sample(n)
        # single-GPU diagnostics without the process group: finish the partial sums locally
        part = diagnostics.ess_partials(want, None, n // 2)
        ess_one = diagnostics.ess_from_partials(part.cpu().numpy(), n, N)
        same = bool(torch.equal(every, want))
        row = {"chains": N, "draws": n, "bit_equal_to_single_gpu": same, "ess_all_ranks": ess_all[0], "ess_single_gpu": ess_one[0],
               "bulk_ess_all_ranks": bulk_all, "bulk_ess_single_gpu": None if n < 8 else diagnostics.bulk_ess(want[:, :, 0].contiguous(), local_only=True),
               "rejections_all_ranks": float(rej_all.item()), "rejections_single_gpu": float(single.n_reject)}
        ok = ok and same and abs(ess_all[0] - ess_one[0]) <= 1e-9 * abs(ess_one[0]) and row["rejections_all_ranks"] == row["rejections_single_gpu"]
        ok = ok and (n < 8 or abs(row["bulk_ess_all_ranks"] - row["bulk_ess_single_gpu"]) <= 1e-6 * row["bulk_ess_single_gpu"])
        report[name] = row
if rank == 0:
    print(json.dumps({"world_size": world, "ok": ok, "cases": report}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
