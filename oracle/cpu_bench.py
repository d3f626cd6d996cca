"""
CPU baseline worker: the float64 oracle port of the reference samplers, one independent chain
per process (oracle; used only by bench.py's cpu_baseline / --impl reference legs).

Follows BASELINE.md section 3: chains seeded from SeedSequence(48385).spawn, initial states
from sample_sphere, BLAS/OpenMP threads pinned to 1, wall-clock with perf_counter.
"""

import os
import time

import numpy as np


def build_target(workload):
    from . import sphere, targets
    name = workload["name"]
    if name == "vmf_mixture":
        modes = sphere.sample_sphere(workload["d"] - 1, workload["K"], seed=1234)
        return targets.VMFMixture(workload["kappa"] * modes)
    if name == "bingham":
        return targets.Bingham(targets.random_bingham_matrix(workload["d"], workload["vmax"], 0, True, 6982))
    if name == "curved_vmf":
        return targets.CurvedVMF(targets.brownian_knots(10, workload["d"], 0.5, seed=4562), workload["kappa"])
    raise ValueError(name)


def run_chain(args):
    """One chain of ``n_transitions`` steps. Returns (transitions, rejections, seconds)."""
    workload, kind, proc_index, n_transitions = args
    from . import samplers, sphere
    pdf = build_target(workload)
    child = np.random.SeedSequence(48385).spawn(proc_index + 1)[proc_index]
    init_seed = np.random.SeedSequence(48485).spawn(proc_index + 1)[proc_index]
    x0 = sphere.sample_sphere(pdf.d - 1, seed=init_seed)
    smp = samplers.SliceSampler(kind, pdf.log_prob, x0, np.random.default_rng(child))
    t0 = time.perf_counter()
    out = smp.sample(n_transitions + 1)
    return n_transitions, smp.n_reject, time.perf_counter() - t0, out[1:, 0].copy()


def run_all_cores(workload, kind, n_transitions, n_procs=None):
    """Independent chains on every host core. Returns dict(value=samples/s, cores, wall, ...)."""
    import multiprocessing as mp
    n_procs = n_procs or os.cpu_count() or 1
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[var] = "1"
    ctx = mp.get_context("spawn")
    with ctx.Pool(n_procs) as pool:
        pool.map(_warm, range(n_procs))  # start-up and imports stay outside the timed region
        t0 = time.perf_counter()
        res = pool.map(run_chain, [(workload, kind, i, n_transitions) for i in range(n_procs)])
        wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    out = {"value": total / wall, "cores": n_procs, "wall_s": wall, "transitions": total,
           "mean_rejections": sum(r[1] for r in res) / total}
    try:  # ESS of the first coordinate over the CPU chains, same estimator as the GPU arm
        from . import ess
        trace = np.stack([r[3] for r in res])
        rel = ess.ess_multichain(trace, min(trace.shape[1] // 2, 400)) / trace.size
        out["relative_ess_x0"] = float(rel)
        out["ess_per_sec"] = float(rel * out["value"])
    except Exception as exc:
        out["ess_error"] = str(exc)
    return out


def _warm(_):
    from . import samplers, sphere, targets  # noqa: F401
    return 0
