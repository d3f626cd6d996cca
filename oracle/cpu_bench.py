"""
CPU baseline worker, one independent chain per process (used only by bench.py's cpu_baseline /
--impl reference legs).  When ``oracle/_ref`` holds the upstream package (``oracle/build_ref.py``, run by
``__graft_entry__.build()`` in the build container) the chains are run by the REFERENCE ITSELF, unmodified:
``geosss.ShrinkageSphericalSliceSampler(pdf, x0, seed).sample(n)`` (geosss/mcmc.py:335-401) on targets built by
the reference's own constructors -- ``kind`` "reference".  Otherwise the float64 oracle port of the same
samplers runs -- ``kind`` "port".

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


def reference_dir():
    here = os.path.dirname(os.path.abspath(__file__))
    ref = os.path.join(here, "_ref")
    return ref if os.path.isdir(os.path.join(ref, "geosss")) and not os.environ.get("GEOSSS_CPU_PORT") else None


def kind_name():
    return "reference" if reference_dir() else "port"


def build_reference_target(gs, workload):
    """The same targets through the upstream constructors (scripts/mixture_vMF.py:406-411, scripts/bingham.py:123-131,
    scripts/curve_vMF.py:577-582, scripts/protein_reg3d3d.py:561-595)."""
    name = workload["name"]
    if name == "vmf_mixture":
        modes = gs.sample_sphere(workload["d"] - 1, workload["K"], seed=1234)
        return gs.MixtureModel([gs.VonMisesFisher(workload["kappa"] * m) for m in modes])
    if name == "bingham":
        return gs.random_bingham(workload["d"], workload["vmax"], 0, workload.get("eigensystem", True), 6982)
    if name == "curved_vmf":
        from geosss.spherical_curve import SlerpCurve, brownian_curve
        from geosss.distributions import CurvedVonMisesFisher
        return CurvedVonMisesFisher(SlerpCurve(brownian_curve(10, workload["d"], 0.5, seed=4562)), workload["kappa"])
    if name == "cpd":
        from geosss.pointcloud import PointCloud
        from geosss.registration import CoherentPointDrift
        data = np.load(os.path.join(reference_dir(), "data", "protein_registration.npz"))
        target = PointCloud(data["target"], weights=None)
        source = PointCloud(data["source"], weights=np.ones(len(data["source"])) / len(data["source"]))
        return CoherentPointDrift(target, source, float(data["sigma"]), 20, omega=float(data["prob_outlier"]))
    raise ValueError(name)


def run_chain_reference(args):
    """One chain run by the upstream package.  Returns (transitions, rejections, seconds, trace of x0)."""
    workload, kind, proc_index, n_transitions = args
    import sys
    ref = reference_dir()
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import geosss as gs
    pdf = build_reference_target(gs, workload)
    child = np.random.SeedSequence(48385).spawn(proc_index + 1)[proc_index]
    init_seed = np.random.SeedSequence(48485).spawn(proc_index + 1)[proc_index]
    d = workload["d"] if workload["name"] != "cpd" else 4
    x0 = gs.sample_sphere(d - 1, seed=init_seed)
    cls = gs.ShrinkageSphericalSliceSampler if kind == 1 else gs.RejectionSphericalSliceSampler
    smp = cls(pdf, x0, seed=child)
    t0 = time.perf_counter()
    out = smp.sample(n_transitions + 1)
    return n_transitions, smp.n_reject, time.perf_counter() - t0, np.asarray(out)[1:, 0].copy()


def run_chain(args):
    """One chain of ``n_transitions`` steps. Returns (transitions, rejections, seconds)."""
    if reference_dir():
        return run_chain_reference(args)
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
           "mean_rejections": sum(r[1] for r in res) / total, "kind": kind_name()}
    try:  # bulk ESS of the first coordinate over the CPU chains, same estimator as the GPU arm
        from . import ess
        trace = np.stack([r[3] for r in res])
        rel = ess.bulk_ess(trace, relative=True)
        out["relative_ess_x0"] = float(rel)
        out["ess_per_sec"] = float(rel * out["value"])
    except Exception as exc:
        out["ess_error"] = str(exc)
    return out


def _warm(_):
    if reference_dir():
        import sys
        if reference_dir() not in sys.path:
            sys.path.insert(0, reference_dir())
        import geosss  # noqa: F401
        return 0
    from . import samplers, sphere, targets  # noqa: F401
    return 0
