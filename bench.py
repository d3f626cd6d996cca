"""
bench.py -- GeoSSS samples/sec on the BASELINE.json workload.

    python bench.py [--gpus N --steps K --warmup W]            # our CUDA path
    python bench.py --impl reference [...]                      # CPU arm: oracle port on all host cores

Workload (BASELINE.json configs[1], scripts/mixture_vMF.py): mixture of K=5 von Mises-Fisher
components on S^9 (d=10), modes = sample_sphere(9, 5, seed=1234), kappa=100, GeoSSS-shrink,
65,536 chains PER GPU (weak scaling: each rank owns a disjoint block of global chain ids).
One "step" = --transitions Markov transitions of every chain, every state retained in HBM
(chain-major (N, T, d) float32), exactly what ``sampler.sample()`` returns for CUDA input.  One step is one
launch of the sampler kernel; its output (T*N*40 B >> L2) is larger than L2.

value  = transitions of all chains of all ranks / device time (CUDA events, max over ranks)
e2e    = same metric through the public Python API with HOST buffers: initial states numpy ->
         device, all samples device -> pinned host, inside the timed region
roofline: retained-sample bytes written per second by the sampler kernel against the measured
         HBM copy bandwidth (the kernel is issue/MUFU-bound, not HBM-bound)
"""

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "GeoSSS-shrink samples/sec"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=65536, help="chains per GPU")
    ap.add_argument("--transitions", type=int, default=500, help="transitions per step")
    ap.add_argument("--kappa", type=float, default=100.0)
    ap.add_argument("--sampler", default="shrink", choices=["shrink", "reject"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-transitions", type=int, default=2500, help="transitions per host core in the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--numa-bind", default="off", choices=["on", "off"],
                    help="pin the process to the GPU's NUMA-local cores (measured on this pool: 3.6x slower end to end at "
                         "1 GPU, +3 %% at 8 GPUs -- gpurun_out/r41_bench_*.json -- so it is off by default)")
    return ap.parse_args()


def workload_dict(a):
    return {"name": "vmf_mixture", "d": 10, "K": 5, "kappa": a.kappa}


def config_dict(a, n_gpus):
    return {"workload": f"C2 vMF mixture d=10 K=5 kappa={a.kappa:g} (scripts/mixture_vMF.py), sss-{a.sampler}",
            "chains_per_gpu": a.chains, "chains_total": a.chains * n_gpus, "transitions_per_step": a.transitions,
            "retention": "every state, (N,T,d) float32 in HBM", "l2": "outputs_larger_than_L2",
            "sharding": f"chains x{n_gpus}"}


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML every few ms while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.gpu, self.samples, self.stop_flag, self.thread, self.max_mhz = gpu_index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:
            self.error = str(exc)
            return

        def loop():
            while not self.stop_flag:
                try:
                    self.samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                                         pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
                except Exception:
                    pass
                time.sleep(0.002)

        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "error", "?")]}
        self.thread.join(timeout=1.0)
        inside = [(c, r) for t, c, r in self.samples if t0 <= t <= t1]
        reasons = set()
        for _, r in inside:
            for bit, name in self.REASONS.items():
                if r & bit:
                    reasons.add(name)
        return {"sm_mhz": statistics.median([c for c, _ in inside]) if inside else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(inside)}


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ CPU arm
def cpu_arm(a, n_transitions):
    from oracle import cpu_bench
    kind = 1 if a.sampler == "shrink" else 0
    return cpu_bench.run_all_cores(workload_dict(a), kind, n_transitions)


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_core = max(200, a.cpu_transitions // 5)
    for _ in range(a.warmup):
        cpu_arm(a, 50)
    total, cores, rej, wall, last = 0, 1, 0.0, 0.0, {}
    for _ in range(a.steps):
        r = last = cpu_arm(a, per_core)
        total += r["transitions"]
        wall += r["wall_s"]  # process start-up and imports are outside the timed region
        cores = r["cores"]
        rej = r["mean_rejections"]
    value = total / wall
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * wall / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(a, a.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"oracle port of the reference samplers, {cores} independent chains x {per_core} transitions per step, "
                                       f"mean rejections/transition {rej:.2f}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ess": {"relative_ess_x0": last.get("relative_ess_x0"), "ess_per_sec": last.get("ess_per_sec"),
                    "chains_used": cores, "draws": per_core, "estimator": "oracle/ess.py ess_multichain on the last step's chains"},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------ our arm
def main_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import geosss_b200 as gs
    from geosss_b200 import _lib  # noqa: F401  (fails loudly if the CUDA library is missing)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from geosss_b200 import distributed as gdist
    numa_cpus = None
    if a.numa_bind == "on":
        numa_cpus = gdist.bind_host_to_gpu(local)  # pinned host buffers next to this GPU's PCIe root
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    N, T, d = a.chains, a.transitions, 10
    modes = gs.sample_sphere(d - 1, 5, seed=1234)
    pdf = gs.MixtureModel([gs.VonMisesFisher(a.kappa * m) for m in modes])
    cls = gs.ShrinkageSphericalSliceSampler if a.sampler == "shrink" else gs.RejectionSphericalSliceSampler
    x0 = gs.sample_sphere(d - 1, N, seed=np.random.SeedSequence(48485).spawn(world)[rank])
    smp = cls(pdf, torch.as_tensor(x0).to(dev), 48385, chain_offset=rank * N)
    buf = torch.empty((N, T, d), dtype=torch.float32, device=dev)  # chain-major: what sample() returns for CUDA input

    # burn-in + warm-up
    smp._run_device(200, None)
    for _ in range(max(a.warmup, 3)):
        smp._run_device(T, buf, chain_major=True)
    rej0 = int(smp.n_reject_per_chain.sum())

    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    t_host0 = time.perf_counter()
    for s, e in ev:
        s.record()
        smp._run_device(T, buf, chain_major=True)  # one kernel launch: T transitions of N chains, all states retained
        e.record()
    barrier()
    t_host1 = time.perf_counter()
    total_ms = ev[0][0].elapsed_time(ev[-1][1])
    launch_ms = [s.elapsed_time(e) for s, e in ev]
    clk = clocks.stop(t_host0, t_host1)
    total_ms = max_over_ranks(total_ms)
    rej = (int(smp.n_reject_per_chain.sum()) - rej0) / (a.steps * T * N)

    samples_per_step = N * T * world
    value = samples_per_step * a.steps / (total_ms * 1e-3)

    # roofline of the sampler kernel: retained bytes written / launch duration vs measured HBM copy bandwidth
    peak, peak_src = measured_peak_hbm()
    avg_launch_s = statistics.mean(launch_ms) * 1e-3
    algo_bytes = N * T * d * 4
    achieved = algo_bytes / avg_launch_s / 1e9

    # DRAM traffic of this launch from the committed ncu capture (profiles/r1_traffic.json), when the workload matches
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            tr = json.load(f)["k_vmf_fast<5,10,SHRINK>"]
        if a.sampler == "shrink" and tr["chains"] == N and tr["transitions_per_step"] == T and a.kappa == 100.0:
            traffic = tr["chain_major_quorum_22_default"]["dram_read"] + tr["chain_major_quorum_22_default"]["dram_write"]
    except Exception:
        traffic = None

    # compute-side view of the same launch: ALGORITHMIC flops (9d + E (3d + 4 + F_logp),
    # E = 1 + measured rejections, F_logp = 2Kd + 7K + 2) against the nominal FP32 FMA peak at the measured clock
    K_mix = 5
    flops_per_transition = 9 * d + (1.0 + rej) * (3 * d + 4 + 2 * K_mix * d + 7 * K_mix + 2)
    try:
        with open(os.path.join(ROOT, "profiles", "r1_micro_peaks.json")) as f:
            fp32_peak, fp32_src = float(json.load(f)["fp32_tflops_sustained"]), "measured FFMA microbenchmark (profiles/r1_micro_peaks.json)"
    except Exception:
        fp32_peak, fp32_src = 148 * 128 * 2 * (clk.get("sm_mhz") or 1965.0) * 1e6 / 1e12, "nominal 148 SM x 128 lanes x 2 x SM clock"
    fp32_ach = flops_per_transition * N * T / avg_launch_s / 1e12
    roof_compute = {"bound": "fp32 FMA issue", "peak_source": fp32_src, "achieved": fp32_ach, "peak": fp32_peak,
                    "unit": "TFLOP/s", "frac": fp32_ach / fp32_peak, "algorithmic_flops_per_transition": flops_per_transition,
                    "note": "the reference algorithm's flop count; the kernel itself also spends ~45% of its instructions on "
                            "Philox, Box-Muller and MUFU work that this count ignores (profiles/*k_vmf_fast*)"}

    # ESS of the first coordinate (scripts/ess_mixture_vMF.py:233) over the last retained step, all chains of
    # all ranks: autocovariances on the GPU (csrc/diag.cuh), max_lag+2 doubles all-reduced, Geyer cut on the host
    from geosss_b200 import diagnostics
    ess = None
    try:
        n_ess, rel = diagnostics.ess(buf, None, min(T // 2, 400))
        ess = {"relative_ess_x0": rel, "ess_per_sec": rel * value, "chains_used": N * world, "draws": T,
               "estimator": "multi-chain autocovariance ESS (Vehtari et al. 2021), device partial sums"}
    except Exception as exc:  # diagnostics must not break the bench line
        ess = {"error": str(exc)}

    # ---- end to end through the public API with host buffers ---------------------------------
    x0_host = np.ascontiguousarray(x0)
    out_host = torch.empty((T + 1, N, d), dtype=torch.float32, pin_memory=True)
    def e2e_step(seed):
        s = cls(pdf, x0_host, seed, chain_offset=rank * N)
        s.sample(T + 1, 0, out=out_host)  # H2D of the initial states, T transitions, D2H of every state
        return s
    e2e_step(1)
    # raw pinned device->host bandwidth into the SAME pinned buffer, all ranks copying at once (the ranks of one host
    # share its PCIe switches and memory controllers): the ceiling of the e2e number, measured in this run
    probe = torch.empty(out_host.numel() * 4, dtype=torch.uint8, device=dev)
    host_probe = out_host.view(-1).view(torch.uint8)
    host_probe.copy_(probe)
    barrier()
    tp = time.perf_counter()
    for _ in range(2):
        host_probe.copy_(probe, non_blocking=True)
    torch.cuda.synchronize()
    d2h_gbs = 2 * probe.numel() / max_over_ranks(time.perf_counter() - tp) / 1e9
    del probe
    barrier()
    t0 = time.perf_counter()
    for i in range(a.e2e_steps):
        e2e_step(2 + i)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = N * T * world * a.e2e_steps / e2e_s
    h2d = N * d * 4
    d2h = T * N * d * 4 + N * 8

    # ---- the bandwidth-bound operator of the path: standalone batched log_prob (fp32), input > L2 ----
    lp_roof = None
    if rank == 0:
        try:
            n_pts = 1 << 24  # 2^24 x 40 B = 671 MB read per launch (L2 is 126 MB)
            xs_lp = torch.nn.functional.normalize(torch.randn(n_pts, d, device=dev), dim=1)
            out_lp = torch.empty(n_pts, device=dev)
            h = pdf._handle(dev)
            run_lp = lambda: _lib.check(_lib.lib().geosss_log_prob(h, _lib.F32, _lib.PATH_FAST, xs_lp.data_ptr(), out_lp.data_ptr(),
                                                                   n_pts, torch.cuda.current_stream(dev).cuda_stream))
            for _ in range(3):
                run_lp()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run_lp()
            e1.record()
            torch.cuda.synchronize()
            sec = e0.elapsed_time(e1) * 1e-3 / 10
            gbs = n_pts * (d + 1) * 4 / sec / 1e9
            lp_roof = {"kernel": "k_logp_tiled<VmfLpConst<K=5,d=10>> (TMA bulk-staged tiles, 8-byte row loads, parameters in the constant bank)", "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
                       "frac": gbs / peak, "points_per_s": n_pts / sec, "algorithmic_bytes_per_launch": n_pts * (d + 1) * 4,
                       "avg_launch_ms": sec * 1e3}
            del xs_lp, out_lp
        except Exception as exc:
            lp_roof = {"error": str(exc)}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_arm(a, a.cpu_transitions)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"oracle port of the reference sampler (float64 numpy/scipy), {r['cores']} independent chains x "
                         f"{a.cpu_transitions} transitions, mean rejections/transition {r['mean_rejections']:.2f}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_dict(a, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": a.e2e_steps, "pinned_d2h_gbs_this_box": d2h_gbs, "pinned_d2h_probe": "per rank, all ranks at once, same buffer and size as the e2e step",
                        "host_cpus_bound_to_gpu_numa": None if numa_cpus is None else len(numa_cpus), "api": "ShrinkageSphericalSliceSampler(pdf, x0_numpy, seed).sample(T+1, out=pinned)"},
                "gpu_launches": a.steps,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src, "kernel": "k_vmf_fast<K=5,d=10,%s>" % a.sampler.upper(),
                             "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": avg_launch_s * 1e3,
                             "note": "kernel is instruction-issue/MUFU bound; HBM fraction = retained-sample write traffic"},
                "roofline_compute": roof_compute, "roofline_log_prob": lp_roof, "cpu_baseline": cpu, "clocks": clk, "mean_rejections_per_transition": rej, "ess": ess,
                "samples_per_sec_per_gpu": value / world}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)
