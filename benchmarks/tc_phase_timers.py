"""Phase timers of k_logp_tc (experiment build only: capi.cu compiled with -DTCP_TIMERS into libgeosss_cuda_timers.so.bin).

Block 0's role warps accumulate clock64 time per phase; this script runs one launch per case and prints cycles per tile.
Usage on the GPU box: cp geosss_b200/libgeosss_cuda_timers.so.bin geosss_b200/libgeosss_cuda.so; python benchmarks/tc_phase_timers.py
"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geosss_b200 as gs  # noqa: E402
from geosss_b200 import _lib  # noqa: E402


def main():
    _lib.lib()
    raw = C.CDLL(_lib.LIB_PATH)
    raw.geosss_debug_tc_timers.argtypes = [C.POINTER(C.c_ulonglong)]
    dev = torch.device("cuda", 0)
    mix = lambda d, K, kappa: gs.MixtureModel([gs.VonMisesFisher(kappa * m) for m in gs.sample_sphere(d - 1, K, seed=1234)])
    cases = [
        ("bingham dense d=50", gs.random_bingham(50, 300, 0, False, 6982)),
        ("bingham dense d=64", gs.random_bingham(64, 300, 0, False, 6982)),
        ("bingham dense d=32", gs.random_bingham(32, 300, 0, False, 6982)),
        ("vmf mixture d=32 K=8", mix(32, 8, 100.0)),
        ("vmf mixture d=64 K=12", mix(64, 12, 100.0)),
        ("vmf mixture d=100 K=8", mix(100, 8, 100.0)),
    ]
    n = 1 << 22
    buf = (C.c_ulonglong * 32)()
    for name, pdf in cases:
        x = torch.nn.functional.normalize(torch.randn(n, pdf.d, device=dev, generator=torch.Generator(dev).manual_seed(2)), dim=1)
        pdf.log_prob(x)
        raw.geosss_debug_tc_timers(buf)  # clear
        pdf.log_prob(x)
        raw.geosss_debug_tc_timers(buf)
        t = [int(v) for v in buf]
        tiles = max(t[24], 1) * 2  # warp 10 counts its own loop bound n_it once; both groups together see n_it tiles
        tiles = t[24]
        per = lambda v, k=1.0: round(v / tiles * k, 1)
        print(json.dumps({
            "case": name, "tiles_of_block_0": tiles,
            "producer": {"wait_landing_stage_free": per(t[0]), "issue": per(t[1])},
            "mma_warp_per_own_tile": {"wait_operands": per(t[4], 2), "wait_accumulator_free": per(t[5], 2), "issue": per(t[6], 2)},
            "split_warp_per_own_tile": {"wait_rows_landed": per(t[8], 2), "wait_stage_mma_done": per(t[9], 2), "load_split_store": per(t[10], 2),
                                        "wait_st": per(t[11], 2), "hand_off": per(t[12], 2)},
            "epilogue_warp_per_own_tile": {"wait_accumulator": per(t[16], 2), "tmem_load": per(t[17], 2), "math": per(t[18], 2), "store": per(t[19], 2)},
        }), flush=True)
        del x


if __name__ == "__main__":
    main()
