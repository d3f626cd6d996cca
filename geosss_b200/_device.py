"""Device plumbing: torch owns device memory and streams, the kernels are libgeosss_cuda's."""

import numpy as np
import torch

from . import _lib


def resolve_device(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("geosss_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda" if device is None else device)
    if dev.type != "cuda":
        raise ValueError("geosss_b200 runs on CUDA devices only")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def dtype_code(dtype):
    """-> (code, torch dtype, numpy dtype)"""
    name = str(dtype).replace("torch.", "").replace("numpy.", "")
    if dtype in (np.float64, torch.float64, float) or name in ("float64", "f64", "double"):
        return _lib.F64, torch.float64, np.float64
    if dtype in (np.float32, torch.float32) or name in ("float32", "f32", "float"):
        return _lib.F32, torch.float32, np.float32
    raise ValueError(f"dtype must be float32 or float64, got {dtype!r}")


def stream_ptr(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def to_device(x, tdtype, dev):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    return t.to(device=dev, dtype=tdtype).contiguous()


def seed_to_u64(seed):
    """int, numpy SeedSequence or None -> 64-bit Philox key."""
    if seed is None:
        seed = np.random.SeedSequence()
    if isinstance(seed, np.random.SeedSequence):
        lo, hi = seed.generate_state(2, np.uint32)
        return (int(hi) << 32) | int(lo)
    if isinstance(seed, np.random.Generator):
        return int(seed.integers(0, 2**63))
    return int(seed) & 0xFFFFFFFFFFFFFFFF
