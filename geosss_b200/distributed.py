"""
Chains over GPUs: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The reference parallelises over chains with joblib processes and pickled results
(scripts/protein_reg3d3d.py:296-349).  Here the global chain ids [0, N) are split into
contiguous blocks, one per rank; each rank runs its block with ``chain_offset`` = first global
id, so every chain draws the same Philox stream whatever the number of GPUs, and the sampling
kernels need no collective at all.  Collectives only follow the path where the caller asks for
them: gathering retained samples (all_gather) and reducing diagnostic partial sums (all_reduce).
"""

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_chains, rank, world):
    """Contiguous block of global chain ids owned by ``rank``: (offset, count)."""
    base, extra = divmod(int(n_chains), int(world))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def shard_initial_states(initial_states, rank, world):
    """This rank's rows of an (N, d) array of initial states plus its chain offset."""
    off, cnt = shard_range(len(initial_states), rank, world)
    return initial_states[off:off + cnt], off


def make_sampler(cls, pdf, initial_states, seed, rank=None, world=None, **kwargs):
    """Sampler over this rank's chain block; ``initial_states`` is the full (N, d) array."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    local, off = shard_initial_states(initial_states, rank, world)
    return cls(pdf, local, seed, chain_offset=off, **kwargs)


def gather_chains(local, n_chains, group=None):
    """all_gather of per-rank (count_r, ...) tensors into the full (N, ...) tensor on every rank.
    Blocks may differ by one chain; they are padded to the largest block for the collective."""
    world = dist.get_world_size(group)
    counts = [shard_range(n_chains, r, world)[1] for r in range(world)]
    biggest = max(counts)
    pad = local
    if local.shape[0] < biggest:
        pad = torch.cat([local, local.new_zeros((biggest - local.shape[0],) + tuple(local.shape[1:]))])
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous(), group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)])


def reduce_moments(samples, group=None):
    """Mean and variance per dimension over all chains of all ranks from (.., d) local samples."""
    flat = samples.reshape(-1, samples.shape[-1]).to(torch.float64)
    acc = torch.cat([flat.sum(0), (flat * flat).sum(0), flat.new_tensor([float(flat.shape[0])])])
    if dist.is_initialized():
        dist.all_reduce(acc, group=group)
    d = samples.shape[-1]
    n = acc[-1]
    mean = acc[:d] / n
    var = acc[d:2 * d] / n - mean * mean
    return mean, var


def reduce_sum(values, group=None):
    """Sum a small tensor / array of partials (e.g. diagnostics.ess_partials, rejection totals)."""
    t = values if isinstance(values, torch.Tensor) else torch.as_tensor(np.asarray(values, dtype=np.float64))
    if dist.is_initialized():
        dist.all_reduce(t, group=group)
    return t


def bind_host_to_gpu(device_index):
    """Pin this process to the CPU cores NVML reports as local to the GPU (same NUMA node), so that the
    pinned buffers of the host-streaming path are allocated next to the GPU's PCIe root and device->host
    copies of different ranks do not cross the socket interconnect.  Returns the CPU list, or None when
    NVML / the affinity call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = device_index
        if vis and all(v.strip().isdigit() for v in vis.split(",")):
            idx = int(vis.split(",")[device_index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
