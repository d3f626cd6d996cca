"""Device->host bandwidth of one 8-GPU host under different copy SCHEDULES and host-buffer kinds.

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/d2h_schedule_probe.py

Every variant: the ranks listed copy the same 512 MB device buffer to their own host buffer three times with plain
cudaMemcpyAsync (one call per chunk), the others stay idle; barrier before, wall clock of the slowest participant.
Rank 0 prints one JSON line per variant: per-rank GB/s and the aggregate.  Buffer kinds: torch pinned (cudaHostAlloc
default), cudaHostAllocWriteCombined, cudaHostRegister of a 2 MB-aligned malloc (transparent huge pages requested).
"""
import ctypes
import json
import mmap
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rt = ctypes.CDLL("libcudart.so.12")
NBYTES = 512 << 20
dbuf = torch.empty(NBYTES, dtype=torch.uint8, device=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def gather(v):
    if world == 1:
        return [v]
    out = [None] * world
    dist.all_gather_object(out, v)
    return out


class HostBuf:
    def __init__(self, kind):
        self.kind, self.ptr, self.keep = kind, ctypes.c_void_p(), None
        if kind == "pinned":
            self.keep = torch.empty(NBYTES, dtype=torch.uint8, pin_memory=True)
            self.keep.zero_()
            self.ptr = ctypes.c_void_p(self.keep.data_ptr())
        elif kind == "write_combined":
            rc = rt.cudaHostAlloc(ctypes.byref(self.ptr), ctypes.c_size_t(NBYTES), ctypes.c_uint(0x04))  # cudaHostAllocWriteCombined
            assert rc == 0, rc
        elif kind == "registered_thp":
            self.keep = mmap.mmap(-1, NBYTES + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
            addr = ctypes.addressof(ctypes.c_char.from_buffer(self.keep))
            addr = (addr + (2 << 20) - 1) & ~((2 << 20) - 1)
            try:
                self.keep.madvise(14)  # MADV_HUGEPAGE
            except Exception:
                pass
            ctypes.memset(addr, 0, NBYTES)
            rc = rt.cudaHostRegister(ctypes.c_void_p(addr), ctypes.c_size_t(NBYTES), ctypes.c_uint(0))
            assert rc == 0, rc
            self.ptr = ctypes.c_void_p(addr)

    def copy(self, chunk):
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for off in range(0, NBYTES, chunk):
            n = min(chunk, NBYTES - off)
            rc = rt.cudaMemcpyAsync(ctypes.c_void_p(self.ptr.value + off), ctypes.c_void_p(dbuf.data_ptr() + off), ctypes.c_size_t(n),
                                    ctypes.c_int(2), st)
            assert rc == 0, rc


def run(tag, buf, ranks, chunk=NBYTES, stagger_ms=0.0):
    active = rank in ranks
    if active:
        buf.copy(chunk)
    barrier()
    t0 = time.perf_counter()
    if active:
        if stagger_ms:
            time.sleep(1e-3 * stagger_ms * ranks.index(rank))
        for _ in range(3):
            buf.copy(chunk)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rows = gather({"rank": rank, "gbs": round(3 * NBYTES / dt / 1e9, 1) if active else None, "s": dt if active else 0.0})
    barrier()
    if rank == 0:
        wall = max(r["s"] for r in rows)
        n_act = sum(1 for r in rows if r["gbs"] is not None)
        print(json.dumps({"variant": tag, "buffer": buf.kind, "ranks": list(ranks), "chunk_mb": chunk >> 20, "stagger_ms": stagger_ms,
                          "per_rank_gbs": [r["gbs"] for r in rows], "aggregate_gbs": round(3 * NBYTES * n_act / wall / 1e9, 1)}), flush=True)


pinned = HostBuf("pinned")
everyone = list(range(world))
run("all ranks at once", pinned, everyone)
for r in everyone:
    run(f"rank {r} alone", pinned, [r])
if world >= 8:
    run("ranks 0-3", pinned, [0, 1, 2, 3])
    run("ranks 4-7", pinned, [4, 5, 6, 7])
    run("even ranks", pinned, [0, 2, 4, 6])
    run("odd ranks", pinned, [1, 3, 5, 7])
    run("ranks 0,1,4,5", pinned, [0, 1, 4, 5])
    for pair in ([0, 1], [0, 4], [2, 3], [4, 5], [6, 7], [3, 4]):
        run(f"pair {pair}", pinned, pair)
    for k in (5, 6, 7):
        run(f"first {k} ranks", pinned, list(range(k)))
run("all ranks, 64 MB chunks", pinned, everyone, chunk=64 << 20)
run("all ranks, 16 MB chunks", pinned, everyone, chunk=16 << 20)
run("all ranks, starts staggered by 5 ms", pinned, everyone, stagger_ms=5.0)
run("all ranks, starts staggered by 20 ms", pinned, everyone, stagger_ms=20.0)
for kind in ("write_combined", "registered_thp"):
    try:
        b = HostBuf(kind)
        run("all ranks at once", b, everyone)
        run("rank 0 alone", b, [0])
    except Exception as exc:
        if rank == 0:
            print(json.dumps({"variant": kind, "error": str(exc)}), flush=True)
if world > 1:
    dist.destroy_process_group()
