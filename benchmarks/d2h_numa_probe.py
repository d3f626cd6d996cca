"""Where does the pinned device->host bandwidth go when all ranks of one host copy at once?
torchrun --nproc-per-node N benchmarks/d2h_numa_probe.py   (diagnostic; prints one JSON line per variant on rank 0)"""
import ctypes, json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
libc = ctypes.CDLL("libc.so.6", use_errno=True)


def page_node(t):
    """NUMA node of the first page of a host tensor (move_pages with nodes = NULL only queries)."""
    page = ctypes.c_void_p(t.data_ptr() & ~4095)
    status = ctypes.c_int(-99)
    rc = libc.syscall(279, 0, ctypes.c_ulong(1), ctypes.byref(page), None, ctypes.byref(status), 0)
    return status.value if rc == 0 else -1000 - ctypes.get_errno()


def gather(v):
    if world == 1:
        return [v]
    out = [None] * world
    dist.all_gather_object(out, v)
    return out


def probe(nbytes, tag):
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.zero_()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    host.copy_(d)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        host.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    gbs = 3 * nbytes / (time.perf_counter() - t0) / 1e9
    rows = gather({"rank": rank, "gbs": round(gbs, 1), "page_node": page_node(host)})
    if rank == 0:
        print(json.dumps({"variant": tag, "per_rank": rows, "total_gbs": round(sum(r["gbs"] for r in rows), 1)}), flush=True)
    del host, d


def info():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    bus = pynvml.nvmlDeviceGetPciInfo(h).busId
    bus = bus.decode() if isinstance(bus, bytes) else bus
    node = "?"
    for cand in (bus.lower(), bus.lower()[4:]):
        p = f"/sys/bus/pci/devices/{cand}/numa_node"
        if os.path.exists(p):
            node = open(p).read().strip()
            break
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
    st = {l.split(":")[0]: l.split(":")[1].strip() for l in open("/proc/self/status") if l.startswith(("Cpus_allowed_list", "Mems_allowed_list"))}
    return {"rank": rank, "bus": bus, "gpu_numa_node": node, "nvml_cpus": f"{cpus[0]}-{cpus[-1]} ({len(cpus)})" if cpus else "", **st}, cpus


meta, cpus = info()
rows = gather(meta)
if rank == 0:
    print(json.dumps({"topology": rows, "nodes": sorted(os.listdir("/sys/devices/system/node")) if os.path.exists("/sys/devices/system/node") else None}), flush=True)
probe(512 << 20, "default placement")
allowed = sorted(set(cpus) & os.sched_getaffinity(0))
if allowed:
    os.sched_setaffinity(0, allowed)
    time.sleep(0.05)
probe(513 << 20, "after sched_setaffinity to NVML-local cpus (new pinned allocation)")
if world > 1:
    dist.destroy_process_group()
