"""What the host side of this box moves: plain cudaMemcpyAsync loops (one copy per call) between pinned host memory
and the GPU, every rank of the job at the same time.  Launch like bench.py:
    python tools/pcie_ceiling.py                                                           (1 GPU)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ceiling.py
Rank 0 prints one JSON record: per-direction GB/s per GPU and in aggregate, alone and duplex (D2H of the PCM share
next to the H2D of the bitstream share, 7:1 as in bench.py's e2e step)."""
import json, os, sys, time
import torch

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def maxr(v):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


n = 2 << 30                                           # 2 GiB per copy
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 7, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n // 7, dtype=torch.uint8, device="cuda")
out = {"n_gpus": world, "bytes_per_copy": n, "how": "torch Tensor.copy_(non_blocking=True) = one cudaMemcpyAsync per call, pinned host memory, "
       "3 copies per direction, all ranks at once, wall clock = max over ranks"}
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    fn(); barrier()
    t = time.perf_counter()
    for _ in range(3):
        fn()
    barrier()
    dt = maxr(time.perf_counter() - t)
    out[name + "_GBps_per_gpu"] = 3 * n / dt / 1e9
    out[name + "_GBps_aggregate"] = world * 3 * n / dt / 1e9
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
barrier()
t = time.perf_counter()
with torch.cuda.stream(s1):
    for _ in range(3):
        h.copy_(d, non_blocking=True)
with torch.cuda.stream(s2):
    for _ in range(3):
        d2.copy_(h2, non_blocking=True)
barrier()
dt = maxr(time.perf_counter() - t)
out["duplex_d2h_GBps_per_gpu"] = 3 * n / dt / 1e9
out["duplex_d2h_GBps_aggregate"] = world * 3 * n / dt / 1e9
out["duplex_total_GBps_aggregate"] = world * 3 * (n + n // 7) / dt / 1e9
try:
    out["host_cores"] = len(os.sched_getaffinity(0))
except AttributeError:
    pass
if rank == 0:
    print(json.dumps(out))
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open("gpurun_out/pcie_ceiling_n%d.json" % world, "w"), indent=1)
    except OSError:
        pass
if world > 1:
    dist.destroy_process_group()
