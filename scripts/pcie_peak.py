"""Host <-> device copy ceiling of the box, alone and with every rank copying at once (the e2e leg stops
scaling at ~130 GB/s aggregate on 8 ranks -- is that the hardware?).
usage: python scripts/pcie_peak.py                               one GPU
       python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_peak.py [--bind-numa]
--bind-numa binds each rank to the CPUs of its GPU's NUMA node before the pinned buffer is allocated (first touch)."""
import json, os, sys, time
import torch

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
numa = None
if "--bind-numa" in sys.argv:
    from bench import bind_to_gpu_numa
    numa = bind_to_gpu_numa(lr)
torch.cuda.set_device(lr)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
GB = 1 << 30
x = torch.empty(GB, dtype=torch.uint8).pin_memory()
x.fill_(1)
d = torch.empty(GB, dtype=torch.uint8, device="cuda")


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def rate(fn, reps=6):
    for _ in range(2):
        fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    sync()
    return reps * GB / (time.perf_counter() - t0) / 1e9


h2d = rate(lambda: d.copy_(x, non_blocking=True))
d2h = rate(lambda: x.copy_(d, non_blocking=True))
vals = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
if world > 1:
    allv = [torch.empty_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    per = [[round(float(v[0]), 1), round(float(v[1]), 1)] for v in allv]
    print(json.dumps({"ranks": world, "numa_bound": numa is not None, "rank0_numa": numa,
                      "h2d_gbs_per_rank": [p[0] for p in per], "d2h_gbs_per_rank": [p[1] for p in per],
                      "h2d_gbs_aggregate": round(sum(p[0] for p in per), 1),
                      "d2h_gbs_aggregate": round(sum(p[1] for p in per), 1),
                      "note": "1 GiB pinned buffers, all ranks copying concurrently between barriers"}))
if world > 1:
    dist.destroy_process_group()
