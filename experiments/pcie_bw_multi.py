"""pcie_bw_multi.py — aggregate pinned-memory copy bandwidth of the box with 1, 2, 4 or 8 GPUs copying AT THE SAME TIME
(the ceiling of bench.py's end-to-end leg at N GPUs: every rank moves its outputs device -> host concurrently).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        experiments/pcie_bw_multi.py [--gib 2] [--reps 5]

One process per GPU (as bench.py runs), gloo for the barriers.  Each rank times, between barriers: D2H alone, H2D alone,
and both directions at once on two streams (what the pipelined host path does).  Rank 0 prints one JSON line with the
per-rank and aggregate GB/s plus the CPU affinity / NUMA node of every rank."""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def numa_of_gpu(index: int):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        p = f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node"
        with open(p) as f:
            return int(f.read().strip())
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=2.0)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    n = int(args.gib * (1 << 30)) // 8
    d = torch.empty(n, dtype=torch.float64, device="cuda"); d.fill_(1.0)
    d2 = torch.empty(n, dtype=torch.float64, device="cuda")
    h = torch.empty(n, dtype=torch.float64, pin_memory=True); h.fill_(2.0)
    h2 = torch.empty(n, dtype=torch.float64, pin_memory=True); h2.fill_(3.0)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def bar():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn):
        fn(); bar()
        best = 1e30
        for _ in range(args.reps):
            bar()
            t = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t
            # the slowest rank defines the concurrent figure
            x = torch.tensor([dt], dtype=torch.float64)
            if world > 1:
                dist.all_reduce(x, op=dist.ReduceOp.MAX)
            best = min(best, float(x.item()))
        return best

    def both():
        with torch.cuda.stream(s1):
            h.copy_(d, non_blocking=True)
        with torch.cuda.stream(s2):
            d2.copy_(h2, non_blocking=True)

    gb = 8 * n / 1e9
    res = {"d2h": gb / timed(lambda: h.copy_(d, non_blocking=True)),
           "h2d": gb / timed(lambda: d.copy_(h, non_blocking=True)),
           "both_each_direction": gb / timed(both)}
    info = {"rank": rank, "gpu": local, "numa_node_of_gpu": numa_of_gpu(local),
            "cpu_affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."] + [len(os.sched_getaffinity(0))]}
    infos = [None] * world
    if world > 1:
        dist.all_gather_object(infos, info)
    else:
        infos = [info]
    if rank == 0:
        print(json.dumps({"gpus": world, "gib_per_copy": args.gib,
                          "per_gpu_GBps_concurrent": res,
                          "aggregate_GBps": {k: v * world for k, v in res.items()},
                          "ranks": infos, "host_cpus": os.cpu_count()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
