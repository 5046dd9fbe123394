#!/usr/bin/env python
"""How much host<->device bandwidth one rank keeps when all ranks of a node copy at once.

The end-to-end leg of bench.py moves ~36 MB up and ~43 MB down per step and rank through pinned host memory.  This probe
times the same two copies (a) on rank 0 alone and (b) on all ranks at the same time, so that the end-to-end number at
N GPUs can be read against what the host can deliver.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_concurrency.py
Prints one JSON line on rank 0.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    bound = None
    if os.environ.get("SZ_BIND_NUMA", "0") == "1":
        import bench
        bound = bench.bind_to_gpu_numa_node(local)
    up_b, dn_b = 36 << 20, 43 << 20
    h_up = torch.empty(up_b, dtype=torch.uint8, pin_memory=True)
    h_dn = torch.empty(dn_b, dtype=torch.uint8, pin_memory=True)
    h_up.fill_(1)
    h_dn.fill_(2)
    d_up = torch.empty(up_b, dtype=torch.uint8, device="cuda")
    d_dn = torch.empty(dn_b, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def one(both):
        with torch.cuda.stream(s_up):
            d_up.copy_(h_up, non_blocking=True)
        with torch.cuda.stream(s_dn if both else s_up):
            h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.synchronize()

    def run(active, both, reps=40):
        if world > 1:
            dist.barrier()
        ms = None
        if active:
            for _ in range(5):
                one(both)
            t0 = time.perf_counter()
            for _ in range(reps):
                one(both)
            ms = (time.perf_counter() - t0) / reps * 1e3
        if world > 1:
            dist.barrier()
        return ms

    res = {}
    res["alone_serial_ms"] = run(rank == 0, False)
    res["alone_overlap_ms"] = run(rank == 0, True)
    res["all_serial_ms"] = run(True, False)
    res["all_overlap_ms"] = run(True, True)
    rows = [None] * world
    if world > 1:
        dist.all_gather_object(rows, {"rank": rank, "bound": bound, **res})
    else:
        rows = [{"rank": 0, "bound": bound, **res}]
    if rank == 0:
        gb = (up_b + dn_b) / 1e6
        out = {"bytes_up": up_b, "bytes_down": dn_b, "world": world,
               "alone_serial_ms": rows[0]["alone_serial_ms"], "alone_overlap_ms": rows[0]["alone_overlap_ms"],
               "all_serial_ms": [r["all_serial_ms"] for r in rows], "all_overlap_ms": [r["all_overlap_ms"] for r in rows],
               "alone_GBps": gb / rows[0]["alone_overlap_ms"], "all_GBps_per_rank": [gb / r["all_overlap_ms"] for r in rows],
               "bound": [r["bound"] for r in rows]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
