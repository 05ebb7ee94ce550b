#!/usr/bin/env python
"""Probe (run under torchrun, >= 2 ranks): what torch.distributed._symmetric_memory offers on this box --
peer pointers, multicast pointer, signal pads, barrier, capture in a CUDA graph -- and how fast a peer copy of
one rank's z block is.  Prints one JSON line per rank 0.  Only used to decide whether the exchange of
z / row statistics can be written as peer-memory stores from this library's own kernels."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
out = {"world": world}
try:
    import torch.distributed._symmetric_memory as sm
    n_local = 65536 // world
    t = sm.empty((65536, 256), dtype=torch.bfloat16, device=dev)
    h = sm.rendezvous(t, dist.group.WORLD)
    out["buffer_ptrs"] = len(h.buffer_ptrs)
    try:
        from torch._C._autograd import DeviceType
        out["multicast"] = bool(type(h).has_multicast_support(DeviceType.CUDA, dev.index))
    except Exception as exc:  # noqa: BLE001
        out["multicast"] = f"{type(exc).__name__}"
    out["multicast_ptr"] = int(getattr(h, "multicast_ptr", 0) or 0) != 0
    out["signal_pad_bytes"] = int(h.signal_pad_size)
    t.zero_()
    h.barrier(channel=0)
    # push my block into every peer's buffer through get_buffer views (plain copies: copy engine or SM kernel)
    mine = torch.full((n_local, 256), float(rank + 1), dtype=torch.bfloat16, device=dev)
    peers = [h.get_buffer(p, (65536, 256), torch.bfloat16) for p in range(world)]
    def push():
        for k in range(world):
            p = (rank + k) % world
            peers[p][rank * n_local:(rank + 1) * n_local].copy_(mine)
    push(); h.barrier(channel=0); torch.cuda.synchronize()
    ok = all(float(t[p * n_local, 0]) == p + 1 for p in range(world))
    out["push_correct"] = ok
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        push(); h.barrier(channel=0)
    e1.record(); torch.cuda.synchronize()
    out["push_plus_barrier_us"] = round(1e3 * e0.elapsed_time(e1) / 20, 1)
    e0.record()
    for _ in range(20):
        h.barrier(channel=0)
    e1.record(); torch.cuda.synchronize()
    out["barrier_us"] = round(1e3 * e0.elapsed_time(e1) / 20, 1)
    # NCCL all-gather of the same data for comparison
    z_all = torch.empty((65536, 256), dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        dist.all_gather_into_tensor(z_all, mine)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        dist.all_gather_into_tensor(z_all, mine)
    e1.record(); torch.cuda.synchronize()
    out["nccl_all_gather_32MB_us"] = round(1e3 * e0.elapsed_time(e1) / 20, 1)
    small = torch.zeros(n_local, 8, device=dev); small_all = torch.empty(65536, 8, device=dev)
    for _ in range(3):
        dist.all_gather_into_tensor(small_all, small)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        dist.all_gather_into_tensor(small_all, small)
    e1.record(); torch.cuda.synchronize()
    out["nccl_all_gather_2MB_us"] = round(1e3 * e0.elapsed_time(e1) / 20, 1)
    # graph capture of push + barrier
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
                push(); h.barrier(channel=0)
        torch.cuda.synchronize(); dist.barrier()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        out["graph_push_plus_barrier_us"] = round(1e3 * e0.elapsed_time(e1) / 20, 1)
    except Exception as exc:  # noqa: BLE001
        out["graph_error"] = f"{type(exc).__name__}: {exc}"[:300]
except Exception as exc:  # noqa: BLE001
    out["error"] = f"{type(exc).__name__}: {exc}"[:500]
if rank == 0:
    print(json.dumps(out), flush=True)
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0)
