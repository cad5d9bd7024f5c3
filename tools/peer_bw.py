"""Bandwidth of orbit_draws_scatter (the survivor-exchange store kernel) to the local buffer and to a peer over NVLink.
torchrun --nproc-per-node 2 tools/peer_bw.py [n_commands]"""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orbit_b200 import multi_gpu  # noqa: E402
from orbit_b200.passes import Context  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    ctx = Context(torch.cuda.current_device())
    pe = multi_gpu.PeerExchange(ctx, n * world)
    src = torch.randint(0, 255, (4 + 28 * n,), dtype=torch.uint8, device=ctx.device)
    src[:4] = torch.tensor([n], dtype=torch.int32, device=ctx.device).view(torch.uint8)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {"rank": rank, "commands": n, "MB": 28 * n / 1e6}
    for name, dst_rank in (("local", rank), ("peer", (rank + 1) % world)):
        ts = []
        for _ in range(6):
            dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pe.lib.orbit_draws_scatter(ctx._h, C.c_void_p(src.data_ptr()), pe.capacity, C.c_void_p(pe.peer_ptrs[dst_rank]), rank * n, world * n, pe.capacity, stream)
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        us = sorted(ts[1:])[len(ts[1:]) // 2]
        out[name + "_us"] = round(us, 1); out[name + "_GBs"] = round(28 * n / us / 1e3, 1)
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        for g in gathered:
            print(json.dumps(g), flush=True)
    dist.barrier()
    pe.close(); ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
