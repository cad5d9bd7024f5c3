"""Two GPUs, one process per GPU over NCCL (skipped on a one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
the sharded view of SURVEY §8e through the CUDA path — entity ranges per rank (draw_begin / draw_end), pyramid broadcast,
survivor exchange in both forms (28-byte commands by peer stores; 16-byte record entries + emission on rank 0) — must give
rank 0 the unsharded oracle lists; an overflowing rank must not corrupt the assembled list."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import oracle_ref as O
    from orbit_b200 import frame, multi_gpu, scenes
    from orbit_b200.passes import Context
    ctx = Context(rank)
    sc, view = scenes.config_c3(scale=0.02)          # ~5k instanced entities, 1M meshlet instances, 4K view
    depth = scenes.make_depth(sc, view)
    sv = multi_gpu.ShardedView(ctx, sc, view, depth, rank, world)
    sv.enable_mask_exchange(sc.n_records_lod0, sc.n_meshlet_instances)
    sv.enable_peer_exchange(sc.n_meshlet_instances)
    res = {}
    hs = O.HostScene(sc) if rank == 0 else None
    for f in range(3):
        sv.clear_gathered(); sv.peer_early.clear(); sv.peer_late.clear()
        dist.barrier()
        if f < 2:
            sv.step_best()
            torch.cuda.synchronize(); dist.barrier()
            if rank == 0:
                early, late = sv.gathered_lists()
        else:
            c_e, c_l = sv.step_overlapped(0)         # the 28-byte form, same frame protocol
            torch.cuda.synchronize(); dist.barrier()
            if rank == 0:
                early, late = sv.peer_early.read(int(c_e.sum())), sv.peer_late.read(int(c_l.sum()))
        if rank == 0:
            o = O.depth_prepass_culling(hs, view, depth)
            for name, g, ob in (("early", early, o["early"][1]), ("late", late, o["late"][1])):
                gn = int(g[:4].view(torch.int32).item()); on, od = O.parse_draws(ob)
                res["f%d_%s" % (f, name)] = bool(gn == on and np.array_equal(g[4:4 + 28 * gn].cpu().numpy(), od.view(np.uint8).reshape(-1)))
                res["f%d_%s_n" % (f, name)] = on
    if rank == 0:
        np.savez(os.path.join(out_dir, "rank0.npz"), **res)
    sv.close()
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


def test_sharded_view_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r = np.load(os.path.join(str(tmp_path), "rank0.npz"))
    for f in range(3):
        for name in ("early", "late"):
            assert bool(r["f%d_%s" % (f, name)]), (f, name)
    assert int(r["f1_early_n"]) > 0
