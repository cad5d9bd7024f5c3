"""CPU, world_size 2 over gloo: the meshlet-range sharding of one view — partition, pyramid
broadcast, survivor all-gather — with the oracle standing in for the per-rank compute. The sharded result must
equal the unsharded one byte for byte (rank-major order == canonical draw order)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_ref as O
    from orbit_b200 import scenes
    from orbit_b200.multi_gpu import broadcast_pyramid, exchange_survivors, partition_draws, views_for_rank
    sc, view = scenes.config_c2(scale=0.04)        # 400 buildings, 80k meshlets
    depth = scenes.make_depth(sc, view)
    lod0 = sc.mesh_infos["mesh_lods"][:, 0, 1][sc.draws["mesh_index"]]
    parts = partition_draws(lod0, world)
    b, e = parts[rank]
    hs = O.HostScene(sc, draw_begin=b, draw_end=e)
    results = {}
    for f in range(2):
        early = O.cull_pass(hs, O.gpu_cull_info(view, "read"))
        # rank 0 owns the depth buffer: builds the pyramid, everyone receives it
        info = O.hiz_geometry(view.width, view.height)
        tex = torch.zeros(info.total_texels, dtype=torch.float32)
        if rank == 0:
            _, t = O.hiz_build(depth)
            tex.copy_(torch.from_numpy(t))
        broadcast_pyramid(tex, src=0)
        hs.hiz_info, hs.hiz_texels, hs.depth_size = info, tex.numpy().copy(), (view.width, view.height)
        late = O.cull_pass(hs, O.gpu_cull_info(view, "write"))
        for name, pair in (("early", early), ("late", late)):
            merged, counts = exchange_survivors(torch.from_numpy(pair[1]))
            results["f%d_%s" % (f, name)] = merged.numpy().copy()
            results["f%d_%s_counts" % (f, name)] = np.array(counts)
    results["entity_vis"] = hs.entity_visibility.copy()
    results["meshlet_vis"] = hs.meshlet_visibility.copy()
    results["range"] = np.array([b, e])
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **results)
    assert views_for_rank(5, rank, world) == [v for v in range(5) if v % world == rank]
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_view_equals_unsharded(tmp_path, oracle):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from orbit_b200 import scenes
    O = oracle
    sc, view = scenes.config_c2(scale=0.04)
    depth = scenes.make_depth(sc, view)
    hs = O.HostScene(sc)
    ranks = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for f in range(2):
        ref = O.depth_prepass_culling(hs, view, depth)
        for name in ("early", "late"):
            n, draws = O.parse_draws(ref[name][1])
            for r in range(world):
                got = ranks[r]["f%d_%s" % (f, name)]
                assert int(got[:4].view(np.uint32)[0]) == n, (f, name, r)
                assert np.array_equal(got[4:], draws.view(np.uint8)), (f, name, r)
            assert int(ranks[0]["f%d_%s_counts" % (f, name)].sum()) == n
    # visibility words: each rank owns the words of its draw range; together they equal the unsharded state
    ev = np.zeros_like(hs.entity_visibility); mv = np.zeros_like(hs.meshlet_visibility)
    words = (200 + 31) // 32
    for r in range(world):
        b, e = ranks[r]["range"]
        ev[b // 32:(e + 31) // 32] = ranks[r]["entity_vis"][b // 32:(e + 31) // 32]
        mv[b * words:e * words] = ranks[r]["meshlet_vis"][b * words:e * words]
    assert np.array_equal(ev, hs.entity_visibility) and np.array_equal(mv, hs.meshlet_visibility)
