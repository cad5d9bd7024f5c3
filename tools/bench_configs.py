#!/usr/bin/env python
"""Measures the BASELINE.json configs beyond the headline one (bench.py covers C2) and checks each against the
oracle at FULL size. One JSON line per config into stdout (and gpurun_out/ when run there).

  C3  instanced 50M-meshlet scene, one 3840x2160 view; meshlet ranges sharded over the ranks
      (torchrun --nproc-per-node N tools/bench_configs.py c3): Hi-Z broadcast + survivor all-gather.
  C4  main view two-pass + 4 orthographic shadow cascades (pass 0) over 10M meshlets + clustered light assignment
      (16x9x24 clusters, 65 536 point lights).
  C5  many-view batch over a 20M-meshlet scene: views v -> rank v mod N, no inter-GPU traffic (default 32 of the
      256 views per run; --views 256 for all).
"""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

from orbit_b200 import frame, multi_gpu, scenes
from orbit_b200 import layouts as L
from orbit_b200.passes import ClusterSettings, Context, OcclusionCullInfo, compute_clusters


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()[:16]


def graph_time(fn, reps=5):
    """us per call of fn(), captured once into a CUDA graph and replayed."""
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def event_time(fn, reps=5):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def cascade_views(cam_view, scene, n=4, res=2048):
    """Orthographic shadow cascades in the spirit of shadow_renderer.rs:466-712: light direction normalize(-1,1,1)
    looking down on the city, cascade i covers a sphere of radius r_i around a point ahead of the camera; all 6
    ortho planes are passed (pass 0, no occlusion). LOD range unchanged (single-LOD scene)."""
    sun = np.array([1.0, -1.0, -1.0]); sun /= np.linalg.norm(sun)
    cam_pos = -cam_view.view[:3, :3].T @ cam_view.view[:3, 3]
    fwd = -cam_view.view[2, :3]
    out = []
    near, far, lam = 0.01, 200.0, 0.8
    for i in range(n):
        t0, t1 = i / n, (i + 1) / n
        split = lambda t: lam * near * (far / near) ** t + (1 - lam) * (near + (far - near) * t)
        d0, d1 = split(t0), split(t1)
        centre = cam_pos + fwd * 0.5 * (d0 + d1)
        r = 0.5 * (d1 - d0) + d1 * 0.8
        eye = centre - sun * (r + 80.0)
        out.append(scenes.orthographic_view(eye, sun, res, res, half_width=r, near=0.0, far=2.0 * r + 80.0, up=(0.0, 1.0, 0.0)))
    return out


def check_against_oracle(name, O, hs, view, kind, gpu_pair, results):
    o = O.cull_pass(hs, O.gpu_cull_info(view, kind))
    ghdr, grecs = frame.read_dispatch(gpu_pair[0]); ohdr, orecs = O.parse_dispatch(o[0])
    gn, gd = frame.read_draws(gpu_pair[1]); on, od = O.parse_draws(o[1])
    ok = ghdr.tolist() == ohdr.tolist() and gn == on and sha(grecs) == sha(orecs) and sha(gd) == sha(od)
    results[name] = {"records": int(ohdr[0]), "draws": int(on), "bit_exact": bool(ok)}
    return ok


def run_c3_timeline(args, rank, world, ctx):
    """Where the sharded frame's time goes: CUDA events at the stage boundaries of ShardedView.step_best, every rank."""
    scene, view = scenes.config_c3(args.scale)
    depth = scenes.make_depth(scene, view) if rank == 0 else np.zeros((view.height, view.width), np.float32)
    sv = multi_gpu.ShardedView(ctx, scene, view, depth, rank, world)
    sv.enable_mask_exchange(scene.n_records_lod0, scene.n_meshlet_instances)
    for _ in range(3):
        sv.step_best()
    torch.cuda.synchronize(); dist.barrier()
    rows = []
    for _ in range(5):
        marks = {}
        dist.barrier(); torch.cuda.synchronize()
        sv.step_best(marks)
        torch.cuda.synchronize()
        rows.append({k: marks["start"].elapsed_time(v) * 1e3 for k, v in marks.items()})
    med = {k: float(np.median([r[k] for r in rows])) for k in rows[0]}
    allm = [None] * world
    dist.all_gather_object(allm, med)
    if rank == 0:
        keys = list(med)
        print("C3 sharded frame, %d GPUs: us from the frame's start (median of 5), per rank" % world)
        for k in keys:
            print("  %-26s %s" % (k, "  ".join("%7.1f" % m[k] for m in allm)))
    sv.close()


def run_c3(args, rank, world, ctx):
    if args.timeline:
        return run_c3_timeline(args, rank, world, ctx)
    import oracle_ref as O
    scene, view = scenes.config_c3(args.scale)
    depth = scenes.make_depth(scene, view)
    sv = multi_gpu.ShardedView(ctx, scene, view, depth, rank, world)
    for _ in range(2):
        sv.step(exchange=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # compute only (no survivor exchange) and full step incl. exchange; max over ranks
    t_compute = event_time(lambda: sv.step(exchange=False))[0]
    t_full = event_time(lambda: sv.step(exchange=True))[0]
    early, late = sv.step(exchange=True)
    torch.cuda.synchronize()
    t_peer, peer_ok = None, None
    if world > 1:
        sv.enable_peer_exchange(scene.n_meshlet_instances)
        sv.step(exchange="peer")
        t_peer = event_time(lambda: sv.step(exchange="peer"))[0]
        n_e, n_l = sv.step(exchange="peer")
        torch.cuda.synchronize()
        pe, pl = sv.peer_early.read(n_e), sv.peer_late.read(n_l)
        torch.cuda.synchronize()
        peer_ok = bool(torch.equal(pe, early) and torch.equal(pl, late))
        # gather: only rank 0 (the GPU that submits the draws) receives the list
        sv.step(exchange="gather")
        t_gather = event_time(lambda: sv.step(exchange="gather"))[0]
        sv.peer_early.clear(); sv.peer_late.clear()
        dist.barrier()
        n_e, n_l = sv.step(exchange="gather")
        torch.cuda.synchronize()
        gather_ok = True
        if rank == 0:
            gather_ok = bool(torch.equal(sv.peer_early.read(n_e), early) and torch.equal(sv.peer_late.read(n_l), late))
        # the same two exchanges with the counts kept on the device (no host round trip inside the frame)
        t_async = {}
        for mode in ("peer_async", "gather_async"):
            sv.step(exchange=mode)
            t_async[mode] = event_time(lambda: sv.step(exchange=mode))[0]
            sv.peer_early.clear(); sv.peer_late.clear()
            dist.barrier()
            c_e, c_l = sv.step(exchange=mode)
            torch.cuda.synchronize()
            if rank == 0 or mode == "peer_async":
                ok = bool(torch.equal(sv.peer_early.read(int(c_e.sum())), early) and torch.equal(sv.peer_late.read(int(c_l.sum())), late))
                t_async[mode + "_ok"] = ok
        # early list's exchange overlapped with Hi-Z + late pass, one closing fence
        sv.step_overlapped(0)
        t_async["gather_overlapped"] = event_time(lambda: sv.step_overlapped(0))[0]
        sv.peer_early.clear(); sv.peer_late.clear()
        dist.barrier()
        c_e, c_l = sv.step_overlapped(0)
        torch.cuda.synchronize()
        if rank == 0:
            t_async["gather_overlapped_ok"] = bool(torch.equal(sv.peer_early.read(int(c_e.sum())), early) and torch.equal(sv.peer_late.read(int(c_l.sum())), late))
        tt = torch.tensor([t_async["gather_overlapped"]], device=ctx.device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_async["gather_overlapped"] = float(tt[0])
        t = torch.tensor([t_compute, t_full, t_peer, t_gather, t_async["peer_async"], t_async["gather_async"]], device=ctx.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_compute, t_full, t_peer, t_gather = float(t[0]), float(t[1]), float(t[2]), float(t[3])
        t_async["peer_async"], t_async["gather_async"] = float(t[4]), float(t[5])
    res = {"config": "C3", "n_gpus": world, "entities": scene.n_entities, "meshlet_instances": scene.n_meshlet_instances,
           "ranges": sv.ranges, "us_compute": t_compute, "us_with_exchange": t_full,
           "gmeshlets_per_s_compute": scene.n_meshlet_instances / t_compute / 1e3,
           "gmeshlets_per_s_with_exchange": scene.n_meshlet_instances / t_full / 1e3,
           "pyramid_bytes": int(sv.vstate.depth_pyramid.texels.numel() * 4),
           "us_with_peer_exchange": t_peer, "peer_exchange_equals_allgather": peer_ok,
           "gmeshlets_per_s_with_peer_exchange": (scene.n_meshlet_instances / t_peer / 1e3) if t_peer else None}
    if world > 1:
        res["us_with_gather_to_rank0"] = t_gather
        res["gather_equals_allgather_on_rank0"] = gather_ok
        res["gmeshlets_per_s_with_gather"] = scene.n_meshlet_instances / t_gather / 1e3
        res["device_side_counts"] = {"us_with_peer_exchange": t_async["peer_async"], "us_with_gather_to_rank0": t_async["gather_async"],
                                     "peer_ok": t_async.get("peer_async_ok"), "gather_ok_on_rank0": t_async.get("gather_async_ok"),
                                     "us_with_gather_overlapped": t_async["gather_overlapped"], "gather_overlapped_ok_on_rank0": t_async.get("gather_overlapped_ok"),
                                     "gmeshlets_per_s_with_gather_overlapped": scene.n_meshlet_instances / t_async["gather_overlapped"] / 1e3,
                                     "gmeshlets_per_s_with_gather": scene.n_meshlet_instances / t_async["gather_async"] / 1e3,
                                     "gmeshlets_per_s_with_peer_exchange": scene.n_meshlet_instances / t_async["peer_async"] / 1e3}
    if rank == 0:
        n_early = int(early[:4].view(torch.int32).item()); n_late = int(late[:4].view(torch.int32).item())
        res["survivors_early"], res["survivors_late"] = n_early, n_late
        if not args.no_oracle:
            hs = O.HostScene(scene)
            t0 = time.perf_counter()
            for _ in range(3):   # frame 0, 1, 2 -> same steady state as the GPU after its 2 + timing frames
                o = O.depth_prepass_culling(hs, view, depth)
            res["oracle_s_per_frame"] = (time.perf_counter() - t0) / 3
            on, od = O.parse_draws(o["early"][1]); ln, ld = O.parse_draws(o["late"][1])
            ge = early.cpu().numpy(); gl = late.cpu().numpy()
            res["bit_exact_vs_oracle"] = bool(on == n_early and ln == n_late and sha(od) == sha(ge[4:]) and sha(ld) == sha(gl[4:]))
        print(json.dumps(res), flush=True)


def run_c4(args, rank, world, ctx):
    import oracle_ref as O
    scene, view = scenes.config_c4(args.scale)
    depth = scenes.make_depth(scene, view)
    lights = scenes.make_lights(scenes.SEEDS["C4"], 65536 if args.scale >= 1.0 else 4096, scene.aabb_min, scene.aabb_max)
    ds = frame.DeviceScene.upload(ctx, scene, lights=lights)
    vs = frame.ViewState(ctx, ds, (view.width, view.height))
    d_depth = torch.from_numpy(depth).to(ctx.device)
    pf = frame.PreparedFrame(ctx, ds, vs, view, d_depth)
    pf.launch(); pf.launch(); torch.cuda.synchronize()
    res = {"config": "C4", "entities": scene.n_entities, "meshlet_instances": scene.n_meshlet_instances, "lights": len(lights)}
    res["main_view_two_pass_us"] = graph_time(pf.launch)[0]
    casc = cascade_views(view, scene)
    hs = O.HostScene(scene)
    checks = {}
    t_c = []
    for i, cv in enumerate(casc):
        info = frame.cull_info_for(cv, OcclusionCullInfo("none"))
        fn = lambda info=info, i=i: frame.cull_pass(ctx, "cascade%d" % i, ds, info)
        pair = fn(); torch.cuda.synchronize()
        t_c.append(graph_time(fn)[0])
        if not args.no_oracle:
            check_against_oracle("cascade%d" % i, O, hs, cv, "none", pair, checks)
    res["cascade_us"] = t_c
    st = ClusterSettings(screen_resolution=(1920, 1080), z_slice_count=24, tile_size_px=120)
    fn = lambda: compute_clusters(ctx, st, view.view, view.projection_matrix, view.near, d_depth, ds.scene)
    info, params = fn(); torch.cuda.synchronize()
    res["clusters_us"] = graph_time(fn)[0]
    n = 16 * 9 * 24
    unique = info.unique_cluster_buffer[:16 + 4 * n].cpu().numpy().view(np.uint32)
    na = int(unique[3])
    total = int(info.light_index_buffer[:4].cpu().numpy().view(np.uint32)[0])
    res["active_clusters"], res["light_indices"] = na, total
    res["sphere_aabb_tests"] = na * len(lights)
    if not args.no_oracle:
        t0 = time.perf_counter()
        ref = O.light_cluster(params, depth, lights)
        res["oracle_clusters_s"] = time.perf_counter() - t0
        idx = info.light_index_buffer[:4 + 4 * total].cpu().numpy().view(np.uint32)
        img = info.light_offset_image[:8 * n].cpu().numpy().view(np.uint32)
        checks["clusters"] = {"bit_exact": bool(int(ref["unique"][3]) == na and int(ref["index"][0]) == total
                                                and np.array_equal(ref["index"][:1 + total], idx) and np.array_equal(ref["image"], img)
                                                and np.array_equal(ref["unique"][:4 + na], unique[:4 + na]))}
    res["checks"] = checks
    frames_us = res["main_view_two_pass_us"] + sum(t_c)
    res["gmeshlets_per_s_5_views"] = 5 * scene.n_meshlet_instances / frames_us / 1e3
    print(json.dumps(res), flush=True)


def run_c5(args, rank, world, ctx):
    import oracle_ref as O
    scene, views = scenes.config_c5(args.scale, n_views=args.views)
    mine = multi_gpu.views_for_rank(len(views), rank, world)
    ds = frame.DeviceScene.upload(ctx, scene)
    hs = O.HostScene(scene)
    checks = {}
    # pass 0 (frustum + cone) for every view of this rank, one graph per view
    fns = []
    for v in mine:
        info = frame.cull_info_for(views[v], OcclusionCullInfo("none"))
        fns.append(lambda info=info: frame.cull_pass(ctx, "view", ds, info))
    pairs = fns[0](); torch.cuda.synchronize()
    if not args.no_oracle and rank == 0:
        check_against_oracle("view%d_pass0" % mine[0], O, hs, views[mine[0]], "none", pairs, checks)
    def all_views():
        for f in fns:
            f()
    all_views(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = event_time(all_views, reps=3)[0]
    if world > 1:
        tt = torch.tensor([t], device=ctx.device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt[0])
    if rank == 0:
        print(json.dumps({"config": "C5", "n_gpus": world, "views": len(views), "entities": scene.n_entities,
                          "meshlet_instances": scene.n_meshlet_instances, "us_all_views_max_rank": t,
                          "us_per_view": t / max(len(mine), 1),
                          "gmeshlets_per_s": len(views) * scene.n_meshlet_instances / t / 1e3, "checks": checks}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--views", type=int, default=32)
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--timeline", action="store_true", help="c3: stage-boundary timeline of the sharded frame")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    {"c3": run_c3, "c4": run_c4, "c5": run_c5}[args.which](args, rank, world, ctx)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
