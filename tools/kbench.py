#!/usr/bin/env python
"""Development micro-benchmark: times the five stage kernels of one C2 steady-state frame with CUDA events,
for a sweep of meshlet-stage tuning knobs (ORBIT_MC_CTAS_PER_SM: CTAs of 8 warps per SM of the meshlet stream test kernel). Not part of the
product or of bench.py's contract; its output goes to gpurun_out/ and summaries to profiles/."""
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from orbit_b200 import frame, scenes
from orbit_b200.passes import Context


def time_frame(ctx, copies, reps=7, per_graph=8):
    """Per-stage device time: a CUDA graph of `per_graph` back-to-back launches of ONE stage rotating over the
    scene copies (no CPU in the loop, inputs out of L2), replayed `reps` times; reports us per launch."""
    stages = {"entity_early": lambda pf, s: pf.entity(False, s), "meshlet_early": lambda pf, s: pf.meshlet(False, s),
              "hiz": lambda pf, s: pf.hiz(s), "entity_late": lambda pf, s: pf.entity(True, s),
              "meshlet_late": lambda pf, s: pf.meshlet(True, s)}
    # pass 0 (frustum + cone only, no Hi-Z) over every meshlet the frustum keeps: the HBM-heaviest use of the stage
    from orbit_b200.passes import OcclusionCullInfo, create_meshlet_dispatch_command, create_meshlet_draw_commands
    p0 = []
    for i, pf in enumerate(copies):
        ci = frame.cull_info_for(pf.view, OcclusionCullInfo("none"))
        _, disp = create_meshlet_dispatch_command(ctx, "p0_%d" % i, pf.dscene.assets, pf.dscene.scene, ci)
        p0.append((ci, disp))
    stages["meshlet_pass0"] = lambda pf, s: create_meshlet_draw_commands(ctx, "p0_%d" % copies.index(pf), pf.dscene.assets, pf.dscene.scene, p0[copies.index(pf)][0], p0[copies.index(pf)][1])
    # the same sweep with the mesh-shading output as well (forward_depth_prepass.task:224-256): 44 B per record
    tp = [torch.zeros(44 * int(pf.dscene.scene.record_capacity), dtype=torch.uint8, device=ctx.device) for pf in copies]
    stages["meshlet_pass0_task_payloads"] = lambda pf, s: create_meshlet_draw_commands(
        ctx, "p0_%d" % copies.index(pf), pf.dscene.assets, pf.dscene.scene, p0[copies.index(pf)][0], p0[copies.index(pf)][1], tp[copies.index(pf)])
    out = {}
    for name, fn in stages.items():
        for pf in copies:          # consistent steady-state inputs for every stage
            pf.launch()
            fn(pf, pf._stream())   # the stage itself once outside the capture: transient buffers are created (and zero-filled) here
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(per_graph):
                pf = copies[i % len(copies)]
                fn(pf, pf._stream())
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3 / per_graph)
        out[name] = (float(np.median(ts[1:])), float(np.min(ts[1:])))
    return out


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "sweep"
    if os.environ.get("KBENCH_CONFIG", "c2") == "c3":      # per-stage times of the 50 M-meshlet instanced 4K view
        scene, view = scenes.config_c3()
    else:
        scene, _ = scenes.config_c2()
        view = bench.c2_view(scenes, scene, 0)
    depth_np = scenes.make_depth(scene, view)
    configs = [(None, None)] if which == "default" else list(itertools.product([0], [2, 3]))
    for rpw, cps in configs:
        if rpw is not None:
            os.environ["ORBIT_MC_CTAS_PER_SM"] = str(cps)
        ctx = Context(0)
        copies = []
        for i in range(int(os.environ.get("KBENCH_COPIES", "4"))):   # 1 = everything L2-resident (latency floor of each stage)
            ds = frame.DeviceScene.upload(ctx, scene)
            vs = frame.ViewState(ctx, ds, (view.width, view.height), name="v%d" % i)
            pf = frame.PreparedFrame(ctx, ds, vs, view, torch.from_numpy(depth_np).to(ctx.device), name="c%d" % i)
            pf.launch(); pf.launch()
            copies.append(pf)
        torch.cuda.synchronize()
        t = time_frame(ctx, copies)
        hdr, recs = frame.read_dispatch(copies[0].context._transients["p0_0_meshlet_dispatch_buffer"])
        n0, _ = frame.read_draws(copies[0].context._transients["p0_0_meshlet_draw_command_buffer"], capacity=0)
        lanes0 = int(recs["meshlet_count"].sum())
        bytes0 = 32 * lanes0 + 16 * len(recs) + 64 * len(np.unique(recs["entity_index"])) + 28 * n0
        t["pass0_GBs"] = (bytes0 / (t["meshlet_pass0"][0] * 1e-6) / 1e9, 0)
        t["pass0_lanes_M"] = (lanes0 / 1e6, 0)
        t["pass0_survivors_M"] = (n0 / 1e6, 0)
        print(json.dumps({"recs_per_warp": rpw, "ctas_per_sm": cps, **{k: round(v[0], 2) for k, v in t.items()},
                          "min_meshlet_late": round(t["meshlet_late"][1], 2)}), flush=True)
        del copies
        ctx.close()


if __name__ == "__main__":
    main()
