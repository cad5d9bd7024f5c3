#!/usr/bin/env python
"""Development micro-benchmark: times the five stage kernels of one C2 steady-state frame with CUDA events,
for a sweep of meshlet-stage tuning knobs (ORBIT_MC_RECS_PER_WARP x ORBIT_MC_CTAS_PER_SM). Not part of the
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


def time_frame(ctx, copies, reps=20):
    names = ["entity_early", "meshlet_early", "hiz", "entity_late", "meshlet_late"]
    acc = {n: [] for n in names}
    for i in range(reps):
        pf = copies[i % len(copies)]
        s = pf._stream()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record(); pf.entity(False, s); ev[1].record(); pf.meshlet(False, s); ev[2].record(); pf.hiz(s)
        ev[3].record(); pf.entity(True, s); ev[4].record(); pf.meshlet(True, s); ev[5].record()
        torch.cuda.synchronize()
        for k, n in enumerate(names):
            acc[n].append(ev[k].elapsed_time(ev[k + 1]) * 1e3)
    return {n: (float(np.median(v[3:])), float(np.min(v[3:]))) for n, v in acc.items()}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "sweep"
    scene, _ = scenes.config_c2()
    view = bench.c2_view(scenes, scene, 0)
    depth_np = scenes.make_depth(scene, view)
    configs = [(None, None)] if which == "default" else list(itertools.product([2, 4, 8], [1, 2, 3]))
    for rpw, cps in configs:
        if rpw is not None:
            os.environ["ORBIT_MC_RECS_PER_WARP"] = str(rpw)
            os.environ["ORBIT_MC_CTAS_PER_SM"] = str(cps)
        ctx = Context(0)
        copies = []
        for i in range(4):
            ds = frame.DeviceScene.upload(ctx, scene)
            vs = frame.ViewState(ctx, ds, (view.width, view.height), name="v%d" % i)
            pf = frame.PreparedFrame(ctx, ds, vs, view, torch.from_numpy(depth_np).to(ctx.device), name="c%d" % i)
            pf.launch(); pf.launch()
            copies.append(pf)
        torch.cuda.synchronize()
        t = time_frame(ctx, copies)
        print(json.dumps({"recs_per_warp": rpw, "ctas_per_sm": cps, **{k: round(v[0], 2) for k, v in t.items()},
                          "min_meshlet_late": round(t["meshlet_late"][1], 2)}), flush=True)
        del copies
        ctx.close()


if __name__ == "__main__":
    main()
