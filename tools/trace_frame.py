#!/usr/bin/env python
"""Development aid: timeline of one C2 steady-state frame from %globaltimer stamps (a -DORBIT_TRACE build of the library:
ORBIT_B200_LIB=_variants/lib_trace.so python tools/trace_frame.py). Every kernel launch of the frame gets its own block of
stamps; printed per launch: start relative to the frame's first stamp, the gap to the previous launch's last stamp, and for
every stamped point min / median / max over the launch's CTAs (us since the launch's first stamp). Not part of the product."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from orbit_b200 import _lib, frame, scenes
from orbit_b200.passes import Context

NAMES = {0: ("entity_cull", {0: "start", 1: "draw words", 2: "loads issued", 3: "tests done", 4: "cta scan", 9: "first poll back", 5: "gather done", 6: "records written", 7: "exit"}),
         1: ("meshlet_test_direct", {0: "start", 1: "count", 2: "first tile staged", 3: "tiles done", 4: "cta done"}),
         2: ("meshlet_test_packed", {0: "start", 1: "count+words here", 5: "vis words here", 6: "queue built", 7: "meshlets here", 9: "alpha here", 10: "warp0 drained", 11: "cta drained", 3: "rounds done", 4: "cta done"}),
         3: ("meshlet_emit", {0: "start", 1: "totals here", 2: "prefix", 4: "chunk found", 5: "group 0 scanned", 6: "warp0 emitted", 3: "cta emitted"}),
         4: ("hiz_build", {0: "start", 1: "level0 loaded", 2: "tile done", 3: "ticket", 4: "top done"})}
VALUE_SLOTS = {0: {8: "poll iterations"}, 2: {8: "CTA candidates"}, 3: {8: "groups walked", 9: "outputs of warp 0"}}


def trace_api(ctx):
    h = _lib.lib()
    h.orbit_debug_trace.restype = C.c_int
    h.orbit_debug_trace.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    ptr, nbytes = C.c_void_p(), C.c_uint64()
    assert h.orbit_debug_trace(ctx._h, C.byref(ptr), C.byref(nbytes)) == 0
    return h, ptr, nbytes.value


def reset_trace(ctx):
    h, ptr, nbytes = trace_api(ctx)
    z = torch.zeros(nbytes // 8, dtype=torch.int64, device=ctx.device)
    h.orbit_device_copy(ptr, C.c_void_p(z.data_ptr()), nbytes, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()


def read_trace(ctx):
    h, ptr, nbytes = trace_api(ctx)
    out = torch.empty(nbytes // 8, dtype=torch.int64, device=ctx.device)
    h.orbit_device_copy(C.c_void_p(out.data_ptr()), ptr, nbytes, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(16, 1024, 16)


def report(t, label):
    print("== %s" % label)
    frame0, prev_end, out = None, None, []
    for b in range(16):
        blk = t[b]
        if blk[0, 15] < 1000:
            continue
        k = int(blk[0, 15] - 1000)
        name, slots = NAMES[k]
        live = blk[:, 0] > 0
        a = blk[live].astype(np.float64)
        stamps = np.concatenate([a[:, s][a[:, s] > 1e12] for s in slots])
        t0, t1 = a[:, 0].min(), stamps.max()
        if frame0 is None:
            frame0 = t0
        print("  [%2d] %-20s %4d CTAs  starts at %7.2f us%s, first start -> last stamp %6.2f us" % (
            b, name, int(live.sum()), (t0 - frame0) / 1e3, "" if prev_end is None else " (%.2f us after the previous launch's last stamp)" % ((t0 - prev_end) / 1e3),
            (t1 - t0) / 1e3))
        pts = {}
        for s, lab in slots.items():
            v = a[:, s]
            v = v[v > 1e12]
            if len(v):
                print("         %-18s min %7.2f  median %7.2f  max %7.2f us   (%d CTAs)" % (lab, (v.min() - t0) / 1e3, (np.median(v) - t0) / 1e3, (v.max() - t0) / 1e3, len(v)))
                pts[lab] = [(v.min() - t0) / 1e3, (np.median(v) - t0) / 1e3, (v.max() - t0) / 1e3]
        for s, lab in VALUE_SLOTS.get(k, {}).items():
            v = a[:, s]
            print("         %-18s min %7d  median %7d  max %7d" % (lab, v.min(), np.median(v), v.max()))
        out.append({"kernel": name, "start_us": (t0 - frame0) / 1e3, "span_us": (t1 - t0) / 1e3, "points": pts})
        prev_end = t1
    if prev_end is not None:
        print("  frame: first stamp -> last stamp %.2f us" % ((prev_end - frame0) / 1e3))
    return out


def main():
    ctx = Context(0)
    scene, _ = scenes.config_c2()
    view = bench.c2_view(scenes, scene, 0)
    depth = scenes.make_depth(scene, view)
    copies = []
    for i in range(4):
        ds = frame.DeviceScene.upload(ctx, scene)
        vs = frame.ViewState(ctx, ds, (view.width, view.height), name="view%d" % i)
        pf = frame.PreparedFrame(ctx, ds, vs, view, torch.from_numpy(depth).to(ctx.device), name="c%d" % i, main_pass=True)
        pf.launch(); pf.launch()
        copies.append(pf)
    torch.cuda.synchronize()
    res = {}
    for rep in range(2):
        for i in (1, 2, 3, 1, 2, 3):
            copies[i].launch()                 # the other copies' frames push copy 0's inputs out of L2
        torch.cuda.synchronize()
        reset_trace(ctx)
        g = torch.cuda.CUDAGraph()             # captured AFTER the reset: the launches get trace blocks 0..9 in order
        with torch.cuda.graph(g):
            copies[0].launch()
        g.replay()
        torch.cuda.synchronize()
        res["frame_%d" % rep] = report(read_trace(ctx), "C2 steady-state frame as one CUDA graph, run %d" % rep)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
