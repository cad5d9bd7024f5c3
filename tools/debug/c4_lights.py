"""Development aid: C4 clustered light assignment, both culling paths against the oracle, first differing cluster printed."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle_ref as O
from orbit_b200 import frame, scenes
from orbit_b200.passes import ClusterSettings, Context, compute_clusters

sc, view = scenes.config_c4(float(os.environ.get("SCALE", "1.0")))
depth = scenes.make_depth(sc, view)
lights = scenes.make_lights(scenes.SEEDS["C4"], 65536, sc.aabb_min, sc.aabb_max)
st = ClusterSettings(screen_resolution=(1920, 1080), z_slice_count=24, tile_size_px=120)
n = 16 * 9 * 24
ref = None
for budget in ("256", "0"):
    os.environ["ORBIT_LIGHT_HITS_BUDGET_MB"] = budget
    ctx = Context(0)
    ds = frame.DeviceScene.upload(ctx, sc, lights=lights)
    d_depth = torch.from_numpy(depth).to(ctx.device)
    for rep in range(int(os.environ.get("REPS", "2"))):
        if rep:
            info.light_index_buffer.fill_(0xEE)
        info, params = compute_clusters(ctx, st, view.view, view.projection_matrix, view.near, d_depth, ds.scene)
        torch.cuda.synchronize()
        if ref is None:
            ref = O.light_cluster(params, depth, lights)
        na, total = int(ref["unique"][3]), int(ref["index"][0])
        image = info.light_offset_image[:8 * n].cpu().numpy().view(np.uint32).reshape(n, 2)
        rimage = ref["image"].reshape(n, 2)
        idx = info.light_index_buffer[:4 + 4 * total].cpu().numpy().view(np.uint32)
        same_img = np.array_equal(image, rimage)
        same_idx = int(idx[0]) == total and np.array_equal(idx[1:], ref["index"][1:1 + total])
        print("budget", budget, "rep", rep, "active", na, "total", total, "gpu total", int(idx[0]), "image", same_img, "index", same_idx)
        if not same_idx:
            bad = 0
            for t, c in enumerate(ref["unique"][4:4 + na]):
                o, k = rimage[c]
                a, b = idx[1 + o:1 + o + k], ref["index"][1 + o:1 + o + k]
                if not np.array_equal(a, b):
                    print("  t", t, "cluster", int(c), "off", int(o), "count", int(k), "\n   gpu", a.tolist(), "\n   ref", b.tolist())
                    bad += 1
                    if bad == 12: break
    ctx.close()
