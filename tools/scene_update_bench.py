"""Device time of orbit_scene_update (scene.rs:404-492 on the GPU) at the C2 and C3 entity counts, against its HBM
roofline and against the oracle's serial loop on one host core. Graph of 8 launches over 4 rotating copies (inputs out
of L2 at C3 scale), median of 7 replays."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from orbit_b200 import frame, layouts as L, scenes  # noqa: E402
from orbit_b200.passes import Context  # noqa: E402
from orbit_b200.scene import SceneData  # noqa: E402


def main():
    import oracle_ref as O
    ctx = Context(0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6548.2) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6548.2
    for name, cfg in (("C2", scenes.config_c2), ("C3", scenes.config_c3)):
        sc, _ = cfg()
        n = sc.n_entities
        ds = frame.DeviceScene.upload(ctx, sc)
        copies = []
        for _ in range(4):
            sd = SceneData(ctx, n)
            sd.set_entities(sc.transforms, sc.draws["mesh_index"])
            sd.update_scene(ds.assets)          # first frame: allocates the visibility ranges
            copies.append(sd)
        torch.cuda.synchronize()
        out = {"config": name, "entities": n}
        for label, reset in (("steady_us", False), ("first_frame_us", True)):
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(8):
                        sd = copies[i % 4]
                        if reset:               # every replay allocates all ranges again (memsets are part of the graph: ~1 us each)
                            sd.visibility_offsets.fill_(-1); sd.visibility_cursor.zero_()
                        sd.update_scene(ds.assets)
                ts = []
                for _ in range(8):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); g.replay(); b.record()
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b) * 1e3 / 8)
            out[label] = round(float(np.median(ts[1:])), 2)
        bytes_alg = n * (48 + 4 + 4 + 128 + 12)
        out["algorithmic_MB"] = round(bytes_alg / 1e6, 2)
        out["achieved_GBs"] = round(bytes_alg / (out["steady_us"] * 1e-6) / 1e9, 1)
        out["frac_of_measured_hbm"] = round(out["achieved_GBs"] / peak, 3)
        out["Mentities_per_s"] = round(n / out["steady_us"], 1)
        vo = np.full(n, L.NO_VISIBILITY_RANGE, np.uint32); cur = np.zeros(1, np.uint32)
        O.scene_update(sc.transforms, sc.draws["mesh_index"].copy(), vo, sc.mesh_infos, cur)
        t0 = time.perf_counter()
        for _ in range(3):
            O.scene_update(sc.transforms, sc.draws["mesh_index"].copy(), vo, sc.mesh_infos, cur)
        out["oracle_1_core_us"] = round((time.perf_counter() - t0) / 3 * 1e6, 1)
        print(json.dumps(out), flush=True)
        del copies
    ctx.close()


if __name__ == "__main__":
    main()
