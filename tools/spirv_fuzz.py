"""Randomised differential test: the reference's shipped entity_cull / meshlet_cull / depth_reduce SPIR-V (interpreted by
oracle/spirv_vm) against the C++ oracle on many small random scenes, cameras, LOD settings, plane sets, passes and
visibility states. Build-container tool (needs /root/reference). Prints one line per mismatch and a summary.
    python tools/spirv_fuzz.py [n_cases] [seed]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "spirv_vm")):
    sys.path.insert(0, p)
import oracle_ref as O  # noqa: E402
import reference_passes as R  # noqa: E402
import spirv_cases as S  # noqa: E402
from orbit_b200 import layouts as L, scenes  # noqa: E402


def random_case(rng):
    n_ent = int(rng.integers(3, 40))
    n_lods = int(rng.integers(1, 5))
    lods = sorted((int(rng.integers(1, 140)) for _ in range(n_lods)), reverse=True)
    if rng.random() < 0.5:
        g = int(np.ceil(np.sqrt(n_ent)))
        sc = scenes.make_scene("fz", int(rng.integers(1, 1 << 30)), n_ent, max(1, n_ent // int(rng.integers(1, 4))), lods, layout="city",
                               grid=(g, g), pitch=float(rng.uniform(4, 20)), instanced=bool(rng.random() < 0.5))
    else:
        g = int(np.ceil(n_ent ** (1 / 3)))
        sc = scenes.make_scene("fz", int(rng.integers(1, 1 << 30)), n_ent, max(1, n_ent // 2), lods, layout="lattice3d", grid=(g, g, g + 1),
                               pitch=float(rng.uniform(4, 20)))
    centre = (sc.aabb_min + sc.aabb_max) / 2
    ext = float(np.linalg.norm(sc.aabb_max - sc.aabb_min))
    d = rng.normal(size=3); d[1] *= 0.3; d /= np.linalg.norm(d)
    eye = centre - d * rng.uniform(0.05, 0.9) * ext + rng.normal(size=3) * 2.0
    w, h = int(rng.integers(24, 200)), int(rng.integers(16, 120))
    if rng.random() < 0.7:
        view = scenes.perspective_view(tuple(eye), tuple(d), w, h, fov_deg=float(rng.uniform(30, 110)), near=float(10 ** rng.uniform(-2.5, 0)))
    else:
        view = scenes.orthographic_view(eye, d, w, h, half_width=float(rng.uniform(5, 0.7 * ext + 6)), near=float(rng.uniform(-20, 1)), far=float(rng.uniform(ext, 3 * ext + 10)))
    if rng.random() < 0.4:     # extra planes, up to 12 in total
        extra = []
        for _ in range(int(rng.integers(1, 12 - len(view.planes) + 1))):
            nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
            extra.append([nrm[0], nrm[1], nrm[2], float(rng.uniform(0, ext))])
        view.planes = np.vstack([view.planes, np.array(extra)])
    if rng.random() < 0.2:
        view.planes = view.planes[:0]                       # no frustum culling at all
    lo = int(rng.integers(0, 4)); view.lod_range = (lo, lo + int(rng.integers(1, 6)))
    view.lod_base, view.lod_step = float(10 ** rng.uniform(-1, 2)), float(rng.uniform(1.05, 3.0))
    view.lod_target_view = tuple(rng.normal(size=3) * 3.0) if rng.random() < 0.3 else (0.0, 0.0, 0.0)
    if rng.random() < 0.3:      # non-affine model matrices
        m = sc.entities["model_matrix"]
        m[:, 3, 3] = rng.uniform(0.6, 1.6, len(m)).astype(np.float32)
        m[:, 0, 3] = rng.uniform(-0.02, 0.02, len(m)).astype(np.float32)
    depth = scenes.make_depth(sc, view) if rng.random() < 0.7 else rng.random((h, w), dtype=np.float32) ** 4
    kind = ["none", "read", "write"][int(rng.integers(0, 3))]
    mocc = bool(rng.random() < 0.7)
    opts = {}
    if rng.random() < 0.4:
        opts = {"alpha_filter": int(rng.integers(1, 8)), "noskip": int(rng.integers(0, 8))}
    return sc, view, depth, kind, mocc, opts


def cluster_fuzz(n, seed, log2f):
    """mark_active / active_cluster_compaction / light_culling (shipped SPIR-V) vs the oracle on random grids and lights."""
    rng = np.random.default_rng(seed)
    bad = 0
    t0 = time.time()
    lists = 0
    for case in range(n):
        w, h = int(rng.integers(16, 120)), int(rng.integers(12, 80))
        tile = int(rng.choice([4, 8, 10, 16, 24, 33]))
        cz = int(rng.integers(1, 33))
        near, far = float(10 ** rng.uniform(-2, 0)), float(10 ** rng.uniform(1, 2.7))
        sc, _ = scenes.config_c1(scale=float(rng.uniform(0.02, 0.08)))
        d = rng.normal(size=3); d[1] *= 0.2; d /= np.linalg.norm(d)
        view = scenes.perspective_view((-6.0 + rng.normal() * 3, 2.0 + abs(rng.normal()) * 3, -6.0 + rng.normal() * 3), tuple(d), w, h,
                                       fov_deg=float(rng.uniform(40, 100)), near=near)
        depth = scenes.make_depth(sc, view) if rng.random() < 0.6 else (rng.random((h, w), dtype=np.float32) ** 3) * (rng.random((h, w)) < 0.8)
        depth = np.ascontiguousarray(depth, np.float32)
        lights = scenes.make_lights(int(rng.integers(1, 1 << 20)), int(rng.integers(1, 600)), sc.aabb_min, sc.aabb_max)
        if rng.random() < 0.3:
            lights["outer_radius"] *= np.float32(rng.uniform(2, 30))
        p = L.ClusterParams()
        cx, cy = -(-w // tile), -(-h // tile)
        p.info.world_to_view_matrix.set(view.view)
        p.info.screen_to_view_matrix.set(np.linalg.inv(np.asarray(view.projection_matrix, np.float64)))
        p.info.cluster_count[0], p.info.cluster_count[1], p.info.cluster_count[2] = cx, cy, cz
        p.info.tile_size_px = tile
        p.info.screen_size[0], p.info.screen_size[1] = w, h
        p.info.z_near, p.info.z_far = near, far
        p.info.global_light_count = len(lights)
        p.z_scale, p.z_bias = scenes.cluster_grid_info(near, far, cz)
        a = S.canon_clusters(R.light_cluster(p, depth, lights, log2f))
        b = S.canon_clusters(O.light_cluster(p, depth, lights))
        ok = a["header"] == b["header"] and a["total"] == b["total"] and all(np.array_equal(a[k], b[k]) for k in ("masks", "bounds", "active", "counts", "lists"))
        lists += int(b["total"])
        if not ok:
            bad += 1
            print(json.dumps({"cluster_case": case, "seed": seed, "size": [w, h], "tile": tile, "cz": cz, "header": [a["header"], b["header"]], "total": [a["total"], b["total"]]}), flush=True)
    print(json.dumps({"cluster_cases": n, "seed": seed, "mismatching_cases": bad, "light_indices_compared": lists, "seconds": round(time.time() - t0)}))


def task_fuzz(n, seed, log2f):
    """The shipped task shaders' payloads (count, entity, offset, ascending lane indices) vs the oracle's payload output."""
    rng = np.random.default_rng(seed)
    bad = tasks = 0
    t0 = time.time()
    for case in range(n):
        sc, view, depth, kind, mocc, opts = random_case(rng)
        shader = S.TASK_SHADERS[int(rng.integers(0, 3))]
        if shader.startswith("shadow") and kind == "write":
            # Pass 2 is never requested from shadow.task (shadow_renderer.rs:693-707 culls with pass 0). Its SHIPPED binary
            # also disagrees there with its own source and with the other two task shaders: with a meshlet visibility buffer
            # it emits the meshlets that were visible last frame as well (found by this sweep: 41 vs 20 tasks etc.), and
            # without one it indexes descriptor 0xFFFFFFFF. Not reproduced, not compared.
            kind = "none"
        hs = O.HostScene(sc)
        hs.entity_visibility[:] = rng.integers(0, 1 << 32, len(hs.entity_visibility), dtype=np.uint64).astype(np.uint32)
        hs.meshlet_visibility[:] = rng.integers(0, 1 << 32, len(hs.meshlet_visibility), dtype=np.uint64).astype(np.uint32)
        ev, mv = hs.entity_visibility.copy(), hs.meshlet_visibility.copy()
        levels = None
        if kind == "write":
            hs.update_pyramid(depth)
            levels = R.hiz_build(depth, O.hiz_geometry(view.width, view.height), log2f)
        g = O.gpu_cull_info(view, kind, mocc)
        if "alpha_filter" in opts: g.alpha_mode_flags = opts["alpha_filter"]
        if "noskip" in opts and kind == "write": g.noskip_alpha_mode = opts["noskip"]
        disp = R.entity_cull(sc, g, ev, mv, levels, sc.n_records_lod0, log2f)
        res = R.task_shader(sc, g, ev, mv, levels, disp, log2f, shader)
        o = O.cull_pass(hs, g, task_payloads=True)
        nrec = int(o[0][:4].view(np.uint32)[0])
        a = S.canon_payloads([(c, p[0], p[1], p[2]) for (_, c, p) in res])
        b = S.canon_payload_buffer(o[2], nrec)
        ascending = all(list(p[2][:c]) == sorted(p[2][:c]) for (_, c, p) in res)
        tasks += int(b["task_count"].sum())
        if not (np.array_equal(a, b) and ascending):
            bad += 1
            print(json.dumps({"task_case": case, "seed": seed, "shader": shader, "kind": kind, "mocc": mocc, "records": [len(a), len(b)],
                              "tasks": [int(a["task_count"].sum()), int(b["task_count"].sum())], "ascending": ascending}), flush=True)
    print(json.dumps({"task_cases": n, "seed": seed, "mismatching_cases": bad, "tasks_compared": tasks, "seconds": round(time.time() - t0)}))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "tasks":
        O.build()
        task_fuzz(int(sys.argv[2]) if len(sys.argv) > 2 else 40, int(sys.argv[3]) if len(sys.argv) > 3 else 1, lambda x: np.float32(O.log2f(float(x))))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "clusters":
        O.build()
        cluster_fuzz(int(sys.argv[2]) if len(sys.argv) > 2 else 20, int(sys.argv[3]) if len(sys.argv) > 3 else 1, lambda x: np.float32(O.log2f(float(x))))
        return
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    O.build()
    log2f = lambda x: np.float32(O.log2f(float(x)))
    rng = np.random.default_rng(seed)
    bad = 0
    tot_r = tot_d = 0
    t0 = time.time()
    for case in range(n):
        sc, view, depth, kind, mocc, opts = random_case(rng)
        hs = O.HostScene(sc)
        hs.entity_visibility[:] = rng.integers(0, 1 << 32, len(hs.entity_visibility), dtype=np.uint64).astype(np.uint32)
        hs.meshlet_visibility[:] = rng.integers(0, 1 << 32, len(hs.meshlet_visibility), dtype=np.uint64).astype(np.uint32)
        ev, mv = hs.entity_visibility.copy(), hs.meshlet_visibility.copy()
        levels = None
        if kind == "write":
            hs.update_pyramid(depth)
            levels = R.hiz_build(depth, O.hiz_geometry(view.width, view.height), log2f)
            if not np.array_equal(np.concatenate([l.reshape(-1) for l in levels]).view(np.uint32), hs.hiz_texels.view(np.uint32)):
                print(json.dumps({"case": case, "what": "hiz", "size": [view.width, view.height]})); bad += 1
        g = O.gpu_cull_info(view, kind, mocc)
        if "alpha_filter" in opts: g.alpha_mode_flags = opts["alpha_filter"]
        if "noskip" in opts and kind == "write": g.noskip_alpha_mode = opts["noskip"]
        disp = R.entity_cull(sc, g, ev, mv, levels, sc.n_records_lod0, log2f)
        draws = R.meshlet_cull(sc, g, ev, mv, levels, disp, sc.n_meshlet_instances, log2f)
        o = O.cull_pass(hs, g)
        vh, vr = S.canon_records(disp); oh, orr = S.canon_records(o[0])
        vn, vd = S.canon_draws(draws); on, od = S.canon_draws(o[1])
        ok = vh == oh and np.array_equal(vr, orr) and vn == on and np.array_equal(vd, od) and np.array_equal(ev, hs.entity_visibility)
        # meshlet visibility: the shaders only write words of dispatched records; both sides start from the same random words
        ok = ok and np.array_equal(mv, hs.meshlet_visibility)
        tot_r += len(orr); tot_d += on
        if not ok:
            bad += 1
            print(json.dumps({"case": case, "seed": seed, "kind": kind, "mocc": mocc, "proj": int(view.projection_type), "planes": len(view.planes),
                              "records": [vh[0], oh[0]], "draws": [vn, on], "ev_equal": bool(np.array_equal(ev, hs.entity_visibility)),
                              "mv_equal": bool(np.array_equal(mv, hs.meshlet_visibility)), "lods": view.lod_range, "opts": opts}), flush=True)
    print(json.dumps({"cases": n, "seed": seed, "mismatching_cases": bad, "records_compared": tot_r, "draws_compared": tot_d, "seconds": round(time.time() - t0)}))


if __name__ == "__main__":
    main()
