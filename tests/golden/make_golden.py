#!/usr/bin/env python
"""Generates the committed golden fixtures from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`). The reference has no golden vectors for this path and cannot be run in
the build image, so these pin the ORACLE (and, through the GPU tests, the CUDA path) against regressions; they
are not outputs of the reference itself (DESIGN.md "parity status").

Fixture = digests + small excerpts of every buffer of two frames of config C1 (inputs are regenerated
deterministically by orbit_b200.scenes; their digests are stored too so a platform that generates different
input bytes is detected instead of mis-reported as a parity failure)."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402

import oracle_ref as O  # noqa: E402
from orbit_b200 import scenes  # noqa: E402
from orbit_b200.passes import ClusterSettings  # noqa: E402
from orbit_b200 import layouts as L  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def c1_case():
    sc, view = scenes.config_c1()
    depth = scenes.make_depth(sc, view)
    return sc, view, depth


def input_digests(sc, depth):
    return {"meshlets": sha(sc.meshlets), "mesh_infos": sha(sc.mesh_infos), "entities": sha(sc.entities),
            "entity_draws": sha(sc.entity_draws), "materials": sha(sc.materials), "depth": sha(depth)}


def run_frames(sc, view, depth, frames=2):
    hs = O.HostScene(sc)
    out = {}
    for f in range(frames):
        st = O.Stats()
        o = O.depth_prepass_culling(hs, view, depth, stats=st)
        m = O.main_pass_culling(hs, view, stats=st)
        for k, pair in (("early", o["early"]), ("late", o["late"]), ("main", m)):
            hdr, recs = O.parse_dispatch(pair[0]); n, draws = O.parse_draws(pair[1])
            out["frame%d_%s" % (f, k)] = {"records": int(hdr[0]), "header": [int(v) for v in hdr], "draws": int(n),
                                           "records_sha": sha(recs), "draws_sha": sha(draws),
                                           "first_draws": draws[:8].view(np.uint32).reshape(-1, 7).tolist(),
                                           "last_draw": draws[-1:].view(np.uint32).reshape(-1, 7).tolist()}
        out["frame%d_state" % f] = {"entity_visibility_sha": sha(hs.entity_visibility), "meshlet_visibility_sha": sha(hs.meshlet_visibility),
                                    "hiz_sha": sha(hs.hiz_texels), "hiz_top_levels": hs.hiz_texels[-21:].view(np.uint32).tolist(),
                                    "near_threshold": {k: v for k, v in st.as_dict().items() if k.startswith("near_")}}
    return out


def cluster_case():
    sc, view = scenes.config_c4(scale=0.01)
    depth = scenes.make_depth(sc, view)
    lights = scenes.make_lights(scenes.SEEDS["C4"], 2048, sc.aabb_min, sc.aabb_max)
    st = ClusterSettings(screen_resolution=(1920, 1080), z_slice_count=24, tile_size_px=120)
    cx, cy, cz = st.cluster_counts()
    p = L.ClusterParams()
    p.info.world_to_view_matrix.set(view.view)
    p.info.screen_to_view_matrix.set(np.linalg.inv(view.projection_matrix))
    p.info.cluster_count[0], p.info.cluster_count[1], p.info.cluster_count[2] = cx, cy, cz
    p.info.tile_size_px = 120
    p.info.screen_size[0], p.info.screen_size[1] = 1920, 1080
    p.info.z_near, p.info.z_far = view.near, st.far_plane
    p.info.global_light_count = len(lights)
    p.z_scale, p.z_bias = st.cluster_grid_info(view.near)
    return sc, view, depth, lights, p


def run_clusters(depth, lights, p):
    r = O.light_cluster(p, depth, lights)
    na, total = int(r["unique"][3]), int(r["index"][0])
    return {"active": na, "total_indices": total, "masks_sha": sha(r["masks"]), "bounds_sha": sha(r["bounds"]),
            "unique_sha": sha(r["unique"][:4 + na]), "image_sha": sha(r["image"]), "index_sha": sha(r["index"][:1 + total]),
            "first_active": r["unique"][4:12].tolist()}


def main():
    sc, view, depth = c1_case()
    fixture = {"c1_inputs": input_digests(sc, depth), "c1": run_frames(sc, view, depth)}
    sc4, view4, depth4, lights, p = cluster_case()
    fixture["clusters_inputs"] = {"depth": sha(depth4), "lights": sha(lights), "params": sha(np.frombuffer(bytes(p), np.uint8))}
    fixture["clusters"] = run_clusters(depth4, lights, p)
    with open(os.path.join(HERE, "c1_and_clusters.json"), "w") as f:
        json.dump(fixture, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "c1_and_clusters.json"))


if __name__ == "__main__":
    main()
