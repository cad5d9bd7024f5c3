"""GPU: the asset-side producers (csrc/asset_bounds.cu) against the oracle's restatement bit for bit, and a scene whose
inputs come from real geometry through them — LOD chains included — culled two-pass on the GPU and by the oracle."""
import numpy as np
import pytest
import torch

from orbit_b200 import assets, scenes
from orbit_b200 import layouts as L

pytestmark = pytest.mark.gpu


def test_bounds_match_the_oracle_and_feed_the_culling_path(gpu_context, oracle):
    from orbit_b200 import frame
    ctx = gpu_context
    builder, placement = assets.city_from_geometry(0x0B17F4, n_meshes=24, n_entities=400, grid=(20, 20), n0=12)
    vertices, data, meshlets0, infos0, ranges = builder.arrays()
    vertices, data, meshlets, infos, ranges = builder.finish(ctx)
    o_ml, o_mi, skipped = oracle.asset_bounds(vertices, data, meshlets0, infos0, ranges)
    assert skipped == 0 and len(meshlets) > 2000
    assert np.array_equal(meshlets.view(np.uint8), o_ml.view(np.uint8)), "meshlet bounds"
    assert np.array_equal(infos.view(np.uint8), o_mi.view(np.uint8)), "mesh bounds"
    assert (meshlets["cone_cutoff"] < 127).sum() > len(meshlets) // 2
    # ---- the culling path on these inputs: several LODs per mesh, two frames, two-pass + main
    sc = assets.scene_from_assets("geometry_city", placement, meshlets, infos)
    view = scenes.perspective_view((-8.0, 3.0, -8.0), (0.6, -0.05, 0.8), 1280, 720)
    view.lod_base = 6.0                      # LOD switches inside this small city
    depth = scenes.make_depth(sc, view)
    ds = frame.DeviceScene.upload(ctx, sc)
    vs = frame.ViewState(ctx, ds, (view.width, view.height), name="geom")
    d_depth = torch.from_numpy(depth).to(ctx.device)
    hs = oracle.HostScene(sc)
    lods_seen = set()
    for f in range(2):
        g = frame.depth_prepass_culling(ctx, ds, vs, view, d_depth)
        g["main"] = frame.main_pass_culling(ctx, ds, vs, view)
        torch.cuda.synchronize()
        o = oracle.depth_prepass_culling(hs, view, depth)
        o["main"] = oracle.main_pass_culling(hs, view)
        for k in ("early", "late", "main"):
            ghdr, grecs = frame.read_dispatch(g[k][0]); ohdr, orecs = oracle.parse_dispatch(o[k][0])
            gn, gd = frame.read_draws(g[k][1]); on, od = oracle.parse_draws(o[k][1])
            assert ghdr.tolist() == ohdr.tolist() and np.array_equal(grecs.view(np.uint32), orecs.view(np.uint32)), (f, k)
            assert gn == on and np.array_equal(gd.view(np.uint32), od.view(np.uint32)), (f, k)
            for r in orecs:
                mesh = sc.draws["mesh_index"][r["entity_index"]]
                table = sc.mesh_infos["mesh_lods"][mesh]
                lods_seen.add(int(np.searchsorted(table[:int(sc.mesh_infos["lod_count"][mesh]), 0], r["meshlet_offset"], side="right") - 1))
        assert np.array_equal(vs.meshlet_visibility.cpu().numpy().view(np.uint32), hs.meshlet_visibility), f
    assert len(lods_seen) >= 2, lods_seen


def test_degenerate_and_oversized_meshlets_on_the_gpu(gpu_context, oracle):
    from orbit_b200 import _lib
    import ctypes as C
    ctx = gpu_context
    v = np.zeros(4, assets.vertex_dtype)
    v["position"] = [[0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1, 0]]
    data = np.zeros(8, np.uint32)
    data[:4] = [0, 1, 2, 3]
    data[4:].view(np.uint8)[:6] = [0, 1, 2, 0, 0, 0]
    ml = np.zeros(3, L.meshlet_dtype)
    ml[0]["vertex_count"], ml[0]["triangle_count"] = 4, 2
    ml[1]["vertex_count"], ml[1]["triangle_count"] = 4, 200
    ml[2]["vertex_count"], ml[2]["triangle_count"] = 4, 1
    data2 = np.concatenate([data, np.array([0, 1, 3, 2], np.uint32), np.zeros(1, np.uint32)])
    data2[12:].view(np.uint8)[:3] = [0, 1, 2]
    ml[2]["data_offset"] = 8
    ml["bounding_sphere"] = 7.0
    ctx.poll_status()
    d_v, d_d, d_m = ctx.upload(v), ctx.upload(data2), ctx.upload(ml)
    p = lambda t: C.c_void_p(t.data_ptr())
    assert _lib.lib().orbit_meshlet_bounds(ctx._h, p(d_v), 32, p(d_d), p(d_m), 3, None) == 0
    torch.cuda.synchronize()
    out = d_m.cpu().numpy().view(L.meshlet_dtype).reshape(-1)
    ref, _, skipped = oracle.asset_bounds(v, data2, ml, np.zeros(1, L.mesh_info_dtype), np.array([0, 4], np.uint32))
    assert skipped == 1 and np.array_equal(out.view(np.uint8), ref.view(np.uint8))
    code, st = ctx.poll_status()
    assert st.asset_error == 1 and code == _lib.ERR_INVALID_ARGUMENT
