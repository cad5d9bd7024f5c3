"""GPU parity at the FULL sizes of BASELINE configs C4 and C5 (C1-C3 at full size: test_gpu_fullsize.py, test_gpu_pipeline.py).

C4: 50 176 entities / 10 M meshlets — the main view two-pass + MAIN, the four shadow cascades built exactly as
    ShadowRenderer::render_cascaded_shadow builds them (scenes.cascade_views restates shadow_renderer.rs:466-712: lambda 0.8,
    max distance 32, 2048^2, 6 light planes + the back-face-filtered camera planes, LOD range 2.. for cascades 2-3), and clustered
    light assignment over 16x9x24 clusters with 65 536 point lights (+ sky + sun).
C5: 99 856 entities / 20 M meshlets, 8 of the 256 cameras: pass 0 (frustum + cone) for all eight, two-pass + MAIN over two
    frames for two of them.
Lists are compared through hashes of the canonical byte streams (the oracle and the CUDA path emit the same canonical order)."""
import hashlib

import numpy as np
import pytest
import torch

from orbit_b200 import scenes

pytestmark = pytest.mark.gpu


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def _same_pass(oracle, g, o, what):
    from orbit_b200.frame import read_dispatch, read_draws
    ghdr, grecs = read_dispatch(g[0]); ohdr, orecs = oracle.parse_dispatch(o[0])
    gn, gd = read_draws(g[1]); on, od = oracle.parse_draws(o[1])
    assert ghdr.tolist() == ohdr.tolist(), what
    assert _sha(grecs) == _sha(orecs), what
    assert gn == on and _sha(gd) == _sha(od), what
    return int(ohdr[0]), on


def test_c4_full_size_main_view_cascades_and_lights(gpu_context, oracle):
    from orbit_b200 import frame
    from orbit_b200.passes import ClusterSettings, OcclusionCullInfo, compute_clusters
    ctx = gpu_context
    sc, view = scenes.config_c4(1.0)
    assert sc.n_meshlet_instances >= 10_000_000
    depth = scenes.make_depth(sc, view)
    lights = scenes.make_lights(scenes.SEEDS["C4"], 65536, sc.aabb_min, sc.aabb_max)
    ds = frame.DeviceScene.upload(ctx, sc, lights=lights)
    vs = frame.ViewState(ctx, ds, (view.width, view.height), name="c4")
    d_depth = torch.from_numpy(depth).to(ctx.device)
    hs = oracle.HostScene(sc)
    # ---- main view: two frames of early -> Hi-Z -> late -> main
    for f in range(2):
        g = frame.depth_prepass_culling(ctx, ds, vs, view, d_depth)
        g["main"] = frame.main_pass_culling(ctx, ds, vs, view)
        torch.cuda.synchronize()
        o = oracle.depth_prepass_culling(hs, view, depth)
        o["main"] = oracle.main_pass_culling(hs, view)
        for k in ("early", "late", "main"):
            _same_pass(oracle, g[k], o[k], ("main view", f, k))
        assert np.array_equal(vs.meshlet_visibility.cpu().numpy().view(np.uint32), hs.meshlet_visibility), f
        assert np.array_equal(vs.entity_visibility.cpu().numpy().view(np.uint32), hs.entity_visibility), f
    # ---- the four cascades: orthographic, pass 0, 9-11 planes
    cascades = scenes.cascade_views(view)
    assert len(cascades) == 4 and all(6 <= len(c.planes) <= 11 for c in cascades)
    assert [c.lod_range for c in cascades] == [(0, 8), (0, 8), (2, 8), (2, 8)]
    totals = []
    for i, cv in enumerate(cascades):
        g = frame.cull_pass(ctx, "c4_cascade%d" % i, ds, frame.cull_info_for(cv, OcclusionCullInfo("none")))
        torch.cuda.synchronize()
        o = oracle.cull_pass(hs, oracle.gpu_cull_info(cv, "none"))
        totals.append(_same_pass(oracle, g, o, ("cascade", i)))
    assert totals[-1][1] > totals[0][1] > 0              # the cascades grow with distance
    # ---- clustered light assignment, 16 x 9 x 24 clusters (tile 120 px), 65 538 lights
    st = ClusterSettings(screen_resolution=(1920, 1080), z_slice_count=24, tile_size_px=120)
    info, params = compute_clusters(ctx, st, view.view, view.projection_matrix, view.near, d_depth, ds.scene)
    torch.cuda.synchronize()
    ref = oracle.light_cluster(params, depth, lights)
    n = 16 * 9 * 24
    na, total = int(ref["unique"][3]), int(ref["index"][0])
    unique = info.unique_cluster_buffer[:16 + 4 * n].cpu().numpy().view(np.uint32)
    assert unique[:4].tolist() == ref["unique"][:4].tolist() and np.array_equal(unique[4:4 + na], ref["unique"][4:4 + na])
    assert np.array_equal(info.tile_depth_slice_mask[:4 * 16 * 9].cpu().numpy().view(np.uint32), ref["masks"])
    assert np.array_equal(info.cluster_depth_bounds[:8 * n].cpu().numpy().view(np.uint32), ref["bounds"])
    assert np.array_equal(info.light_offset_image[:8 * n].cpu().numpy().view(np.uint32), ref["image"])
    idx = info.light_index_buffer[:4 + 4 * total].cpu().numpy().view(np.uint32)
    assert int(idx[0]) == total and np.array_equal(idx[1:], ref["index"][1:1 + total])
    assert na > 50 and total > 0


def test_c5_full_size_eight_cameras(gpu_context, oracle):
    from orbit_b200 import frame
    from orbit_b200.passes import OcclusionCullInfo
    ctx = gpu_context
    sc, views = scenes.config_c5(1.0, n_views=256)
    assert sc.n_meshlet_instances >= 19_000_000 and len(views) == 256
    ds = frame.DeviceScene.upload(ctx, sc)
    hs = oracle.HostScene(sc)
    chosen = [0, 37, 74, 111, 148, 185, 222, 255]
    survivors = []
    for v in chosen:            # pass 0: frustum + cone
        g = frame.cull_pass(ctx, "c5_pass0", ds, frame.cull_info_for(views[v], OcclusionCullInfo("none")))
        torch.cuda.synchronize()
        o = oracle.cull_pass(hs, oracle.gpu_cull_info(views[v], "none"))
        survivors.append(_same_pass(oracle, g, o, ("pass 0", v))[1])
    assert max(survivors) > 1_000_000
    for v in chosen[:2]:        # two-pass occlusion + MAIN, two frames, own visibility state per camera
        depth = scenes.make_depth(sc, views[v])
        d_depth = torch.from_numpy(depth).to(ctx.device)
        vs = frame.ViewState(ctx, ds, (views[v].width, views[v].height), name="c5_v%d" % v)
        hs.reset_visibility()
        for f in range(2):
            g = frame.depth_prepass_culling(ctx, ds, vs, views[v], d_depth)
            g["main"] = frame.main_pass_culling(ctx, ds, vs, views[v])
            torch.cuda.synchronize()
            o = oracle.depth_prepass_culling(hs, views[v], depth)
            o["main"] = oracle.main_pass_culling(hs, views[v])
            for k in ("early", "late", "main"):
                _same_pass(oracle, g[k], o[k], ("two-pass", v, f, k))
            assert np.array_equal(vs.meshlet_visibility.cpu().numpy().view(np.uint32), hs.meshlet_visibility), (v, f)
