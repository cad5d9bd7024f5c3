"""GPU parity of clustered light assignment (mark_active / compaction / light_culling) against the oracle."""
import numpy as np
import pytest
import torch

from orbit_b200 import layouts as L
from orbit_b200 import scenes

pytestmark = pytest.mark.gpu


def _run(ctx, oracle, sc, view, depth, lights, settings):
    from orbit_b200 import frame
    from orbit_b200.passes import compute_clusters
    ds = frame.DeviceScene.upload(ctx, sc, lights=lights)
    d_depth = torch.from_numpy(depth).to(ctx.device)
    info, params = compute_clusters(ctx, settings, view.view, view.projection_matrix, view.near, d_depth, ds.scene)
    torch.cuda.synchronize()
    ref = oracle.light_cluster(params, depth, lights)
    cx, cy, cz = settings.cluster_counts()
    n = cx * cy * cz
    masks = info.tile_depth_slice_mask[:4 * cx * cy].cpu().numpy().view(np.uint32)
    assert np.array_equal(masks, ref["masks"]), "tile masks"
    bounds = info.cluster_depth_bounds[:8 * n].cpu().numpy().view(np.uint32)
    assert np.array_equal(bounds, ref["bounds"]), "depth bounds"
    unique = info.unique_cluster_buffer[:16 + 4 * n].cpu().numpy().view(np.uint32)
    na = int(ref["unique"][3])
    assert unique[:4].tolist() == ref["unique"][:4].tolist()
    assert np.array_equal(unique[4:4 + na], ref["unique"][4:4 + na]), "compacted cluster ids"
    image = info.light_offset_image[:8 * n].cpu().numpy().view(np.uint32)
    assert np.array_equal(image, ref["image"]), "(offset,count) image"
    total = int(ref["index"][0])
    index = info.light_index_buffer[:4 + 4 * total].cpu().numpy().view(np.uint32)
    assert int(index[0]) == total
    assert np.array_equal(index[1:], ref["index"][1:1 + total]), "light index list"
    return na, total


def test_clusters_c4_reduced(gpu_context, oracle):
    from orbit_b200.passes import ClusterSettings
    sc, view = scenes.config_c4(scale=0.02)
    depth = scenes.make_depth(sc, view)
    lights = scenes.make_lights(scenes.SEEDS["C4"], 4096, sc.aabb_min, sc.aabb_max)
    st = ClusterSettings(screen_resolution=(1920, 1080), z_slice_count=24, tile_size_px=120)
    na, total = _run(gpu_context, oracle, sc, view, depth, lights, st)
    assert na > 0 and total > 0


def test_clusters_default_settings(gpu_context, oracle):
    """Reference defaults: 8 px tiles, 32 slices (cluster.rs:23-33) on a small target; hits the 256 cap."""
    from orbit_b200.passes import ClusterSettings
    sc, view = scenes.config_c4(scale=0.005)
    view = scenes.perspective_view((-6.0, 3.0, -6.0), (0.6, 0.0, 0.8), 320, 180)
    depth = scenes.make_depth(sc, view)
    lights = scenes.make_lights(scenes.SEEDS["C4"] + 1, 3000, sc.aabb_min * 0.3, sc.aabb_max * 0.3, intensity=(20.0, 60.0))
    st = ClusterSettings(screen_resolution=(320, 180))
    na, total = _run(gpu_context, oracle, sc, view, depth, lights, st)
    assert na > 0 and total > 0


def test_clusters_no_lights_and_sky_only(gpu_context, oracle):
    from orbit_b200.passes import ClusterSettings
    sc, view = scenes.config_c4(scale=0.005)
    view = scenes.perspective_view((-6.0, 3.0, -6.0), (0.6, 0.0, 0.8), 256, 128)
    depth = np.zeros((128, 256), np.float32)   # all sky: z = inf, slice saturates, nothing active
    lights = scenes.make_lights(1, 16, sc.aabb_min, sc.aabb_max)
    st = ClusterSettings(screen_resolution=(256, 128))
    na, total = _run(gpu_context, oracle, sc, view, depth, lights, st)
    assert na == 0 and total == 0


def test_clusters_cta_per_cluster_fallback(oracle, monkeypatch):
    """Grids whose (cluster, light) bit matrix exceeds the scratch budget take the CTA-per-cluster kernel; forced here with a
    zero budget on a context of its own — same lists as the light-parallel path and the oracle."""
    from orbit_b200.passes import ClusterSettings, Context
    monkeypatch.setenv("ORBIT_LIGHT_HITS_BUDGET_MB", "0")
    ctx = Context(0)
    monkeypatch.delenv("ORBIT_LIGHT_HITS_BUDGET_MB")
    sc, view = scenes.config_c4(scale=0.02)
    depth = scenes.make_depth(sc, view)
    lights = scenes.make_lights(scenes.SEEDS["C4"], 4096, sc.aabb_min, sc.aabb_max)
    st = ClusterSettings(screen_resolution=(1920, 1080), z_slice_count=24, tile_size_px=120)
    na, total = _run(ctx, oracle, sc, view, depth, lights, st)
    assert na > 0 and total > 0
    ctx.close()
