"""CPU: the host-side mirror of the reference pass interface (CullInfo::to_gpu, ClusterSettings, sharding)."""
import ctypes as C

import numpy as np
import pytest

from orbit_b200 import layouts as L
from orbit_b200 import scenes
from orbit_b200.passes import AlphaModeFlags, ClusterSettings, CullInfo, OcclusionCullInfo, Projection


def test_cull_info_to_gpu_perspective_write():
    """draw_gen.rs:121-203: p00 = f/aspect, p11 = f, z_near only in VisibilityWrite; lod range is [start, end-1]."""
    view = scenes.perspective_view((1, 2, 3), (0, 0, -1), 1920, 1080)
    marker = object()
    ci = CullInfo(view.view, view.planes, Projection.perspective(view.fov, 0.01),
                  OcclusionCullInfo("write", marker, marker, marker, noskip_alphamode=0, aspect_ratio=1920 / 1080),
                  lod_range=(0, 8), lod_base=16.0, lod_step=2.0)
    g = ci.to_gpu()
    raw = np.frombuffer(bytes(g), np.uint8)
    assert raw.size == 400
    f32 = raw.view(np.float32); u32 = raw.view(np.uint32)
    assert np.allclose(f32[:16].reshape(4, 4).T, view.view.astype(np.float32))      # column-major view matrix @0
    assert np.all(f32[16:32] == 0)                                                   # reprojection matrix: always zero
    assert np.allclose(f32[32:52].reshape(5, 4), view.planes.astype(np.float32)) and np.all(f32[52:80] == 0)
    assert u32[80] == 5 and u32[81] == (AlphaModeFlags.OPAQUE | AlphaModeFlags.MASKED) and u32[82] == 0 and u32[83] == 2
    assert u32[85] != L.NO_BUFFER and u32[88] == 0
    assert abs(f32[89] - 1.0 / (1920 / 1080)) < 1e-6 and abs(f32[90] - 1.0) < 1e-6 and f32[91] == np.float32(0.01) and f32[92] == 0.0
    assert f32[93] == 16.0 and f32[94] == 2.0 and u32[95] == 0 and u32[99] == 7


def test_cull_info_to_gpu_read_and_none_leave_projection_zero():
    """In passes 0/1 to_gpu leaves p00/p11/z_near/z_far at 0 (draw_gen.rs:170-200 fills them only for VisibilityWrite)."""
    view = scenes.perspective_view((0, 0, 0), (0, 0, -1), 1280, 720)
    for kind, pidx in (("none", 0), ("read", 1)):
        oc = OcclusionCullInfo(kind, object() if kind != "none" else None, None)
        g = CullInfo(view.view, view.planes, Projection.perspective(view.fov, 0.01), oc).to_gpu()
        assert g.occlusion_pass == pidx and g.p00_or_width_recip_x2 == 0.0 and g.z_near == 0.0
        assert g.meshlet_visibility_buffer == L.NO_BUFFER


def test_cull_info_to_gpu_orthographic():
    g = CullInfo(np.eye(4), [], Projection.orthographic(50.0, -10.0, 300.0),
                 OcclusionCullInfo("write", object(), object(), object(), 0, aspect_ratio=2.0)).to_gpu()
    assert g.projection_type == 1 and g.cull_plane_count == 0
    assert abs(g.p00_or_width_recip_x2 - 2.0 / 100.0) < 1e-9 and abs(g.p11_or_height_recip_x2 - 2.0 / 50.0) < 1e-9
    assert g.z_near == -10.0 and g.z_far == 300.0


def test_more_than_12_planes_is_rejected():
    with pytest.raises(AssertionError):
        CullInfo(np.eye(4), np.zeros((13, 4)), Projection.perspective(1.0, 0.01)).to_gpu()


def test_cluster_settings_match_reference_defaults():
    st = ClusterSettings()
    st.set_resolution((1920, 1080))
    assert st.tile_px_size() == 8 and st.tile_counts() == [240, 135] and st.linear_cluster_count() == 240 * 135 * 32
    zs, zb = st.cluster_grid_info(0.01)
    # slice(z) = log2(z)*scale + bias maps near -> 0 and far -> slice_count
    assert abs(np.log2(0.01) * zs + zb) < 1e-3 and abs(np.log2(200.0) * zs + zb - 32) < 1e-3
    c4 = ClusterSettings(screen_resolution=(1920, 1080), z_slice_count=24, tile_size_px=120)
    assert c4.cluster_counts() == [16, 9, 24]


def test_scene_generator_is_deterministic_and_well_formed():
    a, va = scenes.config_c1(scale=0.2)
    b, vb = scenes.config_c1(scale=0.2)
    for k in ("meshlets", "mesh_infos", "entities", "entity_draws", "materials"):
        assert np.array_equal(getattr(a, k).view(np.uint8), getattr(b, k).view(np.uint8)), k
    assert a.entity_draws[:4].view(np.uint32)[0] == a.n_entities
    d = a.draws
    words = (100 + 31) // 32
    assert np.array_equal(d["visibility_offset"], np.arange(a.n_entities, dtype=np.uint32) * words)   # scene.rs:422-431
    lods = a.mesh_infos["mesh_lods"]
    assert np.all(lods[:, 0, 0] + lods[:, 0, 1] <= len(a.meshlets))
    m = a.entities["model_matrix"]
    assert np.all(m[:, :, 3] == np.array([0, 0, 0, 1], np.float32))   # affine: bottom row (0,0,0,1) -> p.w == 1 exactly
    alpha = a.materials.view(np.uint32).reshape(16, 20)[:, 16]
    assert set(alpha.tolist()) <= {0, 1, 2}


def test_partition_draws_balanced_and_aligned():
    from orbit_b200.multi_gpu import partition_draws, views_for_rank
    rng = np.random.default_rng(0)
    counts = rng.integers(1, 400, size=10007)
    for world in (1, 2, 4, 8):
        parts = partition_draws(counts, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(counts)
        for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
            assert e0 == b1 and b1 % 32 == 0
        sums = [counts[b:e].sum() for b, e in parts]
        assert max(sums) - min(sums) <= 2 * 32 * 400
    assert partition_draws([5, 5], 4)[-1][1] == 2
    assert sorted(sum((views_for_rank(256, r, 8) for r in range(8)), [])) == list(range(256))


def test_partition_draws_with_weights():
    """Rank 0 of a sharded view takes a lighter share (it also builds the pyramid and emits both lists): the ranges stay
    contiguous, 32-aligned and complete, and the sums follow the weights."""
    from orbit_b200.multi_gpu import partition_draws
    rng = np.random.default_rng(1)
    counts = rng.integers(1, 400, size=20011)
    total = counts.sum()
    for world, w0 in ((2, 0.9), (4, 0.77), (8, 0.5), (8, 0.0)):
        weights = [w0] + [1.0] * (world - 1)
        parts = partition_draws(counts, world, weights)
        assert parts[0][0] == 0 and parts[-1][1] == len(counts)
        for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
            assert e0 == b1 and b1 % 32 == 0
        sums = np.array([counts[b:e].sum() for b, e in parts], dtype=np.float64)
        want = total * np.array(weights) / sum(weights)
        assert np.all(np.abs(sums - want) <= 2 * 32 * 400), (world, w0)
    assert partition_draws(counts, 3, None) == partition_draws(counts, 3, [1, 1, 1])
