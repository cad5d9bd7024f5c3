"""Golden fixtures (tests/golden/c1_and_clusters.json, made by tests/golden/make_golden.py from the oracle):
CPU: the oracle still reproduces them; GPU: the CUDA path reproduces them through the C ABI."""
import importlib.util
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
G = importlib.util.module_from_spec(spec)
spec.loader.exec_module(G)
FIX = json.load(open(os.path.join(HERE, "golden", "c1_and_clusters.json")))


def _inputs_or_skip(got, key):
    if got != FIX[key]:
        pytest.skip("this platform generates different input bytes (libm last-ulp differences in numpy): "
                    "golden outputs do not apply; parity is still covered by the live oracle comparison tests")


def test_oracle_reproduces_golden_c1(oracle):
    sc, view, depth = G.c1_case()
    _inputs_or_skip(G.input_digests(sc, depth), "c1_inputs")
    got = G.run_frames(sc, view, depth)
    assert got == FIX["c1"]
    # the steady state (frame 1): early pass redraws exactly what frame 0 found, late pass finds nothing new
    assert got["frame1_early"]["draws"] == got["frame0_main"]["draws"] and got["frame1_late"]["draws"] == 0


def test_oracle_reproduces_golden_clusters(oracle):
    sc, view, depth, lights, p = G.cluster_case()
    _inputs_or_skip({"depth": G.sha(depth), "lights": G.sha(lights), "params": G.sha(np.frombuffer(bytes(p), np.uint8))}, "clusters_inputs")
    assert G.run_clusters(depth, lights, p) == FIX["clusters"]


@pytest.mark.gpu
def test_cuda_reproduces_golden_c1(gpu_context):
    import torch
    from orbit_b200 import frame
    sc, view, depth = G.c1_case()
    _inputs_or_skip(G.input_digests(sc, depth), "c1_inputs")
    ctx = gpu_context
    ds = frame.DeviceScene.upload(ctx, sc)
    vs = frame.ViewState(ctx, ds, (view.width, view.height))
    d_depth = torch.from_numpy(depth).to(ctx.device)
    for f in range(2):
        g = frame.depth_prepass_culling(ctx, ds, vs, view, d_depth)
        m = frame.main_pass_culling(ctx, ds, vs, view)
        torch.cuda.synchronize()
        for k, pair in (("early", g["early"]), ("late", g["late"]), ("main", m)):
            want = FIX["c1"]["frame%d_%s" % (f, k)]
            hdr, recs = frame.read_dispatch(pair[0]); n, draws = frame.read_draws(pair[1])
            assert [int(v) for v in hdr] == want["header"] and n == want["draws"]
            assert G.sha(recs) == want["records_sha"] and G.sha(draws) == want["draws_sha"], (f, k)
        st = FIX["c1"]["frame%d_state" % f]
        assert G.sha(vs.entity_visibility.cpu().numpy()) == st["entity_visibility_sha"]
        assert G.sha(vs.meshlet_visibility.cpu().numpy()) == st["meshlet_visibility_sha"]
        assert G.sha(vs.depth_pyramid.texels.cpu().numpy()) == st["hiz_sha"]


@pytest.mark.gpu
def test_cuda_reproduces_golden_clusters(gpu_context):
    import torch
    from orbit_b200 import frame
    from orbit_b200.passes import ClusterSettings, compute_clusters
    sc, view, depth, lights, p = G.cluster_case()
    _inputs_or_skip({"depth": G.sha(depth), "lights": G.sha(lights), "params": G.sha(np.frombuffer(bytes(p), np.uint8))}, "clusters_inputs")
    ctx = gpu_context
    ds = frame.DeviceScene.upload(ctx, sc, lights=lights)
    st = ClusterSettings(screen_resolution=(1920, 1080), z_slice_count=24, tile_size_px=120)
    info, _ = compute_clusters(ctx, st, view.view, view.projection_matrix, view.near, torch.from_numpy(depth).to(ctx.device), ds.scene)
    torch.cuda.synchronize()
    want = FIX["clusters"]
    n = 16 * 9 * 24
    unique = info.unique_cluster_buffer[:16 + 4 * n].cpu().numpy().view(np.uint32)
    assert int(unique[3]) == want["active"]
    index = info.light_index_buffer[:4 + 4 * want["total_indices"]].cpu().numpy().view(np.uint32)
    assert int(index[0]) == want["total_indices"]
    assert G.sha(info.tile_depth_slice_mask[:4 * 16 * 9].cpu().numpy()) == want["masks_sha"]
    assert G.sha(info.cluster_depth_bounds[:8 * n].cpu().numpy()) == want["bounds_sha"]
    assert G.sha(unique[:4 + want["active"]]) == want["unique_sha"]
    assert G.sha(info.light_offset_image[:8 * n].cpu().numpy()) == want["image_sha"]
    assert G.sha(index) == want["index_sha"]
