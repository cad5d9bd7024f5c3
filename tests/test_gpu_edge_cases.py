"""GPU parity on the edge cases of the path: empty inputs, ragged records, non-affine matrices (the p/p.w division),
12 cull planes, LOD clamping, dispatch-record capacity overflow, graph replay."""
import numpy as np
import pytest
import torch

from orbit_b200 import layouts as L
from orbit_b200 import scenes

pytestmark = pytest.mark.gpu


def _cmp(oracle, g, o):
    from orbit_b200.frame import read_dispatch, read_draws
    ghdr, grecs = read_dispatch(g[0]); ohdr, orecs = oracle.parse_dispatch(o[0])
    gn, gd = read_draws(g[1]); on, od = oracle.parse_draws(o[1])
    assert ghdr.tolist() == ohdr.tolist()
    assert np.array_equal(grecs.view(np.uint32), orecs.view(np.uint32))
    assert gn == on and np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    return int(ohdr[0]), on


def _two_pass(ctx, oracle, sc, view, depth, frames=2):
    from orbit_b200 import frame
    ds = frame.DeviceScene.upload(ctx, sc)
    vs = frame.ViewState(ctx, ds, (view.width, view.height))
    hs = oracle.HostScene(sc)
    d_depth = torch.from_numpy(depth).to(ctx.device)
    tot = 0
    for f in range(frames):
        g = frame.depth_prepass_culling(ctx, ds, vs, view, d_depth)
        torch.cuda.synchronize()
        o = oracle.depth_prepass_culling(hs, view, depth)
        for k in ("early", "late"):
            tot += _cmp(oracle, g[k], o[k])[1]
        assert np.array_equal(vs.meshlet_visibility.cpu().numpy().view(np.uint32), hs.meshlet_visibility)
        assert np.array_equal(vs.entity_visibility.cpu().numpy().view(np.uint32), hs.entity_visibility)
    return tot


def test_zero_entities(gpu_context, oracle):
    sc, view = scenes.config_c1(scale=0.1)
    sc.entity_draws[:4] = 0          # device-side count = 0: nothing may be emitted, headers must still be written
    sc_n = sc.n_entities
    depth = scenes.make_depth(sc, view)
    from orbit_b200 import frame
    ds = frame.DeviceScene.upload(gpu_context, sc)
    hs = oracle.HostScene(sc)
    from orbit_b200.passes import OcclusionCullInfo
    g = frame.cull_pass(gpu_context, "zero", ds, frame.cull_info_for(view, OcclusionCullInfo("none")))
    torch.cuda.synchronize()
    o = oracle.cull_pass(hs, oracle.gpu_cull_info(view, "none"))
    assert _cmp(oracle, g, o) == (0, 0)
    ds.scene.entity_draw_count = 0   # host-side count 0 as well (grid of one CTA)
    g = frame.cull_pass(gpu_context, "zero", ds, frame.cull_info_for(view, OcclusionCullInfo("none")))
    torch.cuda.synchronize()
    assert _cmp(oracle, g, o) == (0, 0)


@pytest.mark.parametrize("lods", [(1,), (31,), (32,), (33,), (64, 7), (257, 96, 5)])
def test_ragged_meshlet_counts(gpu_context, oracle, lods):
    """Records of 1..32 meshlets, partial last records, LOD switches that re-use visibility words."""
    sc, view = scenes.config_c1(scale=0.15, lods=lods)
    view.lod_base, view.lod_step = 6.0, 1.7
    depth = scenes.make_depth(sc, view)
    assert _two_pass(gpu_context, oracle, sc, view, depth) > 0


def test_non_affine_model_matrices(gpu_context, oracle):
    """Bottom row != (0,0,0,1): exercises the p / p.w division the kernels skip for affine matrices."""
    sc, view = scenes.config_c1(scale=0.2)
    m = sc.entities["model_matrix"]            # [n][col][row]
    rng = np.random.default_rng(5)
    n = len(m)
    m[:, 3, 3] = rng.uniform(0.7, 1.4, n).astype(np.float32)           # w scale
    m[::3, 0, 3] = rng.uniform(-0.01, 0.01, len(m[::3])).astype(np.float32)   # a little perspective in x
    depth = scenes.make_depth(sc, view)
    assert _two_pass(gpu_context, oracle, sc, view, depth) > 0


def test_twelve_planes_and_lod_clamp(gpu_context, oracle):
    from orbit_b200 import frame
    from orbit_b200.passes import OcclusionCullInfo
    sc, view = scenes.config_c1(scale=0.3, lods=(120, 60, 30, 15))
    extra = []
    rng = np.random.default_rng(11)
    for _ in range(7):                 # 5 frustum planes + 7 random ones = MAX_CULL_PLANES
        nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
        extra.append([nrm[0], nrm[1], nrm[2], 40.0])
    view.planes = np.vstack([view.planes, np.array(extra)])
    assert len(view.planes) == 12
    for lod_range, base, step in (((0, 8), 4.0, 1.3), ((2, 3), 16.0, 2.0), ((0, 1), 16.0, 2.0), ((1, 8), 1e-3, 1.01)):
        view.lod_range, view.lod_base, view.lod_step = lod_range, base, step
        ds = frame.DeviceScene.upload(gpu_context, sc)
        hs = oracle.HostScene(sc)
        g = frame.cull_pass(gpu_context, "p12", ds, frame.cull_info_for(view, OcclusionCullInfo("none")))
        torch.cuda.synchronize()
        o = oracle.cull_pass(hs, oracle.gpu_cull_info(view, "none"))
        nrec, ndraw = _cmp(oracle, g, o)
        assert nrec > 0


def test_dispatch_capacity_overflow(gpu_context, oracle):
    from orbit_b200 import frame
    from orbit_b200.passes import OcclusionCullInfo
    sc, view = scenes.config_c1(scale=0.3)
    ctx = gpu_context
    ds = frame.DeviceScene.upload(ctx, sc)
    info = frame.cull_info_for(view, OcclusionCullInfo("none"))
    full = frame.cull_pass(ctx, "rc_full", ds, info)
    torch.cuda.synchronize()
    hdr, recs = frame.read_dispatch(full[0])
    n_full = int(hdr[0]); assert n_full > 8
    ds.scene.record_capacity = n_full // 2
    part = frame.cull_pass(ctx, "rc_part", ds, info)
    torch.cuda.synchronize()
    hdr2 = part[0][:12].cpu().numpy().view(np.uint32)
    assert int(hdr2[0]) == n_full and hdr2[1] == 1 and hdr2[2] == 1      # exact count even though records were dropped
    kept = part[0][12:12 + 16 * (n_full // 2)].cpu().numpy().view(L.dispatch_dtype)
    assert np.array_equal(kept.view(np.uint32), recs[:n_full // 2].view(np.uint32))
    code, st = ctx.poll_status()
    assert code == -5 and st.dispatch_overflow == 1
    # the meshlet stage clamps the device-side count to the capacity it was given: draws of the kept records only
    hs = oracle.HostScene(sc)
    o_disp, o_draws = oracle.cull_pass(hs, oracle.gpu_cull_info(view, "none"))
    o_disp = o_disp.copy()
    o_disp[:4] = np.frombuffer(np.uint32(n_full // 2).tobytes(), np.uint8)     # what a clamped reader sees
    o_draws2 = np.zeros_like(o_draws)
    import ctypes as C
    g = oracle.gpu_cull_info(view, "none")
    sb = hs.buffers(use_meshlet_vis=False)
    oracle.lib().oracle_meshlet_cull(C.byref(g), C.byref(sb), None, 0, 0, o_disp.ctypes.data_as(C.c_void_p),
                                     o_draws2.ctypes.data_as(C.c_void_p), sc.n_meshlet_instances, None, C.byref(oracle.Stats()))
    gn, gd = frame.read_draws(part[1]); on, od = oracle.parse_draws(o_draws2)
    assert gn == on and np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    ctx.poll_status()


def test_graph_replay_is_stateless(gpu_context, oracle):
    """A captured frame replayed many times keeps producing the oracle's steady state (scan epochs, tickets and
    scratch parities live in device memory)."""
    from orbit_b200 import frame
    sc, view = scenes.config_c1(scale=0.3)
    depth = scenes.make_depth(sc, view)
    ctx = gpu_context
    ds = frame.DeviceScene.upload(ctx, sc)
    vs = frame.ViewState(ctx, ds, (view.width, view.height))
    pf = frame.PreparedFrame(ctx, ds, vs, view, torch.from_numpy(depth).to(ctx.device), name="replay")
    pf.launch()
    pf.capture()               # frame 1 (warm) + frame 2 (captured)
    hs = oracle.HostScene(sc)
    for _ in range(3):
        o = oracle.depth_prepass_culling(hs, view, depth)
    for i in range(7):         # odd number of replays: scratch parity differs from the capture-time one
        pf.replay()
    torch.cuda.synchronize()
    _cmp(oracle, (pf.early_dispatch, pf.early_draws), o["early"])
    _cmp(oracle, (pf.late_dispatch, pf.late_draws), o["late"])
    assert np.array_equal(vs.meshlet_visibility.cpu().numpy().view(np.uint32), hs.meshlet_visibility)


def test_contexts_are_independent_across_threads(oracle):
    """SURVEY §8b threading: pass recording runs on worker threads (context.rs:1392-1423), so the library must be callable
    from any thread with an explicit context + stream and no thread-local state. Three host threads, each with its own
    context, stream and scene, run three frames concurrently; every thread's outputs equal the oracle's."""
    import threading
    from orbit_b200 import frame
    from orbit_b200.passes import Context
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cases = []
    for k, scale in enumerate((0.3, 0.45, 0.6)):
        sc, view = scenes.config_c1(scale=scale, lods=(100, 40) if k % 2 else (100,))
        cases.append((sc, view, scenes.make_depth(sc, view)))
    results, errors = [None] * len(cases), []

    def worker(k):
        try:
            sc, view, depth = cases[k]
            torch.cuda.set_device(0)
            ctx = Context(0)
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                ds = frame.DeviceScene.upload(ctx, sc)
                vs = frame.ViewState(ctx, ds, (view.width, view.height), name="thr%d" % k)
                d_depth = torch.from_numpy(depth).to(ctx.device)
                out = []
                for f in range(3):
                    g = frame.depth_prepass_culling(ctx, ds, vs, view, d_depth, name="thr%d" % k)
                    stream.synchronize()
                    out.append({s: (g[s][0].cpu().numpy().copy(), g[s][1][:4 + 28 * frame.read_draws(g[s][1], capacity=0)[0]].cpu().numpy().copy())
                                for s in ("early", "late")})
                results[k] = (out, vs.meshlet_visibility.cpu().numpy().view(np.uint32).copy())
            ctx.close()
        except Exception as e:   # surfaced in the main thread
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(cases))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k, (sc, view, depth) in enumerate(cases):
        hs = oracle.HostScene(sc)
        out, mvis = results[k]
        for f in range(3):
            o = oracle.depth_prepass_culling(hs, view, depth)
            for s in ("early", "late"):
                ohdr, orecs = oracle.parse_dispatch(o[s][0]); on, od = oracle.parse_draws(o[s][1])
                gd, gdraw = out[f][s]
                assert gd[:12].view(np.uint32).tolist() == ohdr.tolist(), (k, f, s)
                assert np.array_equal(gd[12:12 + 16 * int(ohdr[0])], orecs.view(np.uint8).reshape(-1)), (k, f, s)
                assert int(gdraw[:4].view(np.uint32)[0]) == on and np.array_equal(gdraw[4:], od.view(np.uint8).reshape(-1)), (k, f, s)
        assert np.array_equal(mvis, hs.meshlet_visibility), k


@pytest.mark.parametrize("pattern", ["two_ends", "single_record", "every_97th"])
def test_sparse_survivors_with_long_empty_stretches(gpu_context, oracle, pattern):
    """Survivors only in a few places of a long record list (pass 1 with hand-set visibility words): the emit kernel's
    output-balanced split then hands warps shares that straddle thousands of empty records, which it steps over chunk
    by chunk from the prefix instead of loading them."""
    from orbit_b200 import frame
    ctx = gpu_context
    sc, view = scenes.config_c2(scale=0.25)
    # a camera high above the middle of the city looking straight down: every entity is inside the frustum
    mid = (sc.aabb_min + sc.aabb_max) / 2
    view = scenes.perspective_view((mid[0], 2500.0, mid[2]), (0.0, -1.0, 0.001), 640, 360)
    ds = frame.DeviceScene.upload(ctx, sc)
    vs = frame.ViewState(ctx, ds, (view.width, view.height), name="sparse_" + pattern)
    hs = oracle.HostScene(sc)
    hs.entity_visibility[:] = 0xFFFFFFFF
    words = hs.meshlet_visibility
    words[:] = 0
    if pattern == "two_ends":
        words[:40] = 0xFFFFFFFF; words[-40:] = 0xFFFFFFFF
    elif pattern == "single_record":
        words[len(words) // 2] = 0x80000001
    else:
        words[::97] = 0x00010001
    vs.entity_visibility.copy_(torch.from_numpy(hs.entity_visibility.view(np.int32)).to(ctx.device).view(vs.entity_visibility.dtype))
    vs.meshlet_visibility.copy_(torch.from_numpy(words.view(np.int32)).to(ctx.device).view(vs.meshlet_visibility.dtype))
    g = frame.main_pass_culling(ctx, ds, vs, view)
    torch.cuda.synchronize()
    o = oracle.main_pass_culling(hs, view)
    nrec, ndraw = _cmp(oracle, g, o)
    assert nrec > 3000 and 0 < ndraw < nrec            # long list, few survivors
