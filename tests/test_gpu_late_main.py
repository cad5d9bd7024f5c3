"""GPU parity of the fused LATE + MAIN passes (orbit_entity_cull_late_main / orbit_meshlet_cull_late_main, orbit_cuda.h):
one entity kernel and one pass-2 test kernel must produce, byte for byte, what the four separate stage calls — and the
oracle's separate LATE (forward.rs:266-403) and MAIN (forward.rs:518-548) passes — produce: both dispatch buffers, both
draw lists, both visibility bitmasks. Cases: perspective and orthographic cameras, a camera that moves (the late pass finds
survivors), MAIN-pass alpha filters that differ from the LATE pass's, a noskip alpha mode in the LATE pass, task payloads,
CUDA-graph replay, draw sub-ranges, and the pairs that must be refused."""
import ctypes as C

import numpy as np
import pytest
import torch

from orbit_b200 import scenes

pytestmark = pytest.mark.gpu


def _same(oracle, g, o, what):
    from orbit_b200.frame import read_dispatch, read_draws
    ghdr, grecs = read_dispatch(g[0]); ohdr, orecs = oracle.parse_dispatch(o[0])
    gn, gd = read_draws(g[1]); on, od = oracle.parse_draws(o[1])
    assert ghdr.tolist() == ohdr.tolist(), (what, "dispatch header", ghdr.tolist(), ohdr.tolist())
    assert np.array_equal(grecs.view(np.uint32), orecs.view(np.uint32)), (what, "records")
    assert gn == on, (what, "draw count", gn, on)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32)), (what, "draws")
    return int(ohdr[0]), on


def _bits_equal(vs, hs, what):
    assert np.array_equal(vs.meshlet_visibility.cpu().numpy().view(np.uint32), hs.meshlet_visibility), (what, "meshlet visibility")
    assert np.array_equal(vs.entity_visibility.cpu().numpy().view(np.uint32), hs.entity_visibility), (what, "entity visibility")


@pytest.mark.parametrize("ortho", [False, True])
def test_fused_late_main_follows_the_oracle_over_moving_frames(gpu_context, oracle, ortho):
    from orbit_b200 import frame
    from orbit_b200.passes import OcclusionCullInfo
    ctx = gpu_context
    sc, view_a = scenes.config_c2(scale=0.08)
    if ortho:
        lo, hi = np.asarray(sc.aabb_min), np.asarray(sc.aabb_max)
        c = 0.5 * (lo + hi)
        view_a = scenes.orthographic_view((c[0] - 60.0, c[1] + 40.0, c[2] - 60.0), (0.6, -0.45, 0.66), 1024, 1024, 80.0, 0.5, 400.0)
        view_b = scenes.orthographic_view((c[0] - 40.0, c[1] + 40.0, c[2] - 70.0), (0.5, -0.45, 0.74), 1024, 1024, 80.0, 0.5, 400.0)
    else:
        view_b = scenes.perspective_view((-6.0, 2.0, -6.0), (np.sin(np.radians(52.0)), 0.0, np.cos(np.radians(52.0))), view_a.width, view_a.height)
    depths = {id(v): scenes.make_depth(sc, v) for v in (view_a, view_b)}
    ds = frame.DeviceScene.upload(ctx, sc)
    vs = frame.ViewState(ctx, ds, (view_a.width, view_a.height), name="lm_%d" % ortho)
    hs = oracle.HostScene(sc)
    totals = []
    for f, view in enumerate([view_a, view_a, view_b, view_a, view_b]):
        depth = depths[id(view)]
        d_depth = torch.from_numpy(depth).to(ctx.device)
        early = frame.cull_pass(ctx, "lm_early", ds, frame.cull_info_for(view, OcclusionCullInfo("read", vs.entity_visibility, vs.meshlet_visibility)))
        vs.depth_pyramid.update(d_depth)
        g = frame.late_and_main_culling(ctx, ds, vs, view, name="lm", main_name="lm_main")
        torch.cuda.synchronize()
        o = oracle.depth_prepass_culling(hs, view, depth)
        o["main"] = oracle.main_pass_culling(hs, view)
        _same(oracle, early, o["early"], (f, "early"))
        totals.append((_same(oracle, g["late"], o["late"], (f, "late")), _same(oracle, g["main"], o["main"], (f, "main"))))
        _bits_equal(vs, hs, f)
    # the moved frames' late passes found survivors, the main lists are never empty, and LATE and MAIN dispatch the same records
    assert totals[2][0][1] > 0 and all(t[1][1] > 0 for t in totals)
    assert all(t[0][0] == t[1][0] for t in totals)


def _raw_pair(ctx, ds, vs, g_late, g_main, sb, rcap, dcap, payloads=False, tag="raw"):
    from orbit_b200 import _lib, layouts as L
    lib = _lib.lib()
    mk = ctx.create_transient
    out = {k: (mk("%s_%s_dispatch" % (tag, k), L.DISPATCH_HEADER + 16 * rcap), mk("%s_%s_draws" % (tag, k), L.DRAW_HEADER + 28 * dcap))
           for k in ("late", "main")}
    pay = {k: (mk("%s_%s_payloads" % (tag, k), 44 * rcap) if payloads else None) for k in ("late", "main")}
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.orbit_entity_cull_late_main(ctx._h, C.byref(g_late), C.byref(g_main), C.byref(sb), vs.depth_pyramid._h,
                                           p(out["late"][0]), p(out["main"][0]), rcap, stream) == 0
    assert lib.orbit_meshlet_cull_late_main(ctx._h, C.byref(g_late), C.byref(g_main), C.byref(sb), vs.depth_pyramid._h, p(out["late"][0]), rcap,
                                            p(out["late"][1]), p(out["main"][1]), dcap, p(pay["late"]), p(pay["main"]), stream) == 0
    return out, pay


def test_fused_pair_with_other_alpha_filters_noskip_mode_and_payloads(gpu_context, oracle):
    """MAIN-pass alpha filters that differ from the LATE pass's (the MAIN list is filtered by the MAIN flags), a noskip alpha
    mode in the LATE pass, and task payloads of both lists — against the oracle's separate passes with the same CullInfos."""
    from orbit_b200 import frame
    from orbit_b200.passes import OcclusionCullInfo, _scene_buffers
    ctx = gpu_context
    sc, view = scenes.config_c2(scale=0.06)
    view_b = scenes.perspective_view((-6.0, 2.0, -6.0), (np.sin(np.radians(50.0)), 0.0, np.cos(np.radians(50.0))), view.width, view.height)
    ds = frame.DeviceScene.upload(ctx, sc)
    vs = frame.ViewState(ctx, ds, (view.width, view.height), name="lm_alpha")
    hs = oracle.HostScene(sc)
    rcap, dcap = int(ds.scene.record_capacity), int(ds.scene.draw_capacity)
    cases = [(view, 0b011, 0b000, 0b011), (view, 0b011, 0b000, 0b100), (view_b, 0b011, 0b010, 0b001), (view, 0b111, 0b100, 0b110),
             (view_b, 0b001, 0b001, 0b111)]
    seen_main = set()
    for f, (v, late_flags, noskip, main_flags) in enumerate(cases):
        depth = scenes.make_depth(sc, v)
        d_depth = torch.from_numpy(depth).to(ctx.device)
        oc_r = OcclusionCullInfo("read", vs.entity_visibility, vs.meshlet_visibility)
        oc_w = OcclusionCullInfo("write", vs.entity_visibility, vs.meshlet_visibility, vs.depth_pyramid, noskip_alphamode=noskip, aspect_ratio=v.aspect)
        ci_early, ci_late, ci_main = (frame.cull_info_for(v, oc_r), frame.cull_info_for(v, oc_w), frame.cull_info_for(v, oc_r))
        ci_late.alpha_mode_filter, ci_main.alpha_mode_filter = late_flags, main_flags
        early = frame.cull_pass(ctx, "lma_early", ds, ci_early)
        vs.depth_pyramid.update(d_depth)
        g_late, g_main = ci_late.to_gpu(), ci_main.to_gpu()
        sb = _scene_buffers(ds.assets, ds.scene, ci_late)
        g, pay = _raw_pair(ctx, ds, vs, g_late, g_main, sb, rcap, dcap, payloads=True, tag="lma")
        torch.cuda.synchronize()
        o_early = oracle.cull_pass(hs, ci_early.to_gpu())
        hs.update_pyramid(depth)
        o_late = oracle.cull_pass(hs, g_late, task_payloads=True)
        o_main = oracle.cull_pass(hs, g_main, task_payloads=True)
        _same(oracle, early, o_early[:2], (f, "early"))
        nl = _same(oracle, g["late"], o_late[:2], (f, "late"))
        nm = _same(oracle, g["main"], o_main[:2], (f, "main"))
        for k, o in (("late", o_late), ("main", o_main)):
            n = nl[0]
            assert np.array_equal(pay[k][:44 * n].cpu().numpy().view(np.uint32), np.frombuffer(o[2], np.uint32)[:11 * n]), (f, k, "payloads")
        _bits_equal(vs, hs, f)
        seen_main.add(nm[1])
    assert len(seen_main) > 2          # the MAIN filter really selects different lists


def test_fused_pair_replays_as_a_graph_and_over_draw_ranges(gpu_context, oracle):
    """PreparedFrame takes the fused form by default: direct launches, CUDA-graph replays, and two draw sub-ranges culled one
    after the other against shared bitmasks whose lists concatenate to the unsharded MAIN and LATE lists."""
    from orbit_b200 import frame
    from orbit_b200.multi_gpu import partition_draws
    ctx = gpu_context
    sc, view = scenes.config_c2(scale=0.06)
    depth = scenes.make_depth(sc, view)
    d_depth = torch.from_numpy(depth).to(ctx.device)
    ds = frame.DeviceScene.upload(ctx, sc)
    vs = frame.ViewState(ctx, ds, (view.width, view.height), name="lm_pf")
    pf = frame.PreparedFrame(ctx, ds, vs, view, d_depth, name="lm_pf", main_pass=True)
    unfused = frame.PreparedFrame(ctx, ds, vs, view, d_depth, name="lm_pf_unfused", main_pass=True, fuse_late_main=False)
    assert pf.fused and not unfused.fused
    hs = oracle.HostScene(sc)

    def check(p, what):
        torch.cuda.synchronize()
        o = oracle.depth_prepass_culling(hs, view, depth)
        o["main"] = oracle.main_pass_culling(hs, view)
        for k, g in (("early", (p.early_dispatch, p.early_draws)), ("late", (p.late_dispatch, p.late_draws)), ("main", (p.main_dispatch, p.main_draws))):
            _same(oracle, g, o[k], (what, k))
        _bits_equal(vs, hs, what)
    pf.launch(); check(pf, "fused 0")
    unfused.launch(); check(unfused, "unfused")
    pf.launch(); check(pf, "fused 1")
    pf.capture()
    oracle.depth_prepass_culling(hs, view, depth); oracle.main_pass_culling(hs, view)       # capture() ran one warm frame
    for i in range(3):
        pf.replay(); check(pf, "replay %d" % i)
    # ---- two draw sub-ranges, fused, against fresh shared bitmasks
    vs2 = frame.ViewState(ctx, ds, (view.width, view.height), name="lm_ranges")
    hs2 = oracle.HostScene(sc)
    lod0 = sc.mesh_infos["mesh_lods"][:, 0, 1][sc.draws["mesh_index"]]
    parts = [frame.PreparedFrame(ctx, frame.DeviceScene.upload(ctx, sc, draw_begin=b, draw_end=max(e, b)), vs2, view, d_depth, name="lm_r%d" % i, main_pass=True)
             for i, (b, e) in enumerate(partition_draws(lod0, 2))]
    assert all(p.fused for p in parts)
    for f in range(2):
        for p in parts:
            p.entity(False); p.meshlet(False)
        parts[0].hiz()
        for p in parts:
            p.entity_late_main(); p.meshlet_late_main()
        torch.cuda.synchronize()
        o = oracle.depth_prepass_culling(hs2, view, depth)
        o["main"] = oracle.main_pass_culling(hs2, view)
        for k, bufs in (("late", [(p.late_dispatch, p.late_draws) for p in parts]), ("main", [(p.main_dispatch, p.main_draws) for p in parts])):
            recs = np.concatenate([frame.read_dispatch(b[0])[1] for b in bufs]); draws = np.concatenate([frame.read_draws(b[1])[1] for b in bufs])
            ohdr, orecs = oracle.parse_dispatch(o[k][0]); on, od = oracle.parse_draws(o[k][1])
            assert len(recs) == int(ohdr[0]) and np.array_equal(recs.view(np.uint32), orecs.view(np.uint32)), (f, k)
            assert len(draws) == on and np.array_equal(draws.view(np.uint32), od.view(np.uint32)), (f, k)
        _bits_equal(vs2, hs2, ("ranges", f))


def test_incompatible_pairs_are_refused(gpu_context):
    """orbit_cull_pair_compatible and the error return of the fused calls: another plane, another LOD base, pass numbers the
    wrong way round, meshlet occlusion off."""
    from orbit_b200 import _lib, frame, layouts as L
    from orbit_b200.passes import OcclusionCullInfo, _scene_buffers
    ctx, lib = gpu_context, _lib.lib()
    sc, view = scenes.config_c1(scale=0.2)
    ds = frame.DeviceScene.upload(ctx, sc)
    vs = frame.ViewState(ctx, ds, (view.width, view.height), name="lm_bad")
    oc_r = OcclusionCullInfo("read", vs.entity_visibility, vs.meshlet_visibility)
    oc_w = OcclusionCullInfo("write", vs.entity_visibility, vs.meshlet_visibility, vs.depth_pyramid, aspect_ratio=view.aspect)
    ci_late, ci_main = frame.cull_info_for(view, oc_w), frame.cull_info_for(view, oc_r)
    g_late, g_main = ci_late.to_gpu(), ci_main.to_gpu()
    ok = lambda a, b: lib.orbit_cull_pair_compatible(C.byref(a), C.byref(b))
    assert ok(g_late, g_main) == 1
    assert ok(g_main, g_late) == 0 and ok(g_late, g_late) == 0
    bad = []
    for edit in ("plane", "lod", "mocc", "view"):
        g = ci_main.to_gpu()
        if edit == "plane":
            g.cull_planes[1][3] += 0.5
        elif edit == "lod":
            g.lod_base *= 2.0
        elif edit == "mocc":
            g.meshlet_visibility_buffer = L.NO_BUFFER
        else:
            g.view_matrix.m[3][0] += 1.0
        assert ok(g_late, g) == 0, edit
        bad.append(g)
    g = ci_main.to_gpu(); g.alpha_mode_flags = 0b100
    assert ok(g_late, g) == 1                       # the alpha filter may differ
    sb = _scene_buffers(ds.assets, ds.scene, ci_late)
    rcap, dcap = int(ds.scene.record_capacity), int(ds.scene.draw_capacity)
    a = ctx.create_transient("lm_bad_a", L.DISPATCH_HEADER + 16 * rcap); b = ctx.create_transient("lm_bad_b", L.DISPATCH_HEADER + 16 * rcap)
    da = ctx.create_transient("lm_bad_da", L.DRAW_HEADER + 28 * dcap); db = ctx.create_transient("lm_bad_db", L.DRAW_HEADER + 28 * dcap)
    p = lambda t: C.c_void_p(t.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    INVALID = lib.orbit_entity_cull_late_main(ctx._h, C.byref(g_late), C.byref(bad[0]), C.byref(sb), vs.depth_pyramid._h, p(a), p(b), rcap, stream)
    assert INVALID != 0
    assert lib.orbit_meshlet_cull_late_main(ctx._h, C.byref(g_late), C.byref(bad[1]), C.byref(sb), vs.depth_pyramid._h, p(a), rcap, p(da), p(db), dcap,
                                            None, None, stream) == INVALID
    assert lib.orbit_entity_cull_late_main(ctx._h, C.byref(g_late), C.byref(g_main), C.byref(sb), vs.depth_pyramid._h, p(a), p(a), rcap, stream) == INVALID


def test_random_cameras_through_the_prepared_frame(gpu_context, oracle):
    """Randomised differential run of the whole frame in the form bench.py times (EARLY, Hi-Z, fused LATE + MAIN through
    PreparedFrame): a multi-LOD lattice, twelve random cameras (position, direction, field of view, resolution, LOD base),
    each culled for two frames against visibility bits left by the previous camera; every buffer against the oracle."""
    from orbit_b200 import frame
    ctx = gpu_context
    rng = np.random.default_rng(20260217)
    sc, _ = scenes.config_c1(scale=0.5, lods=(100, 45, 12))
    lo, hi = np.asarray(sc.aabb_min, np.float64), np.asarray(sc.aabb_max, np.float64)
    ds = frame.DeviceScene.upload(ctx, sc)
    hs = oracle.HostScene(sc)
    states = {}
    late_survivors = 0
    for i in range(12):
        w, h = [(640, 360), (1280, 720), (1000, 600), (1920, 1080)][int(rng.integers(4))]
        eye = lo + (hi - lo) * rng.uniform(-0.1, 1.1, 3)
        d = rng.normal(size=3); d[1] *= 0.3; d /= np.linalg.norm(d)
        view = scenes.perspective_view(tuple(eye), tuple(d), w, h, fov_deg=float(rng.uniform(40.0, 110.0)))
        view.lod_base = float(rng.choice([4.0, 16.0, 40.0]))
        depth = scenes.make_depth(sc, view)
        d_depth = torch.from_numpy(depth).to(ctx.device)
        if (w, h) not in states:          # one visibility state per resolution (the pyramid belongs to it); the oracle shares ONE
            states[(w, h)] = frame.ViewState(ctx, ds, (w, h), name="lm_rand_%dx%d" % (w, h))
        vs = states[(w, h)]
        # the oracle keeps one set of bits: copy them into this resolution's state so both sides start from the same history
        vs.meshlet_visibility.copy_(torch.from_numpy(hs.meshlet_visibility.view(np.int32)).to(ctx.device).view(vs.meshlet_visibility.dtype)[:vs.meshlet_visibility.numel()])
        vs.entity_visibility.copy_(torch.from_numpy(hs.entity_visibility.view(np.int32)).to(ctx.device).view(vs.entity_visibility.dtype)[:vs.entity_visibility.numel()])
        pf = frame.PreparedFrame(ctx, ds, vs, view, d_depth, name="lm_rand_%d" % i, main_pass=True)
        assert pf.fused
        for f in range(2):
            pf.launch()
            torch.cuda.synchronize()
            o = oracle.depth_prepass_culling(hs, view, depth)
            o["main"] = oracle.main_pass_culling(hs, view)
            for k, g in (("early", (pf.early_dispatch, pf.early_draws)), ("late", (pf.late_dispatch, pf.late_draws)), ("main", (pf.main_dispatch, pf.main_draws))):
                n = _same(oracle, g, o[k], (i, f, k))
                if k == "late":
                    late_survivors += n[1]
            _bits_equal(vs, hs, (i, f))
    assert late_survivors > 0
