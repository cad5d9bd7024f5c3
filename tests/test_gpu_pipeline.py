"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bar: dispatch records, draw commands, both visibility bitmasks and every Hi-Z texel bit-exact. The CUDA path is
deterministic and uses the oracle's canonical order, so buffers are compared byte for byte (stronger than the
sorted-set equality the reference's atomic appends would allow)."""
import numpy as np
import pytest
import torch

from orbit_b200 import layouts as L
from orbit_b200 import scenes

pytestmark = pytest.mark.gpu


def _cmp_pass(name, gpu_pair, ora_pair, oracle, frame):
    from orbit_b200.frame import read_dispatch, read_draws
    ghdr, grecs = read_dispatch(gpu_pair[0])
    ohdr, orecs = oracle.parse_dispatch(ora_pair[0])
    assert ghdr.tolist() == ohdr.tolist(), (name, "dispatch header")
    assert np.array_equal(grecs.view(np.uint32), orecs.view(np.uint32)), (name, "dispatch records")
    gn, gdraws = read_draws(gpu_pair[1])
    on, odraws = oracle.parse_draws(ora_pair[1])
    assert gn == on, (name, "draw count", gn, on)
    assert np.array_equal(gdraws.view(np.uint32), odraws.view(np.uint32)), (name, "draw commands")
    # sorted-set equality (what the reference's nondeterministic order would permit) follows from the above
    return int(ohdr[0]), on


def _run_two_frames(ctx, oracle, sc, view, depth, meshlet_occlusion=True):
    from orbit_b200 import frame
    ds = frame.DeviceScene.upload(ctx, sc)
    vs = frame.ViewState(ctx, ds, (view.width, view.height))
    hs = oracle.HostScene(sc)
    d_depth = torch.from_numpy(depth).to(ctx.device)
    totals = []
    for f in range(2):
        g = frame.depth_prepass_culling(ctx, ds, vs, view, d_depth, meshlet_occlusion)
        torch.cuda.synchronize()
        o = oracle.depth_prepass_culling(hs, view, depth, meshlet_occlusion)
        for k in ("early", "late"):
            totals.append((f, k) + _cmp_pass("frame%d %s" % (f, k), g[k], o[k], oracle, f))
        assert np.array_equal(vs.depth_pyramid.texels.cpu().numpy().view(np.uint32), hs.hiz_texels.view(np.uint32)), "Hi-Z texels"
        gm = frame.main_pass_culling(ctx, ds, vs, view, meshlet_occlusion)
        torch.cuda.synchronize()
        om = oracle.main_pass_culling(hs, view, meshlet_occlusion)
        totals.append((f, "main") + _cmp_pass("frame%d main" % f, gm, om, oracle, f))
        ev = vs.entity_visibility.cpu().numpy().view(np.uint32)
        assert np.array_equal(ev, hs.entity_visibility), "entity visibility words"
        mv = vs.meshlet_visibility.cpu().numpy().view(np.uint32)
        assert np.array_equal(mv, hs.meshlet_visibility), "meshlet visibility words"
    code, st = ctx.poll_status()
    assert code == 0
    return totals


def test_c1_two_frames_bit_exact(gpu_context, oracle):
    sc, view = scenes.config_c1()
    depth = scenes.make_depth(sc, view)
    totals = _run_two_frames(gpu_context, oracle, sc, view, depth)
    # every rejection path must actually fire on this scene
    late0 = [t for t in totals if t[:2] == (0, "late")][0]
    assert 0 < late0[2] < sc.n_records_lod0 and 0 < late0[3] < sc.n_meshlet_instances


def test_c1_multi_lod_bit_exact(gpu_context, oracle):
    sc, view = scenes.config_c1(lods=(200, 100, 50, 25))
    view.lod_base, view.lod_step = 8.0, 1.5
    depth = scenes.make_depth(sc, view)
    _run_two_frames(gpu_context, oracle, sc, view, depth)


def test_c2_reduced_bit_exact(gpu_context, oracle):
    sc, view = scenes.config_c2(scale=0.09)   # 900 buildings, 180k meshlets
    depth = scenes.make_depth(sc, view)
    _run_two_frames(gpu_context, oracle, sc, view, depth)


def test_no_meshlet_occlusion(gpu_context, oracle):
    sc, view = scenes.config_c1(scale=0.2)
    depth = scenes.make_depth(sc, view)
    _run_two_frames(gpu_context, oracle, sc, view, depth, meshlet_occlusion=False)


def test_orthographic_pass0_and_pass2(gpu_context, oracle):
    """Shadow-cascade style view (pass 0) plus the ortho occlusion branch the reference's callers never reach."""
    from orbit_b200 import frame
    from orbit_b200.passes import OcclusionCullInfo
    sc, _ = scenes.config_c2(scale=0.04)
    centre = (sc.aabb_min + sc.aabb_max) * 0.5
    d = np.array([-0.5, -0.7, 0.5]); d /= np.linalg.norm(d)
    view = scenes.orthographic_view(centre - d * 100.0, d, 2048, 2048, half_width=60.0, near=-20.0, far=250.0)
    depth = scenes.make_depth(sc, view)
    ctx = gpu_context
    ds = frame.DeviceScene.upload(ctx, sc)
    hs = oracle.HostScene(sc)
    g = frame.shadow_pass_culling(ctx, ds, view)
    torch.cuda.synchronize()
    o = oracle.cull_pass(hs, oracle.gpu_cull_info(view, "none"))
    nrec, ndraw = _cmp_pass("ortho pass0", g, o, oracle, 0)
    assert ndraw > 0
    vs = frame.ViewState(ctx, ds, (view.width, view.height))
    d_depth = torch.from_numpy(depth).to(ctx.device)
    gp = frame.depth_prepass_culling(ctx, ds, vs, view, d_depth)
    torch.cuda.synchronize()
    op = oracle.depth_prepass_culling(hs, view, depth)
    for k in ("early", "late"):
        _cmp_pass("ortho " + k, gp[k], op[k], oracle, 0)
    assert np.array_equal(vs.meshlet_visibility.cpu().numpy().view(np.uint32), hs.meshlet_visibility)


@pytest.mark.parametrize("size", [(1920, 1080), (3840, 2160), (1280, 720), (100, 60), (64, 64), (129, 257), (2, 2), (4096, 33)])
def test_hiz_bit_exact(gpu_context, oracle, size):
    from orbit_b200.passes import DepthPyramid
    w, h = size
    rng = np.random.default_rng(w * 7919 + h)
    depth = rng.random((h, w), dtype=np.float32)
    depth[rng.random((h, w)) < 0.3] = 0.0
    pyr = DepthPyramid(gpu_context, "t", (w, h))
    pyr.update(torch.from_numpy(depth).to(gpu_context.device))
    torch.cuda.synchronize()
    info, texels = oracle.hiz_build(depth)
    assert (pyr.info.width, pyr.info.height, pyr.info.levels, pyr.info.total_texels) == (info.width, info.height, info.levels, info.total_texels)
    assert np.array_equal(pyr.texels.cpu().numpy().view(np.uint32), texels.view(np.uint32))
    # size-independent property: the top level is the min of level 0 restricted to what the footprints reach,
    # and every level is <= (farther than or equal to) each of its four children
    for l in range(1, info.levels):
        pw, ph = max(info.width >> (l - 1), 1), max(info.height >> (l - 1), 1)
        cw, ch = max(info.width >> l, 1), max(info.height >> l, 1)
        parent = texels[info.level_offset[l]:info.level_offset[l] + cw * ch].reshape(ch, cw)
        child = texels[info.level_offset[l - 1]:info.level_offset[l - 1] + pw * ph].reshape(ph, pw)
        assert parent.min() == child.min()


def test_task_payloads(gpu_context, oracle):
    from orbit_b200 import frame
    from orbit_b200.passes import OcclusionCullInfo
    sc, view = scenes.config_c1(scale=0.3)
    ctx = gpu_context
    ds = frame.DeviceScene.upload(ctx, sc)
    hs = oracle.HostScene(sc)
    info = frame.cull_info_for(view, OcclusionCullInfo("none"))
    payload = torch.zeros(44 * sc.n_records_lod0, dtype=torch.uint8, device=ctx.device)
    g = frame.cull_pass(ctx, "tp", ds, info, task_payloads=payload)
    torch.cuda.synchronize()
    od, odr, op = oracle.cull_pass(hs, oracle.gpu_cull_info(view, "none"), task_payloads=True)
    nrec, _ = _cmp_pass("task payload pass", g, (od, odr), oracle, 0)
    assert np.array_equal(payload.cpu().numpy()[:44 * nrec], op[:44 * nrec])


def test_capacity_overflow_reports_exact_count(gpu_context, oracle):
    from orbit_b200 import frame
    from orbit_b200.passes import OcclusionCullInfo
    sc, view = scenes.config_c1(scale=0.3)
    ctx = gpu_context
    ds = frame.DeviceScene.upload(ctx, sc)
    full = frame.cull_pass(ctx, "cap_full", ds, frame.cull_info_for(view, OcclusionCullInfo("none")))
    torch.cuda.synchronize()
    n_full, draws_full = frame.read_draws(full[1])
    assert n_full > 10
    ds.scene.draw_capacity = n_full // 2
    small = frame.cull_pass(ctx, "cap_small", ds, frame.cull_info_for(view, OcclusionCullInfo("none")))
    torch.cuda.synchronize()
    n_small, draws_small = frame.read_draws(small[1], capacity=n_full // 2)
    assert n_small == n_full
    assert np.array_equal(draws_small.view(np.uint32), draws_full[:n_full // 2].view(np.uint32))
    code, st = ctx.poll_status()
    assert code == -5 and st.draw_overflow == 1
    code, st = ctx.poll_status()
    assert code == 0
