"""GPU: orbit_scene_update (SceneData::update_scene, scene.rs:404-492) byte-for-byte against the oracle, alone and as
the producer in front of the culling passes."""
import ctypes as C

import numpy as np
import pytest

from orbit_b200 import layouts as L
from orbit_b200 import scenes
from scene_update_cases import random_entities, random_mesh_infos

pytestmark = pytest.mark.gpu


def run_gpu(ctx, t, slots, vo, cursor, mi, capacity=1 << 26):
    import torch
    from orbit_b200 import _lib
    dev = ctx.device
    n = len(t)
    d_t = torch.from_numpy(np.ascontiguousarray(t).view(np.uint8).reshape(-1).copy()).to(dev) if n else torch.zeros(16, dtype=torch.uint8, device=dev)
    d_slots = torch.from_numpy(slots.view(np.int32).copy()).to(dev) if n else torch.zeros(1, dtype=torch.int32, device=dev)
    d_vo = torch.from_numpy(vo.view(np.int32).copy()).to(dev) if n else torch.zeros(1, dtype=torch.int32, device=dev)
    d_cur = torch.from_numpy(cursor.view(np.int32).copy()).to(dev)
    d_mi = torch.from_numpy(mi.view(np.uint8).reshape(-1).copy()).to(dev)
    d_ed = torch.full((max(n, 1) * 128,), 0xCD, dtype=torch.uint8, device=dev)
    d_draws = torch.full((4 + 12 * max(n, 1),), 0xCD, dtype=torch.uint8, device=dev)
    u = L.SceneUpdate()
    u.transforms, u.mesh_slots, u.visibility_offsets = d_t.data_ptr(), d_slots.data_ptr(), d_vo.data_ptr()
    u.mesh_infos, u.visibility_cursor = d_mi.data_ptr(), d_cur.data_ptr()
    u.n_entities, u.visibility_capacity_words = n, capacity
    u.entity_data, u.entity_draws = d_ed.data_ptr(), d_draws.data_ptr()
    _lib.check(_lib.lib().orbit_scene_update(ctx._h, C.byref(u), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "orbit_scene_update")
    torch.cuda.synchronize()
    count = int(d_draws[:4].view(torch.int32).item())
    return (d_ed.cpu().numpy()[:count * 128].view(L.entity_dtype), d_draws.cpu().numpy()[:4 + 12 * count],
            d_vo.cpu().numpy().view(np.uint32)[:n], d_cur.cpu().numpy().view(np.uint32), d_ed.cpu().numpy()[count * 128:])


def check_case(ctx, oracle, t, slots, vo, cursor, mi, capacity=1 << 26):
    o_vo, o_cur = vo.copy(), cursor.copy()
    o_ed, o_draws, o_ovf = oracle.scene_update(t, slots, o_vo, mi, o_cur, capacity)
    g_ed, g_draws, g_vo, g_cur, tail = run_gpu(ctx, t, slots, vo, cursor, mi, capacity)
    assert np.array_equal(g_draws, o_draws)
    assert np.array_equal(g_ed.view(np.uint32), o_ed.view(np.uint32))
    assert np.array_equal(g_vo, o_vo) and g_cur[0] == o_cur[0]
    assert np.all(tail == 0xCD), "wrote past the last instance"
    return o_ovf, g_vo, g_cur


@pytest.mark.parametrize("n", [1, 31, 256, 257, 5000, 70001, 300001])   # the last one grows the per-tile scratch (> 1024 tiles)
def test_scene_update_matches_oracle(gpu_context, oracle, n):
    t, slots, vo, cursor = random_entities(n, 50, 100 + n)
    check_case(gpu_context, oracle, t, slots, vo, cursor, random_mesh_infos(50, n))


def test_scene_update_preallocated_ranges_and_second_frame(gpu_context, oracle):
    t, slots, vo, cursor = random_entities(20000, 64, 7, preallocated_fraction=0.4)
    mi = random_mesh_infos(64, 7)
    _, vo1, cur1 = check_case(gpu_context, oracle, t, slots, vo, cursor, mi)
    # next frame: every entity owns a range now -> nothing is allocated, the offsets are reused (scene.rs:422-423)
    _, vo2, cur2 = check_case(gpu_context, oracle, t, slots, vo1.copy(), cur1.copy(), mi)
    assert np.array_equal(vo1, vo2) and cur1[0] == cur2[0]


def test_scene_update_no_meshes_and_empty(gpu_context, oracle):
    t, slots, vo, cursor = random_entities(1000, 4, 9)
    slots[:] = L.NO_MESH
    check_case(gpu_context, oracle, t, slots, vo, cursor, random_mesh_infos(4, 9))
    g = run_gpu(gpu_context, t[:0], slots[:0], vo[:0], cursor, random_mesh_infos(4, 9))
    assert len(g[1]) == 4 and int(g[1].view(np.uint32)[0]) == 0


def test_scene_update_overflow_status(gpu_context, oracle):
    t, slots, vo, cursor = random_entities(3000, 5, 4, no_mesh_fraction=0.0)
    mi = random_mesh_infos(5, 4)
    mi["mesh_lods"][:, :, 1] = 64
    gpu_context.poll_status()
    ovf, _, _ = check_case(gpu_context, oracle, t, slots, vo, cursor, mi, capacity=5999)
    assert ovf == 1 and gpu_context.poll_status()[1].visibility_overflow == 1
    ovf, _, _ = check_case(gpu_context, oracle, t, slots, vo, cursor, mi, capacity=6000)
    assert ovf == 0 and gpu_context.poll_status()[1].visibility_overflow == 0


def test_scene_update_c3_size(gpu_context, oracle):
    """250 k instanced entities (BASELINE C3): the scale §8f item 3 names."""
    sc, _ = scenes.config_c3()
    vo = np.full(sc.n_entities, L.NO_VISIBILITY_RANGE, np.uint32)
    check_case(gpu_context, oracle, sc.transforms, sc.draws["mesh_index"].copy(), vo, np.zeros(1, np.uint32), sc.mesh_infos)


def test_scene_update_feeds_the_culling_passes(gpu_context, oracle):
    """SceneData.update_scene -> early/Hi-Z/late culling on the buffers it produced == the oracle chain."""
    import torch
    from orbit_b200 import frame
    from orbit_b200.scene import SceneData
    ctx = gpu_context
    sc, view = scenes.config_c1(scale=0.5, lods=(100, 40))
    depth = scenes.make_depth(sc, view)
    ds = frame.DeviceScene.upload(ctx, sc)
    sd = SceneData(ctx, sc.n_entities)
    sd.set_entities(sc.transforms, sc.draws["mesh_index"])
    sd.update_scene(ds.assets)
    ds.scene.entity_buffer, ds.scene.entity_draw_buffer = sd.entity_data_buffer, sd.entity_draw_buffer
    vs = frame.ViewState(ctx, ds, (view.width, view.height))
    # oracle chain: its own scene update produces the entity buffers its culling reads
    vo = np.full(sc.n_entities, L.NO_VISIBILITY_RANGE, np.uint32)
    o_ed, o_draws, _ = oracle.scene_update(sc.transforms, sc.draws["mesh_index"].copy(), vo, sc.mesh_infos, np.zeros(1, np.uint32))
    sc.entities, sc.entity_draws = o_ed.copy(), o_draws.copy()
    hs = oracle.HostScene(sc)
    d_depth = torch.from_numpy(depth).to(ctx.device)
    for f in range(2):
        g = frame.depth_prepass_culling(ctx, ds, vs, view, d_depth)
        torch.cuda.synchronize()
        o = oracle.depth_prepass_culling(hs, view, depth)
        for k in ("early", "late"):
            ghdr, grecs = frame.read_dispatch(g[k][0]); ohdr, orecs = oracle.parse_dispatch(o[k][0])
            assert ghdr.tolist() == ohdr.tolist() and np.array_equal(grecs.view(np.uint32), orecs.view(np.uint32)), (f, k)
            gn, gd = frame.read_draws(g[k][1]); on, od = oracle.parse_draws(o[k][1])
            assert gn == on and np.array_equal(gd.view(np.uint32), od.view(np.uint32)), (f, k)
        assert f == 0 or frame.read_draws(g["early"][1])[0] > 0      # frame 1 draws what frame 0's late pass found


def test_compiled_host_frame_loop_matches_oracle(gpu_context, oracle):
    """orbit_b200/host/frame_driver.cpp (the compiled host loop bench.py's e2e number comes from): HOST transforms + depth
    in, HOST survivor lists out, pipelined over scene copies — the lists of the last step equal the oracle chain's."""
    import torch
    from orbit_b200 import frame
    from orbit_b200.scene import SceneData
    ctx = gpu_context
    sc, view = scenes.config_c1(scale=0.5, lods=(100, 40))
    depth = scenes.make_depth(sc, view)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
    copies, sds = [], []
    for i in range(4):
        ds = frame.DeviceScene.upload(ctx, sc)
        vs = frame.ViewState(ctx, ds, (view.width, view.height), name="hfl%d" % i)
        pf = frame.PreparedFrame(ctx, ds, vs, view, torch.zeros((view.height, view.width), dtype=torch.float32, device=ctx.device), name="hfl%d" % i,
                                 main_pass=True)
        sd = SceneData(ctx, sc.n_entities)
        sd.set_entities(sc.transforms, sc.draws["mesh_index"])
        sd.transforms.zero_()                                           # the loop must bring the transforms in itself
        sd.entity_data_buffer, sd.entity_draw_buffer = ds.scene.entity_buffer, ds.scene.entity_draw_buffer
        copies.append(pf); sds.append(sd)
    h_t, h_d = pin(sc.transforms), torch.from_numpy(depth).pin_memory()
    h_c = torch.zeros(3, dtype=torch.int32).pin_memory()
    h_e = torch.zeros(28 * sc.n_meshlet_instances, dtype=torch.uint8).pin_memory()
    h_l = torch.zeros(28 * sc.n_meshlet_instances, dtype=torch.uint8).pin_memory()
    h_m = torch.zeros(28 * sc.n_meshlet_instances, dtype=torch.uint8).pin_memory()
    rep = frame.host_frame_loop(ctx, copies, sds, h_t, h_d, h_c, h_e, h_l, h_m, steps=6, lookahead=2)
    assert rep["h2d_bytes_per_step"] == h_t.numel() + depth.nbytes
    # step 5 ran on copy 1 and was that copy's second frame
    vo = np.full(sc.n_entities, L.NO_VISIBILITY_RANGE, np.uint32)
    o_ed, o_draws, _ = oracle.scene_update(sc.transforms, sc.draws["mesh_index"].copy(), vo, sc.mesh_infos, np.zeros(1, np.uint32))
    sc.entities, sc.entity_draws = o_ed.copy(), o_draws.copy()
    hs = oracle.HostScene(sc)
    oracle.depth_prepass_culling(hs, view, depth); oracle.main_pass_culling(hs, view)
    o = oracle.depth_prepass_culling(hs, view, depth)
    om = oracle.main_pass_culling(hs, view)                             # MAIN: pass 1 with the bits the late pass wrote
    ne, e = oracle.parse_draws(o["early"][1]); nl, l = oracle.parse_draws(o["late"][1]); nm, m = oracle.parse_draws(om[1])
    assert (int(h_c[0]), int(h_c[1]), int(h_c[2])) == (ne, nl, nm) and ne > 0 and nm >= ne
    assert np.array_equal(h_e.numpy()[:28 * ne], e.view(np.uint8).reshape(-1))
    assert np.array_equal(h_l.numpy()[:28 * nl], l.view(np.uint8).reshape(-1))
    assert np.array_equal(h_m.numpy()[:28 * nm], m.view(np.uint8).reshape(-1))
    assert rep["d2h_bytes_per_step"] == 12 + 28 * (ne + nl + nm)
