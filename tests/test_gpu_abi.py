"""GPU tests of the C ABI's contract beyond parity of one call: contexts used from worker threads (the reference records its
passes from rayon workers, context.rs:1392-1423), contexts on two devices in one process, scratch growth inside a stage
call, orbit_ctx_reserve, entity and meshlet stages of one context overlapping on two streams, and the draw_begin / draw_end
sub-ranges that shard one view driven through the CUDA path on ONE GPU."""
import ctypes as C
import threading

import numpy as np
import pytest
import torch

from orbit_b200 import scenes

pytestmark = pytest.mark.gpu


def _frames_equal(oracle, g, o):
    from orbit_b200.frame import read_dispatch, read_draws
    for k in g:
        ghdr, grecs = read_dispatch(g[k][0]); ohdr, orecs = oracle.parse_dispatch(o[k][0])
        gn, gd = read_draws(g[k][1]); on, od = oracle.parse_draws(o[k][1])
        assert ghdr.tolist() == ohdr.tolist(), k
        assert np.array_equal(grecs.view(np.uint32), orecs.view(np.uint32)), k
        assert gn == on and np.array_equal(gd.view(np.uint32), od.view(np.uint32)), k


def _run_two_frames(device_index, oracle, errors, tag):
    """Creates a context on `device_index` from the calling thread and checks two frames against the oracle."""
    try:
        from orbit_b200 import frame
        from orbit_b200.passes import Context
        ctx = Context(device_index)
        sc, view = scenes.config_c1(scale=0.2)
        depth = scenes.make_depth(sc, view)
        with torch.cuda.device(device_index):
            stream = torch.cuda.Stream(device=device_index)
            with torch.cuda.stream(stream):
                ds = frame.DeviceScene.upload(ctx, sc)
                vs = frame.ViewState(ctx, ds, (view.width, view.height))
                d_depth = torch.from_numpy(depth).to(ctx.device)
                hs = oracle.HostScene(sc)
                for _ in range(2):
                    g = frame.depth_prepass_culling(ctx, ds, vs, view, d_depth)
                    stream.synchronize()
                    o = oracle.depth_prepass_culling(hs, view, depth)
                    _frames_equal(oracle, g, o)
        ctx.close()
    except BaseException as e:      # noqa: BLE001 — reported to the main thread
        errors.append((tag, repr(e)))


def test_context_from_worker_threads(oracle):
    """Two worker threads, each with its own context on device 0, run frames at the same time."""
    errors = []
    ts = [threading.Thread(target=_run_two_frames, args=(0, oracle, errors, "t%d" % i)) for i in range(2)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errors, errors


def test_stage_call_with_another_device_current(oracle):
    """The caller's current device is not the context's: every entry point makes the context's device current for the call and
    restores the caller's (needs 2 GPUs). Driven through the raw C ABI from a worker thread whose current device stays 0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU")
    errors = []

    def worker():
        try:
            from orbit_b200 import _lib
            from orbit_b200.passes import Context, DepthPyramid
            torch.cuda.set_device(0)
            ctx = Context(1)                                   # orbit_ctx_create(1) while device 0 is current
            depth = np.random.default_rng(5).random((270, 480), dtype=np.float32)
            d_depth = torch.from_numpy(depth).to(ctx.device)
            pyr = DepthPyramid(ctx, "guard", (480, 270))
            stream = torch.cuda.Stream(device=1)
            stream.wait_stream(torch.cuda.current_stream(ctx.device))
            rc = _lib.lib().orbit_hiz_build(ctx._h, pyr._h, C.c_void_p(d_depth.data_ptr()), 480, 270, C.c_void_p(stream.cuda_stream))
            assert rc == 0, rc
            stream.synchronize()
            _, ref = oracle.hiz_build(depth)
            assert np.array_equal(pyr.texels.cpu().numpy().view(np.uint32), ref.view(np.uint32)), "pyramid built on device 1"
            assert torch.cuda.current_device() == 0, "current device changed to %d" % torch.cuda.current_device()
            ctx.close()
            assert torch.cuda.current_device() == 0
        except BaseException as e:      # noqa: BLE001
            errors.append(repr(e))
    t = threading.Thread(target=worker); t.start(); t.join()
    assert not errors, errors


def test_contexts_on_two_devices_in_one_process(oracle):
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU")
    errors = []
    ts = [threading.Thread(target=_run_two_frames, args=(d, oracle, errors, "dev%d" % d)) for d in (0, 1)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errors, errors


def test_scratch_grows_inside_a_stage_call_and_reserve(oracle):
    """A fresh context sees a small scene, then a larger one (record scratch, scan descriptors grow inside the calls, no
    synchronisation needed for correctness), then orbit_ctx_reserve for something larger still."""
    from orbit_b200 import _lib, frame
    from orbit_b200.passes import Context
    ctx = Context(0)
    for scale in (0.05, 1.0):
        sc, view = scenes.config_c1(scale=scale)
        depth = scenes.make_depth(sc, view)
        ds = frame.DeviceScene.upload(ctx, sc)
        vs = frame.ViewState(ctx, ds, (view.width, view.height))
        hs = oracle.HostScene(sc)
        d_depth = torch.from_numpy(depth).to(ctx.device)
        for _ in range(2):
            g = frame.depth_prepass_culling(ctx, ds, vs, view, d_depth)
            torch.cuda.synchronize()
            _frames_equal(oracle, g, oracle.depth_prepass_culling(hs, view, depth))
    assert _lib.lib().orbit_ctx_reserve(ctx._h, 300000, 2_000_000, 70000, 4000, 300000) == 0
    assert _lib.lib().orbit_ctx_reserve(None, 1, 1, 1, 1, 1) == _lib.ERR_INVALID_ARGUMENT
    g = frame.depth_prepass_culling(ctx, ds, vs, view, d_depth)
    torch.cuda.synchronize()
    _frames_equal(oracle, g, oracle.depth_prepass_culling(hs, view, depth))
    ctx.close()


def test_main_entity_stage_overlaps_late_meshlet_stage(gpu_context, oracle):
    """PreparedFrame forks the MAIN pass's entity stage onto a side stream beside the late meshlet stage (disjoint scratch of
    one context). Result = the sequential protocol's, directly and replayed as a CUDA graph."""
    from orbit_b200 import frame
    ctx = gpu_context
    sc, view = scenes.config_c2(scale=0.05)
    depth = scenes.make_depth(sc, view)
    ds = frame.DeviceScene.upload(ctx, sc)
    vs = frame.ViewState(ctx, ds, (view.width, view.height), name="ovl")
    pf = frame.PreparedFrame(ctx, ds, vs, view, torch.from_numpy(depth).to(ctx.device), name="ovl", main_pass=True)
    hs = oracle.HostScene(sc)

    def check():
        torch.cuda.synchronize()
        o = oracle.depth_prepass_culling(hs, view, depth)
        o["main"] = oracle.main_pass_culling(hs, view)
        g = {"early": (pf.early_dispatch, pf.early_draws), "late": (pf.late_dispatch, pf.late_draws), "main": (pf.main_dispatch, pf.main_draws)}
        _frames_equal(oracle, g, o)
    pf.launch(overlap_main_entity=True); check()
    pf.launch(overlap_main_entity=False); check()
    pf.capture()
    o = oracle.depth_prepass_culling(hs, view, depth); oracle.main_pass_culling(hs, view)      # capture() ran one warm frame
    for _ in range(3):
        pf.replay(); check()
    g = torch.cuda.CUDAGraph()                       # the forked form captured into a graph
    with torch.cuda.graph(g):
        pf.launch(overlap_main_entity=True)
    for _ in range(2):
        g.replay(); check()


def test_draw_ranges_on_one_gpu_concatenate_to_the_unsharded_result(gpu_context, oracle):
    """orbit_entity_cull / orbit_meshlet_cull with draw_begin / draw_end != 0: two halves of the entity draws culled one after
    the other on ONE GPU against shared visibility bitmasks; the halves' records and commands, concatenated, are the
    unsharded lists, and the bitmasks are the unsharded bitmasks (two frames, two-pass)."""
    from orbit_b200 import frame
    from orbit_b200.multi_gpu import partition_draws
    ctx = gpu_context
    sc, view = scenes.config_c2(scale=0.08)
    depth = scenes.make_depth(sc, view)
    lod0 = sc.mesh_infos["mesh_lods"][:, 0, 1][sc.draws["mesh_index"]]
    for world in (2, 3):
        ranges = partition_draws(lod0, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == sc.n_entities and all(b % 32 == 0 for b, _ in ranges)
        whole = frame.DeviceScene.upload(ctx, sc)
        vs = frame.ViewState(ctx, whole, (view.width, view.height), name="halves%d" % world)
        d_depth = torch.from_numpy(depth).to(ctx.device)
        parts = []
        for i, (b, e) in enumerate(ranges):
            ds = frame.DeviceScene.upload(ctx, sc, draw_begin=b, draw_end=max(e, b))
            parts.append(frame.PreparedFrame(ctx, ds, vs, view, d_depth, name="w%d_part%d" % (world, i)))
        hs = oracle.HostScene(sc)
        for f in range(2):
            for pf in parts:
                if pf.dscene.scene.draw_end > pf.dscene.scene.draw_begin:
                    pf.entity(False); pf.meshlet(False)
            parts[0].hiz()
            for pf in parts:
                if pf.dscene.scene.draw_end > pf.dscene.scene.draw_begin:
                    pf.entity(True); pf.meshlet(True)
            torch.cuda.synchronize()
            o = oracle.depth_prepass_culling(hs, view, depth)
            for k in ("early", "late"):
                recs, draws = [], []
                for pf in parts:
                    if pf.dscene.scene.draw_end <= pf.dscene.scene.draw_begin:
                        continue
                    disp, dr = (pf.early_dispatch, pf.early_draws) if k == "early" else (pf.late_dispatch, pf.late_draws)
                    recs.append(frame.read_dispatch(disp)[1]); draws.append(frame.read_draws(dr)[1])
                recs, draws = np.concatenate(recs), np.concatenate(draws)
                ohdr, orecs = oracle.parse_dispatch(o[k][0]); on, od = oracle.parse_draws(o[k][1])
                assert len(recs) == int(ohdr[0]) and np.array_equal(recs.view(np.uint32), orecs.view(np.uint32)), (world, f, k)
                assert len(draws) == on and np.array_equal(draws.view(np.uint32), od.view(np.uint32)), (world, f, k)
            assert np.array_equal(vs.meshlet_visibility.cpu().numpy().view(np.uint32), hs.meshlet_visibility), (world, f)
            assert np.array_equal(vs.entity_visibility.cpu().numpy().view(np.uint32), hs.entity_visibility), (world, f)


def test_record_mask_exchange_on_one_gpu(gpu_context, oracle):
    """The compact survivor exchange of the sharded view (orbit_meshlet_test -> orbit_record_masks_put ->
    orbit_draws_from_masks) with three 'ranks' played by one GPU: each range is tested into its own entry buffer, every buffer
    is put into its own region (with its count word) of the receiving array, and the list emitted from the regions must be the
    unsharded oracle list (two frames, both lists)."""
    from orbit_b200 import _lib, frame
    from orbit_b200.multi_gpu import partition_draws
    ctx, lib = gpu_context, _lib.lib()
    sc, view = scenes.config_c2(scale=0.08)
    depth = scenes.make_depth(sc, view)
    lod0 = sc.mesh_infos["mesh_lods"][:, 0, 1][sc.draws["mesh_index"]]
    world = 3
    ranges = partition_draws(lod0, world)
    whole = frame.DeviceScene.upload(ctx, sc)
    vs = frame.ViewState(ctx, whole, (view.width, view.height), name="mx")
    d_depth = torch.from_numpy(depth).to(ctx.device)
    parts = [frame.PreparedFrame(ctx, frame.DeviceScene.upload(ctx, sc, draw_begin=b, draw_end=max(e, b)), vs, view, d_depth, name="mx%d" % i)
             for i, (b, e) in enumerate(ranges)]
    rcap, dcap = parts[0].rcap, parts[0].dcap
    masks = [torch.zeros(16 * rcap, dtype=torch.uint8, device=ctx.device) for _ in parts]
    combined = torch.zeros(16 * rcap * world, dtype=torch.uint8, device=ctx.device)   # one region of rcap entries per rank
    counts = torch.zeros(world, dtype=torch.int32, device=ctx.device)
    out = torch.zeros(4 + 28 * dcap, dtype=torch.uint8, device=ctx.device)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    hs = oracle.HostScene(sc)

    def gather(late):
        combined.fill_(0xAB); counts.fill_(-1)
        for i, pf in enumerate(parts):
            disp = pf.late_dispatch if late else pf.early_dispatch
            assert lib.orbit_record_masks_put(ctx._h, p(masks[i]), p(disp), rcap, C.c_void_p(combined.data_ptr() + 16 * rcap * i),
                                              C.c_void_p(counts.data_ptr() + 4 * i), stream) == 0
        sb = parts[0].sb_late if late else parts[0].sb_early
        assert lib.orbit_draws_from_masks(ctx._h, C.byref(sb), p(combined), rcap, p(counts), world, p(out), dcap, stream) == 0
        torch.cuda.synchronize()
        return frame.read_draws(out)
    for f in range(2):
        for i, pf in enumerate(parts):
            pf.entity(False); pf.meshlet_test(False, masks[i])
        n_e, early = gather(False)
        parts[0].hiz()
        for i, pf in enumerate(parts):
            pf.entity(True); pf.meshlet_test(True, masks[i])
        n_l, late = gather(True)
        o = oracle.depth_prepass_culling(hs, view, depth)
        on, od = oracle.parse_draws(o["early"][1]); ln, ld = oracle.parse_draws(o["late"][1])
        assert (n_e, n_l) == (on, ln), (f, n_e, on, n_l, ln)
        assert np.array_equal(early.view(np.uint32), od.view(np.uint32)) and np.array_equal(late.view(np.uint32), ld.view(np.uint32)), f
        assert np.array_equal(vs.meshlet_visibility.cpu().numpy().view(np.uint32), hs.meshlet_visibility), f
    assert n_e > 0
    # a regular call on the same context afterwards is not disturbed by the test-only calls' counters
    g = frame.depth_prepass_culling(ctx, whole, vs, view, d_depth)
    torch.cuda.synchronize()
    _frames_equal(oracle, g, oracle.depth_prepass_culling(hs, view, depth))


def test_peer_put_and_wait_on_one_gpu(gpu_context):
    """orbit_peer_put / orbit_peer_wait with the GPU as its own peer: several transfers in one launch deliver their bytes and
    then their flags; a wait on delivered flags returns at once, a wait on a flag nobody writes gives up and reports
    OrbitStatus::peer_timeout (orbit_ctx_poll_status -> ORBIT_ERR_CUDA) instead of hanging."""
    import time
    from orbit_b200 import _lib, layouts as L
    ctx, lib = gpu_context, _lib.lib()
    dev = ctx.device
    n = 3
    sizes = [16 * 1000, 16 * 70001, 16]
    src = [torch.randint(0, 2 ** 31 - 1, (s // 4,), dtype=torch.int32, device=dev) for s in sizes]
    dst = [torch.zeros(s // 4, dtype=torch.int32, device=dev) for s in sizes]
    flags = torch.zeros(8, dtype=torch.int32, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for epoch in (1, 2):
        for d in dst:
            d.zero_()
        puts = (L.PeerPut * n)()
        for i in range(n):
            puts[i].src, puts[i].dst, puts[i].bytes, puts[i].dst_flag = src[i].data_ptr(), dst[i].data_ptr(), sizes[i], flags.data_ptr() + 4 * (2 * i)
        assert lib.orbit_peer_put(ctx._h, puts, n, epoch, stream) == 0
        assert lib.orbit_peer_wait(ctx._h, C.c_void_p(flags.data_ptr()), n, 2, epoch, stream) == 0
        torch.cuda.synchronize()
        assert flags.cpu().tolist() == [epoch, 0, epoch, 0, epoch, 0, 0, 0]
        assert all(torch.equal(a, b) for a, b in zip(src, dst))
    st = L.Status()
    assert lib.orbit_ctx_poll_status(ctx._h, C.byref(st)) == 0 and st.peer_timeout == 0
    # misaligned / oversized requests are refused
    puts[0].bytes = 24
    assert lib.orbit_peer_put(ctx._h, puts, 1, 3, stream) != 0
    assert lib.orbit_peer_put(ctx._h, puts, 17, 3, stream) != 0
    # a flag that never arrives: bounded wait, status word set
    t0 = time.time()
    assert lib.orbit_peer_wait(ctx._h, C.c_void_p(flags.data_ptr() + 4), 1, 1, 99, stream) == 0
    torch.cuda.synchronize()
    assert time.time() - t0 < 30.0
    rc = lib.orbit_ctx_poll_status(ctx._h, C.byref(st))
    assert st.peer_timeout == 1 and rc != 0
    assert lib.orbit_ctx_poll_status(ctx._h, C.byref(st)) == 0 and st.peer_timeout == 0      # cleared by the poll
