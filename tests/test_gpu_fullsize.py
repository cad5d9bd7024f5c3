"""GPU, BASELINE.json FULL sizes: C2 (10k entities / 2M meshlets, 1080p) and C3 (250k entities / 50M instanced
meshlets, 4K) — byte-exact against the oracle (it finishes a C3 frame in well under a second on the box's cores)
plus the size-independent properties of the two-pass protocol."""
import numpy as np
import pytest
import torch

from orbit_b200 import layouts as L
from orbit_b200 import scenes

pytestmark = pytest.mark.gpu


def _frames(ctx, oracle, sc, view, depth, n_frames=3):
    from orbit_b200 import frame
    ds = frame.DeviceScene.upload(ctx, sc)
    vs = frame.ViewState(ctx, ds, (view.width, view.height))
    hs = oracle.HostScene(sc)
    d_depth = torch.from_numpy(depth).to(ctx.device)
    out = []
    for f in range(n_frames):
        g = frame.depth_prepass_culling(ctx, ds, vs, view, d_depth)
        torch.cuda.synchronize()
        o = oracle.depth_prepass_culling(hs, view, depth)
        rec = {}
        for k in ("early", "late"):
            ghdr, grecs = frame.read_dispatch(g[k][0]); ohdr, orecs = oracle.parse_dispatch(o[k][0])
            gn, gd = frame.read_draws(g[k][1]); on, od = oracle.parse_draws(o[k][1])
            assert ghdr.tolist() == ohdr.tolist(), (f, k)
            assert np.array_equal(grecs.view(np.uint32), orecs.view(np.uint32)), (f, k, "records")
            assert gn == on and np.array_equal(gd.view(np.uint32), od.view(np.uint32)), (f, k, "draws")
            rec[k] = (grecs.copy(), gd.copy())
        mv = vs.meshlet_visibility.cpu().numpy().view(np.uint32)
        ev = vs.entity_visibility.cpu().numpy().view(np.uint32)
        assert np.array_equal(mv, hs.meshlet_visibility) and np.array_equal(ev, hs.entity_visibility), (f, "visibility words")
        assert np.array_equal(vs.depth_pyramid.texels.cpu().numpy().view(np.uint32), hs.hiz_texels.view(np.uint32)), (f, "Hi-Z")
        rec["mv"], rec["ev"] = mv.copy(), ev.copy()
        out.append(rec)
    code, _ = ctx.poll_status()
    assert code == 0
    return out


def _properties(sc, frames):
    f0, f1, f2 = frames
    # frame 0 starts from all-zero visibility: the early pass draws nothing, the late pass finds everything visible
    assert len(f0["early"][1]) == 0 and len(f0["late"][1]) > 0
    # steady state (static camera): the late pass finds nothing new, and frames repeat exactly (idempotence)
    assert len(f1["late"][1]) == 0 and len(f2["late"][1]) == 0
    assert np.array_equal(f1["early"][1].view(np.uint32), f2["early"][1].view(np.uint32))
    assert np.array_equal(f1["mv"], f2["mv"]) and np.array_equal(f1["ev"], f2["ev"])
    # canonical order: dispatch records ascend by (entity draw, meshlet offset); draw commands by (record, lane)
    recs, draws = f1["early"]
    key = recs["entity_index"].astype(np.int64) * (1 << 32) + recs["meshlet_offset"]
    assert np.all(np.diff(key) > 0)
    dkey = draws["cmd_first_instance"].astype(np.int64) * (1 << 32) + draws["meshlet_index"]
    assert np.all(np.diff(dkey) > 0)
    # every early draw of frame 1 was visible at the end of frame 0 (its bit is set), and alpha-filtered:
    # frame-0 late draws (= newly visible, ignores the alpha filter) is a superset of frame-1 early draws
    late0 = f0["late"][1]
    k0 = set((late0["cmd_first_instance"].astype(np.int64) * (1 << 32) + late0["meshlet_index"]).tolist())
    assert set(dkey.tolist()) <= k0
    # visibility popcount == number of frame-0 late draws (every visible meshlet was newly visible in frame 0)
    pop = int(np.unpackbits(f0["mv"].view(np.uint8)).sum())
    assert pop == len(late0)
    # draw command fields: instance_count == 1, index_count multiple of 3 in [96, 192]
    assert np.all(draws["cmd_instance_count"] == 1) and np.all(draws["cmd_index_count"] % 3 == 0)
    assert draws["cmd_index_count"].min() >= 96 and draws["cmd_index_count"].max() <= 192


def test_c2_full_size(gpu_context, oracle):
    sc, view = scenes.config_c2()
    assert sc.n_entities == 10000 and sc.n_meshlet_instances == 2_000_000
    depth = scenes.make_depth(sc, view)
    frames = _frames(gpu_context, oracle, sc, view, depth)
    _properties(sc, frames)


def test_c3_full_size(gpu_context, oracle):
    sc, view = scenes.config_c3()
    assert sc.n_entities == 250000 and sc.n_meshlet_instances == 50_000_000 and (view.width, view.height) == (3840, 2160)
    depth = scenes.make_depth(sc, view)
    frames = _frames(gpu_context, oracle, sc, view, depth)
    _properties(sc, frames)
