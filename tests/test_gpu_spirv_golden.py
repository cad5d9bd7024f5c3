"""GPU: the CUDA path held directly against outputs of the reference's own shipped shaders (tests/golden/
spirv_reference.json, made by tests/golden/make_spirv_golden.py with the SPIR-V interpreter in oracle/spirv_vm) — no
oracle in between. Hi-Z pyramids and visibility words byte for byte; records / commands / light lists as sorted sets."""
import json
import os

import numpy as np
import pytest

import spirv_cases as S
from orbit_b200 import layouts as L

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "spirv_reference.json")) as f:
        return json.load(f)


def test_cuda_hiz_matches_shipped_depth_reduce(gpu_context, golden):
    import torch
    from orbit_b200.passes import DepthPyramid
    for name, depth in S.hiz_cases().items():
        h, w = depth.shape
        pyr = DepthPyramid(gpu_context, "spv_" + name, (w, h))
        pyr.update(torch.from_numpy(depth).to(gpu_context.device))
        torch.cuda.synchronize()
        assert S.matches(golden["hiz"][name], pyr.texels.cpu().numpy()), name


def test_cuda_cull_passes_match_shipped_shaders(gpu_context, golden):
    import torch
    from orbit_b200 import frame
    from orbit_b200.passes import OcclusionCullInfo
    ctx = gpu_context
    for name, (sc, view, depth, mocc, frames, protocol) in S.cull_cases().items():
        g = golden["cull"][name]
        ds = frame.DeviceScene.upload(ctx, sc)
        vs = frame.ViewState(ctx, ds, (view.width, view.height), name="spv_" + name)
        d_depth = torch.from_numpy(depth).to(ctx.device) if depth is not None else None
        mvis = vs.meshlet_visibility if mocc else None
        opts = S.CULL_OPTS.get(name, {})
        k = 0
        for f in range(frames):
            passes = {"two_pass": [("early", "read"), ("late", "write")], "pass0": [("pass0", "none")], "pass2_only": [("late", "write")]}[protocol]
            for label, kind in passes:
                if kind == "write":
                    vs.depth_pyramid.update(d_depth)
                    oc = OcclusionCullInfo("write", vs.entity_visibility, mvis, vs.depth_pyramid, noskip_alphamode=opts.get("noskip", 0), aspect_ratio=view.aspect)
                elif kind == "read":
                    oc = OcclusionCullInfo("read", vs.entity_visibility, mvis)
                else:
                    oc = OcclusionCullInfo("none")
                info = frame.cull_info_for(view, oc)
                if "alpha_filter" in opts:
                    info.alpha_mode_filter = opts["alpha_filter"]
                disp, draws = frame.cull_pass(ctx, "spv_%s_%s" % (name, label), ds, info)
                torch.cuda.synchronize()
                step = g["steps"][k]; k += 1
                hdr, recs = S.canon_records(disp.cpu().numpy())
                n, cmds = S.canon_draws(draws.cpu().numpy())
                assert hdr == step["dispatch_header"], (name, f, label)
                assert S.matches(step["records"], recs), (name, f, label, "records")
                assert n == step["draw_count"] and S.matches(step["draws"], cmds), (name, f, label, "draws")
                assert S.matches(step["entity_visibility"], vs.entity_visibility.cpu().numpy().view(np.uint32)), (name, f, label, "entity visibility")
                assert S.matches(step["meshlet_visibility"], vs.meshlet_visibility.cpu().numpy().view(np.uint32)), (name, f, label, "meshlet visibility")
        if g["hiz"] is not None:
            assert S.matches(g["hiz"], vs.depth_pyramid.texels.cpu().numpy()), (name, "hiz")


def test_cuda_clusters_match_shipped_light_cluster_shaders(gpu_context, golden):
    import ctypes as C
    import torch
    from orbit_b200 import _lib
    ctx = gpu_context
    dev = ctx.device
    for name, (p, depth, lights) in S.cluster_cases().items():
        g = golden["clusters"][name]
        cx, cy, cz = p.info.cluster_count[0], p.info.cluster_count[1], p.info.cluster_count[2]
        n = cx * cy * cz
        cap = L.MAX_LIGHTS_PER_CLUSTER * n
        t = lambda nbytes: torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        masks, bounds, unique, image, index = t(4 * cx * cy), t(8 * n), t(16 + 4 * n), t(8 * n), t(4 + 4 * cap)
        d_depth = torch.from_numpy(np.ascontiguousarray(depth, np.float32)).to(dev)
        d_lights = torch.from_numpy(np.ascontiguousarray(lights).view(np.uint8).reshape(-1)).to(dev)
        ptr = lambda x: C.c_void_p(x.data_ptr())
        _lib.check(_lib.lib().orbit_light_cluster(ctx._h, C.byref(p), ptr(d_depth), ptr(d_lights), ptr(masks), ptr(bounds), ptr(unique), ptr(image),
                                                  ptr(index), cap, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "orbit_light_cluster")
        torch.cuda.synchronize()
        u32 = lambda x: x.cpu().numpy().view(np.uint32)
        c = S.canon_clusters({"masks": u32(masks), "bounds": u32(bounds), "unique": u32(unique), "image": u32(image), "index": u32(index)})
        assert c["header"] == g["header"] and c["total"] == g["total"], name
        for k in ("masks", "bounds", "active", "counts", "lists"):
            assert S.matches(g[k], c[k]), (name, k)


def test_cuda_task_payloads_match_shipped_task_shaders(gpu_context, golden):
    import torch
    from orbit_b200 import frame
    from orbit_b200.passes import OcclusionCullInfo
    ctx = gpu_context
    cases = S.cull_cases()
    for name, kind in (("ortho_pass0", "none"), ("ortho_pass2", "write"), ("persp_two_pass_entity_occlusion_only", "write")):
        sc, view, depth, mocc, _, _ = cases[name]
        ds = frame.DeviceScene.upload(ctx, sc)
        vs = frame.ViewState(ctx, ds, (view.width, view.height), name="spvt_" + name)
        mvis = vs.meshlet_visibility if mocc else None
        if kind == "write":
            vs.depth_pyramid.update(torch.from_numpy(depth).to(ctx.device))
            oc = OcclusionCullInfo("write", vs.entity_visibility, mvis, vs.depth_pyramid, noskip_alphamode=0, aspect_ratio=view.aspect)
        else:
            oc = OcclusionCullInfo("none")
        payloads = torch.zeros(L.TASK_PAYLOAD_STRIDE * sc.n_records_lod0, dtype=torch.uint8, device=ctx.device)
        disp, _ = frame.cull_pass(ctx, "spvt_" + name, ds, frame.cull_info_for(view, oc), task_payloads=payloads)
        torch.cuda.synchronize()
        nrec = int(disp[:4].cpu().numpy().view(np.uint32)[0])
        pl = S.canon_payload_buffer(payloads.cpu().numpy(), nrec)
        for shader, entry in golden["task"][name].items():
            assert S.matches(entry["payloads"], pl), (name, shader)
            if "meshlet_visibility_comp_semantics" in entry:
                assert S.matches(entry["meshlet_visibility_comp_semantics"], vs.meshlet_visibility.cpu().numpy().view(np.uint32)), (name, shader)
