"""Generates tests/golden/spirv_reference.json: outputs of the reference's own SHIPPED shaders
(/root/reference/shaders/{depth_reduce,entity_cull,meshlet_cull}.comp.spv, light_cluster/*.comp.spv) executed by the
SPIR-V interpreter in oracle/spirv_vm on the cases of tests/spirv_cases.py, dispatched the way draw_gen.rs / cluster.rs do.

Run in the build container (needs /root/reference):   python tests/golden/make_spirv_golden.py
The fixtures pin the oracle (tests/test_spirv_golden.py) and the CUDA path (tests/test_gpu_spirv_golden.py) against the
reference's GPU programs themselves. Implementation-defined pieces (OpDot order, Log2, sampler footprint) are fixed as
stated in oracle/spirv_vm/spirv_vm.py — the same choices as DESIGN.md §3."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "spirv_vm")):
    sys.path.insert(0, p)
import oracle_ref as O  # noqa: E402  (only for the contract log2 and GpuCullInfo packing — host logic)
import reference_passes as R  # noqa: E402
import spirv_cases as S  # noqa: E402


def main():
    assert R.available(), "needs /root/reference/shaders/*.spv"
    O.build()
    log2f = lambda x: np.float32(O.log2f(float(x)))
    out = {"generator": "tests/golden/make_spirv_golden.py", "shaders": "Thefefe/orbit shaders/*.comp.spv (shipped binaries)", "hiz": {}, "cull": {}, "clusters": {}}
    t0 = time.time()
    for name, depth in S.hiz_cases().items():
        h, w = depth.shape
        info = O.hiz_geometry(w, h)
        levels = R.hiz_build(depth, info, log2f)
        out["hiz"][name] = S.pack(np.concatenate([l.reshape(-1) for l in levels]))
        print("hiz", name, "%.0fs" % (time.time() - t0), flush=True)
    for name, (sc, view, depth, mocc, frames, protocol) in S.cull_cases().items():
        ev = np.zeros((sc.n_entities + 31) // 32 + 1, np.uint32)
        mv = np.zeros(max(sc.n_visibility_words, 1), np.uint32)
        steps = []
        levels = None
        if depth is not None:
            info = O.hiz_geometry(view.width, view.height)
        for f in range(frames):
            passes = {"two_pass": [("early", "read"), ("late", "write")], "pass0": [("pass0", "none")], "pass2_only": [("late", "write")]}[protocol]
            for label, kind in passes:
                if kind == "write":
                    levels = R.hiz_build(depth, info, log2f)
                g = S.tweak_gpu_cull_info(O.gpu_cull_info(view, kind, mocc), name)
                pyr = levels if kind == "write" else None
                disp = R.entity_cull(sc, g, ev, mv, pyr, sc.n_records_lod0, log2f)
                draws = R.meshlet_cull(sc, g, ev, mv, pyr, disp, sc.n_meshlet_instances, log2f)
                hdr, recs = S.canon_records(disp)
                n, d = S.canon_draws(draws)
                steps.append({"frame": f, "pass": label, "dispatch_header": hdr, "records": S.pack(recs), "draw_count": n, "draws": S.pack(d),
                              "entity_visibility": S.pack(ev), "meshlet_visibility": S.pack(mv)})
                print("cull", name, f, label, "records", hdr[0], "draws", n, "%.0fs" % (time.time() - t0), flush=True)
        out["cull"][name] = {"steps": steps, "hiz": S.pack(np.concatenate([l.reshape(-1) for l in levels])) if levels is not None else None}
    # ---- mesh-shading path: the three shipped task shaders (payload = what orbit_meshlet_cull's task_payloads output restates)
    out["task"] = {}
    cases = S.cull_cases()
    for name, kind in (("ortho_pass0", "none"), ("ortho_pass2", "write"), ("persp_two_pass_entity_occlusion_only", "write")):
        sc, view, depth, mocc, _, _ = cases[name]
        ev = np.zeros((sc.n_entities + 31) // 32 + 1, np.uint32)
        mv = np.zeros(max(sc.n_visibility_words, 1), np.uint32)
        levels = R.hiz_build(depth, O.hiz_geometry(view.width, view.height), log2f) if kind == "write" else None
        g = O.gpu_cull_info(view, kind, mocc)
        disp = R.entity_cull(sc, g, ev, mv, levels, sc.n_records_lod0, log2f)
        recs = S.canon_records(disp)[1]
        out["task"][name] = {}
        for shader in (S.TASK_SHADERS if kind == "none" else S.TASK_SHADERS[:1]):
            mv_t = mv.copy()
            res = R.task_shader(sc, g, ev.copy(), mv_t, levels, disp, log2f, shader)
            pl = S.canon_payloads([(c, p[0], p[1], p[2]) for (_, c, p) in res])
            entry = {"payloads": S.pack(pl), "tasks": int(pl["task_count"].sum())}
            if kind == "write" and mocc:
                # the task shaders leave `visible` = true for lanes >= meshlet_count, so the words they store have the unused
                # high bits set; the compute shader's words have them clear. The bits are never read.
                mv_c = mv.copy()
                R.meshlet_cull(sc, g, ev.copy(), mv_c, levels, disp, sc.n_meshlet_instances, log2f)
                for r in recs:
                    cnt = int(r["meshlet_count"]); used = 0xFFFFFFFF if cnt >= 32 else (1 << cnt) - 1
                    wt, wc = int(mv_t[r["visibility_offset"]]), int(mv_c[r["visibility_offset"]])
                    assert wt & used == wc & used and wt | used == 0xFFFFFFFF and wc & ~used & 0xFFFFFFFF == 0
                entry["meshlet_visibility_comp_semantics"] = S.pack(mv_c)
            out["task"][name][shader] = entry
            print("task", name, shader, "records", len(pl), "tasks", entry["tasks"], "%.0fs" % (time.time() - t0), flush=True)
    for name, (p, depth, lights) in S.cluster_cases().items():
        c = S.canon_clusters(R.light_cluster(p, depth, lights, log2f))
        out["clusters"][name] = {"header": c["header"], "total": c["total"], "masks": S.pack(c["masks"]), "bounds": S.pack(c["bounds"]),
                                 "active": S.pack(c["active"]), "counts": S.pack(c["counts"]), "lists": S.pack(c["lists"])}
        print("clusters", name, c["header"], c["total"], "%.0fs" % (time.time() - t0), flush=True)
    path = os.path.join(ROOT, "tests", "golden", "spirv_reference.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
