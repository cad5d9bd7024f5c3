"""How much do the implementation-defined pieces of SPIR-V matter on this path? Re-runs the reference's shipped shaders
in the interpreter (oracle/spirv_vm) on one fixture case with (a) a*b+c chains of OpDot / OpMatrixTimesVector fused, as
an optimising driver compiler would, and (b) Log2 evaluated in double precision and rounded (≈ a correctly rounded
log2f) instead of the contract's orbit_log2f, and counts the outputs that change against the committed fixtures.
Build-container tool (needs /root/reference); prints one JSON line per variant."""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "spirv_vm")):
    sys.path.insert(0, p)
import oracle_ref as O  # noqa: E402
import reference_passes as R  # noqa: E402
import spirv_cases as S  # noqa: E402


PROTOCOLS = {"two_pass": [("early", "read"), ("late", "write")], "pass0": [("pass0", "none")], "pass2_only": [("late", "write")]}


def run(name, log2f):
    """The shipped shaders through the interpreter on one fixture case, every protocol of tests/spirv_cases.py."""
    sc, view, depth, mocc, frames, protocol = S.cull_cases()[name]
    ev = np.zeros((sc.n_entities + 31) // 32 + 1, np.uint32)
    mv = np.zeros(max(sc.n_visibility_words, 1), np.uint32)
    info = O.hiz_geometry(view.width, view.height) if depth is not None else None
    out = []
    for f in range(frames):
        for label, kind in PROTOCOLS[protocol]:
            levels = R.hiz_build(depth, info, log2f) if kind == "write" else None
            g = S.tweak_gpu_cull_info(O.gpu_cull_info(view, kind, mocc), name)
            disp = R.entity_cull(sc, g, ev, mv, levels, sc.n_records_lod0, log2f)
            draws = R.meshlet_cull(sc, g, ev, mv, levels, disp, sc.n_meshlet_instances, log2f)
            out.append((f, label, S.canon_records(disp)[1], S.canon_draws(draws)[1]))
    return out


def near_threshold_report(name):
    """The north star's 1e-5 report for one case: predicates the ORACLE evaluated whose value lies within 1e-5 (relative) of its
    threshold — the only ones a real driver's arithmetic could legitimately flip."""
    sc, view, depth, mocc, frames, protocol = S.cull_cases()[name]
    hs = O.HostScene(sc)
    st = O.Stats()
    for f in range(frames):
        for label, kind in PROTOCOLS[protocol]:
            if kind == "write":
                hs.update_pyramid(depth)
            g = S.tweak_gpu_cull_info(O.gpu_cull_info(view, kind, mocc), name)
            O.cull_pass(hs, g, stats=st)
    d = st.as_dict()
    return {"case": name, "lanes_tested": d["lanes"], "records": d["records"],
            "within_1e-5_of_threshold": {k: d[k] for k in ("near_plane", "near_cone", "near_cullable", "near_depth", "near_hiz_level", "near_lod")}}


def main():
    names = sys.argv[1:] or list(S.cull_cases())
    O.build()
    contract_log2 = lambda x: np.float32(O.log2f(float(x)))
    exact_log2 = lambda x: np.float32(math.log2(float(x))) if x > 0 else np.float32(-np.inf if x == 0 else np.nan)
    for name in names:
        print(json.dumps(near_threshold_report(name)), flush=True)
        base = run(name, contract_log2)
        for label, fma, lg in (("fused a*b+c chains in OpDot / OpMatrixTimesVector", True, contract_log2),
                               ("Log2 correctly rounded instead of orbit_log2f", False, exact_log2),
                               ("both", True, exact_log2)):
            R.CONTRACT_FMA = fma
            res = run(name, lg)
            R.CONTRACT_FMA = False
            diff_r = diff_d = tot_r = tot_d = 0
            for (f, l, r0, d0), (_, _, r1, d1) in zip(base, res):
                a = set(map(bytes, r0)); b = set(map(bytes, r1)); diff_r += len(a ^ b); tot_r += len(a)
                a = set(map(bytes, d0)); b = set(map(bytes, d1)); diff_d += len(a ^ b); tot_d += len(a)
            print(json.dumps({"case": name, "variant": label, "records": tot_r, "records_changed": diff_r, "draws": tot_d, "draws_changed": diff_d}), flush=True)


if __name__ == "__main__":
    main()
