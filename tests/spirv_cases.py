"""Cases shared by the SPIR-V fixture generator (tests/golden/make_spirv_golden.py, runs the reference's shipped shaders
in oracle/spirv_vm) and by the tests that hold the oracle and the CUDA path against those fixtures."""
import base64
import hashlib
import zlib

import numpy as np

from orbit_b200 import layouts as L
from orbit_b200 import scenes

REC_KEY = ["entity_index", "meshlet_offset"]
DRAW_KEY = ["cmd_first_instance", "meshlet_index"]


def _persp(w, h, lod=(10.0, 1.6)):
    v = scenes.perspective_view((-6.0, 2.0, -6.0), (np.sin(np.radians(30.0)), 0.0, np.cos(np.radians(30.0))), w, h)
    v.lod_base, v.lod_step = lod
    return v


def cull_cases():
    """name -> (scene, view, depth, meshlet_occlusion, frames, protocol). protocol "two_pass" = early / Hi-Z / late per frame,
    "pass0" = one pass without occlusion, "pass2_only" = Hi-Z + one write pass on zeroed visibility."""
    out = {}
    sc, _ = scenes.config_c1(scale=0.12, lods=(100, 40))
    v = _persp(320, 180)
    out["persp_two_pass"] = (sc, v, scenes.make_depth(sc, v), True, 2, "two_pass")
    sc, _ = scenes.config_c1(scale=0.06, lods=(100,))
    v = _persp(256, 144, lod=(16.0, 2.0))
    out["persp_two_pass_entity_occlusion_only"] = (sc, v, scenes.make_depth(sc, v), False, 2, "two_pass")
    sc, _ = scenes.config_c1(scale=0.08, lods=(77, 33, 9))          # ragged record sizes, three LODs
    centre = (sc.aabb_min + sc.aabb_max) / 2
    d = np.array([1.0, -1.0, -1.0]); d /= np.linalg.norm(d)
    v = scenes.orthographic_view(centre - d * 100.0, d, 256, 256, half_width=60.0, near=-20.0, far=250.0)
    out["ortho_pass0"] = (sc, v, None, True, 1, "pass0")
    out["ortho_pass2"] = (sc, v, scenes.make_depth(sc, v), True, 1, "pass2_only")
    # 12 cull planes, LOD range clamped to [1, 2], model matrices with a bottom row != (0,0,0,1) (the p / p.w division)
    sc, _ = scenes.config_c1(scale=0.08, lods=(120, 60, 30, 15))
    m = sc.entities["model_matrix"]            # [n][col][row]
    rng = np.random.default_rng(5)
    m[:, 3, 3] = rng.uniform(0.7, 1.4, len(m)).astype(np.float32)
    m[::3, 0, 3] = rng.uniform(-0.01, 0.01, len(m[::3])).astype(np.float32)
    v = _persp(320, 180, lod=(4.0, 1.3))
    extra = []
    for _ in range(7):
        nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
        extra.append([nrm[0], nrm[1], nrm[2], 40.0])
    v.planes = np.vstack([v.planes, np.array(extra)])
    v.lod_range = (1, 3)
    out["persp_planes12_nonaffine_lodclamp"] = (sc, v, scenes.make_depth(sc, v), True, 2, "two_pass")
    # every alpha mode passes the filter, masked + transparent are "noskip" (drawn in pass 2 even if visible last frame)
    sc, _ = scenes.config_c1(scale=0.06, lods=(100, 40))
    v = _persp(256, 144)
    out["persp_noskip_alpha"] = (sc, v, scenes.make_depth(sc, v), True, 2, "two_pass")
    # BASELINE config C1 at full size (1 k entities / 100 k meshlets; the reference's own CPU-runnable case), 640x360 view
    sc, _ = scenes.config_c1()
    v = _persp(640, 360, lod=(16.0, 2.0))
    out["c1_full_size"] = (sc, v, scenes.make_depth(sc, v), True, 2, "two_pass")
    return out


# per-case CullInfo fields that differ from the callers' defaults (draw_gen.rs:105-119)
CULL_OPTS = {"persp_noskip_alpha": {"alpha_filter": 1 | 2 | 4, "noskip": 2 | 4}}


def tweak_gpu_cull_info(g, name):
    """Applies CULL_OPTS to a packed GpuCullInfo the way CullInfo::to_gpu would have (noskip only exists in pass 2)."""
    o = CULL_OPTS.get(name, {})
    if "alpha_filter" in o:
        g.alpha_mode_flags = o["alpha_filter"]
    if "noskip" in o and g.occlusion_pass == 2:
        g.noskip_alpha_mode = o["noskip"]
    return g


def cluster_cases():
    """name -> (ClusterParams, depth, lights)"""
    out = {}
    sc, _ = scenes.config_c1(scale=0.06)
    for name, (w, h, tile, cz, far, nl, seed) in {"small_grid": (160, 90, 20, 12, 120.0, 300, 7),
                                                   "cap_256": (64, 48, 16, 6, 60.0, 700, 9)}.items():
        view = scenes.perspective_view((-6.0, 2.0, -6.0), (np.sin(np.radians(30.0)), 0.0, np.cos(np.radians(30.0))), w, h)
        depth = scenes.make_depth(sc, view)
        lights = scenes.make_lights(seed, nl, sc.aabb_min, sc.aabb_max)
        if name == "cap_256":
            lights["outer_radius"] = 500.0            # every light reaches every cluster: the 256-per-cluster cap decides
        p = L.ClusterParams()
        cx, cy = -(-w // tile), -(-h // tile)
        p.info.world_to_view_matrix.set(view.view)
        p.info.screen_to_view_matrix.set(np.linalg.inv(np.asarray(view.projection_matrix, np.float64)))
        p.info.cluster_count[0], p.info.cluster_count[1], p.info.cluster_count[2] = cx, cy, cz
        p.info.tile_size_px = tile
        p.info.screen_size[0], p.info.screen_size[1] = w, h
        p.info.z_near, p.info.z_far = float(view.near), far
        p.info.global_light_count = len(lights)
        p.z_scale, p.z_bias = scenes.cluster_grid_info(view.near, far, cz)
        out[name] = (p, depth, lights)
    return out


def hiz_cases():
    rng = np.random.default_rng(11)
    return {"%dx%d" % (w, h): rng.random((h, w), dtype=np.float32) for (w, h) in ((50, 38), (33, 7), (130, 70), (64, 64), (3, 200))}


# ---- canonical forms (the reference appends with atomicAdd: order is arbitrary, so lists are compared sorted) ----
def pack(a, keep_bytes_below=12288):
    """count + SHA-256 of the array's bytes; the (compressed) bytes themselves too when they are small, so that a
    mismatch on the small arrays can be diagnosed without the reference."""
    b = np.ascontiguousarray(a).view(np.uint8).tobytes()
    d = {"n": int(len(a)), "sha256": hashlib.sha256(b).hexdigest()}
    z = base64.b64encode(zlib.compress(b, 9)).decode()
    if len(z) <= keep_bytes_below:
        d["z"] = z
    return d


def matches(d, a):
    """True when array `a` has exactly the bytes the fixture entry `d` was made from."""
    b = np.ascontiguousarray(a).view(np.uint8).tobytes()
    return int(len(a)) == d["n"] and hashlib.sha256(b).hexdigest() == d["sha256"]


def unpack(d, dtype):
    return np.frombuffer(zlib.decompress(base64.b64decode(d["z"])), dtype=dtype)


def canon_records(dispatch_bytes):
    hdr = np.frombuffer(bytes(dispatch_bytes[:12]), np.uint32)
    recs = np.frombuffer(bytes(dispatch_bytes[12:12 + 16 * int(hdr[0])]), L.dispatch_dtype)
    return hdr.tolist(), np.sort(recs, order=REC_KEY)


def canon_draws(draw_bytes):
    n = int(np.frombuffer(bytes(draw_bytes[:4]), np.uint32)[0])
    d = np.frombuffer(bytes(draw_bytes[4:4 + 28 * n]), L.draw_command_dtype)
    return n, np.sort(d, order=DRAW_KEY)


def canon_clusters(res):
    """{masks, bounds, unique, image, index} -> order-independent form."""
    na = int(res["unique"][3])
    active = np.sort(np.asarray(res["unique"][4:4 + na], np.uint32))
    img = np.asarray(res["image"], np.uint32).reshape(-1, 2)
    lists = []
    for idx in active:
        off, cnt = int(img[idx][0]), int(img[idx][1])
        lists.append(np.asarray(res["index"][1 + off:1 + off + cnt], np.uint32))
    flat = np.concatenate(lists) if lists else np.zeros(0, np.uint32)
    counts = np.array([len(x) for x in lists], np.uint32)
    return {"header": [int(x) for x in res["unique"][:4]], "masks": np.asarray(res["masks"], np.uint32), "bounds": np.asarray(res["bounds"], np.uint32),
            "active": active, "counts": counts, "lists": flat, "total": int(res["index"][0])}


TASK_SHADERS = ("forward/forward_depth_prepass.task.spv", "forward/forward.task.spv", "shadow/shadow.task.spv")


def canon_payloads(entries):
    """entries: iterable of (task_count, entity_index, meshlet_offset, indices) -> task_payload_dtype array sorted by
    (entity_index, meshlet_offset), index bytes beyond the count zeroed (the shader leaves them undefined)."""
    out = np.zeros(len(entries), L.task_payload_dtype)
    for i, (c, e, o, idx) in enumerate(entries):
        out[i]["task_count"], out[i]["entity_index"], out[i]["meshlet_offset"] = c, e, o
        out[i]["meshlet_indices"][:c] = np.asarray(idx[:c], np.uint8)
    return np.sort(out, order=["entity_index", "meshlet_offset"])


def canon_payload_buffer(buf, nrec):
    tp = np.frombuffer(bytes(buf), L.task_payload_dtype)[:nrec]
    return canon_payloads([(int(t["task_count"]), int(t["entity_index"]), int(t["meshlet_offset"]), t["meshlet_indices"].tolist()) for t in tp])
