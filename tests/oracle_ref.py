"""ctypes wrapper of the CPU oracle (oracle/_build/liborbit_oracle.so) — TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker /
CPU baseline; never by orbit_b200. Works on numpy arrays in the reference layouts.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from orbit_b200 import layouts as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liborbit_oracle.so")


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("near_plane", "near_cone", "near_cullable", "near_depth", "near_hiz_level",
                                          "near_lod", "lanes", "records", "survivors", "visible")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def build(force=False):
    src = os.path.join(ORACLE_DIR, "orbit_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        res = subprocess.run(["make", "-C", ORACLE_DIR], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        h = C.CDLL(LIB_PATH)
        vp, u32, u64, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float
        h.oracle_threads.restype = C.c_int
        h.oracle_set_threads.restype = C.c_int; h.oracle_set_threads.argtypes = [C.c_int]
        h.oracle_hiz_geometry.argtypes = [u32, u32, C.POINTER(L.HizInfo)]
        h.oracle_log2f.restype = f32; h.oracle_log2f.argtypes = [f32]
        h.oracle_hiz_level.restype = u32; h.oracle_hiz_level.argtypes = [f32, u32]
        h.oracle_hiz_build.argtypes = [vp, u32, u32, vp]
        h.oracle_hiz_sample.restype = f32; h.oracle_hiz_sample.argtypes = [vp, u32, u32, f32, f32, f32]
        h.oracle_entity_cull.restype = u64
        h.oracle_entity_cull.argtypes = [C.POINTER(L.CullInfo), C.POINTER(L.SceneBuffers), vp, u32, u32, vp, u64, C.POINTER(Stats)]
        h.oracle_meshlet_cull.restype = u64
        h.oracle_meshlet_cull.argtypes = [C.POINTER(L.CullInfo), C.POINTER(L.SceneBuffers), vp, u32, u32, vp, vp, u64, vp, C.POINTER(Stats)]
        h.oracle_meshlet_bounds.restype = C.c_int; h.oracle_meshlet_bounds.argtypes = [vp, u32, vp, vp, u32]
        h.oracle_mesh_bounds.restype = None; h.oracle_mesh_bounds.argtypes = [vp, u32, vp, vp, u32]
        h.oracle_mark_active.argtypes = [C.POINTER(L.ClusterParams), vp, vp, vp]
        h.oracle_compact_clusters.restype = u32; h.oracle_compact_clusters.argtypes = [C.POINTER(L.ClusterParams), vp, vp]
        h.oracle_light_culling.restype = u64
        h.oracle_light_culling.argtypes = [C.POINTER(L.ClusterParams), vp, vp, vp, vp, vp, u64]
        h.oracle_scene_update.restype = C.c_int
        h.oracle_scene_update.argtypes = [vp, vp, vp, vp, vp, u32, u32, vp, vp]
        _lib = h
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def hiz_geometry(w, h):
    info = L.HizInfo()
    lib().oracle_hiz_geometry(w, h, C.byref(info))
    return info


def hiz_build(depth):
    depth = np.ascontiguousarray(depth, np.float32)
    h, w = depth.shape
    info = hiz_geometry(w, h)
    texels = np.zeros(info.total_texels, np.float32)
    lib().oracle_hiz_build(_p(depth), w, h, _p(texels))
    return info, texels


def log2f(x):
    return float(lib().oracle_log2f(C.c_float(x)))


class HostScene:
    """numpy-side twin of frame.DeviceScene + ViewState: the arrays the oracle reads / writes in place."""

    def __init__(self, s, draw_begin=0, draw_end=0):
        self.s = s
        self.entity_visibility = np.zeros((s.n_entities + 31) // 32 + 1, np.uint32)
        self.meshlet_visibility = np.zeros(max(s.n_visibility_words, 1), np.uint32)
        self.draw_begin, self.draw_end = draw_begin, draw_end
        self.hiz_info = None
        self.hiz_texels = None
        self.depth_size = None

    def buffers(self, use_entity_vis=True, use_meshlet_vis=True):
        s = self.s
        sb = L.SceneBuffers()
        sb.entity_draws = s.entity_draws.ctypes.data
        sb.mesh_infos = s.mesh_infos.ctypes.data
        sb.entities = s.entities.ctypes.data
        sb.meshlets = s.meshlets.ctypes.data
        sb.materials = s.materials.ctypes.data
        sb.entity_visibility = self.entity_visibility.ctypes.data if use_entity_vis else 0
        sb.meshlet_visibility = self.meshlet_visibility.ctypes.data if use_meshlet_vis else 0
        sb.entity_draw_count = s.n_entities
        sb.draw_begin, sb.draw_end = self.draw_begin, self.draw_end
        return sb

    def reset_visibility(self):
        """Frame 0 of another view: both bitmasks back to zero (the harness's definition of the never-initialised masks)."""
        self.entity_visibility[:] = 0
        self.meshlet_visibility[:] = 0

    def update_pyramid(self, depth):
        self.hiz_info, self.hiz_texels = hiz_build(depth)
        self.depth_size = (depth.shape[1], depth.shape[0])


def cull_pass(hs, g, record_capacity=None, draw_capacity=None, stats=None, task_payloads=False):
    """entity stage + meshlet stage with GpuCullInfo `g` (a layouts.CullInfo). Returns (dispatch bytes, draw bytes[, payloads])."""
    s = hs.s
    rcap = record_capacity if record_capacity is not None else s.n_records_lod0
    dcap = draw_capacity if draw_capacity is not None else s.n_meshlet_instances
    stats = stats if stats is not None else Stats()
    mocc = g.meshlet_visibility_buffer != L.NO_BUFFER
    sb = hs.buffers(use_meshlet_vis=mocc)
    dispatch = np.zeros(12 + 16 * rcap, np.uint8)
    draws = np.zeros(4 + 28 * dcap, np.uint8)
    dw, dh = hs.depth_size if hs.depth_size else (0, 0)
    lib().oracle_entity_cull(C.byref(g), C.byref(sb), _p(hs.hiz_texels), dw, dh, _p(dispatch), rcap, C.byref(stats))
    payload = np.zeros(44 * max(rcap, 1), np.uint8) if task_payloads else None
    lib().oracle_meshlet_cull(C.byref(g), C.byref(sb), _p(hs.hiz_texels), dw, dh, _p(dispatch), _p(draws), dcap, _p(payload), C.byref(stats))
    if task_payloads:
        return dispatch, draws, payload
    return dispatch, draws


def parse_dispatch(buf):
    hdr = buf[:12].view(np.uint32).copy()
    return hdr, buf[12:12 + 16 * int(hdr[0])].view(L.dispatch_dtype)


def parse_draws(buf, capacity=None):
    n = int(buf[:4].view(np.uint32)[0])
    m = n if capacity is None else min(n, capacity)
    return n, buf[4:4 + 28 * m].view(L.draw_command_dtype)


def light_cluster(params, depth, lights):
    """mark_active + compaction + light_culling. Returns dict of numpy outputs."""
    ci = params.info
    cx, cy, cz = ci.cluster_count[0], ci.cluster_count[1], ci.cluster_count[2]
    n = cx * cy * cz
    depth = np.ascontiguousarray(depth, np.float32)
    masks = np.zeros(cx * cy, np.uint32)
    bounds = np.zeros(2 * n, np.uint32)
    lib().oracle_mark_active(C.byref(params), _p(depth), _p(masks), _p(bounds))
    unique = np.zeros(4 + n, np.uint32)
    lib().oracle_compact_clusters(C.byref(params), _p(masks), _p(unique))
    image = np.zeros(2 * n, np.uint32)
    cap = L.MAX_LIGHTS_PER_CLUSTER * n
    index = np.zeros(1 + cap, np.uint32)
    lights = np.ascontiguousarray(lights)
    lib().oracle_light_culling(C.byref(params), _p(lights), _p(bounds), _p(unique), _p(image), _p(index), cap)
    return {"masks": masks, "bounds": bounds, "unique": unique, "image": image, "index": index}


# ---- the caller protocol on the CPU (twin of orbit_b200.frame) ---------------------------------------------
def gpu_cull_info(view, kind, meshlet_occlusion=True, frustum_culling=True):
    """GpuCullInfo bytes for one pass of `view` through the product's own CullInfo.to_gpu (host logic)."""
    from orbit_b200.frame import cull_info_for
    from orbit_b200.passes import OcclusionCullInfo
    marker = object()
    oc = OcclusionCullInfo(kind, None if kind == "none" else marker,
                           marker if (meshlet_occlusion and kind != "none") else None,
                           marker if kind == "write" else None, 0, view.aspect)
    return cull_info_for(view, oc, frustum_culling).to_gpu()


def depth_prepass_culling(hs, view, depth, meshlet_occlusion=True, stats=None):
    out = {"early": cull_pass(hs, gpu_cull_info(view, "read", meshlet_occlusion), stats=stats)}
    hs.update_pyramid(depth)
    out["late"] = cull_pass(hs, gpu_cull_info(view, "write", meshlet_occlusion), stats=stats)
    return out


def main_pass_culling(hs, view, meshlet_occlusion=True, stats=None):
    return cull_pass(hs, gpu_cull_info(view, "read", meshlet_occlusion), stats=stats)


def scene_update(transforms, mesh_slots, visibility_offsets, mesh_infos, cursor, capacity_words=1 << 26):
    """SceneData::update_scene (scene.rs:404-492), mesh part. visibility_offsets (uint32[n]) and cursor (uint32[1]) are
    updated in place. Returns (entity data array, EntityDrawBuffer bytes, overflow flag)."""
    n = len(transforms)
    assert transforms.dtype == L.transform_dtype and visibility_offsets.dtype == np.uint32 and cursor.dtype == np.uint32
    t = np.ascontiguousarray(transforms)
    slots = np.ascontiguousarray(mesh_slots, dtype=np.uint32)
    entity_data = np.zeros(n, L.entity_dtype)
    draws = np.zeros(4 + 12 * n, np.uint8)
    ovf = lib().oracle_scene_update(_p(t), _p(slots), _p(visibility_offsets), _p(mesh_infos), _p(cursor), n, capacity_words,
                                    _p(entity_data), _p(draws))
    count = int(draws[:4].view(np.uint32)[0])
    return entity_data[:count], draws[:4 + 12 * count], int(ovf)


def asset_bounds(vertices, meshlet_data, meshlets, mesh_infos, vertex_ranges):
    """compute_meshlets' bounds (mesh.rs:321-338 = meshopt::compute_meshlet_bounds) + MeshData::compute_bounds (mesh.rs:192-215)
    on copies of `meshlets` / `mesh_infos`. Returns (meshlets, mesh_infos, meshlets skipped for > 128 triangles)."""
    ml, mi = np.ascontiguousarray(meshlets).copy(), np.ascontiguousarray(mesh_infos).copy()
    v, d, r = np.ascontiguousarray(vertices), np.ascontiguousarray(meshlet_data, dtype=np.uint32), np.ascontiguousarray(vertex_ranges, dtype=np.uint32)
    skipped = lib().oracle_meshlet_bounds(v.ctypes.data, v.dtype.itemsize, d.ctypes.data, ml.ctypes.data, len(ml))
    lib().oracle_mesh_bounds(v.ctypes.data, v.dtype.itemsize, r.ctypes.data, mi.ctypes.data, len(mi))
    return ml, mi, int(skipped)

