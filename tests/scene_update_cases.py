"""Shared inputs for the scene-update tests (CPU oracle tests and GPU parity tests)."""
import numpy as np

from orbit_b200 import layouts as L
from orbit_b200.scenes import Stream


def random_entities(n, n_meshes, seed, no_mesh_fraction=0.2, preallocated_fraction=0.0, negative_scale=True):
    """Random Transforms (unit quaternions rounded to f32, non-uniform and partly negative scales), mesh slots with holes,
    and optionally entities that already own a visibility range (as after an earlier frame)."""
    s = Stream(seed, "scene_update")
    t = np.zeros(n, L.transform_dtype)
    t["position"] = np.stack([s.uniform(n, -500.0, 500.0) for _ in range(3)], axis=1).astype(np.float32)
    g = np.stack([s.uniform(n) for _ in range(3)], axis=1)
    q = np.stack([np.sqrt(1 - g[:, 0]) * np.sin(2 * np.pi * g[:, 1]), np.sqrt(1 - g[:, 0]) * np.cos(2 * np.pi * g[:, 1]),
                  np.sqrt(g[:, 0]) * np.sin(2 * np.pi * g[:, 2]), np.sqrt(g[:, 0]) * np.cos(2 * np.pi * g[:, 2])], axis=1)
    t["orientation"] = q.astype(np.float32)
    scl = np.stack([s.uniform(n, 0.25, 4.0) for _ in range(3)], axis=1)
    if negative_scale:
        scl *= np.where(np.stack([s.uniform(n) for _ in range(3)], axis=1) < 0.1, -1.0, 1.0)
    t["scale"] = scl.astype(np.float32)
    slots = s.integers(n, 0, n_meshes).astype(np.uint32)
    slots[s.uniform(n) < no_mesh_fraction] = L.NO_MESH
    vo = np.full(n, L.NO_VISIBILITY_RANGE, np.uint32)
    cursor = np.zeros(1, np.uint32)
    if preallocated_fraction > 0.0:
        pre = (s.uniform(n) < preallocated_fraction) & (slots != L.NO_MESH)
        k = int(pre.sum())
        vo[pre] = (np.arange(k, dtype=np.uint32) * 7)          # ranges handed out earlier, in some earlier order
        cursor[0] = 7 * k
    return t, slots, vo, cursor


def random_mesh_infos(n_meshes, seed):
    s = Stream(seed, "scene_update_meshes")
    mi = np.zeros(n_meshes, L.mesh_info_dtype)
    counts = s.integers(n_meshes, 0, 400)
    counts[::7] = 32 * s.integers(len(counts[::7]), 0, 5)    # exact multiples of 32 and zero
    mi["mesh_lods"][:, :, 1] = counts[:, None].astype(np.uint32)
    mi["lod_count"] = 1
    return mi
