"""CPU: the oracle's cluster AABB (light_culling.comp + cluster_common.glsl restated) against the reference's own CPU twin
`compute_cluster_aabb` (src/passes/cluster.rs:132-184, used for its debug draw). The twin takes the NOMINAL slice planes
z_near*(z_far/z_near)^(k/n); the shader takes the measured depth bounds of the cluster. Feeding the oracle depth bounds
equal to the nominal planes must therefore give the twin's box: probed with point lights placed just inside / outside
each face of the twin's box (float64 restatement of the twin below)."""
import ctypes as C

import numpy as np

from orbit_b200 import layouts as L
from orbit_b200 import scenes


def twin_aabb(inv_proj, screen, tile, counts, z_near, z_far, cid):
    """cluster.rs:150-184 in float64."""
    def screen_to_view(px, py):
        tex = np.array([px / screen[0], py / screen[1]])
        ndc = np.array([tex[0], 1.0 - tex[1]]) * 2.0 - 1.0
        v = inv_proj @ np.array([ndc[0], ndc[1], 1.0, 1.0])
        return v[:3] / v[3]

    def to_z_plane(b, zd):            # line from the eye through b, plane normal (0,0,-1)
        return b * (zd / -b[2])
    mn = np.array([cid[0] * tile, cid[1] * tile], float)
    mx = np.minimum(mn + tile, screen)
    a, b = screen_to_view(*mn), screen_to_view(*mx)
    near = z_near * (z_far / z_near) ** (cid[2] / counts[2])
    far = z_near * (z_far / z_near) ** ((cid[2] + 1) / counts[2])
    pts = np.array([to_z_plane(a, near), to_z_plane(a, far), to_z_plane(b, near), to_z_plane(b, far)])
    return pts.min(axis=0), pts.max(axis=0), near, far


def test_cluster_aabb_matches_reference_twin(oracle):
    w, h, tile, cz, z_near, z_far = 1920, 1080, 120, 24, 0.1, 200.0
    cx, cy = -(-w // tile), -(-h // tile)
    P = scenes.perspective_infinite_reverse_rh(np.radians(70.0), w / h, z_near)
    inv = np.linalg.inv(np.asarray(P, np.float64))
    p = L.ClusterParams()
    p.info.world_to_view_matrix.set(np.eye(4))
    p.info.screen_to_view_matrix.set(inv)
    p.info.cluster_count[0], p.info.cluster_count[1], p.info.cluster_count[2] = cx, cy, cz
    p.info.tile_size_px = tile
    p.info.screen_size[0], p.info.screen_size[1] = w, h
    p.info.z_near, p.info.z_far = z_near, z_far
    n = cx * cy * cz
    cids = [(0, 0, 0), (5, 3, 7), (15, 8, 23), (8, 4, 12), (15, 0, 1)]
    bounds = np.zeros(2 * n, np.uint32)
    unique = np.zeros(4 + n, np.uint32)
    lights, expect = [], []
    r = 0.05
    for k, cid in enumerate(cids):
        lo, hi, near, far = twin_aabb(inv, np.array([w, h], float), tile, (cx, cy, cz), z_near, z_far, cid)
        idx = cid[0] + cx * (cid[1] + cy * cid[2])
        unique[4 + k] = idx
        # reverse-Z infinite projection: depth = z_near / distance; the shader keeps bits(1 - min depth) and bits(max depth)
        bounds[2 * idx + 0] = np.float32(1.0 - np.float32(z_near / far)).view(np.uint32)
        bounds[2 * idx + 1] = np.float32(z_near / near).view(np.uint32)
        ctr = (lo + hi) / 2
        for axis in range(3):
            for side, face in ((-1, lo), (1, hi)):
                for sign, want in ((-1, True), (1, False)):     # just inside the face -> hit; outside by > r -> miss
                    c = ctr.copy()
                    ext = max(hi[axis] - lo[axis], 1e-6)
                    margin = r * 0.5 if want else r * 1.5 + 2e-3 * ext   # float32 depth bounds move the z faces a little
                    c[axis] = face[axis] + side * sign * margin if not want else face[axis] + side * (r * 0.5)
                    lights.append(c); expect.append((k, want))
    unique[3] = len(cids)
    ld = np.zeros(len(lights), L.light_dtype)
    ld["light_type"] = L.LIGHT_POINT
    ld["position"] = np.array(lights, np.float32)
    ld["outer_radius"] = r
    p.info.global_light_count = len(lights)
    image = np.zeros(2 * n, np.uint32)
    cap = L.MAX_LIGHTS_PER_CLUSTER * len(cids)
    index = np.zeros(1 + cap, np.uint32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    oracle.lib().oracle_light_culling(C.byref(p), vp(ld), vp(bounds), vp(unique), vp(image), vp(index), cap)
    for k, cid in enumerate(cids):
        idx = cid[0] + cx * (cid[1] + cy * cid[2])
        off, cnt = int(image[2 * idx]), int(image[2 * idx + 1])
        got = set(index[1 + off:1 + off + cnt].tolist())
        for j, (kk, want) in enumerate(expect):
            if kk == k:
                assert (j in got) == want, (cid, j, want, lights[j])
