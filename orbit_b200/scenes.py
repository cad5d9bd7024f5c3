"""Procedural scenes, cameras and depth buffers for the five BASELINE.json configs.

glTF assets are unavailable offline, so inputs are generated deterministically (splitmix64 streams seeded per
array) directly in the reference's buffer layouts (orbit_b200.layouts). The same bytes feed the CUDA path, the
oracle and the committed golden fixtures.

Host math restated from the reference (these only PRODUCE inputs; they are not on the GPU hot path):
  perspective_infinite_reverse_rh / orthographic_rh   src/camera.rs:85-98 (glam 0.24 formulas)
  frustum_planes_from_matrix / normalize_plane        src/math.rs:72-89
  mip_levels_from_size                                src/math.rs:18-20
  ClusterSettings::cluster_grid_info                  src/passes/cluster.rs:63-72
  visibility-word allocation                          src/scene.rs:422-431
"""
from dataclasses import dataclass, field

import numpy as np

from . import layouts as L

_MASK = (1 << 64) - 1


def _fnv1a(name):
    h = 0xCBF29CE484222325
    for b in name.encode():
        h = ((h ^ b) * 0x100000001B3) & _MASK
    return h


class Stream:
    """Counter-based splitmix64: element i of the stream = mix(seed + (i+1)*golden). Vectorised, order-free."""

    def __init__(self, seed, name):
        self.seed = np.uint64((seed ^ _fnv1a(name)) & _MASK)
        self.pos = 0

    def bits(self, n):
        n = int(n)
        with np.errstate(over="ignore"):
            i = np.arange(self.pos + 1, self.pos + 1 + n, dtype=np.uint64)
            z = self.seed + i * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        self.pos += n
        return z

    def uniform(self, n, lo=0.0, hi=1.0):
        u = (self.bits(n) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        return lo + (hi - lo) * u

    def integers(self, n, lo, hi):
        """uniform integers in [lo, hi)"""
        return (lo + np.floor(self.uniform(n) * (hi - lo))).astype(np.int64)


# ------------------------------------------------------------------------------------------------------------
# host math (float64, rounded once to f32 where stored)
# ------------------------------------------------------------------------------------------------------------
def perspective_infinite_reverse_rh(fov_y, aspect, z_near):
    f = 1.0 / np.tan(0.5 * fov_y)
    m = np.zeros((4, 4))
    m[0, 0] = f / aspect
    m[1, 1] = f
    m[3, 2] = -1.0
    m[2, 3] = z_near
    return m  # math convention m[row, col]


def orthographic_rh(left, right, bottom, top, near, far):
    rcp_w, rcp_h, r = 1.0 / (right - left), 1.0 / (top - bottom), 1.0 / (near - far)
    m = np.zeros((4, 4))
    m[0, 0] = rcp_w + rcp_w
    m[1, 1] = rcp_h + rcp_h
    m[2, 2] = r
    m[0, 3] = -(left + right) * rcp_w
    m[1, 3] = -(top + bottom) * rcp_h
    m[2, 3] = r * near
    m[3, 3] = 1.0
    return m


def look_to_rh(eye, direction, up=(0.0, 1.0, 0.0)):
    eye = np.asarray(eye, np.float64)
    f = np.asarray(direction, np.float64)
    f = f / np.linalg.norm(f)
    s = np.cross(f, np.asarray(up, np.float64))
    s = s / np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -np.dot(eye, s), -np.dot(eye, u), np.dot(eye, f)
    return m


def frustum_planes_from_matrix(m):
    """math.rs:72-80. m in math convention; returns 6 planes (row3 +- row0, +-row1, +-row2)."""
    return np.array([m[3] + m[0], m[3] - m[0], m[3] + m[1], m[3] - m[1], m[3] + m[2], m[3] - m[2]])


def normalize_plane(p):
    return p / np.linalg.norm(p[:3])


def mip_levels_from_size(max_size):
    return max(1, int(np.floor(np.log2(np.float32(max_size)))) + 1)


def hiz_geometry(depth_w, depth_h):
    npot = lambda v: 1 << (int(v) - 1).bit_length()
    w, h = npot(depth_w) // 2, npot(depth_h) // 2
    levels = mip_levels_from_size(max(w, h))
    offs, off = [], 0
    for l in range(levels):
        offs.append(off)
        off += max(w >> l, 1) * max(h >> l, 1)
    return w, h, levels, offs, off


# ------------------------------------------------------------------------------------------------------------
@dataclass
class Scene:
    name: str
    seed: int
    meshlets: np.ndarray        # layouts.meshlet_dtype
    mesh_infos: np.ndarray      # layouts.mesh_info_dtype
    materials: np.ndarray       # uint8[n_materials*80]
    entities: np.ndarray        # layouts.entity_dtype
    entity_draws: np.ndarray    # uint8[4 + 12*N]  (EntityDrawBuffer)
    n_entities: int
    n_meshlet_instances: int    # sum over entity draws of LOD-0 meshlet count
    n_records_lod0: int         # sum of ceil(lod0/32)
    n_visibility_words: int
    aabb_min: np.ndarray
    aabb_max: np.ndarray
    entity_pos: np.ndarray = field(default=None, repr=False)     # float64 [N,3] world centre of the bounding sphere
    entity_radius: np.ndarray = field(default=None, repr=False)  # float64 [N] world radius
    entity_occluder: np.ndarray = field(default=None, repr=False)  # float64 [N] world radius of the depth-splat disc
    transforms: np.ndarray = field(default=None, repr=False)     # layouts.transform_dtype [N]: what scene.rs:404-492 starts from

    @property
    def draws(self):
        return self.entity_draws[4:].view(L.entity_draw_dtype)

    def bytes_summary(self):
        return {"meshlets": self.meshlets.nbytes, "mesh_infos": self.mesh_infos.nbytes, "entities": self.entities.nbytes,
                "entity_draws": self.entity_draws.nbytes, "materials": self.materials.nbytes}


def _quat_to_mat(q):
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    m = np.empty((len(q), 3, 3))
    m[:, 0, 0] = 1 - 2 * (y * y + z * z); m[:, 0, 1] = 2 * (x * y - z * w); m[:, 0, 2] = 2 * (x * z + y * w)
    m[:, 1, 0] = 2 * (x * y + z * w); m[:, 1, 1] = 1 - 2 * (x * x + z * z); m[:, 1, 2] = 2 * (y * z - x * w)
    m[:, 2, 0] = 2 * (x * z - y * w); m[:, 2, 1] = 2 * (y * z + x * w); m[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return m


def make_scene(name, seed, n_entities, n_meshes, lod_meshlets, layout="city", grid=None, pitch=12.0,
               half_extent_xz=2.0, height_range=(1.0, 8.0), instanced=False):
    """lod_meshlets: meshlet count per LOD, e.g. [200] or [200,100,50,25].
    layout: "city" (ground lattice grid=(nx,nz), yaw-only rotation, boxes standing on y=0) or
            "lattice3d" (grid=(nx,ny,nz), uniform random rotation, boxes centred on the node)."""
    lod_meshlets = list(lod_meshlets)
    per_mesh = int(sum(lod_meshlets))
    M = n_meshes * per_mesh
    # ---- meshes: box of half extent (hx, h, hx); meshlet bounds on its surface
    s_mesh = Stream(seed, "mesh")
    if layout == "city":
        h = s_mesh.uniform(n_meshes, *height_range)
        centre_y = h.copy()
    else:
        h = np.full(n_meshes, half_extent_xz)
        centre_y = np.zeros(n_meshes)
    hx = half_extent_xz
    s = Stream(seed, "meshlet")
    hm = np.repeat(h, per_mesh)
    cym = np.repeat(centre_y, per_mesh)
    area_side = 4.0 * hx * hm          # (2h)(2hx)
    area_top = np.full(M, 4.0 * hx * hx)
    cum = np.stack([area_side, area_side, area_side, area_side, area_top, area_top], axis=1).cumsum(axis=1)
    pick = s.uniform(M) * cum[:, -1]
    face = (pick[:, None] >= cum).sum(axis=1).clip(0, 5)   # 0:+x 1:-x 2:+z 3:-z 4:+y 5:-y
    a, b = s.uniform(M, -1.0, 1.0), s.uniform(M, -1.0, 1.0)
    pos = np.zeros((M, 3)); nrm = np.zeros((M, 3)); t1 = np.zeros((M, 3)); t2 = np.zeros((M, 3))
    for f, (axis, sign) in enumerate([(0, 1), (0, -1), (2, 1), (2, -1), (1, 1), (1, -1)]):
        sel = face == f
        ext = np.stack([np.full(M, hx), hm, np.full(M, hx)], axis=1)
        o = [i for i in range(3) if i != axis]
        pos[sel, axis] = sign * ext[sel, axis]
        pos[sel, o[0]] = a[sel] * ext[sel, o[0]]
        pos[sel, o[1]] = b[sel] * ext[sel, o[1]]
        nrm[sel, axis] = sign
        t1[sel, o[0]] = 1.0
        t2[sel, o[1]] = 1.0
    pos[:, 1] += cym
    theta = np.radians(s.uniform(M, 0.0, 20.0)); phi = s.uniform(M, 0.0, 2 * np.pi)
    ax = nrm + np.tan(theta)[:, None] * (np.cos(phi)[:, None] * t1 + np.sin(phi)[:, None] * t2)
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    radius = s.uniform(M, 0.15, 0.45)
    never = s.uniform(M) < 0.15
    cutoff = np.where(never, 127, np.round(127 * s.uniform(M, 0.1, 0.9))).astype(np.int8)
    vcount = s.integers(M, 32, 65).astype(np.uint8)
    tcount = s.integers(M, 32, 65).astype(np.uint8)
    meshlets = np.zeros(M, L.meshlet_dtype)
    meshlets["bounding_sphere"][:, :3] = pos.astype(np.float32)
    meshlets["bounding_sphere"][:, 3] = radius.astype(np.float32)
    meshlets["cone_axis"] = np.round(127 * ax).astype(np.int8)
    meshlets["cone_cutoff"] = cutoff
    meshlets["vertex_count"] = vcount
    meshlets["triangle_count"] = tcount
    meshlets["vertex_offset"] = (np.cumsum(vcount.astype(np.uint64)) - vcount).astype(np.uint32)
    per = vcount.astype(np.uint64) + tcount.astype(np.uint64)
    meshlets["data_offset"] = ((np.cumsum(per) - per) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    meshlets["material_index"] = s.integers(M, 0, 16).astype(np.uint16)

    mesh_infos = np.zeros(n_meshes, L.mesh_info_dtype)
    sphere_r = np.sqrt(hx * hx + h * h + hx * hx) + 0.45
    mesh_infos["bounding_sphere"][:, 1] = centre_y.astype(np.float32)
    mesh_infos["bounding_sphere"][:, 3] = sphere_r.astype(np.float32)
    mesh_infos["aabb_min"][:, :3] = np.stack([np.full(n_meshes, -hx), centre_y - h, np.full(n_meshes, -hx)], 1)
    mesh_infos["aabb_max"][:, :3] = np.stack([np.full(n_meshes, hx), centre_y + h, np.full(n_meshes, hx)], 1)
    mesh_infos["lod_count"] = len(lod_meshlets)
    base = np.arange(n_meshes, dtype=np.uint64) * per_mesh
    off = 0
    for k in range(L.MAX_MESH_LODS):
        kk = min(k, len(lod_meshlets) - 1)
        if k < len(lod_meshlets):
            mesh_infos["mesh_lods"][:, k, 0] = (base + off).astype(np.uint32)
            mesh_infos["mesh_lods"][:, k, 1] = lod_meshlets[k]
            off += lod_meshlets[k]
        else:  # unused slots repeat the last LOD (never indexed: min(lod, lod_count-1))
            mesh_infos["mesh_lods"][:, k] = mesh_infos["mesh_lods"][:, kk]
    mesh_infos["meshlet_data_offset"] = 0

    # ---- materials: 16 entries, alpha_mode 70% opaque / 20% masked / 10% transparent
    sm = Stream(seed, "material")
    u = sm.uniform(16)
    alpha = np.where(u < 0.7, 0, np.where(u < 0.9, 1, 2)).astype(np.uint32)
    materials = np.zeros(16 * L.MATERIAL_STRIDE, np.uint8)
    mview = materials.view(np.uint32).reshape(16, L.MATERIAL_STRIDE // 4)
    mview[:, L.MATERIAL_ALPHA_OFFSET // 4] = alpha
    mview[:, 0:4] = np.float32(1.0).view(np.uint32)  # base_color = 1

    # ---- entities
    se = Stream(seed, "entity")
    N = n_entities
    if layout == "city":
        nx, nz = grid
        assert nx * nz >= N
        ix = np.arange(N) % nx
        iz = np.arange(N) // nx
        T = np.stack([ix * pitch, np.zeros(N), iz * pitch], axis=1).astype(np.float64)
        yaw = se.uniform(N, 0.0, 2 * np.pi)
        q = np.stack([np.zeros(N), np.sin(yaw / 2), np.zeros(N), np.cos(yaw / 2)], axis=1)
    else:
        nx, ny, nz = grid
        assert nx * ny * nz >= N
        i = np.arange(N)
        T = np.stack([(i % nx) * pitch, ((i // nx) % ny) * pitch, (i // (nx * ny)) * pitch], axis=1).astype(np.float64)
        g = np.stack([se.uniform(N) for _ in range(3)], axis=1)   # Shoemake uniform quaternion
        q = np.stack([np.sqrt(1 - g[:, 0]) * np.sin(2 * np.pi * g[:, 1]), np.sqrt(1 - g[:, 0]) * np.cos(2 * np.pi * g[:, 1]),
                      np.sqrt(g[:, 0]) * np.sin(2 * np.pi * g[:, 2]), np.sqrt(g[:, 0]) * np.cos(2 * np.pi * g[:, 2])], axis=1)
    scale = se.uniform(N, 0.5, 2.0)
    R = _quat_to_mat(q)
    model = np.zeros((N, 4, 4))           # math convention [row, col]
    model[:, :3, :3] = R * scale[:, None, None]
    model[:, :3, 3] = T
    model[:, 3, 3] = 1.0
    entities = np.zeros(N, L.entity_dtype)
    entities["model_matrix"] = model.transpose(0, 2, 1).astype(np.float32)   # stored [col][row]
    nm = np.zeros((N, 4, 4)); nm[:, :3, :3] = R / scale[:, None, None]; nm[:, 3, 3] = 1.0
    entities["normal_matrix"] = nm.transpose(0, 2, 1).astype(np.float32)

    mesh_index = se.integers(N, 0, n_meshes) if instanced else (np.arange(N) % n_meshes)
    words = (lod_meshlets[0] + 31) // 32
    draws = np.zeros(N, L.entity_draw_dtype)
    draws["entity_index"] = np.arange(N, dtype=np.uint32)
    draws["mesh_index"] = mesh_index.astype(np.uint32)
    draws["visibility_offset"] = (np.arange(N, dtype=np.uint64) * words).astype(np.uint32)   # scene.rs:422-431
    entity_draws = np.zeros(4 + 12 * N, np.uint8)
    entity_draws[:4] = np.frombuffer(np.uint32(N).tobytes(), np.uint8)
    entity_draws[4:] = draws.view(np.uint8)

    transforms = np.zeros(N, L.transform_dtype)   # the Transform each model matrix came from (input of orbit_scene_update)
    transforms["position"] = T.astype(np.float32)
    transforms["orientation"] = q.astype(np.float32)
    transforms["scale"] = scale.astype(np.float32)[:, None]

    centre_local = np.stack([np.zeros(N), centre_y[mesh_index], np.zeros(N)], axis=1)
    entity_pos = np.einsum("nij,nj->ni", model[:, :3, :3], centre_local) + T
    entity_radius = sphere_r[mesh_index] * scale
    occ = hx * scale * 0.9
    pad = float(entity_radius.max())
    return Scene(name=name, seed=seed, meshlets=meshlets, mesh_infos=mesh_infos, materials=materials, entities=entities,
                 entity_draws=entity_draws, n_entities=N, n_meshlet_instances=N * lod_meshlets[0],
                 n_records_lod0=N * words, n_visibility_words=N * words,
                 aabb_min=entity_pos.min(axis=0) - pad, aabb_max=entity_pos.max(axis=0) + pad,
                 entity_pos=entity_pos, entity_radius=entity_radius, entity_occluder=occ, transforms=transforms)


# ------------------------------------------------------------------------------------------------------------
@dataclass
class View:
    """What a caller of the culling passes knows about one view (forward.rs:261-284, shadow_renderer.rs:693-707)."""
    width: int
    height: int
    view: np.ndarray                 # 4x4 math convention
    projection_matrix: np.ndarray
    planes: np.ndarray               # [n,4] view-space normalised planes actually passed
    projection_type: int
    fov: float = 0.0
    near: float = 0.01
    far: float = 0.0
    half_width: float = 0.0
    lod_target_view: tuple = (0.0, 0.0, 0.0)
    lod_range: tuple = (0, 8)        # min, max+1 (Range<usize>, app.rs:352-355)
    lod_base: float = 16.0
    lod_step: float = 2.0

    @property
    def aspect(self):
        return self.width / self.height


def perspective_view(eye, direction, width, height, fov_deg=90.0, near=0.01, up=(0.0, 1.0, 0.0)):
    fov = np.radians(fov_deg)
    P = perspective_infinite_reverse_rh(fov, width / height, near)
    planes = np.array([normalize_plane(p) for p in frustum_planes_from_matrix(P)])[0:5]   # forward.rs:268
    return View(width=width, height=height, view=look_to_rh(eye, direction, up), projection_matrix=P, planes=planes,
                projection_type=L.PROJ_PERSPECTIVE, fov=fov, near=near)


def orthographic_view(eye, direction, width, height, half_width, near, far, up=(0.0, 1.0, 0.0)):
    hh = half_width * (height / width)
    P = orthographic_rh(-half_width, half_width, -hh, hh, far, near)   # reversed: camera.rs:91-96
    planes = np.array([normalize_plane(p) for p in frustum_planes_from_matrix(P)])
    return View(width=width, height=height, view=look_to_rh(eye, direction, up), projection_matrix=P, planes=planes,
                projection_type=L.PROJ_ORTHOGRAPHIC, near=near, far=far, half_width=half_width)


def _quat_from_rotation_arc(a, b):
    """glam Quat::from_rotation_arc for unit vectors (general branch): (cross(a,b), 1 + dot(a,b)) normalised; (x,y,z,w)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    q = np.concatenate([np.cross(a, b), [1.0 + float(np.dot(a, b))]])
    return q / np.linalg.norm(q)


def frustum_split(near, far, lam, ratio):
    """math.rs:64-69."""
    uniform = near + (far - near) * ratio
    log = near * (far / near) ** ratio
    return log * lam + (1.0 - lam) * uniform


def perspective_corners(fovy, aspect, near, far):
    """math.rs:149-168: the 8 view-space corners of a sub-frustum."""
    th, tv = np.tan(fovy / 2.0) * aspect, np.tan(fovy / 2.0)
    xn, yn, xf, yf = near * th, near * tv, far * th, far * tv
    return np.array([[-xn, -yn, -near, 1.0], [xn, -yn, -near, 1.0], [xn, yn, -near, 1.0], [-xn, yn, -near, 1.0],
                     [-xf, -yf, -far, 1.0], [xf, -yf, -far, 1.0], [xf, yf, -far, 1.0], [-xf, yf, -far, 1.0]])


SUN_ORIENTATION = _quat_from_rotation_arc((0.0, 0.0, 1.0), np.array([-1.0, 1.0, 1.0]) / np.sqrt(3.0))   # app.rs:593


def cascade_views(camera, n_cascades=4, lam=0.8, max_shadow_distance=32.0, resolution=2048, orientation=SUN_ORIENTATION,
                  min_mesh_lod=0, max_mesh_lod=7):
    """The orthographic shadow-cascade views of ShadowRenderer::render_cascaded_shadow (shadow_renderer.rs:466-712) for a
    perspective `camera` View, restated step by step (host-side data generation: evaluated in float64, rounded to f32 where
    the CullInfo is packed): split distances (lambda 0.8, max distance 32, :469-476), sub-frustum corners in light space
    (:478-486), bounding sphere of the corners (:488-505), forward offset (:510-524), texel snapping (:526-532), light matrix =
    translate(-centre') * rotate(inverse orientation) (:540), near = -radius - 80 / far = radius (:542-543), cull planes = the 6
    planes of the NON-reversed orthographic matrix (:622-630) followed by those of the first 5 camera planes, taken in light
    space, whose normal has z >= 0 (:632-640) — 6 to 11 planes; pass 0, LOD range min..max+1 for cascades 0-1 and 2..max+1 for
    cascades 2-3 (:697-700), LOD target = light_matrix * camera position (:703)."""
    assert camera.projection_type == L.PROJ_PERSPECTIVE
    qi = np.array([-orientation[0], -orientation[1], -orientation[2], orientation[3]])    # direction.inverse() of a unit quaternion
    light_rot = np.eye(4); light_rot[:3, :3] = _quat_to_mat(qi[None])[0]
    view_to_world = np.linalg.inv(camera.view)
    cam_pos = view_to_world[:3, 3]
    view_to_light = light_rot @ view_to_world
    cam_vp = camera.projection_matrix @ camera.view
    out = []
    for i in range(n_cascades):
        near = frustum_split(camera.near, max_shadow_distance, lam, i / n_cascades)
        far = frustum_split(camera.near, max_shadow_distance, lam, (i + 1) / n_cascades)
        corners = (view_to_light @ perspective_corners(camera.fov, camera.aspect, near, far).T).T
        corners = corners / corners[:, 3:4]
        centre = corners.sum(axis=0) / 8.0
        cmin, cmax = corners[:, :3].min(axis=0), corners[:, :3].max(axis=0)
        radius = float(np.sqrt(max(np.sum((corners[:, :3] - centre[:3]) ** 2, axis=1))))
        fwd = view_to_light[:3, 2]                                           # z_axis of view_to_light
        fa = (fwd + 1.0) / 2.0
        lo, hi = cmin - centre[:3], cmax - centre[:3]
        offset = lo + (hi - lo) * fa - radius * fwd                          # lerp_element_wise(min, max, a) - radius * sign
        texel = radius * 2.0 / resolution
        centre2 = np.floor((centre[:3] + offset) / texel) * texel
        T = np.eye(4); T[:3, 3] = -centre2
        light_matrix = T @ light_rot
        near_clip, far_clip = -radius - 80.0, radius
        P = orthographic_rh(-radius, radius, -radius, radius, far_clip, near_clip)          # reverse z (:545-552)
        light_planes = [normalize_plane(pl) for pl in frustum_planes_from_matrix(
            orthographic_rh(-radius, radius, -radius, radius, near_clip, far_clip))]         # non-reversed on purpose (:622-630)
        clip_to_light = cam_vp @ np.linalg.inv(light_matrix)
        cam_planes = [normalize_plane(pl) for pl in frustum_planes_from_matrix(clip_to_light)[:5]]
        cam_planes = [pl for pl in cam_planes if pl[2] >= 0.0]                               # dot(plane.xyz, Z) >= 0 (:636)
        planes = np.array(light_planes + cam_planes)
        lod_range = (min_mesh_lod, max_mesh_lod + 1) if i < 2 else (2, max_mesh_lod + 1)
        target = (light_matrix @ np.append(cam_pos, 1.0))[:3]
        out.append(View(width=resolution, height=resolution, view=light_matrix, projection_matrix=P, planes=planes,
                        projection_type=L.PROJ_ORTHOGRAPHIC, near=near_clip, far=far_clip, half_width=radius,
                        lod_target_view=tuple(float(v) for v in target), lod_range=lod_range))
    return out


def make_depth(scene, view, max_entities=None):
    """Reverse-Z depth buffer: every entity in front of the camera is splatted as a screen-space disc at its
    centre depth (max-blend = nearest wins), sky = 0.0. Deterministic; generated once per (scene, view)."""
    W, H = view.width, view.height
    depth = np.zeros((H, W), np.float32)
    c = (view.view[:3, :3] @ scene.entity_pos.T).T + view.view[:3, 3]
    zp = -c[:, 2]
    R = scene.entity_occluder
    if view.projection_type == L.PROJ_PERSPECTIVE:
        P00, P11 = view.projection_matrix[0, 0], view.projection_matrix[1, 1]
        ok = zp > (R + view.near)
        with np.errstate(divide="ignore", invalid="ignore"):
            px = (0.5 + 0.5 * P00 * c[:, 0] / zp) * W
            py = (0.5 - 0.5 * P11 * c[:, 1] / zp) * H
            rp = 0.5 * P11 * R / zp * H
            d = (view.near / zp)
    else:
        P00, P11 = view.projection_matrix[0, 0], view.projection_matrix[1, 1]
        ok = np.ones(len(c), bool)
        px = (0.5 + 0.5 * P00 * c[:, 0]) * W
        py = (0.5 - 0.5 * P11 * c[:, 1]) * H
        rp = 0.5 * P11 * R * H
        k = 1.0 / (view.far - view.near)
        d = (c[:, 2] + view.far) * k
    ok &= (px + rp >= 0) & (px - rp < W) & (py + rp >= 0) & (py - rp < H) & (d > 0) & (d <= 1.0)
    idx = np.nonzero(ok)[0]
    if max_entities is not None:
        idx = idx[:max_entities]
    d32 = d.astype(np.float32)
    for i in idx:
        r = rp[i]
        x0, x1 = int(max(np.floor(px[i] - r), 0)), int(min(np.ceil(px[i] + r), W - 1))
        y0, y1 = int(max(np.floor(py[i] - r), 0)), int(min(np.ceil(py[i] + r), H - 1))
        if x1 < x0 or y1 < y0:
            continue
        if r < 1.0:
            xi, yi = int(min(max(px[i], 0), W - 1)), int(min(max(py[i], 0), H - 1))
            if d32[i] > depth[yi, xi]:
                depth[yi, xi] = d32[i]
            continue
        yy, xx = np.ogrid[y0:y1 + 1, x0:x1 + 1]
        m = (xx + 0.5 - px[i]) ** 2 + (yy + 0.5 - py[i]) ** 2 <= r * r
        blk = depth[y0:y1 + 1, x0:x1 + 1]
        np.maximum(blk, np.where(m, d32[i], np.float32(0)), out=blk)
    return depth


def make_lights(seed, n_point, aabb_min, aabb_max, intensity=(1.0, 6.0), cutoff=0.25, with_sun=True):
    """GpuLightData array (scene.rs:278-291). outer_radius = sqrt(intensity / cutoff) (scene.rs:273-275)."""
    s = Stream(seed, "light")
    extra = 2 if with_sun else 0
    lights = np.zeros(n_point + extra, L.light_dtype)
    k = 0
    if with_sun:
        lights[0]["light_type"] = L.LIGHT_SKY
        lights[1]["light_type"] = L.LIGHT_DIRECTIONAL
        lights[1]["direction"] = (-0.57735, 0.57735, 0.57735)
        lights[0]["shadow_data_index"] = lights[1]["shadow_data_index"] = 0xFFFFFFFF
        k = 2
    p = np.stack([s.uniform(n_point, aabb_min[i], aabb_max[i]) for i in range(3)], axis=1)
    inten = s.uniform(n_point, *intensity)
    lights["light_type"][k:] = L.LIGHT_POINT
    lights["shadow_data_index"][k:] = 0xFFFFFFFF
    lights["position"][k:] = p.astype(np.float32)
    lights["intensity"][k:] = inten.astype(np.float32)
    lights["color"][k:] = 1.0
    lights["outer_radius"][k:] = np.sqrt(inten / cutoff).astype(np.float32)
    lights["inner_radius"][k:] = 0.1
    return lights


def cluster_grid_info(near, far, slices):
    """ClusterSettings::cluster_grid_info (cluster.rs:63-72), evaluated in f32 like the reference host code."""
    near, far, slices = np.float32(near), np.float32(far), np.float32(slices)
    log_f_n = np.log2(far / near, dtype=np.float32)
    z_scale = np.float32(slices / log_f_n)
    z_bias = np.float32(-((slices * np.log2(near, dtype=np.float32)) / log_f_n))
    return float(z_scale), float(z_bias)


# ------------------------------------------------------------------------------------------------------------
# BASELINE.json configs. `scale` < 1 shrinks entity counts for CPU-sized tests; 1.0 = the named size.
# ------------------------------------------------------------------------------------------------------------
SEEDS = {"C1": 0x0B170001, "C2": 0x0B170002, "C3": 0x0B170003, "C4": 0x0B170004, "C5": 0x0B170005}


def config_c1(scale=1.0, lods=(100,)):
    """1k entities / 100k meshlets lattice, camera inside the lattice, 1920x1080."""
    n = max(2, int(round(10 * scale ** (1 / 3))))
    scene = make_scene("C1", SEEDS["C1"], n ** 3, n ** 3, list(lods), layout="lattice3d", grid=(n, n, n), pitch=6.0,
                       half_extent_xz=1.0)
    f = (n - 1) * 6.0 / 54.0
    view = perspective_view((25.3 * f, 28.1 * f, 26.4 * f), (1.0, -0.1, 0.35), 1920, 1080)
    return scene, view


def config_c2(scale=1.0):
    """10k entities / 2M meshlets city, one 1920x1080 view from a corner at street level, yaw 30 degrees."""
    n = max(4, int(round(100 * np.sqrt(scale))))
    scene = make_scene("C2", SEEDS["C2"], n * n, n * n, [200], layout="city", grid=(n, n), pitch=12.0)
    yaw = np.radians(30.0)
    view = perspective_view((-6.0, 2.0, -6.0), (np.sin(yaw), 0.0, np.cos(yaw)), 1920, 1080)
    return scene, view


def config_c3(scale=1.0):
    """250k entities x 200 instanced meshlets (1024 unique meshes) = 50M meshlet instances, one 3840x2160 view."""
    n = max(4, int(round(500 * np.sqrt(scale))))
    scene = make_scene("C3", SEEDS["C3"], n * n, min(1024, n * n), [200], layout="city", grid=(n, n), pitch=12.0,
                       instanced=True)
    yaw = np.radians(40.0)
    view = perspective_view((-6.0, 30.0, -6.0), (np.sin(yaw), -0.12, np.cos(yaw)), 3840, 2160)
    return scene, view


def config_c4(scale=1.0):
    """50k entities x 200 = 10M meshlets; main view; clusters 16x9x24 over 1920x1080; 64k point lights."""
    n = max(4, int(round(np.sqrt(50000 * scale))))
    scene = make_scene("C4", SEEDS["C4"], n * n, n * n, [200], layout="city", grid=(n, n), pitch=12.0)
    yaw = np.radians(35.0)
    view = perspective_view((-6.0, 3.0, -6.0), (np.sin(yaw), 0.0, np.cos(yaw)), 1920, 1080)
    return scene, view


def config_c5(scale=1.0, n_views=256):
    """100k entities x 200 = 20M meshlets, 256 cameras at 1920x1080."""
    n = max(4, int(round(np.sqrt(100000 * scale))))
    scene = make_scene("C5", SEEDS["C5"], n * n, n * n, [200], layout="city", grid=(n, n), pitch=12.0)
    s = Stream(SEEDS["C5"], "camera")
    lo, hi = scene.aabb_min, scene.aabb_max
    x = s.uniform(n_views, lo[0], hi[0]); y = s.uniform(n_views, 2.0, 40.0); z = s.uniform(n_views, lo[2], hi[2])
    yaw = s.uniform(n_views, 0.0, 2 * np.pi); pitch = np.radians(s.uniform(n_views, -30.0, 10.0))
    views = [perspective_view((x[i], y[i], z[i]),
                              (np.sin(yaw[i]) * np.cos(pitch[i]), np.sin(pitch[i]), np.cos(yaw[i]) * np.cos(pitch[i])),
                              1920, 1080) for i in range(n_views)]
    return scene, views
