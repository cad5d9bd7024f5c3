"""Asset-side producers of the culling path's inputs from real GEOMETRY.

The reference turns a mesh into the arrays the culling shaders read in `GpuAssets::add_mesh` (src/assets/mod.rs:325-470):
for every LOD the index list is cut into meshlets (`compute_meshlets`, src/assets/mesh.rs:292-338: meshopt::build_meshlets +
meshopt::compute_meshlet_bounds), `mesh_lods[lod] = {meshlet_offset, meshlet_count}` records where the LOD's meshlets lie
(made absolute in `MeshInfo::to_gpu`, mod.rs:80-85), and `MeshData::compute_bounds` (mesh.rs:192-215) gives the mesh's AABB and
sphere. This module is the host side of that path for procedurally generated geometry (glTF assets are not available offline):

  * `building_mesh` — a tessellated, bumpy box with per-vertex normals, in the reference's 32-byte `GpuMeshVertex` layout
    (mesh.rs:12-20); coarser LODs = coarser tessellations whose index counts shrink by ~0.8 per level (mod.rs:390);
  * `partition_meshlets` — a deterministic stand-in for meshopt::build_meshlets (third-party, not restated): triangles in index
    order, a meshlet closes when the next triangle would exceed 64 vertices or 64 triangles (mesh.rs:8-9);
  * `AssetBuilder` — accumulates vertices / meshlet data / meshlets / mesh infos exactly in `add_mesh`'s layouts;
  * the BOUNDS — per-meshlet sphere + normal cone, per-mesh sphere + AABB — are computed on the GPU by
    `orbit_meshlet_bounds` / `orbit_mesh_bounds` (csrc/asset_bounds.cu) through `AssetBuilder.finish(context)`.
"""
import ctypes as C

import numpy as np

from . import layouts as L

MAX_MESHLET_VERTICES = 64       # assets/mesh.rs:8
MAX_MESHLET_TRIANGLES = 64      # assets/mesh.rs:9
vertex_dtype = np.dtype([("position", "<f4", (3,)), ("normal", "i1", (4,)), ("uv_coord", "<f4", (2,)), ("tangent", "i1", (4,)),
                         ("_padding", "<u4")])          # GpuMeshVertex, repr(C, align(16)): 32 bytes
assert vertex_dtype.itemsize == 32


def _snorm8(v):
    """math::pack_f32_to_snorm_u8 (math.rs:201-203): clamp, scale by 127, truncate."""
    return (np.clip(v, -1.0, 1.0) * 127.0).astype(np.int8)


def building_mesh(half_x, half_y, half_z, n, seed=0, bump=0.06):
    """A box of half extents (half_x, half_y, half_z) standing on y = 0, every face an n x n grid of quads (two triangles each,
    emitted row by row, so consecutive triangles are neighbours), displaced along the face normal by a smooth deterministic
    bump so that normal cones are not degenerate. Returns (vertices[vertex_dtype], indices[uint32])."""
    verts, idx = [], []
    ext = np.array([half_x, half_y, half_z])
    for f, (axis, sign) in enumerate([(0, 1), (0, -1), (2, 1), (2, -1), (1, 1), (1, -1)]):
        o = [i for i in range(3) if i != axis]
        u, v = np.meshgrid(np.linspace(-1.0, 1.0, n + 1), np.linspace(-1.0, 1.0, n + 1), indexing="xy")
        p = np.zeros((n + 1, n + 1, 3))
        p[..., o[0]] = u * ext[o[0]]
        p[..., o[1]] = v * ext[o[1]]
        ph = 0.7 * f + 0.37 * seed
        h = bump * (np.sin(3.1 * u + ph) * np.cos(2.3 * v - ph) + 0.5 * np.sin(5.7 * (u + v) + ph)) * (1 - u * u) * (1 - v * v)
        p[..., axis] = sign * (ext[axis] + h)
        # normal of the displaced surface: cross of the parametric derivatives (finite differences), oriented outwards
        du = np.gradient(p, axis=1); dv = np.gradient(p, axis=0)
        nrm = np.cross(du, dv)
        nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
        nrm *= np.sign(nrm[..., axis:axis + 1] * sign + 1e-30)
        p[..., 1] += half_y                                   # stand on the ground
        base = sum(len(a) for a in verts)
        fv = np.zeros((n + 1) * (n + 1), vertex_dtype)
        fv["position"] = p.reshape(-1, 3).astype(np.float32)
        fv["normal"][:, :3] = _snorm8(nrm.reshape(-1, 3))
        fv["uv_coord"] = np.stack([(u.reshape(-1) + 1) / 2, (v.reshape(-1) + 1) / 2], axis=1).astype(np.float32)
        fv["tangent"][:, 3] = 127
        verts.append(fv)
        i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
        a = (j * (n + 1) + i).reshape(-1); b = a + 1; c = a + (n + 1); d = c + 1
        flip = (np.cross(np.eye(3)[o[0]], np.eye(3)[o[1]])[axis] * sign) < 0        # keep counter-clockwise winding seen from outside
        tri = np.stack([a, b, d, a, d, c], axis=1) if not flip else np.stack([a, d, b, a, c, d], axis=1)
        idx.append((tri + base).reshape(-1))
    return np.concatenate(verts), np.concatenate(idx).astype(np.uint32)


def partition_meshlets(indices, max_vertices=MAX_MESHLET_VERTICES, max_triangles=MAX_MESHLET_TRIANGLES):
    """Triangles in index order into meshlets of at most `max_vertices` distinct vertices and `max_triangles` triangles.
    Returns a list of (vertex ids in order of first use [uint32], local triangle indices [uint8, 3 per triangle])."""
    out = []
    local, order, tris = {}, [], []
    for t in range(len(indices) // 3):
        tri = [int(v) for v in indices[3 * t:3 * t + 3]]
        new = len({v for v in tri if v not in local})
        if tris and (len(order) + new > max_vertices or len(tris) // 3 + 1 > max_triangles):
            out.append((np.array(order, np.uint32), np.array(tris, np.uint8)))
            local, order, tris = {}, [], []
        for v in tri:
            if v not in local:
                local[v] = len(order); order.append(v)
            tris.append(local[v])
    if tris:
        out.append((np.array(order, np.uint32), np.array(tris, np.uint8)))
    return out


class AssetBuilder:
    """GpuAssets::add_mesh (assets/mod.rs:325-470) for a sequence of meshes: one vertex array, one meshlet-data array, one
    meshlet array, one mesh-info array, ranges handed out consecutively (the reference's allocators do the same while nothing
    has been freed)."""

    def __init__(self):
        self.vertices, self.meshlet_data, self.meshlets, self.mesh_infos, self.vertex_ranges = [], [], [], [], []
        self._n_vertices = self._n_data = self._n_meshlets = 0

    def add_mesh(self, lods, material_index=0):
        """lods: list of (vertices, indices), LOD 0 first, every LOD with its own vertex array (the reference simplifies LOD 0's
        indices over ONE vertex array; separately tessellated LODs are concatenated into one here, which the layouts allow).
        Returns the mesh index."""
        assert 1 <= len(lods) <= L.MAX_MESH_LODS
        mesh_vertex_first = self._n_vertices
        info = np.zeros(1, L.mesh_info_dtype)
        info["vertex_offset"] = mesh_vertex_first
        info["meshlet_data_offset"] = self._n_data
        info["lod_count"] = len(lods)
        lod_vertex_offset = 0
        for k, (verts, indices) in enumerate(lods):
            meshlet_offset = self._n_meshlets                            # absolute: MeshInfo::to_gpu adds the mesh's base (mod.rs:80-85)
            parts = partition_meshlets(indices)
            ml = np.zeros(len(parts), L.meshlet_dtype)
            for i, (vids, tris) in enumerate(parts):
                data_offset = self._n_data                               # compute_meshlets, mesh.rs:305-316
                words = np.zeros(len(vids) + (len(tris) + 3) // 4, np.uint32)
                words[:len(vids)] = vids
                words[len(vids):].view(np.uint8)[:len(tris)] = tris
                self.meshlet_data.append(words); self._n_data += len(words)
                ml[i]["vertex_offset"] = mesh_vertex_first + lod_vertex_offset     # submesh.vertex_offset + vertex_range.start
                ml[i]["data_offset"] = data_offset
                ml[i]["material_index"] = material_index
                ml[i]["vertex_count"] = len(vids)
                ml[i]["triangle_count"] = len(tris) // 3
            self.meshlets.append(ml); self._n_meshlets += len(ml)
            info["mesh_lods"][0, k] = (meshlet_offset, len(ml))
            self.vertices.append(verts); self._n_vertices += len(verts); lod_vertex_offset += len(verts)
        for k in range(len(lods), L.MAX_MESH_LODS):                       # unused slots: never indexed (min(lod, lod_count-1))
            info["mesh_lods"][0, k] = info["mesh_lods"][0, len(lods) - 1]
        self.vertex_ranges.append((mesh_vertex_first, self._n_vertices - mesh_vertex_first))
        self.mesh_infos.append(info)
        return len(self.mesh_infos) - 1

    def arrays(self):
        """(vertices, meshlet_data, meshlets, mesh_infos, vertex_ranges) with every bound still zero."""
        return (np.concatenate(self.vertices), np.concatenate(self.meshlet_data), np.concatenate(self.meshlets),
                np.concatenate(self.mesh_infos), np.array(self.vertex_ranges, np.uint32).reshape(-1))

    def finish(self, context):
        """Uploads the arrays and computes every bound on the GPU (orbit_meshlet_bounds + orbit_mesh_bounds). Returns the
        arrays as numpy, bounds filled in."""
        import torch
        from . import _lib
        lib = _lib.lib()
        vertices, data, meshlets, infos, ranges = self.arrays()
        d_v, d_d, d_m = context.upload(vertices), context.upload(data), context.upload(meshlets)
        d_i, d_r = context.upload(infos), context.upload(ranges)
        s = C.c_void_p(torch.cuda.current_stream(context.device).cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(lib.orbit_meshlet_bounds(context._h, p(d_v), vertex_dtype.itemsize, p(d_d), p(d_m), len(meshlets), s), "orbit_meshlet_bounds")
        _lib.check(lib.orbit_mesh_bounds(context._h, p(d_v), vertex_dtype.itemsize, p(d_r), p(d_i), len(infos), s), "orbit_mesh_bounds")
        torch.cuda.synchronize(context.device)
        return (vertices, data, d_m.cpu().numpy().view(L.meshlet_dtype).reshape(-1).copy(),
                d_i.cpu().numpy().view(L.mesh_info_dtype).reshape(-1).copy(), ranges)


def lod_chain(half_x, half_y, half_z, n0, seed=0, max_lods=L.MAX_MESH_LODS):
    """Tessellation levels whose index counts shrink by ~0.8 per LOD (the reference's index_count_scale *= 0.8, mod.rs:390),
    until the grid stops getting coarser."""
    lods, n_prev = [], None
    for k in range(max_lods):
        n = max(1, int(round(n0 * np.sqrt(0.8 ** k))))
        if n == n_prev:
            break
        lods.append(building_mesh(half_x, half_y, half_z, n, seed=seed))
        n_prev = n
    return lods


def city_from_geometry(seed, n_meshes, n_entities, grid, n0=12, pitch=12.0):
    """A small city whose meshes are real tessellated geometry with LOD chains. Returns (AssetBuilder, entity placement dict);
    `scene_from_assets` turns the finished arrays into a scenes.Scene."""
    from .scenes import Stream
    s = Stream(seed, "geometry")
    h = s.uniform(n_meshes, 1.0, 8.0)
    builder = AssetBuilder()
    for m in range(n_meshes):
        builder.add_mesh(lod_chain(2.0, float(h[m]), 2.0, n0, seed=seed + m), material_index=m % 16)
    return builder, {"heights": h, "n_entities": n_entities, "grid": grid, "pitch": pitch, "seed": seed}


def scene_from_assets(name, placement, meshlets, mesh_infos):
    """scenes.Scene over finished asset arrays: entities on a ground lattice with yaw rotation and uniform scale, as scenes.make_scene."""
    from . import scenes
    from .scenes import Scene, Stream, _quat_to_mat
    seed, N = placement["seed"], placement["n_entities"]
    nx, nz = placement["grid"]
    n_meshes = len(mesh_infos)
    se = Stream(seed, "entity")
    T = np.stack([(np.arange(N) % nx) * placement["pitch"], np.zeros(N), (np.arange(N) // nx) * placement["pitch"]], axis=1).astype(np.float64)
    yaw = se.uniform(N, 0.0, 2 * np.pi)
    q = np.stack([np.zeros(N), np.sin(yaw / 2), np.zeros(N), np.cos(yaw / 2)], axis=1)
    scale = se.uniform(N, 0.5, 2.0)
    R = _quat_to_mat(q)
    model = np.zeros((N, 4, 4)); model[:, :3, :3] = R * scale[:, None, None]; model[:, :3, 3] = T; model[:, 3, 3] = 1.0
    entities = np.zeros(N, L.entity_dtype)
    entities["model_matrix"] = model.transpose(0, 2, 1).astype(np.float32)
    nm = np.zeros((N, 4, 4)); nm[:, :3, :3] = R / scale[:, None, None]; nm[:, 3, 3] = 1.0
    entities["normal_matrix"] = nm.transpose(0, 2, 1).astype(np.float32)
    mesh_index = np.arange(N) % n_meshes
    lod0 = mesh_infos["mesh_lods"][:, 0, 1].astype(np.int64)[mesh_index]
    words = (lod0 + 31) // 32
    draws = np.zeros(N, L.entity_draw_dtype)
    draws["entity_index"] = np.arange(N, dtype=np.uint32)
    draws["mesh_index"] = mesh_index.astype(np.uint32)
    draws["visibility_offset"] = (np.cumsum(words) - words).astype(np.uint32)       # scene.rs:422-431: consecutive ranges
    entity_draws = np.zeros(4 + 12 * N, np.uint8)
    entity_draws[:4] = np.frombuffer(np.uint32(N).tobytes(), np.uint8)
    entity_draws[4:] = draws.view(np.uint8)
    sm = Stream(seed, "material")
    u = sm.uniform(16)
    alpha = np.where(u < 0.7, 0, np.where(u < 0.9, 1, 2)).astype(np.uint32)
    materials = np.zeros(16 * L.MATERIAL_STRIDE, np.uint8)
    materials.view(np.uint32).reshape(16, L.MATERIAL_STRIDE // 4)[:, L.MATERIAL_ALPHA_OFFSET // 4] = alpha
    transforms = np.zeros(N, L.transform_dtype)
    transforms["position"] = T.astype(np.float32); transforms["orientation"] = q.astype(np.float32)
    transforms["scale"] = scale.astype(np.float32)[:, None]
    sphere = mesh_infos["bounding_sphere"].astype(np.float64)[mesh_index]
    entity_pos = np.einsum("nij,nj->ni", model[:, :3, :3], sphere[:, :3]) + T
    entity_radius = sphere[:, 3] * scale
    pad = float(entity_radius.max())
    return Scene(name=name, seed=seed, meshlets=meshlets, mesh_infos=mesh_infos, materials=materials, entities=entities,
                 entity_draws=entity_draws, n_entities=N, n_meshlet_instances=int(lod0.sum()), n_records_lod0=int(words.sum()),
                 n_visibility_words=int(words.sum()), aabb_min=entity_pos.min(axis=0) - pad, aabb_max=entity_pos.max(axis=0) + pad,
                 entity_pos=entity_pos, entity_radius=entity_radius, entity_occluder=2.0 * scale * 0.9, transforms=transforms)
