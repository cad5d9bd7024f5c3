"""CPU: the asset-side producers — the host builder's layouts follow GpuAssets::add_mesh
(assets/mod.rs:325-470) and compute_meshlets (assets/mesh.rs:292-338), and the oracle's restatement of
meshopt::compute_meshlet_bounds / MeshData::compute_bounds satisfies what those bounds exist for: every vertex of a meshlet
lies in its sphere, every triangle normal lies in its cone (the property the shader's coneCull relies on:
meshlet_cull.comp:104-106), the mesh sphere / AABB contain every vertex."""
import numpy as np

from orbit_b200 import assets
from orbit_b200 import layouts as L


def _meshlet_geometry(vertices, data, m):
    vo, do, vc, tc = int(m["vertex_offset"]), int(m["data_offset"]), int(m["vertex_count"]), int(m["triangle_count"])
    vids = data[do:do + vc]
    tris = data[do + vc:do + vc + (3 * tc + 3) // 4].view(np.uint8)[:3 * tc].reshape(-1, 3)
    pos = vertices["position"][vo + vids].astype(np.float64)
    return pos, tris


def test_builder_layouts_and_meshlet_limits():
    b = assets.AssetBuilder()
    lods = assets.lod_chain(2.0, 3.0, 2.0, 10, seed=3)
    assert 2 <= len(lods) <= L.MAX_MESH_LODS
    counts = [len(i) for _, i in lods]
    assert all(counts[k + 1] < counts[k] for k in range(len(counts) - 1))           # every LOD is coarser
    m0 = b.add_mesh(lods, material_index=5)
    m1 = b.add_mesh(lods[:1], material_index=6)
    vertices, data, meshlets, infos, ranges = b.arrays()
    assert (m0, m1) == (0, 1) and len(infos) == 2 and ranges.tolist() == [0, sum(len(v) for v, _ in lods), sum(len(v) for v, _ in lods), len(lods[0][0])]
    assert meshlets["vertex_count"].max() <= 64 and meshlets["triangle_count"].max() <= 64 and meshlets["triangle_count"].min() >= 1
    # LOD table: consecutive, absolute offsets; counts add up to the mesh's meshlets (mod.rs:398-400, 80-85)
    lt = infos["mesh_lods"][0]
    assert lt[0, 0] == 0 and all(lt[k + 1, 0] == lt[k, 0] + lt[k, 1] for k in range(len(lods) - 1))
    assert infos["mesh_lods"][1][0, 0] == lt[len(lods) - 1, 0] + lt[len(lods) - 1, 1]
    assert infos["lod_count"].tolist() == [len(lods), 1]
    # every meshlet reproduces its triangles: the partition covers LOD 0's index list in order
    tri_total = 0
    rebuilt = []
    for m in meshlets[:int(lt[0, 1])]:
        pos, tris = _meshlet_geometry(vertices, data, m)
        assert tris.max() < m["vertex_count"]
        vids = data[int(m["data_offset"]):int(m["data_offset"]) + int(m["vertex_count"])]
        rebuilt.append(vids[tris].reshape(-1))
        tri_total += len(tris)
    assert tri_total * 3 == len(lods[0][1]) and np.array_equal(np.concatenate(rebuilt), lods[0][1])


def test_oracle_bounds_contain_the_geometry(oracle):
    b = assets.AssetBuilder()
    for k in range(3):
        b.add_mesh(assets.lod_chain(2.0, 1.5 + k, 2.0, 9, seed=k), material_index=k)
    vertices, data, meshlets, infos, ranges = b.arrays()
    ml, mi, skipped = oracle.asset_bounds(vertices, data, meshlets, infos, ranges)
    assert skipped == 0
    useful = 0
    for m in ml:
        pos, tris = _meshlet_geometry(vertices, data, m)
        c, r = m["bounding_sphere"][:3].astype(np.float64), float(m["bounding_sphere"][3])
        assert np.linalg.norm(pos - c, axis=1).max() <= r * (1 + 1e-5) + 1e-6
        n = np.cross(pos[tris[:, 1]] - pos[tris[:, 0]], pos[tris[:, 2]] - pos[tris[:, 0]])
        n = n[np.linalg.norm(n, axis=1) > 0]
        n /= np.linalg.norm(n, axis=1, keepdims=True)
        axis = m["cone_axis"].astype(np.float64) / 127.0
        cutoff = float(m["cone_cutoff"]) / 127.0
        if m["cone_cutoff"] < 127:
            useful += 1
            # the quantised cone stays conservative: for every triangle normal, dot(normal, axis) >= sqrt(1 - cutoff^2) up to
            # the quantisation slack folded into the cutoff
            dp = n @ (axis / max(np.linalg.norm(axis), 1e-9))
            assert dp.min() >= np.sqrt(max(0.0, 1.0 - cutoff * cutoff)) - 2e-2
    assert useful > len(ml) // 2                          # the bumpy faces give real cones, not only the trivial-accept case
    for k in range(len(mi)):
        first, count = int(ranges[2 * k]), int(ranges[2 * k + 1])
        p = vertices["position"][first:first + count].astype(np.float64)
        assert np.all(p >= mi["aabb_min"][k][:3] - 1e-6) and np.all(p <= mi["aabb_max"][k][:3] + 1e-6)
        c, r = mi["bounding_sphere"][k][:3].astype(np.float64), float(mi["bounding_sphere"][k][3])
        assert np.linalg.norm(p - c, axis=1).max() <= r * (1 + 1e-6)
        assert np.allclose(c, (p.min(axis=0) + p.max(axis=0)) / 2, atol=1e-5)


def test_degenerate_and_oversized_meshlets(oracle):
    v = np.zeros(4, assets.vertex_dtype)
    v["position"] = [[0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1, 0]]
    data = np.zeros(8, np.uint32)
    data[:4] = [0, 1, 2, 3]
    data[4:].view(np.uint8)[:6] = [0, 1, 2, 0, 0, 0]      # two degenerate triangles (collinear, repeated vertex)
    ml = np.zeros(2, L.meshlet_dtype)
    ml[0]["vertex_count"], ml[0]["triangle_count"] = 4, 2
    ml[1]["vertex_count"], ml[1]["triangle_count"] = 4, 200          # more than the producers stage: skipped, reported
    ml["bounding_sphere"] = 7.0
    infos = np.zeros(1, L.mesh_info_dtype)
    out, _, skipped = oracle.asset_bounds(v, data, ml, infos, np.array([0, 4], np.uint32))
    assert skipped == 1
    assert out[0]["bounding_sphere"].tolist() == [0, 0, 0, 0] and out[0]["cone_cutoff"] == 0     # zeroed bounds: trivial reject
    assert out[1]["bounding_sphere"].tolist() == [7, 7, 7, 7]
