"""CPU: the oracle's restatement of SceneData::update_scene (scene.rs:404-492) against independent restatements:
float64 matrices, and a pure-Python transcription of the reference loop driving a restated FreeListAllocator
(collections/freelist_alloc.rs:40-72, best fit + split) — which shows that with nothing freed the allocator is a bump
pointer, the property the CUDA prefix scan relies on."""
import numpy as np

from orbit_b200 import layouts as L
from scene_update_cases import random_entities, random_mesh_infos


class FreeListAllocator:
    """collections/freelist_alloc.rs: blocks = [(free, start, end)], allocate = smallest free block that fits, split at its start."""

    def __init__(self, size):
        self.blocks = [[True, 0, size]]

    def allocate(self, size):
        best = None
        for i, (free, a, b) in enumerate(self.blocks):
            if free and b - a >= size and (best is None or b - a < self.blocks[best][2] - self.blocks[best][1]):
                best = i
        if best is None:
            return None
        free, a, b = self.blocks[best]
        if b - a == size:
            self.blocks[best][0] = False
            return a
        self.blocks.insert(best, [False, a, a + size])
        self.blocks[best + 1][1] = a + size
        return a


def reference_loop(slots, vo, mesh_infos, capacity):
    """scene.rs:419-444 transcribed; starts from an allocator in which the preallocated ranges are already taken."""
    alloc = FreeListAllocator(capacity)
    taken = int(vo[vo != L.NO_VISIBILITY_RANGE].max()) + 7 if (vo != L.NO_VISIBILITY_RANGE).any() else 0
    if taken:
        assert alloc.allocate(taken) == 0
    draws, vo = [], vo.copy()
    for e in range(len(slots)):
        if slots[e] == L.NO_MESH:
            continue
        if vo[e] == L.NO_VISIBILITY_RANGE:
            mc = int(mesh_infos["mesh_lods"][slots[e], 0, 1])
            vo[e] = alloc.allocate(-(-mc // 32))
        draws.append((len(draws), int(slots[e]), int(vo[e])))
    return draws, vo


def test_draws_and_visibility_ranges_match_reference_loop(oracle):
    for seed, pre in ((1, 0.0), (2, 0.3)):
        t, slots, vo, cursor = random_entities(1500, 40, seed, preallocated_fraction=pre)
        mi = random_mesh_infos(40, seed)
        want_draws, want_vo = reference_loop(slots, vo, mi, 1 << 26)
        ed, draws, ovf = oracle.scene_update(t, slots, vo, mi, cursor)
        got = draws[4:].view(L.entity_draw_dtype)
        assert ovf == 0 and len(got) == len(want_draws) == len(ed)
        assert [(int(d["entity_index"]), int(d["mesh_index"]), int(d["visibility_offset"])) for d in got] == want_draws
        assert np.array_equal(vo, want_vo)
        # a second update allocates nothing
        c2 = cursor.copy()
        _, draws2, _ = oracle.scene_update(t, slots, vo, mi, c2)
        assert np.array_equal(draws2, draws) and c2[0] == cursor[0]


def test_matrices_against_float64(oracle):
    t, slots, vo, cursor = random_entities(2000, 8, 3, no_mesh_fraction=0.0)
    mi = random_mesh_infos(8, 3)
    ed, _, _ = oracle.scene_update(t, slots, vo, mi, cursor)
    q = t["orientation"].astype(np.float64); s = t["scale"].astype(np.float64); p = t["position"].astype(np.float64)
    x, y, z, w = q.T
    R = np.empty((len(q), 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    M = R * s[:, None, :]                                    # column j scaled by s_j
    model = ed["model_matrix"].astype(np.float64).transpose(0, 2, 1)      # -> [row, col]
    assert np.abs(model[:, :3, :3] - M).max() < 4e-6
    assert np.array_equal(model[:, :3, 3], p) and np.all(model[:, 3, 3] == 1.0) and np.all(model[:, 3, :3] == 0.0)
    normal = ed["normal_matrix"].astype(np.float64).transpose(0, 2, 1)
    want = np.linalg.inv(M).transpose(0, 2, 1)
    assert np.abs(normal[:, :3, :3] - want).max() < 2e-5 * np.abs(want).max()
    assert np.all(normal[:, 3, :3] == 0.0) and np.all(normal[:, :3, 3] == 0.0) and np.all(normal[:, 3, 3] == 1.0)


def test_negative_scale_keeps_signed_zero_row(oracle):
    # glam multiplies the whole Vec4 axis (w lane = 0.0) by the scale component: a negative scale stores -0.0
    t = np.zeros(1, L.transform_dtype)
    t["orientation"][0] = (0, 0, 0, 1); t["scale"][0] = (-2.0, 1.0, 3.0)
    mi = random_mesh_infos(1, 1)
    ed, _, _ = oracle.scene_update(t, np.zeros(1, np.uint32), np.full(1, L.NO_VISIBILITY_RANGE, np.uint32), mi, np.zeros(1, np.uint32))
    m = ed["model_matrix"][0]
    assert m[0, 0] == -2.0 and np.signbit(m[0, 3]) and not np.signbit(m[1, 3])


def test_visibility_overflow_is_reported(oracle):
    t, slots, vo, cursor = random_entities(300, 5, 4, no_mesh_fraction=0.0)
    mi = random_mesh_infos(5, 4)
    mi["mesh_lods"][:, :, 1] = 64
    _, _, ovf = oracle.scene_update(t, slots, vo.copy(), mi, cursor.copy(), capacity_words=599)
    assert ovf == 1
    _, _, ovf = oracle.scene_update(t, slots, vo.copy(), mi, cursor.copy(), capacity_words=600)
    assert ovf == 0
