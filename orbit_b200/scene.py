"""Host mirror of SceneData (src/scene.rs:358-502) for the part that feeds the culling path: the per-frame
`update_scene` that turns entity transforms + mesh handles into the entity-data / entity-draw buffers and hands out
meshlet-visibility ranges. Device memory through torch; all arithmetic in orbit_scene_update (CUDA).

Differences from the reference, all on the host-bookkeeping side: entities live in device arrays
(transforms / mesh slots / visibility offsets) instead of a Vec<EntityData>; the visibility allocator is the
reference's FreeListAllocator restricted to what update_scene uses (allocate only => a bump pointer); lights are not
produced here (scene.rs:455-474 stays host code)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import layouts as L
from .passes import SceneGraphData, _ptr, _stream

MESHLET_VISIBILITY_BUFFER_CHUNK_COUNT = 1 << 26   # words; the reference sizes its buffer the same way (scene.rs:392)


class SceneData:
    def __init__(self, context, max_entities, visibility_capacity_words=MESHLET_VISIBILITY_BUFFER_CHUNK_COUNT):
        dev = context.device
        self.context = context
        self.max_entities = int(max_entities)
        self.n_entities = 0
        self.visibility_capacity_words = int(visibility_capacity_words)
        self.transforms = torch.zeros(self.max_entities * L.transform_dtype.itemsize, dtype=torch.uint8, device=dev)
        self.mesh_slots = torch.full((self.max_entities,), -1, dtype=torch.int32, device=dev)            # 0xFFFFFFFF = no mesh
        self.visibility_offsets = torch.full((self.max_entities,), -1, dtype=torch.int32, device=dev)    # 0xFFFFFFFF = no range
        self.visibility_cursor = torch.zeros(1, dtype=torch.int32, device=dev)
        self.entity_data_buffer = torch.zeros(self.max_entities * 128, dtype=torch.uint8, device=dev)
        self.entity_draw_buffer = torch.zeros(L.ENTITY_DRAW_HEADER + 12 * self.max_entities, dtype=torch.uint8, device=dev)

    def set_entities(self, transforms, mesh_slots):
        """transforms: numpy array of layouts.transform_dtype; mesh_slots: uint32 (layouts.NO_MESH = no mesh). Entities
        keep their index, and with it their visibility range (scene.rs:422-423), across calls."""
        n = len(transforms)
        assert n <= self.max_entities and len(mesh_slots) == n and transforms.dtype == L.transform_dtype
        self.transforms[:n * 48] = torch.from_numpy(np.ascontiguousarray(transforms).view(np.uint8).reshape(-1)).to(self.context.device)
        self.mesh_slots[:n] = torch.from_numpy(np.ascontiguousarray(mesh_slots, dtype=np.uint32).view(np.int32)).to(self.context.device)
        self.n_entities = n

    def update_scene(self, assets):
        """scene.rs:404-492 (mesh part). Asynchronous on the current stream."""
        key = (assets.mesh_info_buffer.data_ptr(), self.entity_data_buffer.data_ptr(), self.entity_draw_buffer.data_ptr(), self.n_entities)
        if getattr(self, "_packed_key", None) != key:      # the argument block only changes when a buffer does
            u = L.SceneUpdate()
            u.transforms = self.transforms.data_ptr(); u.mesh_slots = self.mesh_slots.data_ptr()
            u.visibility_offsets = self.visibility_offsets.data_ptr(); u.mesh_infos = key[0]
            u.visibility_cursor = self.visibility_cursor.data_ptr()
            u.n_entities = self.n_entities; u.visibility_capacity_words = self.visibility_capacity_words
            u.entity_data = key[1]; u.entity_draws = key[2]
            self._packed, self._packed_key = u, key
        _lib.check(_lib.lib().orbit_scene_update(self.context._h, C.byref(self._packed), _stream()), "orbit_scene_update")

    def import_to_graph(self, meshlet_visibility_buffer=None, record_capacity=0, draw_capacity=0):
        """scene.rs:494-502. entity_draw_count is the host's upper bound (every entity); the culling kernels clamp to the
        device-side count the update wrote."""
        return SceneGraphData(entity_draw_count=self.n_entities, entity_draw_buffer=self.entity_draw_buffer,
                              entity_buffer=self.entity_data_buffer, meshlet_visibility_buffer=meshlet_visibility_buffer,
                              record_capacity=record_capacity, draw_capacity=draw_capacity)
