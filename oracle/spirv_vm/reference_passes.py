"""Drives the reference's SHIPPED compute shaders (/root/reference/shaders/*.comp.spv) through spirv_vm the way
src/passes/draw_gen.rs and src/passes/cluster.rs dispatch them: bindless descriptor indices in push constants,
GpuCullInfo in a storage buffer, indirect dispatch sizes read back from the buffers. TEST INFRASTRUCTURE."""
import ctypes as C
import os
import struct

import numpy as np

from spirv_vm import VM, Buffer, Image, Sampler

SHADERS = "/root/reference/shaders"
CONTRACT_FMA = False            # sensitivity experiments (tools/spirv_sensitivity.py) flip this; fixtures are made with False
REDUCE_MIN_SAMPLER = 6          # shaders/include/common.glsl:13
# descriptor indices handed out to the resources of one run (any distinct numbers do)
D = {"entity_draws": 3, "mesh_infos": 4, "dispatch": 5, "entities": 6, "cull_info": 7, "meshlets": 8, "draws": 9, "materials": 10,
     "entity_vis": 11, "meshlet_vis": 12, "pyramid": 13}


def available():
    return os.path.exists(os.path.join(SHADERS, "meshlet_cull.comp.spv"))


class Samplers(dict):
    def __missing__(self, k):
        return Sampler(reduce_min=(k == REDUCE_MIN_SAMPLER))


def hiz_build(depth, info, log2f=None):
    """DepthPyramid::update (draw_gen.rs:538-564): one depth_reduce dispatch per mip, each sampling the previous one."""
    vm = VM(os.path.join(SHADERS, "depth_reduce.comp.spv"), log2f=log2f, contract_fma=CONTRACT_FMA)
    vm.resources[(1, 0)] = Samplers()
    levels, src = [], np.ascontiguousarray(depth, np.float32)
    for l in range(info.levels):
        w, h = max(info.width >> l, 1), max(info.height >> l, 1)
        dst = np.zeros((h, w), np.float32)
        vm.resources[(1, 7)] = {1: Image([src])}
        vm.resources[(2, 0)] = {2: Image([dst])}
        vm.push = Buffer(struct.pack("<4I", w, h, 1, 2))
        vm.dispatch(((w + 15) // 16, (h + 15) // 16, 1), (16, 16, 1))     # draw_gen.rs:560 dispatch(ceil(w/16), ceil(h/16), 1)
        levels.append(dst)
        src = dst
    return levels


def _bind_cull(vm, scene, cull_info, entity_vis, meshlet_vis, pyramid_levels):
    g = type(cull_info).from_buffer_copy(bytes(cull_info))
    NO = 0xFFFFFFFF
    g.visibility_buffer = D["entity_vis"] if g.visibility_buffer != NO else NO
    g.meshlet_visibility_buffer = D["meshlet_vis"] if g.meshlet_visibility_buffer != NO else NO
    g.depth_pyramid = D["pyramid"] if g.depth_pyramid != NO else NO
    bufs = {D["entity_draws"]: Buffer(scene.entity_draws), D["mesh_infos"]: Buffer(scene.mesh_infos), D["entities"]: Buffer(scene.entities),
            D["meshlets"]: Buffer(scene.meshlets), D["materials"]: Buffer(scene.materials), D["cull_info"]: Buffer(bytes(g)),
            D["entity_vis"]: Buffer(entity_vis), D["meshlet_vis"]: Buffer(meshlet_vis)}
    vm.resources[(0, 0)] = bufs
    vm.resources[(1, 0)] = Samplers()
    vm.resources[(1, 7)] = {D["pyramid"]: Image(pyramid_levels)} if pyramid_levels is not None else {}
    return bufs


def entity_cull(scene, cull_info, entity_vis, meshlet_vis, pyramid_levels, record_capacity, log2f):
    """create_meshlet_dispatch_command (draw_gen.rs:327-380). Returns the MeshletDispatchBuffer bytes (records in the order
    the VM's invocations appended them); entity_vis is updated in place in pass 2."""
    vm = VM(os.path.join(SHADERS, "entity_cull.comp.spv"), spec={0: 32}, log2f=log2f, contract_fma=CONTRACT_FMA)    # MESHLET_DISPATCH_SIZE = task workgroup size 32
    bufs = _bind_cull(vm, scene, cull_info, entity_vis, meshlet_vis, pyramid_levels)
    out = np.zeros(12 + 16 * record_capacity, np.uint8)
    out[:12] = np.frombuffer(struct.pack("<3I", 0, 1, 1), np.uint8)                        # fill_buffer + {.,1,1}: draw_gen.rs:356-363
    bufs[D["dispatch"]] = Buffer(out)
    vm.push = Buffer(struct.pack("<5I", D["entity_draws"], D["mesh_infos"], D["dispatch"], D["entities"], D["cull_info"]))
    n = int(np.frombuffer(bytes(scene.entity_draws[:4]), np.uint32)[0])
    vm.dispatch(((n + 255) // 256, 1, 1), (256, 1, 1))                                      # draw_gen.rs:377
    entity_vis[:] = bufs[D["entity_vis"]].data.view(entity_vis.dtype)
    return bufs[D["dispatch"]].data.copy()


def meshlet_cull(scene, cull_info, entity_vis, meshlet_vis, pyramid_levels, dispatch, draw_capacity, log2f):
    """create_meshlet_draw_commands (draw_gen.rs:382-435): dispatch_indirect over the record count."""
    vm = VM(os.path.join(SHADERS, "meshlet_cull.comp.spv"), spec={0: 32}, log2f=log2f, contract_fma=CONTRACT_FMA)
    bufs = _bind_cull(vm, scene, cull_info, entity_vis, meshlet_vis, pyramid_levels)
    bufs[D["dispatch"]] = Buffer(dispatch)
    out = np.zeros(4 + 28 * draw_capacity, np.uint8)
    bufs[D["draws"]] = Buffer(out)
    vm.push = Buffer(struct.pack("<6I", D["dispatch"], D["meshlets"], D["draws"], D["entities"], D["cull_info"], D["materials"]))
    gx, gy, gz = struct.unpack("<3I", bytes(dispatch[:12]))
    vm.dispatch((gx, gy, gz), (32, 1, 1))
    meshlet_vis[:] = bufs[D["meshlet_vis"]].data.view(meshlet_vis.dtype)
    return bufs[D["draws"]].data.copy()


def light_cluster(params, depth, lights, log2f):
    """compute_clusters (cluster.rs:368-591): mark_active -> active_cluster_compaction -> light_culling (dispatch_indirect).
    params: layouts.ClusterParams (info + z_scale, z_bias). Returns dict of numpy outputs in the reference's buffers."""
    ci = params.info
    cx, cy, cz = ci.cluster_count[0], ci.cluster_count[1], ci.cluster_count[2]
    n = cx * cy * cz
    w, h = ci.screen_size[0], ci.screen_size[1]
    DD = {"masks": 3, "bounds": 4, "unique": 5, "index": 6, "lights": 7, "info": 8, "depth": 9, "image": 10}
    lights = np.ascontiguousarray(lights)
    masks = Buffer(np.zeros(cx * cy, np.uint32)); bounds = Buffer(np.zeros(2 * n, np.uint32))      # cluster.rs:438-455 fill_buffer 0
    unique = Buffer(np.zeros(4 + n, np.uint32))                                                       # cluster.rs:490-499
    cap = 256 * n
    index = Buffer(np.zeros(1 + cap, np.uint32))
    image = Image([np.zeros((cz, cy, cx, 2), np.uint32)])
    # ---- mark_active (cluster.rs:399-477)
    vm = VM(os.path.join(SHADERS, "light_cluster/mark_active.comp.spv"), log2f=log2f, contract_fma=CONTRACT_FMA)
    vm.resources[(0, 0)] = {DD["masks"]: masks, DD["bounds"]: bounds}
    vm.resources[(1, 0)] = Samplers()
    vm.resources[(1, 7)] = {DD["depth"]: Image([np.ascontiguousarray(depth, np.float32)])}
    vm.push = Buffer(struct.pack("<3I I 2I 4f 4I", cx, cy, cz, ci.tile_size_px, w, h, ci.z_near, ci.z_far, params.z_scale, params.z_bias,
                                 DD["depth"], 1, DD["masks"], DD["bounds"]))
    vm.dispatch(((w + 7) // 8, (h + 7) // 8, 1), (8, 8, 1))
    # ---- compaction (cluster.rs:479-517)
    vm = VM(os.path.join(SHADERS, "light_cluster/active_cluster_compaction.comp.spv"), log2f=log2f, contract_fma=CONTRACT_FMA)
    vm.resources[(0, 0)] = {DD["masks"]: masks, DD["unique"]: unique}
    vm.push = Buffer(struct.pack("<5I", cx, cy, cz, DD["masks"], DD["unique"]))
    vm.dispatch(((cx + 3) // 4, (cy + 3) // 4, (cz + 3) // 4), (4, 4, 4))
    # ---- light culling (cluster.rs:519-591)
    info = type(ci).from_buffer_copy(bytes(ci))
    info.unique_cluster_buffer, info.cluster_offset_image, info.light_index_buffer = DD["unique"], DD["image"], DD["index"]
    info.depth_bounds_buffer, info.global_light_list = DD["bounds"], DD["lights"]
    vm = VM(os.path.join(SHADERS, "light_cluster/light_culling.comp.spv"), log2f=log2f, contract_fma=CONTRACT_FMA)
    vm.resources[(0, 0)] = {DD["unique"]: unique, DD["bounds"]: bounds, DD["index"]: index, DD["lights"]: Buffer(lights), DD["info"]: Buffer(bytes(info))}
    vm.resources[(2, 0)] = {DD["image"]: image}
    vm.push = Buffer(struct.pack("<I", DD["info"]))
    gx, gy, gz = struct.unpack("<3I", unique.read(0, 12))
    vm.dispatch((gx, gy, gz), (256, 1, 1))
    return {"masks": masks.data.view(np.uint32).copy(), "bounds": bounds.data.view(np.uint32).copy(), "unique": unique.data.view(np.uint32).copy(),
            "image": image.levels[0].reshape(-1).copy(), "index": index.data.view(np.uint32).copy()}


def task_shader(scene, cull_info, entity_vis, meshlet_vis, pyramid_levels, dispatch, log2f, shader="forward/forward_depth_prepass.task.spv"):
    """The mesh-shading path's task stage (context.rs:1093-1099: vkCmdDrawMeshTasksIndirectEXT over the dispatch buffer):
    one 32-lane workgroup per dispatch record. Returns [(record index, emitted task count, payload)] with payload =
    [entity_index, meshlet_offset, [32 x u8 meshlet indices]] as the shader left it."""
    vm = VM(os.path.join(SHADERS, shader), spec={0: 32}, log2f=log2f, contract_fma=CONTRACT_FMA)
    bufs = _bind_cull(vm, scene, cull_info, entity_vis, meshlet_vis, pyramid_levels)
    bufs[D["dispatch"]] = Buffer(dispatch)
    # push constants by member name (the three task shaders lay the block out differently); matrices stay identity,
    # they are only read by the mesh stage
    values = {"draw_command_buffer": D["dispatch"], "cull_info_buffer": D["cull_info"], "meshlet_buffer": D["meshlets"],
              "entity_buffer": D["entities"], "materials_buffer": D["materials"]}
    push = bytearray(256)
    m = vm.m
    for vid, inst in vm.globals.items():
        pt = vm.types[inst.words[0]]
        if pt[1] != 9:
            continue
        for i, mt in enumerate(vm.types[pt[2]][1]):
            name, off = m.member_names.get((pt[2], i)), m.member_decor[(pt[2], i)][35][0]
            if name in values:
                push[off:off + 4] = struct.pack("<I", values[name])
            elif vm.types[mt][0] == "mat":
                push[off:off + 64] = np.eye(4, dtype=np.float32).tobytes()
    vm.push = Buffer(bytes(push))
    gx, gy, gz = struct.unpack("<3I", bytes(dispatch[:12]))
    vm.dispatch((gx, gy, gz), (32, 1, 1))
    meshlet_vis[:] = bufs[D["meshlet_vis"]].data.view(meshlet_vis.dtype)
    out = []
    for (g, emitted) in vm.emitted:
        assert len(emitted) >= 1, "workgroup %s emitted nothing" % (g,)
        counts, payload = emitted[0]
        assert all(e[0] == counts for e in emitted)
        out.append((g[0], counts[0], payload))
    return out
