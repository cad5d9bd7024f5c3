"""GPU: the compiled-language host side (include/orbit_passes.hpp, C++) drives two frames of the reference's
depth-prepass protocol with plain cudaMalloc'd buffers — no Python, no torch on the data path — and every output
file is byte-identical to the oracle's."""
import os
import struct
import subprocess

import numpy as np
import pytest

from orbit_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("with_scene_update", [False, True])
def test_cpp_host_mirror_two_frames(tmp_path, oracle, with_scene_update):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = tmp_path / "orbit_host_frame"
    cuda = "/usr/local/cuda"
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", cuda + "/include",
                           os.path.join(ROOT, "tests", "cpp", "orbit_host_frame.cpp"), "-o", str(exe),
                           "-L", os.path.join(ROOT, "orbit_b200", "lib"), "-lorbit_b200", "-L", cuda + "/lib64", "-lcudart",
                           "-Wl,-rpath," + os.path.join(ROOT, "orbit_b200", "lib"), "-Wl,-rpath," + cuda + "/lib64"])
    sc, view = scenes.config_c1(scale=0.4, lods=(100, 40))
    view.lod_base, view.lod_step = 10.0, 1.6
    depth = scenes.make_depth(sc, view)
    d = str(tmp_path)
    for name, arr in (("meshlets", sc.meshlets), ("mesh_infos", sc.mesh_infos), ("materials", sc.materials), ("entities", sc.entities),
                      ("entity_draws", sc.entity_draws), ("depth", depth)):
        np.ascontiguousarray(arr).view(np.uint8).tofile(os.path.join(d, name + ".bin"))
    planes = np.zeros((12, 4), np.float32); planes[:len(view.planes)] = view.planes
    meta = struct.pack("<9I4f", view.width, view.height, sc.n_entities, sc.n_records_lod0, sc.n_meshlet_instances, sc.n_visibility_words,
                       len(view.planes), view.lod_range[0], view.lod_range[1], np.float32(view.fov), np.float32(view.near),
                       np.float32(view.lod_base), np.float32(view.lod_step))
    meta += view.view.astype(np.float32).T.tobytes() + planes.tobytes()       # view matrix column-major
    open(os.path.join(d, "meta.bin"), "wb").write(meta)
    if with_scene_update:       # the C++ host builds the entity buffers with SceneData::update_scene; the oracle chain does the same
        from orbit_b200 import layouts as L
        sc.transforms.view(np.uint8).tofile(os.path.join(d, "transforms.bin"))
        sc.draws["mesh_index"].astype(np.uint32).tofile(os.path.join(d, "mesh_slots.bin"))
        vo = np.full(sc.n_entities, L.NO_VISIBILITY_RANGE, np.uint32)
        o_ed, o_draws, _ = oracle.scene_update(sc.transforms, sc.draws["mesh_index"].copy(), vo, sc.mesh_infos, np.zeros(1, np.uint32))
        sc.entities, sc.entity_draws = o_ed.copy(), o_draws.copy()
    out = subprocess.run([str(exe), d], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    if with_scene_update:
        assert np.array_equal(np.fromfile(os.path.join(d, "entity_data_out.bin"), np.uint8), sc.entities.view(np.uint8).reshape(-1))
        assert np.array_equal(np.fromfile(os.path.join(d, "entity_draws_out.bin"), np.uint8), sc.entity_draws)
    hs = oracle.HostScene(sc)
    for f in range(3):                 # frame 2 ends with LATE + MAIN through the fused wrapper (create_late_and_main_commands)
        o = oracle.depth_prepass_culling(hs, view, depth)
        if f == 2:
            o["main"] = oracle.main_pass_culling(hs, view)
        for k in (("early", "late") if f < 2 else ("early", "late", "main")):
            ohdr, orecs = oracle.parse_dispatch(o[k][0]); on, od = oracle.parse_draws(o[k][1])
            g_disp = np.fromfile(os.path.join(d, "f%d_%s_dispatch.bin" % (f, k)), np.uint8)
            g_draw = np.fromfile(os.path.join(d, "f%d_%s_draws.bin" % (f, k)), np.uint8)
            assert g_disp[:12].view(np.uint32).tolist() == ohdr.tolist(), (f, k)
            assert np.array_equal(g_disp[12:], orecs.view(np.uint8)), (f, k, "records")
            assert int(g_draw[:4].view(np.uint32)[0]) == on and np.array_equal(g_draw[4:], od.view(np.uint8)), (f, k, "draws")
    assert np.array_equal(np.fromfile(os.path.join(d, "entity_vis.bin"), np.uint32), hs.entity_visibility)
    assert np.array_equal(np.fromfile(os.path.join(d, "meshlet_vis.bin"), np.uint32), hs.meshlet_visibility)
    assert np.array_equal(np.fromfile(os.path.join(d, "hiz.bin"), np.uint32), hs.hiz_texels.view(np.uint32))
