// orbit_host_frame.cpp — drives two frames of the reference's depth-prepass culling protocol
// (forward.rs:266-403: EARLY pass 1 -> Hi-Z update -> LATE pass 2) through the C++ host mirror
// include/orbit_passes.hpp, i.e. the compiled-language host side above the C ABI, with plain cudaMalloc'd buffers.
// Test program: tests/test_gpu_cpp_host.py writes the inputs, runs this, and compares every output with the oracle.
//
//   orbit_host_frame <dir>     reads <dir>/{meta,meshlets,mesh_infos,materials,entities,entity_draws,depth}.bin
//                              and, when present, <dir>/{transforms,mesh_slots}.bin: the entity buffers are then produced on
//                              the GPU by SceneData::update_scene (orbit_scene_update) instead of being uploaded
//                              writes <dir>/f<k>_{early,late}_{dispatch,draws}.bin (k = 0, 1), f2_{early,late,main}_* (LATE + MAIN fused),
//                              entity_vis.bin, meshlet_vis.bin, hiz.bin
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "orbit_passes.hpp"

using namespace orbit_host;

struct Meta {            // written by the Python side, little-endian, 4-byte fields
    uint32_t width, height, n_entities, record_capacity, draw_capacity, n_vis_words, n_planes, lod_start, lod_end;
    float fov, near_clip, lod_base, lod_step;
    float view[16];      // column-major
    float planes[12][4];
};

static std::vector<unsigned char> read_file(const std::string& p) {
    FILE* f = std::fopen(p.c_str(), "rb");
    if (!f) { std::fprintf(stderr, "cannot open %s\n", p.c_str()); std::exit(2); }
    std::fseek(f, 0, SEEK_END); long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> v((size_t)n);
    if (n && std::fread(v.data(), 1, (size_t)n, f) != (size_t)n) std::exit(2);
    std::fclose(f);
    return v;
}
static void write_file(const std::string& p, const void* d, size_t n) {
    FILE* f = std::fopen(p.c_str(), "wb");
    if (!f || (n && std::fwrite(d, 1, n, f) != n)) { std::fprintf(stderr, "cannot write %s\n", p.c_str()); std::exit(2); }
    std::fclose(f);
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(3); } } while (0)
static void dump(const std::string& p, const void* d, size_t n);
static void* upload(const std::vector<unsigned char>& h) {
    void* d = nullptr; CU(cudaMalloc(&d, h.size() ? h.size() : 16)); CU(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice)); return d;
}
static void dump(const std::string& p, const void* d, size_t n) {
    std::vector<unsigned char> h(n); CU(cudaMemcpy(h.data(), d, n, cudaMemcpyDeviceToHost)); write_file(p, h.data(), n);
}

int main(int argc, char** argv) {
    if (argc < 2) return 1;
    const std::string dir = argv[1];
    auto mb = read_file(dir + "/meta.bin");
    Meta m; std::memcpy(&m, mb.data(), sizeof(m));
    orbit_ctx* ctx = nullptr;
    check(orbit_ctx_create(0, &ctx), "orbit_ctx_create");
    cudaStream_t stream; CU(cudaStreamCreate(&stream));
    AssetGraphData assets{upload(read_file(dir + "/mesh_infos.bin")), upload(read_file(dir + "/meshlets.bin")), upload(read_file(dir + "/materials.bin"))};
    SceneGraphData scene{m.n_entities, upload(read_file(dir + "/entity_draws.bin")), upload(read_file(dir + "/entities.bin"))};
    {   // optional: build the entity buffers with the scene update (scene.rs:404-492) from Transforms + mesh slots
        FILE* probe = std::fopen((dir + "/transforms.bin").c_str(), "rb");
        if (probe) {
            std::fclose(probe);
            SceneData sd{};
            std::vector<unsigned char> vo((size_t)m.n_entities * 4, 0xFF);            // no visibility ranges yet
            std::vector<unsigned char> cursor(4, 0);
            sd.buffers.transforms = (const OrbitTransform*)upload(read_file(dir + "/transforms.bin"));
            sd.buffers.mesh_slots = (const uint32_t*)upload(read_file(dir + "/mesh_slots.bin"));
            sd.buffers.visibility_offsets = (uint32_t*)upload(vo);
            sd.buffers.visibility_cursor = (uint32_t*)upload(cursor);
            sd.buffers.n_entities = m.n_entities; sd.buffers.visibility_capacity_words = 1u << 26;
            void *ed = nullptr, *dr = nullptr;
            CU(cudaMalloc(&ed, (size_t)m.n_entities * 128)); CU(cudaMalloc(&dr, 4 + (size_t)m.n_entities * 12));
            sd.buffers.entity_data = ed; sd.buffers.entity_draws = dr;
            sd.update_scene(ctx, assets, stream);
            scene = sd.import_to_graph();
            CU(cudaStreamSynchronize(stream));
            dump(dir + "/entity_data_out.bin", ed, (size_t)m.n_entities * 128);
            dump(dir + "/entity_draws_out.bin", dr, 4 + (size_t)m.n_entities * 12);
        }
    }
    float* depth = (float*)upload(read_file(dir + "/depth.bin"));
    uint32_t *entity_vis = nullptr, *meshlet_vis = nullptr;
    const size_t ev_words = (m.n_entities + 31) / 32 + 1, mv_words = m.n_vis_words ? m.n_vis_words : 1;
    CU(cudaMalloc(&entity_vis, ev_words * 4)); CU(cudaMemset(entity_vis, 0, ev_words * 4));     // frame 0: all zero
    CU(cudaMalloc(&meshlet_vis, mv_words * 4)); CU(cudaMemset(meshlet_vis, 0, mv_words * 4));
    void *dispatch = nullptr, *draws = nullptr;
    const size_t dispatch_bytes = 12 + 16 * (size_t)m.record_capacity, draw_bytes = 4 + 28 * (size_t)m.draw_capacity;
    CU(cudaMalloc(&dispatch, dispatch_bytes)); CU(cudaMalloc(&draws, draw_bytes));
    CU(cudaMemset(dispatch, 0, dispatch_bytes));   // the meshlet stage reads ahead of the record count (orbit_cuda.h, conventions)
    DepthPyramid pyramid(ctx, m.width, m.height);

    CullInfo cull;
    std::memcpy(cull.view_matrix, m.view, 64);
    for (uint32_t i = 0; i < m.n_planes; ++i) for (int k = 0; k < 4; ++k) cull.view_space_cull_planes.push_back(m.planes[i][k]);
    cull.projection = Projection::perspective(m.fov, m.near_clip);
    cull.lod_range_start = m.lod_start; cull.lod_range_end = m.lod_end; cull.lod_base = m.lod_base; cull.lod_step = m.lod_step;

    for (int f = 0; f < 2; ++f) {
        for (int late = 0; late < 2; ++late) {
            cull.occlusion_culling = OcclusionCullInfo{};
            cull.occlusion_culling.kind = late ? OcclusionCullInfo::VisibilityWrite : OcclusionCullInfo::VisibilityRead;
            cull.occlusion_culling.visibility_buffer = entity_vis;
            cull.occlusion_culling.meshlet_visibility_buffer = meshlet_vis;
            cull.occlusion_culling.aspect_ratio = (float)m.width / (float)m.height;
            if (late) {
                pyramid.update(depth, stream);                                  // forward.rs:362-367
                cull.occlusion_culling.depth_pyramid = pyramid.get_current();
            }
            create_meshlet_dispatch_command(ctx, assets, scene, cull, dispatch, m.record_capacity, stream);
            create_meshlet_draw_commands(ctx, assets, scene, cull, dispatch, m.record_capacity, draws, m.draw_capacity, stream);
            CU(cudaStreamSynchronize(stream));
            const std::string tag = dir + "/f" + std::to_string(f) + (late ? "_late" : "_early");
            uint32_t nrec = 0, ndraw = 0;
            CU(cudaMemcpy(&nrec, dispatch, 4, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(&ndraw, draws, 4, cudaMemcpyDeviceToHost));
            dump(tag + "_dispatch.bin", dispatch, 12 + 16 * (size_t)nrec);
            dump(tag + "_draws.bin", draws, 4 + 28 * (size_t)ndraw);
        }
    }
    {   // ---- frame 2: EARLY as before, then LATE + MAIN (forward.rs:518-548) through the fused wrapper
        void *main_dispatch = nullptr, *main_draws = nullptr;
        CU(cudaMalloc(&main_dispatch, dispatch_bytes)); CU(cudaMalloc(&main_draws, draw_bytes));
        CU(cudaMemset(main_dispatch, 0, dispatch_bytes));
        auto dump_pass = [&](const std::string& tag, void* disp, void* drw) {
            uint32_t nrec = 0, ndraw = 0;
            CU(cudaMemcpy(&nrec, disp, 4, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(&ndraw, drw, 4, cudaMemcpyDeviceToHost));
            dump(tag + "_dispatch.bin", disp, 12 + 16 * (size_t)nrec);
            dump(tag + "_draws.bin", drw, 4 + 28 * (size_t)ndraw);
        };
        CullInfo early = cull, late = cull;
        early.occlusion_culling = OcclusionCullInfo{};
        early.occlusion_culling.kind = OcclusionCullInfo::VisibilityRead;
        early.occlusion_culling.visibility_buffer = entity_vis;
        early.occlusion_culling.meshlet_visibility_buffer = meshlet_vis;
        early.occlusion_culling.aspect_ratio = (float)m.width / (float)m.height;
        create_meshlet_dispatch_command(ctx, assets, scene, early, dispatch, m.record_capacity, stream);
        create_meshlet_draw_commands(ctx, assets, scene, early, dispatch, m.record_capacity, draws, m.draw_capacity, stream);
        CU(cudaStreamSynchronize(stream));
        dump_pass(dir + "/f2_early", dispatch, draws);
        late.occlusion_culling = early.occlusion_culling;
        late.occlusion_culling.kind = OcclusionCullInfo::VisibilityWrite;
        pyramid.update(depth, stream);
        late.occlusion_culling.depth_pyramid = pyramid.get_current();
        const bool fused = create_late_and_main_commands(ctx, assets, scene, late, early /* MAIN = pass 1 again */, dispatch, main_dispatch,
                                                         m.record_capacity, draws, main_draws, m.draw_capacity, stream);
        if (!fused) { std::fprintf(stderr, "LATE / MAIN pair unexpectedly incompatible\n"); return 5; }
        CU(cudaStreamSynchronize(stream));
        dump_pass(dir + "/f2_late", dispatch, draws);
        dump_pass(dir + "/f2_main", main_dispatch, main_draws);
    }
    dump(dir + "/entity_vis.bin", entity_vis, ev_words * 4);
    dump(dir + "/meshlet_vis.bin", meshlet_vis, mv_words * 4);
    OrbitHizInfo info; check(orbit_hiz_info(pyramid.get_current(), &info), "orbit_hiz_info");
    dump(dir + "/hiz.bin", info.texels, (size_t)info.total_texels * 4);
    OrbitStatus st; int rc = orbit_ctx_poll_status(ctx, &st);
    std::printf("ok status=%d launches=%llu\n", rc, (unsigned long long)orbit_ctx_launch_count(ctx));
    orbit_ctx_destroy(ctx);
    return rc == ORBIT_OK ? 0 : 4;
}
