// meshlet_cull.cu — per-meshlet frustum / normal-cone / two-pass Hi-Z culling + ordered compaction (sm_100a).
//
// Stands in for shaders/meshlet_cull.comp:108-255 driven by create_meshlet_draw_commands
// (src/passes/draw_gen.rs:382-435) and, for the optional task payload output, the task-shader twins
// (shaders/forward/forward_depth_prepass.task:224-256).
//
// B200 design (not the reference's one-workgroup-per-record + atomicAdd):
//   * persistent CTAs pull tiles of kRecsPerTile dispatch records through an atomic ticket; the record count is
//     read on the device (the reference's dispatch_indirect) — no host round trip;
//   * a warp owns kRecsPerWarp consecutive records; lane = meshlet, exactly the reference's 32-lane group, so
//     the ballot-based visibility word is the same bit pattern;
//   * all meshlet loads of a warp's records are issued up front as 2 x 128-bit non-coherent loads per lane
//     (32 B meshlet, 1 KB contiguous per record) to keep >= 4 KB per warp in flight;
//   * view*model is computed once per record by 16 lanes and broadcast through shared memory (the reference
//     recomputes the 4x4 product in every lane), and reused while consecutive records share an entity;
//   * in pass 1 lanes whose visibility bit is clear never load their meshlet (their result is "not drawn"
//     whatever the meshlet is), so the early pass touches only last frame's visible meshlets;
//   * survivors are ranked by ballot+popc inside the warp, staged as 7-word commands in shared memory, ordered
//     across warps by a CTA scan and across CTAs by decoupled look-back (scan.cuh), then streamed out as
//     contiguous 4-byte-coalesced words: draw order = (record index, lane), independent of scheduling.
#include "params.cuh"

namespace orbit {

constexpr int kMcWarps = 8;
constexpr int kMcThreads = kMcWarps * 32;


template <int kRecsPerWarp>
__global__ void __launch_bounds__(kMcThreads) meshlet_cull_kernel(const __grid_constant__ MeshletCullParams p) {
    constexpr int kRecsPerTile = kMcWarps * kRecsPerWarp;
    __shared__ float s_view[16];
    __shared__ __align__(16) float s_mv[kMcWarps][16];
    __shared__ uint32_t s_stage[kMcWarps][kRecsPerWarp * 32 * 7];
    __shared__ uint32_t s_warp_total[kMcWarps];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_base;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const OrbitCullInfo& ci = p.cull;
    if (tid < 16) s_view[tid] = ci.view_matrix.m[tid >> 2][tid & 3];

    const unsigned int epoch = scan_epoch(p.scan);
    uint32_t nrec = __ldcg(p.dispatch_words);  // workgroup_count_x written by the entity stage
    if ((uint64_t)nrec > p.capacity_records) nrec = (uint32_t)p.capacity_records;
    const uint32_t ntiles = (nrec + kRecsPerTile - 1) / kRecsPerTile;
    const uint32_t pass = ci.occlusion_pass;
    const bool mocc = ci.meshlet_visibility_buffer != ORBIT_NO_BUFFER;
    const bool use_vis = (pass == 1u || pass == 2u) && mocc;
    const bool pass2 = (pass == 2u) && mocc;
    const float K = 0.007874015718698502f;

    while (true) {
        __syncthreads();  // s_tile / staging reuse
        if (tid == 0) s_tile = atomicAdd(p.scan.ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) {
            if (tile == 0u && tid == 0) p.draw_words[0] = 0u;  // empty dispatch: count = 0 (fill_buffer, draw_gen.rs:411-417)
            break;
        }
        const uint32_t rec0 = tile * kRecsPerTile + warp * kRecsPerWarp;

        // ---- records of this warp: kRecsPerWarp x 4 words, one coalesced load, then register broadcast
        uint32_t my_word = 0u;
        {
            const uint32_t w = lane;  // kRecsPerWarp*4 <= 32
            const uint32_t r = rec0 + (w >> 2);
            if (w < (uint32_t)kRecsPerWarp * 4u && r < nrec) my_word = __ldcg(p.dispatch_words + 3u + (size_t)rec0 * 4u + w);
        }
        uint32_t r_entity[kRecsPerWarp], r_offset[kRecsPerWarp], r_count[kRecsPerWarp], r_vo[kRecsPerWarp];
#pragma unroll
        for (int r = 0; r < kRecsPerWarp; ++r) {
            r_entity[r] = __shfl_sync(0xFFFFFFFFu, my_word, r * 4 + 0);
            r_offset[r] = __shfl_sync(0xFFFFFFFFu, my_word, r * 4 + 1);
            r_count[r] = (rec0 + r < nrec) ? __shfl_sync(0xFFFFFFFFu, my_word, r * 4 + 2) : 0u;
            r_vo[r] = __shfl_sync(0xFFFFFFFFu, my_word, r * 4 + 3);
        }
        // ---- visibility words (one per record), then all meshlet loads up front
        uint32_t vis_word[kRecsPerWarp];
#pragma unroll
        for (int r = 0; r < kRecsPerWarp; ++r)
            vis_word[r] = (use_vis && r_count[r] != 0u) ? __ldcg(p.meshlet_visibility + r_vo[r]) : 0xFFFFFFFFu;
        uint4 ma[kRecsPerWarp], mb[kRecsPerWarp];
        bool active[kRecsPerWarp], need[kRecsPerWarp];
#pragma unroll
        for (int r = 0; r < kRecsPerWarp; ++r) {
            active[r] = lane < r_count[r];
            // pass 1: a lane whose bit is clear is invisible and never drawn regardless of its meshlet
            need[r] = active[r] && !(pass == 1u && ((vis_word[r] >> lane) & 1u) == 0u);
            ma[r] = make_uint4(0, 0, 0, 0); mb[r] = make_uint4(0, 0, 0, 0);
            if (need[r]) {
                const uint4* m = p.meshlets + 2u * ((size_t)r_offset[r] + lane);
                ma[r] = __ldg(m); mb[r] = __ldg(m + 1);
            }
        }

        uint32_t warp_count = 0u;
        uint32_t prev_entity = 0xFFFFFFFFu;
        ModelView mv;
#pragma unroll
        for (int r = 0; r < kRecsPerWarp; ++r) {
            if (r_count[r] == 0u) continue;  // warp-uniform
            const uint32_t any_need = __ballot_sync(0xFFFFFFFFu, need[r]);
            if (any_need != 0u && r_entity[r] != prev_entity) {
                // view * model, one element per lane (col = lane/4, row = lane%4), broadcast through smem
                __syncwarp();
                if (lane < 16u) {
                    const float4 b = __ldg(p.entities + (size_t)r_entity[r] * 8u + (lane >> 2));
                    const uint32_t row = lane & 3u;
                    s_mv[warp][lane] = add(add(add(mul(s_view[0 + row], b.x), mul(s_view[4 + row], b.y)),
                                               mul(s_view[8 + row], b.z)), mul(s_view[12 + row], b.w));
                }
                __syncwarp();
                const float4* sm = reinterpret_cast<const float4*>(s_mv[warp]);
                float4 c0 = sm[0], c1 = sm[1], c2 = sm[2], c3 = sm[3];
                mv.m[0] = c0.x; mv.m[1] = c0.y; mv.m[2] = c0.z; mv.m[3] = c0.w;
                mv.m[4] = c1.x; mv.m[5] = c1.y; mv.m[6] = c1.z; mv.m[7] = c1.w;
                mv.m[8] = c2.x; mv.m[9] = c2.y; mv.m[10] = c2.z; mv.m[11] = c2.w;
                mv.m[12] = c3.x; mv.m[13] = c3.y; mv.m[14] = c3.z; mv.m[15] = c3.w;
                mv.scale = largest_scale(mv.m);
                prev_entity = r_entity[r];
            }
            bool visible = false, should_draw = false;
            const bool vib = ((vis_word[r] >> lane) & 1u) != 0u;  // vis_word = all ones when !use_vis
            if (need[r]) {
                Sphere s = transform_sphere(mv, __uint_as_float(ma[r].x), __uint_as_float(ma[r].y),
                                            __uint_as_float(ma[r].z), __uint_as_float(ma[r].w));
                visible = (pass == 1u) ? vib : true;  // need[] already implies vib in pass 1
                if (visible) visible = frustum_test(ci, s);
                if (visible) {
                    const uint32_t cone = mb[r].x;
                    const float kx = mul((float)(int)(int8_t)(cone & 0xFFu), K);
                    const float ky = mul((float)(int)(int8_t)((cone >> 8) & 0xFFu), K);
                    const float kz = mul((float)(int)(int8_t)((cone >> 16) & 0xFFu), K);
                    const float cutoff = mul((float)(int)(int8_t)(cone >> 24), K);
                    const float axx = mat_row(mv.m, 0, kx, ky, kz, 0.0f);
                    const float axy = mat_row(mv.m, 1, kx, ky, kz, 0.0f);
                    const float axz = mat_row(mv.m, 2, kx, ky, kz, 0.0f);
                    if (ci.projection_type == 0u) {
                        const float lhs = dot3(s.x, s.y, s.z, axx, axy, axz);
                        const float len = fsqrt(dot3(s.x, s.y, s.z, s.x, s.y, s.z));
                        visible = !(lhs >= fma_(cutoff, len, s.r));
                    } else if (ci.projection_type == 1u) {
                        const float camx = sub(s.x, 0.0f), camy = sub(s.y, 0.0f), camz = sub(s.z, -1.0f);
                        const float qx = sub(s.x, camx), qy = sub(s.y, camy), qz = sub(s.z, camz);
                        const float lhs = dot3(qx, qy, qz, axx, axy, axz);
                        const float len = fsqrt(dot3(qx, qy, qz, qx, qy, qz));
                        visible = !(lhs >= fma_(cutoff, len, s.r));
                    }
                }
                if (pass2 && visible) visible = occlusion_test(ci, s, p.hiz);
                const uint32_t material_index = mb[r].w & 0xFFFFu;
                const uint32_t alpha = __ldg(reinterpret_cast<const uint32_t*>(
                    p.materials + (size_t)material_index * ORBIT_MATERIAL_STRIDE_BYTES + ORBIT_MATERIAL_ALPHA_MODE_OFFSET));
                should_draw = visible && (shl1(alpha) & ci.alpha_mode_flags) != 0u;
                if (pass2 && (shl1(alpha) & ci.noskip_alpha_mode) == 0u) should_draw = visible && !vib;
            }
            const uint32_t draw_mask = __ballot_sync(0xFFFFFFFFu, should_draw);
            if (pass2) {
                const uint32_t vis_mask = __ballot_sync(0xFFFFFFFFu, visible);
                if (lane == 0u) p.meshlet_visibility[r_vo[r]] = vis_mask;
            }
            if (should_draw) {
                const uint32_t rank = warp_count + __popc(draw_mask & ((1u << lane) - 1u));
                uint32_t* c = &s_stage[warp][rank * 7u];
                const uint32_t data_offset = mb[r].z;
                const uint32_t packed = mb[r].w;
                c[0] = (packed >> 24) * 3u;                                 // triangle_count * 3
                c[1] = 1u;
                c[2] = (data_offset + ((packed >> 16) & 0xFFu)) * 4u;       // (data_offset + vertex_count) * 4
                c[3] = data_offset;
                c[4] = r_entity[r];
                c[5] = mb[r].y;                                             // meshlet vertex_offset
                c[6] = r_offset[r] + lane;
            }
            if (p.task_payloads != nullptr) {
                // MeshTaskPayload + emitted task count for this record, indices ascending by lane
                uint32_t* tp = p.task_payloads + (size_t)(rec0 + r) * 11u;
                if (lane < 8u) {
                    uint32_t packed_idx = 0u;
                    uint32_t m = draw_mask;
                    // bytes 4*lane .. 4*lane+3 of the index array = lanes of set bits number 4*lane..4*lane+3
                    for (uint32_t k = 0; k < 4u * lane && m; ++k) m &= m - 1u;
                    for (uint32_t k = 0; k < 4u && m; ++k) { packed_idx |= (uint32_t)(__ffs((int)m) - 1) << (8u * k); m &= m - 1u; }
                    tp[3u + lane] = packed_idx;
                }
                if (lane == 8u) tp[0] = __popc(draw_mask);
                if (lane == 9u) tp[1] = r_entity[r];
                if (lane == 10u) tp[2] = r_offset[r];
            }
            warp_count += __popc(draw_mask);
        }

        // ---- order the warps of the tile, then the tile among all tiles
        if (lane == 0u) s_warp_total[warp] = warp_count;
        __syncthreads();
        if (warp == 0u) {
            uint32_t v = lane < (uint32_t)kMcWarps ? s_warp_total[lane] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int d = 1; d < kMcWarps; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (uint32_t)d) incl += t;
            }
            const uint32_t tile_total = __shfl_sync(0xFFFFFFFFu, incl, kMcWarps - 1);
            const uint32_t base = lookback_exclusive(p.scan, epoch, tile, tile_total);
            if (lane < (uint32_t)kMcWarps) s_warp_total[lane] = base + incl - v;  // global exclusive offset of the warp
            if (lane == 0u) {
                s_base = base;
                if (tile == ntiles - 1u) {
                    p.draw_words[0] = base + tile_total;  // exact count even when it exceeds capacity
                    if ((uint64_t)base + tile_total > p.capacity_draws) *p.overflow_flag = 1u;
                }
            }
        }
        __syncthreads();
        // ---- stream the staged commands out: contiguous words, 128 B per warp instruction
        {
            const uint64_t first = s_warp_total[warp];
            uint64_t n = warp_count;
            if (first >= p.capacity_draws) n = 0; else if (first + n > p.capacity_draws) n = p.capacity_draws - first;
            uint32_t* dst = p.draw_words + 1u + first * 7u;
            const uint32_t nwords = (uint32_t)n * 7u;
            for (uint32_t j = lane; j < nwords; j += 32u) dst[j] = s_stage[warp][j];
        }
    }
    if (tid == 0) scan_cta_exit(p.scan, epoch);
}


cudaError_t launch_meshlet_cull(const MeshletCullParams& p, int recs_per_warp, int grid, cudaStream_t stream) {
    switch (recs_per_warp) {
        case 1: meshlet_cull_kernel<1><<<grid, kMcThreads, 0, stream>>>(p); break;
        case 2: meshlet_cull_kernel<2><<<grid, kMcThreads, 0, stream>>>(p); break;
        default: meshlet_cull_kernel<4><<<grid, kMcThreads, 0, stream>>>(p); break;
    }
    return cudaGetLastError();
}

int meshlet_cull_max_ctas_per_sm(int recs_per_warp) {
    int n = 0;
    switch (recs_per_warp) {
        case 1: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, meshlet_cull_kernel<1>, kMcThreads, 0); break;
        case 2: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, meshlet_cull_kernel<2>, kMcThreads, 0); break;
        default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, meshlet_cull_kernel<4>, kMcThreads, 0); break;
    }
    return n;
}

}  // namespace orbit
