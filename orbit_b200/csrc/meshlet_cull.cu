// meshlet_cull.cu — per-meshlet frustum / normal-cone / two-pass Hi-Z culling + ordered compaction (sm_100a).
//
// Stands in for shaders/meshlet_cull.comp:108-255 driven by create_meshlet_draw_commands
// (src/passes/draw_gen.rs:382-435) and, for the optional task payload output, the task-shader twins
// (shaders/forward/forward_depth_prepass.task:224-256).
//
// B200 design (not the reference's one-workgroup-per-record + atomicAdd). Measured on C2 the stage is bound by
// instruction issue before HBM (IEEE-exact division / square root of the projection math), so the design
// minimises executed warp-instructions and never blocks a CTA:
//   * one co-resident grid; the record count is read on the device (the reference's dispatch_indirect, no host
//     round trip) and split STATICALLY into one contiguous range per CTA and per warp, so ordered compaction
//     needs no serial prefix chain: phase 1 tests every lane and counts, each CTA publishes ONE aggregate, sums
//     the aggregates of all lower CTAs itself (a flat gather, not a hop-by-hop look-back — measured: look-back
//     over thousands of small tiles propagates only 32 tiles per L2 round trip and dominated the runtime), and
//     phase 2 emits the commands of its range at the now-known offset;
//   * lane = meshlet of a record, exactly the reference's 32-lane group, so ballots are the reference's
//     visibility words; all meshlet loads of a tile are issued up front as 2 x 128-bit non-coherent loads per
//     lane (1 KB contiguous per record, R KB in flight per warp);
//   * view*model is computed once per record by 16 lanes (two records per step) and broadcast through shared
//     memory — the reference recomputes the 4x4 product in every lane;
//   * pass 1: only lanes whose visibility bit is set can be visible, so they are PACKED across the tile's
//     records before any meshlet is loaded or tested (the early pass touches only last frame's survivors);
//   * pass 2: lanes surviving frustum + cone are PACKED into a per-warp queue and the expensive Hi-Z projection
//     runs on full warps of survivors instead of once per record at ~17% lane occupancy;
//   * survivors are ranked by ballot + popc; phase 1 leaves one draw mask per record in an L2-resident scratch
//     array, phase 2 re-reads the 16 B it needs of each surviving meshlet (L2 hits) and stores the command:
//     draw order = (record index, lane), independent of scheduling.
#include "params.cuh"

namespace orbit {

constexpr int kMcWarps = 8;
constexpr int kMcThreads = kMcWarps * 32;
constexpr int kMvStride = 20;   // 16 matrix entries + scale, padded

struct ItemTest {
    Sphere s;
    bool pre_visible;   // passed frustum + cone
};

// frustum + cone for one meshlet against the model-view matrix `mv` (17 floats in shared memory)
__device__ __forceinline__ ItemTest test_item(const OrbitCullInfo& ci, const float* __restrict__ mv, const uint4 ma, const uint32_t cone) {
    ItemTest out;
    const float4 c0 = *reinterpret_cast<const float4*>(mv + 0);
    const float4 c1 = *reinterpret_cast<const float4*>(mv + 4);
    const float4 c2 = *reinterpret_cast<const float4*>(mv + 8);
    const float4 c3 = *reinterpret_cast<const float4*>(mv + 12);
    const float scale = mv[16];
    const float cx = __uint_as_float(ma.x), cy = __uint_as_float(ma.y), cz = __uint_as_float(ma.z);
    // (M * (c,1))[row]; m3*1.0f == m3 exactly
    float px = add(add(add(mul(c0.x, cx), mul(c1.x, cy)), mul(c2.x, cz)), c3.x);
    float py = add(add(add(mul(c0.y, cx), mul(c1.y, cy)), mul(c2.y, cz)), c3.y);
    float pz = add(add(add(mul(c0.z, cx), mul(c1.z, cy)), mul(c2.z, cz)), c3.z);
    const float pw = add(add(add(mul(c0.w, cx), mul(c1.w, cy)), mul(c2.w, cz)), c3.w);
    if (pw != 1.0f) { px = fdiv(px, pw); py = fdiv(py, pw); pz = fdiv(pz, pw); }   // x/1 == x exactly
    Sphere& s = out.s;
    s.x = px; s.y = py; s.z = pz;
    s.r_model = __uint_as_float(ma.w); s.s = scale;
    s.r = mul(s.r_model, scale);
    const float nr = -s.r;
    bool visible = true;
    const uint32_t n = ci.cull_plane_count;
#pragma unroll
    for (uint32_t i = 0; i < ORBIT_MAX_CULL_PLANES; ++i) {
        if (i >= n) break;   // uniform
        const float d = add(dot3(ci.cull_planes[i][0], ci.cull_planes[i][1], ci.cull_planes[i][2], px, py, pz), ci.cull_planes[i][3]);
        visible = visible && (d > nr);
    }
    if (visible) {
        const float K = 0.007874015718698502f;
        const float kx = mul((float)(int)(int8_t)(cone & 0xFFu), K);
        const float ky = mul((float)(int)(int8_t)((cone >> 8) & 0xFFu), K);
        const float kz = mul((float)(int)(int8_t)((cone >> 16) & 0xFFu), K);
        const float cutoff = mul((float)((int)cone >> 24), K);
        // (M * (k,0)).xyz: the w column contributes m3*0.0f (kept: +-0 / NaN propagate as in the oracle)
        const float axx = add(add(add(mul(c0.x, kx), mul(c1.x, ky)), mul(c2.x, kz)), mul(c3.x, 0.0f));
        const float axy = add(add(add(mul(c0.y, kx), mul(c1.y, ky)), mul(c2.y, kz)), mul(c3.y, 0.0f));
        const float axz = add(add(add(mul(c0.z, kx), mul(c1.z, ky)), mul(c2.z, kz)), mul(c3.z, 0.0f));
        if (ci.projection_type == 0u) {
            const float lhs = dot3(px, py, pz, axx, axy, axz);
            const float len = fsqrt(dot3(px, py, pz, px, py, pz));
            visible = !(lhs >= fma_(cutoff, len, s.r));
        } else if (ci.projection_type == 1u) {
            const float camx = sub(px, 0.0f), camy = sub(py, 0.0f), camz = sub(pz, -1.0f);
            const float qx = sub(px, camx), qy = sub(py, camy), qz = sub(pz, camz);
            const float lhs = dot3(qx, qy, qz, axx, axy, axz);
            const float len = fsqrt(dot3(qx, qy, qz, qx, qy, qz));
            visible = !(lhs >= fma_(cutoff, len, s.r));
        }
    }
    out.pre_visible = visible;
    return out;
}

__device__ __forceinline__ bool draw_rule(const OrbitCullInfo& ci, const MeshletCullParams& p, bool visible, bool vib, bool pass2,
                                          uint32_t packed) {
    if (!visible) return false;
    const uint32_t material_index = packed & 0xFFFFu;
    const uint32_t alpha = __ldg(reinterpret_cast<const uint32_t*>(
        p.materials + (size_t)material_index * ORBIT_MATERIAL_STRIDE_BYTES + ORBIT_MATERIAL_ALPHA_MODE_OFFSET));
    bool should_draw = (shl1(alpha) & ci.alpha_mode_flags) != 0u;
    if (pass2 && (shl1(alpha) & ci.noskip_alpha_mode) == 0u) should_draw = !vib;   // overrides the alpha filter (meshlet_cull.comp:210-213)
    return should_draw;
}

__device__ __forceinline__ void store_command(uint32_t* __restrict__ dst, uint32_t vertex_offset, uint32_t data_offset, uint32_t packed,
                                              uint32_t entity, uint32_t meshlet_index) {
    dst[0] = (packed >> 24) * 3u;                              // triangle_count * 3
    dst[1] = 1u;
    dst[2] = (data_offset + ((packed >> 16) & 0xFFu)) * 4u;    // (data_offset + vertex_count) * 4
    dst[3] = data_offset;
    dst[4] = entity;
    dst[5] = vertex_offset;
    dst[6] = meshlet_index;
}

template <int R>
__global__ void __launch_bounds__(kMcThreads) meshlet_cull_kernel(const __grid_constant__ MeshletCullParams p) {
    static_assert(R == 2 || R == 4 || R == 8, "records per warp tile");
    __shared__ __align__(16) float s_mv[kMcWarps][R][kMvStride];
    __shared__ uint32_t s_items[kMcWarps][R * 32];
    __shared__ float s_q[kMcWarps][6][64];
    __shared__ uint32_t s_qid[kMcWarps][64];
    __shared__ uint32_t s_mask[kMcWarps][R];
    __shared__ uint32_t s_warp_total[kMcWarps];

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const OrbitCullInfo& ci = p.cull;
    const unsigned int epoch = scan_epoch(p.scan);
    uint32_t nrec = __ldcg(p.dispatch_words);  // workgroup_count_x written by the entity stage
    if ((uint64_t)nrec > p.capacity_records) nrec = (uint32_t)p.capacity_records;
    const uint32_t pass = ci.occlusion_pass;
    const bool mocc = ci.meshlet_visibility_buffer != ORBIT_NO_BUFFER;
    const bool use_vis = (pass == 1u || pass == 2u) && mocc;
    const bool pass2 = (pass == 2u) && mocc;
    const bool packed_mode = (pass == 1u) && use_vis;
    // static contiguous partition: CTA range, then warp range (multiples of R so tiles never straddle warps)
    const uint32_t tiles_total = (nrec + R - 1) / R;
    const uint32_t cta_t0 = (uint32_t)(((uint64_t)tiles_total * blockIdx.x) / gridDim.x);
    const uint32_t cta_t1 = (uint32_t)(((uint64_t)tiles_total * (blockIdx.x + 1u)) / gridDim.x);
    const uint32_t w_t0 = cta_t0 + (uint32_t)(((uint64_t)(cta_t1 - cta_t0) * warp) / kMcWarps);
    const uint32_t w_t1 = cta_t0 + (uint32_t)(((uint64_t)(cta_t1 - cta_t0) * (warp + 1u)) / kMcWarps);
    // row `lane&3` of the view matrix, for the 16-lane view*model product
    const uint32_t vrow_i = lane & 3u;
    const float v0 = ci.view_matrix.m[0][vrow_i], v1 = ci.view_matrix.m[1][vrow_i];
    const float v2 = ci.view_matrix.m[2][vrow_i], v3 = ci.view_matrix.m[3][vrow_i];
    float* const mv_base = &s_mv[warp][0][0];

    // =========================================== phase 1: test + count ===========================================
    uint32_t warp_count = 0u;
    uint32_t next_word = 0u;
    if (w_t0 < w_t1 && lane < 4u * R && w_t0 * R + (lane >> 2) < nrec) next_word = __ldcg(p.dispatch_words + 3u + (size_t)w_t0 * R * 4u + lane);
    for (uint32_t tile = w_t0; tile < w_t1; ++tile) {
        const uint32_t rec0 = tile * R;
        const uint32_t my_word = next_word;
        // prefetch the next tile's records
        next_word = 0u;
        if (tile + 1u < w_t1 && lane < 4u * R && rec0 + R + (lane >> 2) < nrec) next_word = __ldcg(p.dispatch_words + 3u + (size_t)(rec0 + R) * 4u + lane);
        uint32_t r_offset[R], r_count[R], r_vo[R], r_entity[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            r_entity[r] = __shfl_sync(0xFFFFFFFFu, my_word, r * 4 + 0);
            r_offset[r] = __shfl_sync(0xFFFFFFFFu, my_word, r * 4 + 1);
            r_count[r] = __shfl_sync(0xFFFFFFFFu, my_word, r * 4 + 2);   // 0 for records past the end
            r_vo[r] = __shfl_sync(0xFFFFFFFFu, my_word, r * 4 + 3);
        }
        // ---- visibility words: lane r loads the word of record r
        uint32_t vw = 0xFFFFFFFFu;
        {
            const uint32_t cnt = __shfl_sync(0xFFFFFFFFu, my_word, (lane * 4u + 2u) & 31u);
            const uint32_t vo = __shfl_sync(0xFFFFFFFFu, my_word, (lane * 4u + 3u) & 31u);
            if (use_vis && lane < (uint32_t)R && cnt != 0u) vw = __ldcg(p.meshlet_visibility + vo);
        }
        // ---- direct mode: all meshlet loads up front
        uint4 ma[R];
        uint32_t cone[R], packed[R];
        if (!packed_mode) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                ma[r] = make_uint4(0, 0, 0, 0); cone[r] = 0u; packed[r] = 0u;
                if (lane < r_count[r]) {
                    const uint4* m = p.meshlets + 2u * ((size_t)r_offset[r] + lane);
                    ma[r] = __ldg(m);
                    const uint4 b = __ldg(m + 1);
                    cone[r] = b.x; packed[r] = b.w;
                }
            }
        }
        // ---- view * model for every record of the tile: 16 lanes per record, two records per step
        __syncwarp();
#pragma unroll
        for (int st = 0; st < R / 2; ++st) {
            const uint32_t half = lane >> 4, e = lane & 15u;
            const uint32_t ent = half ? r_entity[2 * st + 1] : r_entity[2 * st];
            const uint32_t cnt = half ? r_count[2 * st + 1] : r_count[2 * st];
            if (cnt != 0u) {
                const float4 b = __ldg(p.entities + (size_t)ent * 8u + (e >> 2));
                mv_base[(2 * st + half) * kMvStride + e] = add(add(add(mul(v0, b.x), mul(v1, b.y)), mul(v2, b.z)), mul(v3, b.w));
            }
        }
        __syncwarp();
        if (lane < (uint32_t)R) mv_base[lane * kMvStride + 16] = largest_scale(mv_base + lane * kMvStride);
        uint32_t vis_word[R];
#pragma unroll
        for (int r = 0; r < R; ++r) vis_word[r] = __shfl_sync(0xFFFFFFFFu, vw, r);
        __syncwarp();

        uint32_t my_draw_mask = 0u;   // lane r keeps the draw mask of record r

        if (!packed_mode) {
            // ------------------------------- direct mode (pass 0 / pass 2) -------------------------------
            uint32_t vis_mask[R];
            uint32_t qn = 0u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                vis_mask[r] = 0u;
                if (r_count[r] == 0u) continue;   // warp-uniform
                bool pre = false;
                ItemTest t;
                if (lane < r_count[r]) {
                    t = test_item(ci, mv_base + r * kMvStride, ma[r], cone[r]);
                    pre = t.pre_visible;
                }
                const uint32_t pre_mask = __ballot_sync(0xFFFFFFFFu, pre);
                vis_mask[r] = pre_mask;
                if (pass2) {
                    // queue the survivors for the Hi-Z test; run it whenever a full warp of them is waiting
                    if (lane == 0u) s_mask[warp][r] = pre_mask;
                    if (pre) {
                        const uint32_t slot = qn + __popc(pre_mask & lt);
                        s_q[warp][0][slot] = t.s.x; s_q[warp][1][slot] = t.s.y; s_q[warp][2][slot] = t.s.z;
                        s_q[warp][3][slot] = t.s.r; s_q[warp][4][slot] = t.s.r_model; s_q[warp][5][slot] = t.s.s;
                        s_qid[warp][slot] = ((uint32_t)r << 5) | lane;
                    }
                    qn += __popc(pre_mask);
                    __syncwarp();
                    if (qn >= 32u) {
                        Sphere s;
                        s.x = s_q[warp][0][lane]; s.y = s_q[warp][1][lane]; s.z = s_q[warp][2][lane];
                        s.r = s_q[warp][3][lane]; s.r_model = s_q[warp][4][lane]; s.s = s_q[warp][5][lane];
                        const uint32_t id = s_qid[warp][lane];
                        if (!occlusion_test(ci, s, p.hiz)) atomicAnd(&s_mask[warp][id >> 5], ~(1u << (id & 31u)));
                        const uint32_t rem = qn - 32u;
                        float tx = 0, ty = 0, tz = 0, tr = 0, tm = 0, ts = 0; uint32_t tid2 = 0;
                        if (lane < rem) {
                            tx = s_q[warp][0][32 + lane]; ty = s_q[warp][1][32 + lane]; tz = s_q[warp][2][32 + lane];
                            tr = s_q[warp][3][32 + lane]; tm = s_q[warp][4][32 + lane]; ts = s_q[warp][5][32 + lane];
                            tid2 = s_qid[warp][32 + lane];
                        }
                        __syncwarp();
                        if (lane < rem) {
                            s_q[warp][0][lane] = tx; s_q[warp][1][lane] = ty; s_q[warp][2][lane] = tz;
                            s_q[warp][3][lane] = tr; s_q[warp][4][lane] = tm; s_q[warp][5][lane] = ts;
                            s_qid[warp][lane] = tid2;
                        }
                        qn = rem;
                        __syncwarp();
                    }
                }
            }
            if (pass2) {
                if (lane < qn) {
                    Sphere s;
                    s.x = s_q[warp][0][lane]; s.y = s_q[warp][1][lane]; s.z = s_q[warp][2][lane];
                    s.r = s_q[warp][3][lane]; s.r_model = s_q[warp][4][lane]; s.s = s_q[warp][5][lane];
                    const uint32_t id = s_qid[warp][lane];
                    if (!occlusion_test(ci, s, p.hiz)) atomicAnd(&s_mask[warp][id >> 5], ~(1u << (id & 31u)));
                }
                __syncwarp();
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r_count[r] == 0u) continue;
                    vis_mask[r] = s_mask[warp][r];
                    if (lane == 0u) p.meshlet_visibility[r_vo[r]] = vis_mask[r];   // ballot(visible), meshlet_cull.comp:235-242
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool visible = ((vis_mask[r] >> lane) & 1u) != 0u;
                const bool vib = ((vis_word[r] >> lane) & 1u) != 0u;   // all ones when visibility is not in use
                const bool sd = draw_rule(ci, p, visible, vib, pass2, packed[r]);
                const uint32_t dm = __ballot_sync(0xFFFFFFFFu, sd);
                if (lane == (uint32_t)r) my_draw_mask = dm;
                warp_count += __popc(dm);
            }
        } else {
            // ------------------------------- packed mode (pass 1) -------------------------------
            // Only lanes whose visibility bit is set can be visible (visible = visible_in_buffer,
            // meshlet_cull.comp:137): pack them across the tile's records, then load + test only those.
            uint32_t n_items = 0u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t act = r_count[r] >= 32u ? 0xFFFFFFFFu : ((1u << r_count[r]) - 1u);
                const uint32_t m = vis_word[r] & act;
                if ((m >> lane) & 1u) s_items[warp][n_items + __popc(m & lt)] = ((uint32_t)r << 5) | lane;
                n_items += __popc(m);
            }
            if (lane < (uint32_t)R) s_mask[warp][lane] = 0u;
            __syncwarp();
            for (uint32_t k = 0; k * 32u < n_items; ++k) {
                const uint32_t i = k * 32u + lane;
                const uint32_t id = i < n_items ? s_items[warp][i] : 0u;
                const uint32_t moff = __shfl_sync(0xFFFFFFFFu, my_word, (id >> 5) * 4u + 1u);
                if (i < n_items) {
                    const uint32_t r = id >> 5, j = id & 31u;
                    const uint4* m = p.meshlets + 2u * ((size_t)moff + j);
                    const uint4 a = __ldg(m), b = __ldg(m + 1);
                    const ItemTest t = test_item(ci, mv_base + r * kMvStride, a, b.x);
                    if (draw_rule(ci, p, t.pre_visible, true, false, b.w)) atomicOr(&s_mask[warp][r], 1u << j);
                }
            }
            __syncwarp();
            if (lane < (uint32_t)R) my_draw_mask = s_mask[warp][lane];
            warp_count += __reduce_add_sync(0xFFFFFFFFu, __popc(my_draw_mask));
            __syncwarp();
        }
        // one draw mask per record, kept L2-resident for phase 2
        if (lane < (uint32_t)R && rec0 + lane < nrec) p.draw_masks[rec0 + lane] = my_draw_mask;
    }

    // =========================================== order the ranges ===========================================
    if (lane == 0u) s_warp_total[warp] = warp_count;
    __syncthreads();
    if (warp == 0u) {
        const uint32_t v = lane < (uint32_t)kMcWarps ? s_warp_total[lane] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < kMcWarps; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        const uint32_t cta_total = __shfl_sync(0xFFFFFFFFu, incl, kMcWarps - 1);
        if (lane == 0u) publish(p.scan.status + blockIdx.x, pack_status(epoch, kFlagAggregate, cta_total));
        // flat gather of every lower CTA's aggregate (all CTAs are co-resident and finish phase 1 together)
        uint32_t sum = 0u;
        for (uint32_t i = lane; i < blockIdx.x; i += 32u) {
            while (true) {
                const unsigned long long w = peek(p.scan.status + i);
                if ((unsigned int)(w >> 34) == epoch) { sum += (unsigned int)w; break; }
                __nanosleep(64);
            }
        }
        const uint32_t base = __reduce_add_sync(0xFFFFFFFFu, sum);
        if (lane < (uint32_t)kMcWarps) s_warp_total[lane] = base + incl - v;   // global exclusive offset of each warp
        if (blockIdx.x == gridDim.x - 1u && lane == 0u) {
            p.draw_words[0] = base + cta_total;   // exact count even when it exceeds capacity
            if ((uint64_t)base + cta_total > p.capacity_draws) *p.overflow_flag = 1u;
        }
    }
    __syncthreads();

    // =========================================== phase 2: emit ===========================================
    uint32_t off = s_warp_total[warp];
    const bool want_payload = p.task_payloads != nullptr;
    if (warp_count != 0u || want_payload) {
        for (uint32_t tile = w_t0; tile < w_t1; ++tile) {
            const uint32_t rec0 = tile * R;
            uint32_t my_word = 0u;
            if (lane < 4u * R && rec0 + (lane >> 2) < nrec) my_word = __ldcg(p.dispatch_words + 3u + (size_t)rec0 * 4u + lane);
            uint32_t dmask = 0u;
            if (lane < (uint32_t)R && rec0 + lane < nrec) dmask = __ldcg(p.draw_masks + rec0 + lane);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t dm = __shfl_sync(0xFFFFFFFFu, dmask, r);
                const uint32_t entity = __shfl_sync(0xFFFFFFFFu, my_word, r * 4 + 0);
                const uint32_t moff = __shfl_sync(0xFFFFFFFFu, my_word, r * 4 + 1);
                if ((dm >> lane) & 1u) {
                    const uint4 b = __ldg(p.meshlets + 2u * ((size_t)moff + lane) + 1);
                    const uint64_t idx = (uint64_t)off + __popc(dm & lt);
                    if (idx < p.capacity_draws) store_command(p.draw_words + 1u + idx * 7u, b.y, b.z, b.w, entity, moff + lane);
                }
                off += __popc(dm);
                if (want_payload && rec0 + r < nrec) {
                    // MeshTaskPayload + emitted task count, indices ascending by lane (the task shader's atomicAdd
                    // order is arbitrary): lane q packs index bytes 4q..4q+3
                    uint32_t* tp = p.task_payloads + (size_t)(rec0 + r) * 11u;
                    if (lane < 8u) {
                        uint32_t packed_idx = 0u, m = dm;
                        for (uint32_t k = 0; k < 4u * lane && m; ++k) m &= m - 1u;
                        for (uint32_t k = 0; k < 4u && m; ++k) { packed_idx |= (uint32_t)(__ffs((int)m) - 1) << (8u * k); m &= m - 1u; }
                        tp[3u + lane] = packed_idx;
                    }
                    if (lane == 8u) tp[0] = __popc(dm);
                    if (lane == 9u) tp[1] = entity;
                    if (lane == 10u) tp[2] = moff;
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) scan_cta_exit(p.scan, epoch);
}

cudaError_t launch_meshlet_cull(const MeshletCullParams& p, int recs_per_warp, int grid, cudaStream_t stream) {
    switch (recs_per_warp) {
        case 2: meshlet_cull_kernel<2><<<grid, kMcThreads, 0, stream>>>(p); break;
        case 8: meshlet_cull_kernel<8><<<grid, kMcThreads, 0, stream>>>(p); break;
        default: meshlet_cull_kernel<4><<<grid, kMcThreads, 0, stream>>>(p); break;
    }
    return cudaGetLastError();
}

int meshlet_cull_max_ctas_per_sm(int recs_per_warp) {
    int n = 0;
    switch (recs_per_warp) {
        case 2: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, meshlet_cull_kernel<2>, kMcThreads, 0); break;
        case 8: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, meshlet_cull_kernel<8>, kMcThreads, 0); break;
        default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, meshlet_cull_kernel<4>, kMcThreads, 0); break;
    }
    return n;
}

int meshlet_cull_tile_records(int recs_per_warp) { return (recs_per_warp == 2 || recs_per_warp == 8) ? recs_per_warp : 4; }

}  // namespace orbit
