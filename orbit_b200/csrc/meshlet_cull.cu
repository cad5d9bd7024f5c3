// meshlet_cull.cu — per-meshlet frustum / normal-cone / two-pass Hi-Z culling + ordered compaction (sm_100a).
//
// Stands in for shaders/meshlet_cull.comp:108-255 driven by create_meshlet_draw_commands
// (src/passes/draw_gen.rs:382-435) and, for the optional task payload output, the task-shader twins
// (shaders/forward/forward_depth_prepass.task:224-256).
//
// B200 design (not the reference's one-workgroup-per-record + atomicAdd). Measured on C2 the stage is bound by
// instruction issue before HBM (IEEE-exact division / square root of the projection math), so the design
// minimises executed warp-instructions and never blocks a CTA:
//   * TWO launches. `meshlet_test_kernel` is embarrassingly parallel: persistent warps take tiles of R records
//     CYCLICALLY (tile = global warp id + k * total warps), test every lane and leave one draw mask per record in
//     an L2-resident scratch array — no inter-CTA ordering at all, so hot regions of the scene are spread over
//     all SMs (a contiguous static split measured 2-3x slower: the CTAs owning the visible part of the city made
//     everyone wait). `meshlet_emit_kernel` then orders and writes the survivors: contiguous record ranges per
//     CTA, popcount sums, ONE published aggregate per CTA, a flat gather of all lower aggregates (no hop-by-hop
//     look-back: that propagates only 32 tiles per L2 round trip and dominated v1/v2), and the commands are
//     rebuilt from 16 B of each surviving meshlet (L2 hits). The record count is read on the device (the
//     reference's dispatch_indirect) — no host round trip;
//   * lane = meshlet of a record, exactly the reference's 32-lane group, so ballots are the reference's
//     visibility words; the meshlets of a tile (R records x 1 KB contiguous) are staged into shared memory by
//     TMA bulk copies (cp.async.bulk + mbarrier, one copy per record issued by R lanes), so R KB per warp are in
//     flight without holding registers, and the next tile's copy is issued as soon as the current one is consumed;
//   * view*model is computed once per record by 16 lanes (two records per step) and broadcast through shared
//     memory — the reference recomputes the 4x4 product in every lane;
//   * pass 1: only lanes whose visibility bit is set can be visible, so they are PACKED across the tile's
//     records before any meshlet is loaded or tested (the early pass touches only last frame's survivors);
//   * pass 2: lanes surviving frustum + cone are PACKED into a per-warp queue and the expensive Hi-Z projection
//     runs on full warps of survivors instead of once per record at ~17% lane occupancy;
//   * survivors are ranked by ballot + popc; draw order = (record index, lane), independent of scheduling.
#include "params.cuh"

namespace orbit {

constexpr int kMcWarps = 8;
constexpr int kMcThreads = kMcWarps * 32;
constexpr int kMvStride = 20;   // 16 matrix entries + scale, padded
#ifndef ORBIT_DIRECT_MIN_CTAS
#define ORBIT_DIRECT_MIN_CTAS 3
#endif
constexpr int kDirectMinCtas = ORBIT_DIRECT_MIN_CTAS;

// ---- TMA bulk copy + mbarrier plumbing (PTX; SASS: UBLKCP / SYNCS) --------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct ItemTest {
    Sphere s;
    bool pre_visible;   // passed frustum + cone
};

// frustum + cone for one meshlet against the model-view matrix `mv` (17 floats in shared memory)
// kProj: 0 perspective, 1 orthographic, -1 decided at run time from ci.projection_type
template <int kProj>
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
    if (n == 5u) {
        // the main-view case (forward.rs:268 passes planes[0..5]): straight-line code, plane coefficients as constant-bank
        // operands, no loop control (the generic loop below costs ~25 more instructions per record)
#pragma unroll
        for (uint32_t i = 0; i < 5u; ++i) {
            const float d = add(dot3(ci.cull_planes[i][0], ci.cull_planes[i][1], ci.cull_planes[i][2], px, py, pz), ci.cull_planes[i][3]);
            visible = visible && (d > nr);
        }
    } else {
#pragma unroll 1
        for (uint32_t i = 0; i < n; ++i) {
            const float d = add(dot3(ci.cull_planes[i][0], ci.cull_planes[i][1], ci.cull_planes[i][2], px, py, pz), ci.cull_planes[i][3]);
            visible = visible && (d > nr);
        }
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
        const uint32_t proj = kProj >= 0 ? (uint32_t)kProj : ci.projection_type;
        if (proj == 0u) {
            const float lhs = dot3(px, py, pz, axx, axy, axz);
            const float len = fsqrt(dot3(px, py, pz, px, py, pz));
            visible = !(lhs >= fma_(cutoff, len, s.r));
        } else if (proj == 1u) {
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
struct __align__(128) WarpSmem {
    uint4 meshlets[R * 64];        // R records x 32 meshlets x 2 x 16 B, filled by TMA (packed mode: item list aliases this)
    float mv[R][kMvStride];        // view*model + scale per record
    float q[6][64];                // ring buffer of Hi-Z candidates: x, y, z, r, r_model, scale
    uint32_t qid[64];              //   (record << 5) | lane
    uint32_t mask[R];              // per-record visible masks under construction
    uint32_t aok[R], nsk[R];       // per-record "alpha mode passes the filter" / "alpha mode is noskip" lane masks
    unsigned long long bar;        // mbarrier of the TMA copies
};

// Survivor counts are accumulated per CHUNK of consecutive records by the test kernel (integer atomics: the sums are
// order-independent) so that the emit kernel can order its output with a shared-memory scan instead of an
// inter-CTA exchange. Chunk size: a multiple of 32 records such that there are at most kMaxChunks chunks.
constexpr uint32_t kMaxChunks = 2048u;
__device__ __forceinline__ uint32_t chunk_records(uint32_t nrec) {
    const uint32_t per = (nrec + kMaxChunks - 1u) / kMaxChunks;
    return max(32u, (per + 31u) & ~31u);
}

// view * model (+ largest column scale) for every record of a tile: 16 lanes per record, two records per step
template <int R>
__device__ __forceinline__ void tile_model_view(const MeshletCullParams& p, float* mv_base, uint32_t my_word, uint32_t lane,
                                                float v0, float v1, float v2, float v3) {
#pragma unroll
    for (int st = 0; st < R / 2; ++st) {
        const uint32_t half = lane >> 4, e = lane & 15u;
        const uint32_t ent = __shfl_sync(0xFFFFFFFFu, my_word, (2 * st + half) * 4 + 0);
        const uint32_t cnt = __shfl_sync(0xFFFFFFFFu, my_word, (2 * st + half) * 4 + 2);
        if (cnt != 0u) {
            const float4 b = __ldg(p.entities + (size_t)ent * 8u + (e >> 2));
            mv_base[(2 * st + half) * kMvStride + e] = add(add(add(mul(v0, b.x), mul(v1, b.y)), mul(v2, b.z)), mul(v3, b.w));
        }
    }
    __syncwarp();
    if (lane < (uint32_t)R) mv_base[lane * kMvStride + 16] = largest_scale(mv_base + lane * kMvStride);
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------
// Direct mode: pass 0, pass 2, and pass 1 without a meshlet visibility buffer — every lane of every record is
// tested. kPass2 = (occlusion_pass == 2 && meshlet occlusion culling on); kProj as in test_item.
template <int R, bool kPass2, int kProj>
__global__ void __launch_bounds__(kMcThreads, kDirectMinCtas) meshlet_test_direct_kernel(const __grid_constant__ MeshletCullParams p) {
    static_assert(R == 2 || R == 4 || R == 8, "records per warp tile");
    extern __shared__ __align__(128) unsigned char s_raw[];
    WarpSmem<R>* const all = reinterpret_cast<WarpSmem<R>*>(s_raw);

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    WarpSmem<R>& ws = all[warp];
    const OrbitCullInfo& ci = p.cull;
    pdl_wait();
    // The record words of this warp's first two tiles are requested BEFORE the record count is known (bounded by the
    // dispatch buffer's capacity, masked by the count afterwards): one dependent round trip less in the prologue.
    uint32_t spec_word0 = 0u, spec_word1 = 0u;
    {
        const uint32_t t0 = blockIdx.x * kMcWarps + warp, t1 = t0 + gridDim.x * kMcWarps;
        const uint64_t ra = (uint64_t)t0 * R + (lane >> 2), rb = (uint64_t)t1 * R + (lane >> 2);
        if (lane < 4u * R && ra < p.capacity_records) spec_word0 = __ldcg(p.dispatch_words + 3u + (size_t)t0 * R * 4u + lane);
        if (lane < 4u * R && rb < p.capacity_records) spec_word1 = __ldcg(p.dispatch_words + 3u + (size_t)t1 * R * 4u + lane);
    }
    uint32_t nrec = __ldcg(p.dispatch_words);  // workgroup_count_x written by the entity stage
    if ((uint64_t)nrec > p.capacity_records) nrec = (uint32_t)p.capacity_records;
    const uint32_t chunk_rec = chunk_records(nrec);
    // Scratch is double-buffered by a parity that lives in device memory (CUDA-graph replays must see fresh state).
    // Word A is read by test kernels and written by emit kernels; word B the other way round: a kernel never writes
    // a word that CTAs of the same launch read, so the stream order of the launches is the only synchronisation.
    const uint32_t half = __ldcg(p.chunk_parity) & 1u;
    if (blockIdx.x == 0 && threadIdx.x == 0) p.chunk_parity[1] = half;
    uint32_t* const chunk_counts = p.chunk_counts + half * kMaxChunks;
    uint32_t* const draw_total = p.draw_total + half;
    // cyclic tile assignment over all warps of the grid: balances hot and cold regions of the record list
    const uint32_t tiles_total = (nrec + R - 1) / R;
    const uint32_t w_stride = gridDim.x * kMcWarps;
    const uint32_t w_t0 = blockIdx.x * kMcWarps + warp;
    const uint32_t w_t1 = tiles_total;
    // row `lane&3` of the view matrix, for the 16-lane view*model product
    const uint32_t vrow_i = lane & 3u;
    const float v0 = ci.view_matrix.m[0][vrow_i], v1 = ci.view_matrix.m[1][vrow_i];
    const float v2 = ci.view_matrix.m[2][vrow_i], v3 = ci.view_matrix.m[3][vrow_i];
    float* const mv_base = &ws.mv[0][0];
    const uint32_t bar = smem_addr(&ws.bar);
    const uint32_t buf = smem_addr(&ws.meshlets[0]);
    if (lane == 0u) { mbar_init(bar, R); mbar_fence_init(); }
    __syncwarp();
    uint32_t parity = 0u;

    // record words of a tile: lane l < 4R holds word l (entity, meshlet_offset, meshlet_count, visibility_offset per record)
    auto load_words = [&](uint32_t tile) -> uint32_t {
        uint32_t w = 0u;
        if (tile < w_t1 && lane < 4u * R && tile * R + (lane >> 2) < nrec) w = __ldcg(p.dispatch_words + 3u + (size_t)tile * R * 4u + lane);
        return w;
    };
    // lane r < R issues the bulk copy of record r (count x 32 B) and arrives on the barrier
    auto issue_tma = [&](uint32_t words) {
        const uint32_t off = __shfl_sync(0xFFFFFFFFu, words, (lane * 4u + 1u) & 31u);
        const uint32_t cnt = __shfl_sync(0xFFFFFFFFu, words, (lane * 4u + 2u) & 31u);
        if (lane < (uint32_t)R) {
            if (cnt != 0u) {
                const uint32_t bytes = min(cnt, 32u) * 32u;
                mbar_arrive_expect_tx(bar, bytes);
                tma_load_1d(buf + lane * 1024u, p.meshlets + 2u * (size_t)off, bytes, bar);
            } else {
                mbar_arrive(bar);
            }
        }
    };

    uint32_t warp_total = 0u;   // survivors found by this warp (lets the emit kernel skip everything when zero)
    uint32_t cur_word = spec_word0, next_word = spec_word1;
    {   // the two speculative loads were bounded by the buffer capacity; now apply the real record count
        const uint32_t rec_a = w_t0 * R + (lane >> 2), rec_b = (w_t0 + w_stride) * R + (lane >> 2);
        if (!(lane < 4u * R && rec_a < nrec)) cur_word = 0u;
        if (!(lane < 4u * R && rec_b < nrec)) next_word = 0u;
    }
    if (w_t0 < w_t1) issue_tma(cur_word);
    uint32_t qhead = 0u;   // ring-buffer read position (entries [qhead, qhead+qn) mod 64 are pending)
    for (uint32_t tile = w_t0; tile < w_t1; tile += w_stride) {
        const uint32_t rec0 = tile * R;
        const uint32_t my_word = cur_word;
        cur_word = next_word;
        next_word = load_words(tile + 2u * w_stride);
        // ---- lane r keeps count / visibility offset / visibility word of record r
        const uint32_t my_cnt = __shfl_sync(0xFFFFFFFFu, my_word, (lane * 4u + 2u) & 31u);
        const uint32_t my_vo = __shfl_sync(0xFFFFFFFFu, my_word, (lane * 4u + 3u) & 31u);
        uint32_t vw = 0xFFFFFFFFu;   // pass 2: last frame's visibility (decides should_draw); otherwise unused
        if (kPass2 && lane < (uint32_t)R && my_cnt != 0u) vw = __ldcg(p.meshlet_visibility + my_vo);
        tile_model_view<R>(p, mv_base, my_word, lane, v0, v1, v2, v3);

        uint32_t qn = 0u;
        mbar_wait(bar, parity);
        parity ^= 1u;
        // The record loop is deliberately NOT unrolled: four inlined copies of the test (~600 SASS instructions each)
        // overflowed the instruction cache (9% of issue stalls were "no instruction"); per-record results live in
        // shared memory as warp-uniform 32-bit masks instead of register arrays.
#pragma unroll 1
        for (uint32_t r = 0; r < (uint32_t)R; ++r) {
            const uint32_t cnt = __shfl_sync(0xFFFFFFFFu, my_word, r * 4u + 2u);   // 0 for records past the end
            if (cnt == 0u) { if (lane == 0u) { ws.mask[r] = 0u; ws.aok[r] = 0u; ws.nsk[r] = 0u; } continue; }   // warp-uniform
            bool pre = false, alpha_ok = false, noskip = false;
            ItemTest t;
            if (lane < cnt) {
                const uint4 a = ws.meshlets[r * 64u + lane * 2u];
                const uint4 b = ws.meshlets[r * 64u + lane * 2u + 1u];
                const uint32_t alpha = __ldg(reinterpret_cast<const uint32_t*>(
                    p.materials + (size_t)(b.w & 0xFFFFu) * ORBIT_MATERIAL_STRIDE_BYTES + ORBIT_MATERIAL_ALPHA_MODE_OFFSET));
                alpha_ok = (shl1(alpha) & ci.alpha_mode_flags) != 0u;
                noskip = (shl1(alpha) & ci.noskip_alpha_mode) != 0u;
                t = test_item<kProj>(ci, mv_base + r * kMvStride, a, b.x);
                pre = t.pre_visible;
            }
            const uint32_t pre_mask = __ballot_sync(0xFFFFFFFFu, pre);
            const uint32_t aok_mask = __ballot_sync(0xFFFFFFFFu, alpha_ok);
            const uint32_t nsk_mask = kPass2 ? __ballot_sync(0xFFFFFFFFu, noskip) : 0u;
            if (lane == 0u) { ws.mask[r] = pre_mask; ws.aok[r] = aok_mask; ws.nsk[r] = nsk_mask; }
            if (kPass2) {
                // queue the survivors for the Hi-Z test; run it whenever a full warp of them is waiting
                if (pre) {
                    const uint32_t slot = (qhead + qn + __popc(pre_mask & lt)) & 63u;
                    ws.q[0][slot] = t.s.x; ws.q[1][slot] = t.s.y; ws.q[2][slot] = t.s.z;
                    ws.q[3][slot] = t.s.r; ws.q[4][slot] = t.s.r_model; ws.q[5][slot] = t.s.s;
                    ws.qid[slot] = (r << 5) | lane;
                }
                qn += __popc(pre_mask);
                __syncwarp();
                if (qn >= 32u) {
                    const uint32_t slot = (qhead + lane) & 63u;
                    Sphere s;
                    s.x = ws.q[0][slot]; s.y = ws.q[1][slot]; s.z = ws.q[2][slot];
                    s.r = ws.q[3][slot]; s.r_model = ws.q[4][slot]; s.s = ws.q[5][slot];
                    const uint32_t id = ws.qid[slot];
                    if (!occlusion_test<kProj>(ci, s, p.hiz)) atomicAnd(&ws.mask[id >> 5], ~(1u << (id & 31u)));
                    qhead = (qhead + 32u) & 63u;
                    qn -= 32u;
                    __syncwarp();
                }
            }
        }
        // every lane has read its meshlets: the staging buffer is free -> start the next tile's copy now
        __syncwarp();
        if (tile + w_stride < w_t1) issue_tma(cur_word);
        if (kPass2) {
            if (lane < qn) {
                const uint32_t slot = (qhead + lane) & 63u;
                Sphere s;
                s.x = ws.q[0][slot]; s.y = ws.q[1][slot]; s.z = ws.q[2][slot];
                s.r = ws.q[3][slot]; s.r_model = ws.q[4][slot]; s.s = ws.q[5][slot];
                const uint32_t id = ws.qid[slot];
                if (!occlusion_test<kProj>(ci, s, p.hiz)) atomicAnd(&ws.mask[id >> 5], ~(1u << (id & 31u)));
            }
            qhead = (qhead + qn) & 63u;
            __syncwarp();
        }
        // lane r < R finishes record r with warp-uniform bit logic (meshlet_cull.comp:207-213):
        //   should_draw = visible && alpha passes the filter; in pass 2, unless the alpha mode is "noskip",
        //   should_draw = visible && !visible_last_frame (this overrides the alpha filter)
        uint32_t my_draw_mask = 0u;
        if (lane < (uint32_t)R) {
            const uint32_t vis = ws.mask[lane], aok = ws.aok[lane], nsk = ws.nsk[lane];
            if (kPass2) {
                if (my_cnt != 0u) p.meshlet_visibility[my_vo] = vis;   // ballot(visible), meshlet_cull.comp:235-242
                my_draw_mask = vis & ((nsk & aok) | (~nsk & ~vw));
            } else {
                my_draw_mask = vis & aok;
            }
        }
        __syncwarp();
        // one {draw mask, entity, meshlet offset} per record, kept L2-resident for the emit kernel (so it needs one
        // load per record and never touches the dispatch buffer); survivors counted per chunk
        {
            const uint32_t ent = __shfl_sync(0xFFFFFFFFu, my_word, (lane * 4u + 0u) & 31u);
            const uint32_t mof = __shfl_sync(0xFFFFFFFFu, my_word, (lane * 4u + 1u) & 31u);
            if (lane < (uint32_t)R && rec0 + lane < nrec) p.draw_masks[rec0 + lane] = make_uint4(my_draw_mask, ent, mof, 0u);
        }
        const uint32_t tile_total = __reduce_add_sync(0xFFFFFFFFu, __popc(my_draw_mask));
        if (lane == 0u && tile_total != 0u) atomicAdd(chunk_counts + rec0 / chunk_rec, tile_total);
        warp_total += tile_total;
    }
    pdl_launch_dependents();
    if (lane == 0u && warp_total != 0u) atomicAdd(draw_total, warp_total);
}

// ---------------------------------------------------------------------------------------------------------
// Packed mode: pass 1 with a meshlet visibility buffer. Only lanes whose visibility bit is set can be visible
// (visible = visible_in_buffer, meshlet_cull.comp:137), so they are packed across the tile's records and only
// those meshlets are loaded and tested. Latency-bound (little work per tile): small shared-memory footprint so
// that many warps are resident, and the visibility words and model matrices are fetched in parallel.
template <int R>
struct __align__(16) PackedSmem {
    float mv[2][R][kMvStride];     // double-buffered: the next tile's matrices are built while this tile is tested
    uint32_t items[R * 32];
    uint32_t mask[R];
};

// Loads of one tile that depend only on its record words, kept in registers until the tile's turn comes:
// model-matrix rows (16 lanes per record, two records per step) and last frame's visibility word (lane r < R).
template <int R>
struct PackedPrefetch {
    float4 b[R / 2];
    uint32_t vw;
};

template <int R>
__device__ __forceinline__ PackedPrefetch<R> packed_issue_loads(const MeshletCullParams& p, uint32_t words, uint32_t lane) {
    PackedPrefetch<R> f;
    const uint32_t my_cnt = __shfl_sync(0xFFFFFFFFu, words, (lane * 4u + 2u) & 31u);
    const uint32_t my_vo = __shfl_sync(0xFFFFFFFFu, words, (lane * 4u + 3u) & 31u);
    f.vw = 0u;
    if (lane < (uint32_t)R && my_cnt != 0u) f.vw = __ldcg(p.meshlet_visibility + my_vo) & (my_cnt >= 32u ? 0xFFFFFFFFu : ((1u << my_cnt) - 1u));
#pragma unroll
    for (int st = 0; st < R / 2; ++st) {
        const uint32_t half = lane >> 4, e = lane & 15u;
        const uint32_t ent = __shfl_sync(0xFFFFFFFFu, words, (2 * st + half) * 4 + 0);
        const uint32_t cnt = __shfl_sync(0xFFFFFFFFu, words, (2 * st + half) * 4 + 2);
        f.b[st] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cnt != 0u) f.b[st] = __ldg(p.entities + (size_t)ent * 8u + (e >> 2));
    }
    return f;
}

// view * model (+ largest column scale) of the tile's records from the prefetched rows (same arithmetic as tile_model_view)
template <int R>
__device__ __forceinline__ void packed_finish_model_view(const PackedPrefetch<R>& f, float* mv_base, uint32_t lane,
                                                         float v0, float v1, float v2, float v3) {
#pragma unroll
    for (int st = 0; st < R / 2; ++st) {
        const uint32_t half = lane >> 4, e = lane & 15u;
        const float4 b = f.b[st];
        mv_base[(2 * st + half) * kMvStride + e] = add(add(add(mul(v0, b.x), mul(v1, b.y)), mul(v2, b.z)), mul(v3, b.w));
    }
    __syncwarp();
    if (lane < (uint32_t)R) mv_base[lane * kMvStride + 16] = largest_scale(mv_base + lane * kMvStride);
    __syncwarp();
}

template <int R>
__global__ void __launch_bounds__(kMcThreads, 4) meshlet_test_packed_kernel(const __grid_constant__ MeshletCullParams p) {
    __shared__ PackedSmem<R> s_all[kMcWarps];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    PackedSmem<R>& ws = s_all[warp];
    const OrbitCullInfo& ci = p.cull;
    pdl_wait();
    const uint32_t w_stride = gridDim.x * kMcWarps;
    const uint32_t tile0 = blockIdx.x * kMcWarps + warp;
    // Latency-bound kernel (a warp sees one or two tiles, each a chain of dependent loads: record words -> visibility
    // word + model matrices -> meshlets -> material): the chain of tile t+1 is started before tile t is tested, and
    // the first two tiles' record words are requested before the record count is known (bounded by the buffer's
    // capacity, masked afterwards).
    auto spec_words = [&](uint32_t tile) -> uint32_t {
        const uint64_t rec = (uint64_t)tile * R + (lane >> 2);
        return (lane < 4u * R && rec < p.capacity_records) ? __ldcg(p.dispatch_words + 3u + (size_t)tile * R * 4u + lane) : 0u;
    };
    uint32_t cur_word = spec_words(tile0), next_word = spec_words(tile0 + w_stride);
    uint32_t nrec = __ldcg(p.dispatch_words);
    if ((uint64_t)nrec > p.capacity_records) nrec = (uint32_t)p.capacity_records;
    const uint32_t chunk_rec = chunk_records(nrec);
    const uint32_t half = __ldcg(p.chunk_parity) & 1u;   // see meshlet_test_direct_kernel
    if (blockIdx.x == 0 && threadIdx.x == 0) p.chunk_parity[1] = half;
    uint32_t* const chunk_counts = p.chunk_counts + half * kMaxChunks;
    uint32_t* const draw_total = p.draw_total + half;
    const uint32_t tiles_total = (nrec + R - 1) / R;
    if (!(lane < 4u * R && (uint64_t)tile0 * R + (lane >> 2) < nrec)) cur_word = 0u;
    if (!(lane < 4u * R && (uint64_t)(tile0 + w_stride) * R + (lane >> 2) < nrec)) next_word = 0u;
    const uint32_t vrow_i = lane & 3u;
    const float v0 = ci.view_matrix.m[0][vrow_i], v1 = ci.view_matrix.m[1][vrow_i];
    const float v2 = ci.view_matrix.m[2][vrow_i], v3 = ci.view_matrix.m[3][vrow_i];
    uint32_t warp_total = 0u, buf = 0u;
    PackedPrefetch<R> cur = packed_issue_loads<R>(p, cur_word, lane);
    // A tile none of whose records had a visible meshlet last frame (most of the list: occluded entities) needs neither
    // matrices nor packing — that arithmetic, not memory, was what the kernel spent its issue slots on.
    if (tile0 < tiles_total && __ballot_sync(0xFFFFFFFFu, cur.vw != 0u) != 0u)
        packed_finish_model_view<R>(cur, &ws.mv[0][0][0], lane, v0, v1, v2, v3);
    for (uint32_t tile = tile0; tile < tiles_total; tile += w_stride) {
        const uint32_t rec0 = tile * R;
        const uint32_t my_word = cur_word;
        float* const mv_base = &ws.mv[buf][0][0];
        const uint32_t vw = cur.vw;
        // ---- start the next tile's chain (its words arrived while the previous tile was tested)
        const bool has_next = tile + w_stride < tiles_total;
        PackedPrefetch<R> nxt = packed_issue_loads<R>(p, next_word, lane);
        uint32_t next2_word = 0u;
        {
            const uint32_t t2 = tile + 2u * w_stride;
            if (t2 < tiles_total && lane < 4u * R && t2 * R + (lane >> 2) < nrec) next2_word = __ldcg(p.dispatch_words + 3u + (size_t)t2 * R * 4u + lane);
        }
        // ---- this tile: pack the lanes whose visibility bit is set, test them
        uint32_t my_draw_mask = 0u;
        if (__ballot_sync(0xFFFFFFFFu, vw != 0u) != 0u) {      // warp-uniform
            uint32_t n_items = 0u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t m = __shfl_sync(0xFFFFFFFFu, vw, r);
                if ((m >> lane) & 1u) ws.items[n_items + __popc(m & lt)] = ((uint32_t)r << 5) | lane;
                n_items += __popc(m);
            }
            if (lane < (uint32_t)R) ws.mask[lane] = 0u;
            __syncwarp();
            for (uint32_t k = 0; k * 32u < n_items; ++k) {
                const uint32_t i = k * 32u + lane;
                const uint32_t id = i < n_items ? ws.items[i] : 0u;
                const uint32_t moff = __shfl_sync(0xFFFFFFFFu, my_word, (id >> 5) * 4u + 1u);
                if (i < n_items) {
                    const uint32_t r = id >> 5, j = id & 31u;
                    const uint4* m = p.meshlets + 2u * ((size_t)moff + j);
                    const uint4 a = __ldg(m), b = __ldg(m + 1);
                    const ItemTest t = test_item<-1>(ci, mv_base + r * kMvStride, a, b.x);
                    if (draw_rule(ci, p, t.pre_visible, true, false, b.w)) atomicOr(&ws.mask[r], 1u << j);
                }
            }
            __syncwarp();
            if (lane < (uint32_t)R) my_draw_mask = ws.mask[lane];
            __syncwarp();
        }
        {
            const uint32_t ent = __shfl_sync(0xFFFFFFFFu, my_word, (lane * 4u + 0u) & 31u);
            const uint32_t mof = __shfl_sync(0xFFFFFFFFu, my_word, (lane * 4u + 1u) & 31u);
            if (lane < (uint32_t)R && rec0 + lane < nrec) p.draw_masks[rec0 + lane] = make_uint4(my_draw_mask, ent, mof, 0u);
        }
        const uint32_t tile_total = __reduce_add_sync(0xFFFFFFFFu, __popc(my_draw_mask));
        if (lane == 0u && tile_total != 0u) atomicAdd(chunk_counts + rec0 / chunk_rec, tile_total);
        warp_total += tile_total;
        // ---- the next tile's matrices into the other buffer; rotate
        if (has_next && __ballot_sync(0xFFFFFFFFu, nxt.vw != 0u) != 0u) packed_finish_model_view<R>(nxt, &ws.mv[buf ^ 1u][0][0], lane, v0, v1, v2, v3);
        buf ^= 1u;
        cur = nxt;
        cur_word = next_word;
        next_word = next2_word;
    }
    pdl_launch_dependents();
    if (lane == 0u && warp_total != 0u) atomicAdd(draw_total, warp_total);
}

// Position of the n-th (0-based) set bit of m (n < popc(m)): branch-free binary search on popcounts of halves.
__device__ __forceinline__ uint32_t select_set_bit(uint32_t m, uint32_t n) {
    uint32_t pos = 0u;
#pragma unroll
    for (uint32_t w = 16u; w != 0u; w >>= 1) {
        const uint32_t c = __popc(m & ((1u << w) - 1u));
        const bool hi = n >= c;
        n -= hi ? c : 0u;
        pos += hi ? w : 0u;
        m = hi ? (m >> w) : m;
    }
    return pos;
}

// ---------------------------------------------------------------------------------------------------------
// Second launch: ordered, OUTPUT-BALANCED emission with no inter-CTA exchange.
//   1. every CTA loads the (at most 2048) per-chunk survivor counts the test kernel accumulated and scans them in
//      shared memory: chunk prefix P. (Earlier versions published per-CTA aggregates and gathered them — three
//      more dependent global round trips in a kernel whose whole runtime is a handful of round trips.)
//   2. the T outputs are split evenly over all warps of the grid; a warp binary-searches P for the chunk holding
//      its first output, walks that chunk's draw masks 32 records at a time, and writes its outputs. Survivors
//      cluster in the visible part of the scene, so splitting by RECORDS leaves a few CTAs with 10-16x the average
//      work; splitting by OUTPUTS gives every warp the same number.
//   3. the counts of the OTHER parity are zeroed for the next call and the parity word the next test kernel reads is
//      flipped — no done-counter, fence or atomic on the way out (that exit chain was ~1/3 of this kernel's samples).
constexpr int kEmitWarps = 8;
__global__ void __launch_bounds__(kEmitWarps * 32) meshlet_emit_kernel(const __grid_constant__ MeshletCullParams p) {
    __shared__ uint32_t s_prefix[kMaxChunks];            // inclusive survivor count up to chunk c
    __shared__ uint32_t s_warp_total[kEmitWarps];
    __shared__ uint32_t s_rec[kEmitWarps][4][32];
    __shared__ uint32_t s_payload[kEmitWarps][32 * 11];   // task-payload staging (only used when requested)
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool want_payload = p.task_payloads != nullptr;
    pdl_wait();
    // both halves of the chunk counts are requested before the parity is known: one round trip instead of two
    uint32_t v0[8], v1[8];
    {
        const uint4* c0 = reinterpret_cast<const uint4*>(p.chunk_counts) + tid * 2u;
        const uint4* c1 = reinterpret_cast<const uint4*>(p.chunk_counts + kMaxChunks) + tid * 2u;
        const uint4 a0 = __ldcg(c0), a1 = __ldcg(c0 + 1), b0 = __ldcg(c1), b1 = __ldcg(c1 + 1);
        v0[0] = a0.x; v0[1] = a0.y; v0[2] = a0.z; v0[3] = a0.w; v0[4] = a1.x; v0[5] = a1.y; v0[6] = a1.z; v0[7] = a1.w;
        v1[0] = b0.x; v1[1] = b0.y; v1[2] = b0.z; v1[3] = b0.w; v1[4] = b1.x; v1[5] = b1.y; v1[6] = b1.z; v1[7] = b1.w;
    }
    // both parities' totals in ONE 8-byte load (two 4-byte loads get the second one predicated on the parity by ptxas,
    // even as volatile asm: a dependent round trip in a kernel that is only a handful of round trips long)
    const uint2 t01 = __ldcg(reinterpret_cast<const uint2*>(p.draw_total));
    const uint32_t t0 = t01.x, t1 = t01.y;
    uint32_t nrec = __ldcg(p.dispatch_words);
    const uint32_t parity = __ldcg(p.chunk_parity + 1) & 1u;   // word B: the half the test kernel of this call used
    const uint32_t grand_total = parity ? t1 : t0;
    if ((uint64_t)nrec > p.capacity_records) nrec = (uint32_t)p.capacity_records;
    // zero the other parity's counters for the next call (it runs after this kernel in stream order)
    for (uint32_t i = blockIdx.x * blockDim.x + tid; i < kMaxChunks; i += gridDim.x * blockDim.x) p.chunk_counts[(parity ^ 1u) * kMaxChunks + i] = 0u;
    if (blockIdx.x == 0 && tid == 0) { p.draw_total[parity ^ 1u] = 0u; p.chunk_parity[0] = parity ^ 1u; }   // word A for the next call
    const uint32_t gw = blockIdx.x * kEmitWarps + warp, GW = gridDim.x * kEmitWarps;
    if (grand_total != 0u) {
        const uint32_t chunk_rec = chunk_records(nrec);
        const uint32_t nchunks = (nrec + chunk_rec - 1u) / chunk_rec;
        // ---- 1. chunk counts -> inclusive prefix in shared memory (8 consecutive chunks per thread)
        uint32_t v[8], local = 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = parity ? v1[k] : v0[k]; local += v[k]; }
        uint32_t incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        if (lane == 31u) s_warp_total[warp] = incl;
        __syncthreads();
        uint32_t run = incl - local;
#pragma unroll
        for (int w = 0; w < kEmitWarps; ++w) if ((uint32_t)w < warp) run += s_warp_total[w];
#pragma unroll
        for (int k = 0; k < 8; ++k) { run += v[k]; s_prefix[tid * 8u + (uint32_t)k] = run; }
        __syncthreads();
        const uint32_t total = s_prefix[kMaxChunks - 1u];
        if (blockIdx.x == 0 && tid == 0) {
            p.draw_words[0] = total;   // exact count even when it exceeds capacity
            if ((uint64_t)total > p.capacity_draws) *p.overflow_flag = 1u;
        }
        // ---- 2. my share of the outputs
        const uint32_t o_begin = (uint32_t)(((uint64_t)total * gw) / GW);
        const uint32_t o_end = (uint32_t)(((uint64_t)total * (gw + 1u)) / GW);
        uint32_t* const sr = &s_rec[warp][0][0];
        if (o_begin < o_end) {
            // chunk holding output o_begin: first c with P[c] > o_begin
            uint32_t lo = 0u, hi = nchunks - 1u;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_prefix[mid] > o_begin) hi = mid; else lo = mid + 1u;
            }
            uint32_t running = lo ? s_prefix[lo - 1u] : 0u;               // outputs before record `rec`
            uint32_t rec = lo * chunk_rec;
            // At a chunk boundary the prefix tells whether the chunk holds any survivor: empty chunks are stepped over
            // without touching memory (a warp whose share of the outputs straddles an invisible stretch of the record
            // list would otherwise pay one dependent load per 32 records of it), and the masks of the next group are
            // requested before the current group is processed.
            uint32_t cur_chunk = lo, in_chunk = 0u;                        // rec == cur_chunk * chunk_rec + in_chunk
            auto advance = [&]() -> uint32_t {                               // next group of 32 records worth loading
                in_chunk += 32u;
                if (in_chunk >= chunk_rec) {
                    in_chunk = 0u;
                    ++cur_chunk;
                    while (cur_chunk < nchunks && s_prefix[cur_chunk] == s_prefix[cur_chunk - 1u]) ++cur_chunk;
                }
                return cur_chunk < nchunks ? cur_chunk * chunk_rec + in_chunk : nrec;
            };
            auto load_masks = [&](uint32_t r) -> uint4 {
                const uint32_t my = r + lane;
                return my < nrec ? __ldcg(p.draw_masks + my) : make_uint4(0u, 0u, 0u, 0u);
            };
            uint4 e = load_masks(rec);
            while (running < o_end && rec < nrec) {
                const uint32_t rec_next = advance();
                const uint4 e_next = load_masks(rec_next);                // speculative: unused when this group ends the share
                const uint32_t dm = e.x, entity = e.y, moff = e.z;
                const uint32_t pc = __popc(dm);
                uint32_t inc = pc;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                    if (lane >= (uint32_t)d) inc += t;
                }
                const uint32_t step_total = __shfl_sync(0xFFFFFFFFu, inc, 31);
                if (running + step_total > o_begin) {
                    __syncwarp();
                    sr[lane] = inc; sr[32 + lane] = dm; sr[64 + lane] = entity; sr[96 + lane] = moff;
                    __syncwarp();
                    const uint32_t l0 = o_begin > running ? o_begin - running : 0u;            // first local output of mine
                    const uint32_t l1 = min(step_total, o_end - running);                      // one past my last local output
                    for (uint32_t ol = l0 + lane; ol < l1; ol += 32u) {
                        uint32_t a = 0u, b = 31u;
#pragma unroll
                        for (int it = 0; it < 5; ++it) {
                            const uint32_t mid = (a + b) >> 1;
                            if (sr[mid] > ol) b = mid; else a = mid + 1u;
                        }
                        const uint32_t r = a;
                        const uint32_t excl = r ? sr[r - 1u] : 0u;
                        const uint32_t j = select_set_bit(sr[32 + r], ol - excl);   // (ol-excl)-th survivor of the record
                        const uint32_t midx = sr[96 + r] + j;
                        const uint4 mb = __ldg(p.meshlets + 2u * (size_t)midx + 1);
                        const uint64_t idx = (uint64_t)running + ol;
                        if (idx < p.capacity_draws) store_command(p.draw_words + 1u + idx * 7u, mb.y, mb.z, mb.w, sr[64 + r], midx);
                    }
                }
                running += step_total;
                rec = rec_next;
                e = e_next;
            }
        }
    } else if (blockIdx.x == 0 && tid == 0) {
        p.draw_words[0] = 0u;   // nothing survived (the steady-state late pass)
    }
    pdl_launch_dependents();   // after the emission: an early trigger measured 20% slower when there is a lot to emit
    if (want_payload) {
        // MeshTaskPayload + emitted task count per record (indices ascending by lane — the task shader's atomicAdd
        // order is arbitrary). One LANE per record packs the 11 words; the warp stages its 32 consecutive records
        // (1408 contiguous bytes of the payload array) in shared memory and writes them out coalesced.
        uint32_t* const st = &s_payload[warp][0];
        for (uint32_t base = gw * 32u; base < nrec; base += GW * 32u) {
            const uint32_t r = base + lane;
            uint4 e = make_uint4(0u, 0u, 0u, 0u);
            if (r < nrec) e = __ldcg(p.draw_masks + r);
            uint32_t m = e.x;
            __syncwarp();
            st[lane * 11u] = __popc(m); st[lane * 11u + 1u] = e.y; st[lane * 11u + 2u] = e.z;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint32_t packed_idx = 0u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t b = (uint32_t)__ffs((int)m);          // 0 when m is empty
                    packed_idx |= (b ? b - 1u : 0u) << (8 * k);
                    m &= m - 1u;                                         // 0 & 0xFFFFFFFF stays 0
                }
                st[lane * 11u + 3u + (uint32_t)q] = packed_idx;
            }
            __syncwarp();
            const uint32_t nwords = min(32u, nrec - base) * 11u;
            uint32_t* const tp = p.task_payloads + (size_t)base * 11u;
            for (uint32_t i = lane; i < nwords; i += 32u) tp[i] = st[i];
        }
    }
}

// ---- launch plumbing ---------------------------------------------------------------------------------------
struct TestVariant { bool packed; bool pass2; int proj; };
static TestVariant variant_of(const OrbitCullInfo& ci) {
    const bool mocc = ci.meshlet_visibility_buffer != ORBIT_NO_BUFFER;
    TestVariant v;
    v.packed = ci.occlusion_pass == 1u && mocc;
    v.pass2 = ci.occlusion_pass == 2u && mocc;
    v.proj = ci.projection_type <= 1u ? (int)ci.projection_type : -1;
    return v;
}

template <int R, bool kPass2, int kProj>
static cudaError_t launch_direct(const MeshletCullParams& p, int grid, cudaStream_t stream, int* occupancy) {
    const size_t smem = sizeof(WarpSmem<R>) * kMcWarps;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(meshlet_test_direct_kernel<R, kPass2, kProj>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (occupancy) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(occupancy, meshlet_test_direct_kernel<R, kPass2, kProj>, kMcThreads, smem);
        return cudaSuccess;
    }
    return launch_kernel(meshlet_test_direct_kernel<R, kPass2, kProj>, dim3(grid), dim3(kMcThreads), smem, stream, p);
}

template <int R>
static cudaError_t launch_test(const MeshletCullParams& p, int grid, cudaStream_t stream, int* occupancy) {
    const TestVariant v = variant_of(p.cull);
    if (v.packed) {
        // the packed kernel fills its 32-lane test batches from a whole tile: with the default R = 4 a batch is 44 % full
        // on C2, so it takes 8-record tiles instead (measured 8.9 -> 8.4 us); ORBIT_MC_RECS_PER_WARP=2 or 8 are taken as given
        constexpr int RP = R == 4 ? 8 : R;
        if (occupancy) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(occupancy, meshlet_test_packed_kernel<RP>, kMcThreads, 0); return cudaSuccess; }
        return launch_kernel(meshlet_test_packed_kernel<RP>, dim3(grid), dim3(kMcThreads), 0, stream, p);
    }
    if (v.pass2) {
        if (v.proj == 0) return launch_direct<R, true, 0>(p, grid, stream, occupancy);
        if (v.proj == 1) return launch_direct<R, true, 1>(p, grid, stream, occupancy);
        return launch_direct<R, true, -1>(p, grid, stream, occupancy);
    }
    if (v.proj == 0) return launch_direct<R, false, 0>(p, grid, stream, occupancy);
    if (v.proj == 1) return launch_direct<R, false, 1>(p, grid, stream, occupancy);
    return launch_direct<R, false, -1>(p, grid, stream, occupancy);
}

static cudaError_t launch_test_r(const MeshletCullParams& p, int recs_per_warp, int grid, cudaStream_t stream, int* occupancy) {
    switch (recs_per_warp) {
        case 2: return launch_test<2>(p, grid, stream, occupancy);
        case 8: return launch_test<8>(p, grid, stream, occupancy);
        default: return launch_test<4>(p, grid, stream, occupancy);
    }
}

// Index of the kernel variant a CullInfo selects (api.cu caches one occupancy per variant).
int meshlet_cull_variant_index(const OrbitCullInfo& ci) {
    const TestVariant v = variant_of(ci);
    return v.packed ? 0 : 1 + (v.pass2 ? 3 : 0) + (v.proj + 1);
}

int meshlet_cull_max_ctas_per_sm(const MeshletCullParams& p, int recs_per_warp) {
    int n = 0;
    launch_test_r(p, recs_per_warp, 0, nullptr, &n);
    return n;
}

cudaError_t launch_meshlet_cull(const MeshletCullParams& p, int recs_per_warp, int grid, int emit_grid, cudaStream_t stream) {
    if (grid > 0) {
        cudaError_t e = launch_test_r(p, recs_per_warp, grid, stream, nullptr);
        if (e != cudaSuccess) return e;
    }
    if (emit_grid > 0) return launch_kernel(meshlet_emit_kernel, dim3(emit_grid), dim3(kEmitWarps * 32), 0, stream, p);
    return cudaSuccess;
}

int meshlet_emit_max_ctas_per_sm() {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, meshlet_emit_kernel, kEmitWarps * 32, 0);
    return n;
}

}  // namespace orbit
