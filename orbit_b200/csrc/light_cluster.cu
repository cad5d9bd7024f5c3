// light_cluster.cu — clustered light assignment (sm_100a).
//
// Stands in for compute_clusters (src/passes/cluster.rs:368-591) and its three compute shaders:
//   mark_active.comp:27-57               -> mark_active_kernel
//   active_cluster_compaction.comp:17-44 -> compact_clusters_kernel
//   light_culling.comp:34-151            -> light_view_kernel + light_culling_kernel
//
// B200 design:
//   * mark_active: 2-3 global atomics per PIXEL in the reference; here lanes of a warp that hit the same
//     cluster are merged with a ballot loop over the distinct keys + full-mask redux (max / or), so one lane per
//     distinct cluster issues the atomics.
//     atomicMax / atomicOr are order-independent, so the result is deterministic; an L2 read first skips the
//     atomic when the stored value already covers ours (same-address atomics serialise, reads do not).
//   * compaction: ballot + CTA scan + decoupled look-back instead of atomicAdd: cluster ids come out ascending.
//   * light culling: the reference recomputes view*light_position for every (cluster, light) pair and walks
//     the light list twice; here light view-space spheres are computed once (16 B each, L2-resident), a 1024-thread
//     CTA owns one active cluster, each of its 32 warps tests a contiguous stripe of the lights 32 at a time and
//     ballot-compacts hits in ascending order into shared memory; stripes are concatenated in order up to the
//     reference's cap of 256, and per-cluster ranges are packed in compacted-list order by a look-back scan
//     instead of atomicAdd.
#include "params.cuh"

namespace orbit {


// ---------------------------------------------------------------------------------------------------------
// Merge, inside one warp, the lanes that hit the same key, with one REDUX per distinct key instead of match.any +
// partial-mask redux (ncu: the partial-mask `__reduce_or_sync` is a ~32-instruction software loop and was the top
// line of this kernel). Lanes with key == kNoKey take no part. fn(leader_lane_is_me, same_mask) is called by the
// lanes of one key at a time.
constexpr uint32_t kNoKey = 0xFFFFFFFFu;

__global__ void __launch_bounds__(256) mark_active_kernel(const __grid_constant__ ClusterParams p) {
    const OrbitClusterCullInfo& ci = p.info;
    const uint32_t W = ci.screen_size[0], H = ci.screen_size[1];
    const uint32_t cx = ci.cluster_count[0], cy = ci.cluster_count[1], cz = ci.cluster_count[2];
    const uint32_t tile_px = ci.tile_size_px;
    // a warp covers 32 consecutive pixels of kRows consecutive rows per step (kRows independent loads in flight)
    constexpr uint32_t kRows = 4u;
    const uint32_t warps_per_row = (W + 31u) / 32u;
    const uint32_t row_groups = (H + kRows - 1u) / kRows;
    const uint32_t total_warps = warps_per_row * row_groups;   // < 2^32: screen sizes are bounded by the API
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t wi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wi < total_warps; wi += gridDim.x * (blockDim.x >> 5)) {
        const uint32_t row_group = wi / warps_per_row;           // 32-bit: a 64-bit division here cost ~100 instructions per step
        const uint32_t y0 = row_group * kRows;
        const uint32_t x0 = (wi - row_group * warps_per_row) * 32u;
        const uint32_t x = x0 + lane;
        float dv[kRows];
#pragma unroll
        for (uint32_t k = 0; k < kRows; ++k) dv[k] = (x < W && y0 + k < H) ? __ldg(p.depth + (size_t)(y0 + k) * W + x) : 0.0f;
        // tile column of this lane: one (warp-uniform) integer division per step instead of one per pixel
        uint32_t tx = x0 / tile_px;
        {
            uint32_t t = x0 - tx * tile_px + lane;
            if (tile_px >= 8u) { while (t >= tile_px) { t -= tile_px; ++tx; } } else { tx += t / tile_px; }
        }
        const uint32_t ty0 = y0 / tile_px;
        uint32_t ry = y0 - ty0 * tile_px;   // row offset inside the tile row, advanced per k
        uint32_t ty = ty0;
        // ---- per-lane math for the kRows pixels of this lane's column (no warp-level operations yet)
        uint32_t cl[kRows], tl[kRows], mk[kRows], bn[kRows], bx[kRows];
#pragma unroll
        for (uint32_t k = 0; k < kRows; ++k) {
            const uint32_t y = y0 + k;
            cl[k] = kNoKey; tl[k] = kNoKey; mk[k] = 0u; bn[k] = 0u; bx[k] = 0u;
            if (x < W && y < H) {
                const float d = dv[k];
                // sky pixels (d == +0, the clear value of the reverse-Z buffer) would send the IEEE division into its
                // ~50-instruction slow path for the whole warp; z_near / +0 is +inf exactly when z_near > 0
                const float z = (__float_as_uint(d) == 0u && ci.z_near > 0.0f) ? __uint_as_float(0x7F800000u) : fdiv(ci.z_near, d);
                const uint32_t slice = f2u(fma_(orbit_log2f(z), p.z_scale, p.z_bias));
                mk[k] = shl1(slice);
                if (mk[k] != 0u) tl[k] = tx + ty * cx;
                if (slice < cz) {
                    cl[k] = tx + ty * cx + slice * cx * cy;
                    bn[k] = __float_as_uint(sub(1.0f, d));
                    bx[k] = __float_as_uint(d);
                }
            }
            if (++ry == tile_px) { ry = 0u; ++ty; }
        }
        // ---- lane-local merge: vertically adjacent pixels mostly fall into the same cluster / tile, so rows 1..3 are
        //      folded into the first row with the same key and the warp-level merge below usually runs once, not kRows times
#pragma unroll
        for (uint32_t k = 1; k < kRows; ++k) {
#pragma unroll
            for (uint32_t j = 0; j < k; ++j) {
                if (cl[k] != kNoKey && cl[k] == cl[j]) { bn[j] = max(bn[j], bn[k]); bx[j] = max(bx[j], bx[k]); cl[k] = kNoKey; }
                if (tl[k] != kNoKey && tl[k] == tl[j]) { mk[j] |= mk[k]; tl[k] = kNoKey; }
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < kRows; ++k) {
            // ---- cluster depth bounds: one pair of atomics per distinct cluster of the warp
            uint32_t todo = __ballot_sync(0xFFFFFFFFu, cl[k] != kNoKey);
            while (todo) {
                const uint32_t key = __shfl_sync(0xFFFFFFFFu, cl[k], __ffs((int)todo) - 1);
                const bool mine = cl[k] == key;
                const uint32_t same = __ballot_sync(0xFFFFFFFFu, mine);
                const uint32_t mn = __reduce_max_sync(0xFFFFFFFFu, mine ? bn[k] : 0u);
                const uint32_t mx = __reduce_max_sync(0xFFFFFFFFu, mine ? bx[k] : 0u);
                if (lane == (uint32_t)(__ffs((int)same) - 1)) {
                    // values only grow: a (possibly stale) read that is already >= ours makes the atomic a no-op
                    uint32_t* b = p.depth_bounds + 2u * (size_t)key;
                    if (__ldcg(b) < mn) atomicMax(b, mn);
                    if (__ldcg(b + 1) < mx) atomicMax(b + 1, mx);
                }
                todo &= ~same;
            }
            // ---- tile slice masks: one atomicOr per distinct tile of the warp
            todo = __ballot_sync(0xFFFFFFFFu, tl[k] != kNoKey);
            while (todo) {
                const uint32_t key = __shfl_sync(0xFFFFFFFFu, tl[k], __ffs((int)todo) - 1);
                const bool mine = tl[k] == key;
                const uint32_t same = __ballot_sync(0xFFFFFFFFu, mine);
                const uint32_t orm = __reduce_or_sync(0xFFFFFFFFu, mine ? mk[k] : 0u);
                if (lane == (uint32_t)(__ffs((int)same) - 1) && (__ldcg(p.tile_masks + key) & orm) != orm) atomicOr(p.tile_masks + key, orm);
                todo &= ~same;
            }
        }
    }
}

struct Aabb3 { float lo[3], hi[3]; };

__device__ __forceinline__ void unproject(const OrbitClusterCullInfo& ci, float px, float py, float* out) {
    const float tx = fdiv(px, (float)ci.screen_size[0]), ty = fdiv(py, (float)ci.screen_size[1]);
    const float clx = sub(mul(tx, 2.0f), 1.0f), cly = sub(mul(sub(1.0f, ty), 2.0f), 1.0f);
    const float* m = &ci.screen_to_view_matrix.m[0][0];
    const float vx = add(add(add(mul(m[0], clx), mul(m[4], cly)), mul(m[8], 1.0f)), mul(m[12], 1.0f));
    const float vy = add(add(add(mul(m[1], clx), mul(m[5], cly)), mul(m[9], 1.0f)), mul(m[13], 1.0f));
    const float vz = add(add(add(mul(m[2], clx), mul(m[6], cly)), mul(m[10], 1.0f)), mul(m[14], 1.0f));
    const float vw = add(add(add(mul(m[3], clx), mul(m[7], cly)), mul(m[11], 1.0f)), mul(m[15], 1.0f));
    out[0] = fdiv(vx, vw); out[1] = fdiv(vy, vw); out[2] = fdiv(vz, vw);
}

__device__ __forceinline__ void z_plane_point(const float* v, float zd, float* out) {
    const float dn = add(add(mul(0.0f, v[0]), mul(0.0f, v[1])), mul(-1.0f, v[2]));  // dot((0,0,-1), v)
    const float t = fdiv(zd, dn);
    out[0] = mul(v[0], t); out[1] = mul(v[1], t); out[2] = mul(v[2], t);
}

__device__ __forceinline__ Aabb3 cluster_volume(const ClusterParams& p, uint32_t idx) {
    const OrbitClusterCullInfo& ci = p.info;
    const uint32_t cx = ci.cluster_count[0], cy = ci.cluster_count[1];
    const uint32_t z = idx / (cx * cy);
    const uint32_t rem = idx - z * cx * cy;
    const uint32_t y = rem / cx, x = rem - y * cx;
    const float minx = (float)(x * ci.tile_size_px), miny = (float)(y * ci.tile_size_px);
    const float maxx = fminf(add(minx, (float)ci.tile_size_px), (float)ci.screen_size[0]);
    const float maxy = fminf(add(miny, (float)ci.tile_size_px), (float)ci.screen_size[1]);
    float vmin[3], vmax[3];
    unproject(ci, minx, miny, vmin);
    unproject(ci, maxx, maxy, vmax);
    const float min_d = sub(1.0f, __uint_as_float(__ldcg(p.depth_bounds + 2u * (size_t)idx)));
    const float max_d = __uint_as_float(__ldcg(p.depth_bounds + 2u * (size_t)idx + 1u));
    const float cnear = fdiv(ci.z_near, max_d), cfar = fdiv(ci.z_near, min_d);
    float p0[3], p1[3], p2[3], p3[3];
    z_plane_point(vmin, cnear, p0); z_plane_point(vmin, cfar, p1);
    z_plane_point(vmax, cnear, p2); z_plane_point(vmax, cfar, p3);
    Aabb3 a;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a.lo[k] = fminf(fminf(p0[k], p1[k]), fminf(p2[k], p3[k]));
        a.hi[k] = fmaxf(fmaxf(p0[k], p1[k]), fmaxf(p2[k], p3[k]));
    }
    return a;
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) compact_clusters_kernel(const __grid_constant__ ClusterParams p) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_tile, s_base;
    const OrbitClusterCullInfo& ci = p.info;
    const uint32_t cx = ci.cluster_count[0], cy = ci.cluster_count[1], cz = ci.cluster_count[2];
    const uint32_t total = cx * cy * cz;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned int epoch = scan_epoch(p.scan);
    if (tid == 0) s_tile = atomicAdd(p.scan.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t idx = tile * 256u + tid;
    bool active = false;
    if (idx < total) {
        const uint32_t z = idx / (cx * cy);
        const uint32_t t = idx - z * cx * cy;   // tile index = x + y*cx
        active = (__ldcg(p.tile_masks + t) & shl1(z)) != 0u;
    }
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, active);
    if (lane == 0u) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t warp_base = 0u, tile_total = 0u;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint32_t v = s_warp[w]; if ((uint32_t)w < warp) warp_base += v; tile_total += v; }
    if (warp == 0u) {
        const uint32_t base = lookback_exclusive(p.scan, epoch, tile, tile_total);
        if (lane == 0u) {
            s_base = base;
            if (tile == gridDim.x - 1u) {
                const uint32_t n = base + tile_total;
                p.unique_clusters[0] = (n + 255u) / 256u;   // div_ceil(cluster_count, 256)
                p.unique_clusters[1] = 1u;
                p.unique_clusters[2] = 1u;
                p.unique_clusters[3] = n;
            }
        }
    }
    __syncthreads();
    if (active) {
        const uint32_t slot = s_base + warp_base + __popc(bal & ((1u << lane) - 1u));
        p.unique_clusters[4u + slot] = idx;
        // the light-parallel culling path reads every active cluster's view-space box from many CTAs: computed once, here
        // (ten IEEE divisions per cluster, light_culling.comp:68-119), 32 bytes per compacted-list slot
        if (p.cluster_boxes != nullptr) {
            const Aabb3 b = cluster_volume(p, idx);
            p.cluster_boxes[2u * (size_t)slot] = make_float4(b.lo[0], b.lo[1], b.lo[2], 0.0f);
            p.cluster_boxes[2u * (size_t)slot + 1u] = make_float4(b.hi[0], b.hi[1], b.hi[2], 0.0f);
        }
    }
    if (tid == 0) scan_cta_exit(p.scan, epoch);
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) light_view_kernel(const __grid_constant__ ClusterParams p) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.info.global_light_count) return;
    const uint8_t* l = p.lights + (size_t)j * 64u;
    const uint32_t type = __ldg(reinterpret_cast<const uint32_t*>(l));
    const float4 pos = __ldg(reinterpret_cast<const float4*>(l + 32));      // position xyz, inner_radius
    const float radius = __ldg(reinterpret_cast<const float*>(l + 60));
    const float* m = &p.info.world_to_view_matrix.m[0][0];
    float4 o;
    o.x = add(add(add(mul(m[0], pos.x), mul(m[4], pos.y)), mul(m[8], pos.z)), mul(m[12], 1.0f));
    o.y = add(add(add(mul(m[1], pos.x), mul(m[5], pos.y)), mul(m[9], pos.z)), mul(m[13], 1.0f));
    o.z = add(add(add(mul(m[2], pos.x), mul(m[6], pos.y)), mul(m[10], pos.z)), mul(m[14], 1.0f));
    o.w = (type == ORBIT_LIGHT_POINT) ? radius : __uint_as_float(0x7F800000u);  // non-point lights always hit
    p.light_view[j] = o;
}

constexpr int kLcWarps = 32;

// One CTA of kLcWarps warps per active cluster (persistent over clusters through a ticket). The light list is cut
// into kLcWarps contiguous stripes, one per warp; a warp tests 32 lights per step and ballot-compacts its hits
// in ascending order into its own shared-memory list (at most 256 entries can ever be used). The stripes are then
// concatenated in warp order up to the reference's cap of 256, and the cluster's range in the global list comes
// from a look-back scan over clusters in compacted-list order. With few active clusters (typical: ~100 of 3456)
// this spreads each cluster's L tests over 1024 threads instead of one warp (measured 716 us -> see profiles/).
__global__ void __launch_bounds__(kLcWarps * 32) light_culling_kernel(const __grid_constant__ ClusterParams p) {
    extern __shared__ uint32_t s_dyn[];                       // kLcWarps x 256 hit lists
    __shared__ uint32_t s_warp[kLcWarps];
    __shared__ uint32_t s_tile, s_off;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t* const my_list = s_dyn + warp * ORBIT_MAX_LIGHTS_PER_CLUSTER;
    const unsigned int epoch = scan_epoch(p.scan);
    const uint32_t nactive = __ldcg(p.unique_clusters + 3);
    const uint32_t L = p.info.global_light_count;
    const uint32_t per = (L + kLcWarps - 1) / kLcWarps;
    const uint32_t j_begin = min(warp * per, L), j_end = min(j_begin + per, L);
    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(p.scan.ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= nactive) {
            if (tile == 0u && tid == 0) p.light_index_words[0] = 0u;
            break;
        }
        const uint32_t idx = __ldcg(p.unique_clusters + 4u + tile);
        const Aabb3 box = cluster_volume(p, idx);
        uint32_t count = 0u;   // hits of this warp's stripe (uncapped)
        for (uint32_t j0 = j_begin; j0 < j_end; j0 += 128u) {
            float4 sv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {   // four independent 512-byte loads in flight per warp
                const uint32_t j = j0 + (uint32_t)u * 32u + lane;
                sv[u] = j < j_end ? __ldg(p.light_view + j) : make_float4(0.f, 0.f, 0.f, -1.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t j = j0 + (uint32_t)u * 32u + lane;
                bool hit = false;
                if (j < j_end) {
                    const float4 s = sv[u];
                    float acc = 0.0f;
                    const float c[3] = {s.x, s.y, s.z};
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        // lo <= hi, so at most one of the reference's two branches adds a term; d is that term's
                        // base (or 0, and fma(0,0,acc) == acc): same value, no divergence
                        const float v = c[k];
                        const float d = fmaxf(fmaxf(sub(box.lo[k], v), sub(v, box.hi[k])), 0.0f);
                        acc = fma_(d, d, acc);
                    }
                    hit = acc <= mul(s.w, s.w);
                }
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, hit);
                if (hit) {
                    const uint32_t r = count + __popc(bal & ((1u << lane) - 1u));
                    if (r < ORBIT_MAX_LIGHTS_PER_CLUSTER) my_list[r] = j;
                }
                count += __popc(bal);
            }
        }
        if (lane == 0u) s_warp[warp] = min(count, ORBIT_MAX_LIGHTS_PER_CLUSTER);
        __syncthreads();
        // exclusive prefix of the (capped) stripe counts in warp order; everything past 256 is dropped
        uint32_t before = 0u, total = 0u;
#pragma unroll
        for (int w = 0; w < kLcWarps; ++w) { const uint32_t v = s_warp[w]; if ((uint32_t)w < warp) before += v; total += v; }
        total = min(total, ORBIT_MAX_LIGHTS_PER_CLUSTER);
        if (warp == 0u) {
            const uint32_t off = lookback_exclusive(p.scan, epoch, tile, total);
            if (lane == 0u) {
                s_off = off;
                p.offset_count_image[2u * (size_t)idx] = off;
                p.offset_count_image[2u * (size_t)idx + 1u] = total;
                if (tile == nactive - 1u) {
                    p.light_index_words[0] = off + total;
                    if ((uint64_t)off + total > p.capacity_indices) *p.overflow_flag = 1u;
                }
            }
        }
        __syncthreads();
        const uint32_t off = s_off;
        const uint32_t mine = min(count, ORBIT_MAX_LIGHTS_PER_CLUSTER);
        for (uint32_t k = lane; k < mine; k += 32u) {
            const uint32_t r = before + k;
            if (r < ORBIT_MAX_LIGHTS_PER_CLUSTER && (uint64_t)off + r < p.capacity_indices) p.light_index_words[1u + off + r] = my_list[k];
        }
    }
    if (tid == 0) scan_cta_exit(p.scan, epoch);
}

// ---------------------------------------------------------------------------------------------------------
// Light-parallel formulation of light_culling.comp:121-151 (the path taken whenever its bit matrix fits the scratch budget;
// the CTA-per-cluster kernel above remains for grids with ~10^6 clusters). The CTA-per-cluster kernel re-streams the whole
// light array once per active cluster and occupies one SM per cluster (C4: 99 clusters -> 99 of 148 SMs, 232 G tests/s).
// Here the work is cut the other way:
//   1. light_hits_kernel: a CTA owns 512 consecutive lights (two per thread, view-space sphere computed in registers —
//      no light_view pass) and a chunk of 32 active clusters whose view-space boxes sit in shared memory; every
//      (light, cluster) test is one broadcast shared-memory read of the box + ~17 FP32 instructions, a ballot per
//      (warp, cluster) yields 32 hit bits, and the CTA leaves a 16-word row per cluster in a bit matrix
//      [active cluster][light / 32] (coalesced 64-byte stores). Grid = light blocks x cluster chunks: every SM busy.
//   2. light_lists_kernel: a warp per active cluster reads its row of the matrix (2049 words at C4), counts, caps at 256
//      (the reference keeps the FIRST 256 hits in ascending light order: light_culling.comp:121-132), the clusters' ranges
//      are packed in compacted-list order by the look-back scan, and the warp expands its row's first 256 set bits into
//      ascending light indices. Same arithmetic per test as the kernel above, so the lists are bit-identical.
constexpr uint32_t kLhLightsPerCta = 512u, kLhClusterChunk = 32u, kLhWordsPerCta = kLhLightsPerCta / 32u;

// hits:   [active cluster][words_per_cluster]   one bit per light; words_per_cluster = 16 x light blocks (rows are 64-byte aligned)
// counts: [active cluster][light_blocks]        hits of the cluster among the 512 lights of one light block
__global__ void __launch_bounds__(256) light_hits_kernel(const __grid_constant__ ClusterParams p, uint32_t* __restrict__ hits,
                                                         uint32_t* __restrict__ counts, uint32_t words_per_cluster) {
    __shared__ float s_box[kLhClusterChunk][8];                  // lo xyz, hi xyz (padded to 32 bytes: two 16-byte broadcast reads)
    __shared__ uint32_t s_out[kLhClusterChunk][kLhWordsPerCta];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t nactive = __ldcg(p.unique_clusters + 3);
    if (blockIdx.y * kLhClusterChunk >= nactive) return;        // chunk rows beyond the active clusters: nothing to do, before any load
    const uint32_t L = p.info.global_light_count;
    const uint32_t light0 = blockIdx.x * kLhLightsPerCta;
    // this thread's two lights: light0 + tid and light0 + 256 + tid (word = block * 16 + k * 8 + warp, bit = lane)
    float lx[2], ly[2], lz[2], lr2[2];
    bool live[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const uint32_t j = light0 + (uint32_t)k * 256u + tid;
        live[k] = j < L;
        lx[k] = ly[k] = lz[k] = 0.0f; lr2[k] = -1.0f;
        if (live[k]) {
            const uint8_t* l = p.lights + (size_t)j * 64u;
            const uint32_t type = __ldg(reinterpret_cast<const uint32_t*>(l));
            const float4 pos = __ldg(reinterpret_cast<const float4*>(l + 32));      // position xyz, inner_radius
            const float radius = __ldg(reinterpret_cast<const float*>(l + 60));
            const float* m = &p.info.world_to_view_matrix.m[0][0];
            lx[k] = add(add(add(mul(m[0], pos.x), mul(m[4], pos.y)), mul(m[8], pos.z)), mul(m[12], 1.0f));
            ly[k] = add(add(add(mul(m[1], pos.x), mul(m[5], pos.y)), mul(m[9], pos.z)), mul(m[13], 1.0f));
            lz[k] = add(add(add(mul(m[2], pos.x), mul(m[6], pos.y)), mul(m[10], pos.z)), mul(m[14], 1.0f));
            const float w = (type == ORBIT_LIGHT_POINT) ? radius : __uint_as_float(0x7F800000u);   // non-point lights always hit
            lr2[k] = mul(w, w);
        }
    }
    for (uint32_t chunk0 = blockIdx.y * kLhClusterChunk; chunk0 < nactive; chunk0 += gridDim.y * kLhClusterChunk) {
        const uint32_t nc = min(kLhClusterChunk, nactive - chunk0);
        __syncthreads();                                         // previous chunk's boxes and rows are no longer read
        if (tid < 2u * nc) *reinterpret_cast<float4*>(&s_box[tid >> 1][(tid & 1u) * 4u]) = __ldcg(p.cluster_boxes + 2u * (size_t)chunk0 + tid);
        __syncthreads();
        for (uint32_t c = 0; c < nc; ++c) {
            const float4 lo = *reinterpret_cast<const float4*>(&s_box[c][0]);
            const float4 hi = *reinterpret_cast<const float4*>(&s_box[c][4]);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                // per axis: v < lo adds (lo-v)^2, v > hi adds (v-hi)^2 (light_culling.comp:52-66); lo <= hi, so at most one
                // applies and d is that term's base (or 0, and fma(0,0,acc) == acc): same value, no divergence
                float acc = 0.0f;
                float d = fmaxf(fmaxf(sub(lo.x, lx[k]), sub(lx[k], hi.x)), 0.0f); acc = fma_(d, d, acc);
                d = fmaxf(fmaxf(sub(lo.y, ly[k]), sub(ly[k], hi.y)), 0.0f); acc = fma_(d, d, acc);
                d = fmaxf(fmaxf(sub(lo.z, lz[k]), sub(lz[k], hi.z)), 0.0f); acc = fma_(d, d, acc);
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, live[k] && acc <= lr2[k]);
                if (lane == 0u) s_out[c][(uint32_t)k * 8u + warp] = bal;
            }
        }
        __syncthreads();
        // rows out: 16 consecutive words per cluster + their popcount (16 lanes per cluster: half-warp reduction)
        for (uint32_t i0 = 0; i0 < kLhClusterChunk * kLhWordsPerCta; i0 += 256u) {
            const uint32_t i = i0 + tid, c = i >> 4, w = i & 15u;
            const uint32_t v = c < nc ? s_out[c][w] : 0u;
            const uint32_t word = blockIdx.x * kLhWordsPerCta + w;
            if (c < nc) hits[(size_t)(chunk0 + c) * words_per_cluster + word] = v;
            uint32_t n = (uint32_t)__popc(v);
            n += __shfl_xor_sync(0xFFFFFFFFu, n, 1); n += __shfl_xor_sync(0xFFFFFFFFu, n, 2);
            n += __shfl_xor_sync(0xFFFFFFFFu, n, 4); n += __shfl_xor_sync(0xFFFFFFFFu, n, 8);
            if (c < nc && w == 0u) counts[(size_t)(chunk0 + c) * gridDim.x + blockIdx.x] = n;
        }
    }
}

// One warp per active cluster, 32 clusters per CTA (a tile of the look-back scan: with ~100 active clusters the chain is four
// tiles long). The cluster's hits per light block are scanned (ascending light order = ascending block order), the total is
// capped at the reference's 256, the tile's range comes from the look-back over tiles in compacted-list order, and every lane
// expands the blocks it owns at their ranks — only blocks that hold a hit are ever read from the bit matrix.
constexpr int kLlWarps = 32;
__global__ void __launch_bounds__(kLlWarps * 32) light_lists_kernel(const __grid_constant__ ClusterParams p, const uint32_t* __restrict__ hits,
                                                                    const uint32_t* __restrict__ counts, uint32_t words_per_cluster, uint32_t light_blocks) {
    __shared__ uint32_t s_cnt[kLlWarps];
    __shared__ uint32_t s_tile, s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned int epoch = scan_epoch(p.scan);
    const uint32_t nactive = __ldcg(p.unique_clusters + 3);
    const uint32_t ntiles = (nactive + kLlWarps - 1u) / kLlWarps;
    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(p.scan.ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) {
            if (tile == 0u && tid == 0) p.light_index_words[0] = 0u;      // no active cluster at all
            break;
        }
        const uint32_t t = tile * kLlWarps + warp;                         // this warp's active cluster (compacted-list position)
        const bool have = t < nactive;
        const uint32_t* crow = counts + (size_t)t * light_blocks;
        // ---- 1. hits of the cluster, capped at the reference's 256
        uint32_t count = 0u;
        constexpr int kKeep = 8;                                          // block counts kept in registers: 8 x 32 blocks = 131 072 lights
        uint32_t kept[kKeep];
#pragma unroll
        for (int k = 0; k < kKeep; ++k) {
            const uint32_t b = (uint32_t)k * 32u + lane;
            kept[k] = (have && b < light_blocks) ? __ldcg(crow + b) : 0u;
            count += kept[k];
        }
        if (have) for (uint32_t b = (uint32_t)kKeep * 32u + lane; b < light_blocks; b += 32u) count += __ldcg(crow + b);
        count = min(__reduce_add_sync(0xFFFFFFFFu, count), (uint32_t)ORBIT_MAX_LIGHTS_PER_CLUSTER);
        if (lane == 0u) s_cnt[warp] = count;
        __syncthreads();
        uint32_t before = 0u, total = 0u;
#pragma unroll
        for (int w = 0; w < kLlWarps; ++w) { const uint32_t v = s_cnt[w]; if ((uint32_t)w < warp) before += v; total += v; }
        // ---- 2. the tile's range in the global list: look-back over tiles in compacted-list order
        if (warp == 0u) {
            const uint32_t off = lookback_exclusive(p.scan, epoch, tile, total);
            if (lane == 0u) {
                s_base = off;
                if (tile == ntiles - 1u) {
                    p.light_index_words[0] = off + total;
                    if ((uint64_t)off + total > p.capacity_indices) *p.overflow_flag = 1u;
                }
            }
        }
        __syncthreads();
        if (have) {
            const uint32_t off = s_base + before;
            const uint32_t idx = __ldcg(p.unique_clusters + 4u + t);
            if (lane == 0u) { p.offset_count_image[2u * (size_t)idx] = off; p.offset_count_image[2u * (size_t)idx + 1u] = count; }
            // ---- 3. blocks in ascending order, 32 per step: rank of a block's first hit = hits of all earlier blocks
            const uint32_t* row = hits + (size_t)t * words_per_cluster;
            uint32_t running = 0u;
            // one step = 32 blocks: n = this lane's block count; warp-uniform control flow (b0, running, count are uniform)
            auto expand = [&](uint32_t b0, uint32_t n) {
                const uint32_t b = b0 + lane;
                uint32_t inc = n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                    if (lane >= (uint32_t)d) inc += v;
                }
                uint32_t pos = running + inc - n;
                if (n != 0u && pos < count) {
                    // the block's 16 words in four 16-byte loads, all in flight at once (one word at a time made every hit block
                    // a chain of 16 dependent L2 round trips: 32 us for the 99 clusters of C4)
                    const uint4* r4 = reinterpret_cast<const uint4*>(row + (size_t)b * kLhWordsPerCta);
                    const uint4 v0 = __ldcg(r4), v1 = __ldcg(r4 + 1), v2 = __ldcg(r4 + 2), v3 = __ldcg(r4 + 3);
                    const uint32_t words[kLhWordsPerCta] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
#pragma unroll
                    for (uint32_t w = 0; w < kLhWordsPerCta; ++w) {
                        uint32_t word = words[w];
                        while (word != 0u && pos < count) {
                            const uint32_t bit = (uint32_t)__ffs((int)word) - 1u;
                            if ((uint64_t)off + pos < p.capacity_indices) p.light_index_words[1u + off + pos] = (b * kLhWordsPerCta + w) * 32u + bit;
                            word &= word - 1u;
                            ++pos;
                        }
                    }
                }
                running += __shfl_sync(0xFFFFFFFFu, inc, 31);
            };
            // the kept counts are indexed by compile-time constants only: with a run-time select over kept[] inside a rolled
            // loop the step at b0 = 32 received kept[0] on the device (C4: every hit in light blocks 32.. of some clusters was
            // dropped; caught by the full-size C4 test, tools/debug/c4_lights.py shows the case)
#pragma unroll
            for (int k = 0; k < kKeep; ++k)
                if ((uint32_t)k * 32u < light_blocks && running < count) expand((uint32_t)k * 32u, kept[k]);
#pragma unroll 1
            for (uint32_t b0 = (uint32_t)kKeep * 32u; b0 < light_blocks && running < count; b0 += 32u)
                expand(b0, b0 + lane < light_blocks ? __ldcg(crow + b0 + lane) : 0u);
        }
    }
    if (tid == 0) scan_cta_exit(p.scan, epoch);
}

uint32_t light_hits_blocks(uint32_t n_lights) { return (n_lights + kLhLightsPerCta - 1u) / kLhLightsPerCta; }

cudaError_t launch_light_hits(const ClusterParams& p, uint32_t* hits, uint32_t* counts, uint32_t words_per_cluster, uint32_t max_clusters, cudaStream_t s) {
    const uint32_t gx = light_hits_blocks(p.info.global_light_count);
    uint32_t gy = (max_clusters + kLhClusterChunk - 1u) / kLhClusterChunk;
    if (gy > 16u) gy = 16u;                                      // chunks beyond that are strided over by the same CTAs
    if (gx == 0u) return cudaSuccess;
    light_hits_kernel<<<dim3(gx, gy ? gy : 1u), 256, 0, s>>>(p, hits, counts, words_per_cluster);
    return cudaGetLastError();
}
cudaError_t launch_light_lists(const ClusterParams& p, const uint32_t* hits, const uint32_t* counts, uint32_t words_per_cluster, int grid, cudaStream_t s) {
    light_lists_kernel<<<grid, kLlWarps * 32, 0, s>>>(p, hits, counts, words_per_cluster, light_hits_blocks(p.info.global_light_count));
    return cudaGetLastError();
}

cudaError_t launch_mark_active(const ClusterParams& p, int grid, cudaStream_t s) {
    mark_active_kernel<<<grid, 256, 0, s>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_compact_clusters(const ClusterParams& p, cudaStream_t s) {
    const uint32_t total = p.info.cluster_count[0] * p.info.cluster_count[1] * p.info.cluster_count[2];
    compact_clusters_kernel<<<(total + 255u) / 256u, 256, 0, s>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_light_view(const ClusterParams& p, cudaStream_t s) {
    const uint32_t L = p.info.global_light_count;
    if (L == 0) return cudaSuccess;
    light_view_kernel<<<(L + 255u) / 256u, 256, 0, s>>>(p);
    return cudaGetLastError();
}
static constexpr size_t light_culling_smem_bytes() { return (size_t)kLcWarps * ORBIT_MAX_LIGHTS_PER_CLUSTER * sizeof(uint32_t); }

// More than 48 KB of dynamic shared memory: opt in once per DEVICE (called from orbit_ctx_create with the context's device current).
cudaError_t light_cluster_configure_device() {
    return cudaFuncSetAttribute(light_culling_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)light_culling_smem_bytes());
}

cudaError_t launch_light_culling(const ClusterParams& p, int grid, cudaStream_t s) {
    light_culling_kernel<<<grid, kLcWarps * 32, light_culling_smem_bytes(), s>>>(p);
    return cudaGetLastError();
}

}  // namespace orbit
