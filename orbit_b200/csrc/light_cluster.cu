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
    if (active) p.unique_clusters[4u + s_base + warp_base + __popc(bal & ((1u << lane) - 1u))] = idx;
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
