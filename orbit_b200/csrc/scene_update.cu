// scene_update.cu — SceneData::update_scene (src/scene.rs:404-492) on the GPU (sm_100a): the per-frame producer of the
// entity buffers the culling path reads. SURVEY §8f item 3.
//
// Per entity with a mesh, in entity order: compacted instance index, model matrix
// (glam Mat4::from_scale_rotation_translation), normal matrix (upper-left 3x3 of the transposed glam Mat4::inverse),
// GpuEntityDraw {instance_index, mesh slot, visibility offset}; entities without a visibility range receive
// ceil(lod0 meshlets / 32) consecutive words (the reference's FreeListAllocator with nothing freed is a bump pointer:
// collections/freelist_alloc.rs:40-72).
//
// B200 design. The reference's serial loop carries two running sums (instance count, visibility cursor). Here:
//   launch 1  per-CTA sums of both over 256 entities -> one 64-bit word per tile (count | words << 32);
//   launch 2  every CTA adds up the words of the tiles below it (at 250 k entities: < 1000 coalesced 8-byte loads, no
//             spinning, no co-residency requirement), scans its own 256 entities, and produces the outputs. Each lane
//             builds its 128-byte GpuEntityData in registers and parks it in shared memory (row stride 144 B: float4
//             accesses are conflict-free); the warp's rows are consecutive in the output, so they leave as full
//             coalesced 16-byte stores. HBM-bound: 48 + 8 B read, 128 + 12 (+4) B written per entity.
// Arithmetic: the contract of DESIGN.md §3 (individually rounded binary32 ops, IEEE division) in glam's operation order.
#include "params.cuh"

namespace orbit {

constexpr int kSuThreads = 256;
constexpr int kSuWarps = kSuThreads / 32;
constexpr int kRowF4 = 9;   // 8 float4 of payload + 1 of padding per staged entity

struct EntityNeed { bool has; bool need; uint32_t slot; uint32_t vo; uint32_t words; };

__device__ __forceinline__ EntityNeed entity_need(const SceneUpdateParams& p, uint32_t gid) {
    EntityNeed e{false, false, ORBIT_NO_MESH, ORBIT_NO_VISIBILITY_RANGE, 0u};
    if (gid < p.n_entities) {
        e.slot = __ldg(p.mesh_slots + gid);
        e.vo = __ldcg(p.visibility_offsets + gid);
        e.has = e.slot != ORBIT_NO_MESH;
        e.need = e.has && e.vo == ORBIT_NO_VISIBILITY_RANGE;
        if (e.need) {
            const uint32_t mc = __ldg(reinterpret_cast<const uint32_t*>(p.mesh_infos + (size_t)e.slot * 128u + 64u + 4u));   // mesh_lods[0].meshlet_count
            e.words = (mc >> 5) + ((mc & 31u) ? 1u : 0u);
        }
    }
    return e;
}

__global__ void __launch_bounds__(kSuThreads) scene_update_sums_kernel(const __grid_constant__ SceneUpdateParams p) {
    __shared__ uint32_t s_cnt[kSuWarps], s_wrd[kSuWarps];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const EntityNeed e = entity_need(p, blockIdx.x * kSuThreads + tid);
    const uint32_t cnt = __popc(__ballot_sync(0xFFFFFFFFu, e.has));
    const uint32_t wrd = __reduce_add_sync(0xFFFFFFFFu, e.words);
    if (lane == 0u) { s_cnt[warp] = cnt; s_wrd[warp] = wrd; }
    __syncthreads();
    if (tid == 0u) {
        uint32_t c = 0u, w = 0u;
#pragma unroll
        for (int i = 0; i < kSuWarps; ++i) { c += s_cnt[i]; w += s_wrd[i]; }
        p.tile_sums[blockIdx.x] = (unsigned long long)c | ((unsigned long long)w << 32);
        if (blockIdx.x == 0u) *p.cursor_snapshot = *p.visibility_cursor;   // launch 2 reads the snapshot and rewrites the cursor
    }
}

// glam Mat4::from_scale_rotation_translation (quat_to_axes, each axis times its scale component)
__device__ __forceinline__ void model_from_srt(const float4 pos, const float4 q, const float4 scl, float m[16]) {
    const float x = q.x, y = q.y, z = q.z, w = q.w;
    const float x2 = add(x, x), y2 = add(y, y), z2 = add(z, z);
    const float xx = mul(x, x2), xy = mul(x, y2), xz = mul(x, z2), yy = mul(y, y2), yz = mul(y, z2), zz = mul(z, z2);
    const float wx = mul(w, x2), wy = mul(w, y2), wz = mul(w, z2);
    m[0] = mul(sub(1.0f, add(yy, zz)), scl.x); m[1] = mul(add(xy, wz), scl.x); m[2] = mul(sub(xz, wy), scl.x); m[3] = mul(0.0f, scl.x);
    m[4] = mul(sub(xy, wz), scl.y); m[5] = mul(sub(1.0f, add(xx, zz)), scl.y); m[6] = mul(add(yz, wx), scl.y); m[7] = mul(0.0f, scl.y);
    m[8] = mul(add(xz, wy), scl.z); m[9] = mul(sub(yz, wx), scl.z); m[10] = mul(sub(1.0f, add(xx, yy)), scl.z); m[11] = mul(0.0f, scl.z);
    m[12] = pos.x; m[13] = pos.y; m[14] = pos.z; m[15] = 1.0f;
}

__device__ __forceinline__ float det2(float a, float b, float c, float d) { return sub(mul(a, b), mul(c, d)); }
__device__ __forceinline__ float cof3(float a, float fa, float b, float fb, float c, float fc) { return add(sub(mul(a, fa), mul(b, fb)), mul(c, fc)); }

// Upper-left 3x3 of transpose(glam Mat4::inverse(m)) as a Mat4 (Mat4::from_mat3). Only the nine adjugate entries
// that survive are formed, each with exactly the operations glam performs for it, plus the determinant and 1/det.
__device__ __forceinline__ void normal_from_model(const float m[16], float n[16]) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m03 = m[3], m10 = m[4], m11 = m[5], m12 = m[6], m13 = m[7];
    const float m20 = m[8], m21 = m[9], m22 = m[10], m23 = m[11], m30 = m[12], m31 = m[13], m32 = m[14], m33 = m[15];
    const float c00 = det2(m22, m33, m32, m23), c02 = det2(m12, m33, m32, m13), c03 = det2(m12, m23, m22, m13);
    const float c04 = det2(m21, m33, m31, m23), c06 = det2(m11, m33, m31, m13), c07 = det2(m11, m23, m21, m13);
    const float c08 = det2(m21, m32, m31, m22), c10 = det2(m11, m32, m31, m12), c11 = det2(m11, m22, m21, m12);
    const float c12 = det2(m20, m33, m30, m23), c14 = det2(m10, m33, m30, m13), c15 = det2(m10, m23, m20, m13);
    const float c16 = det2(m20, m32, m30, m22), c18 = det2(m10, m32, m30, m12), c19 = det2(m10, m22, m20, m12);
    const float c20 = det2(m20, m31, m30, m21), c22 = det2(m10, m31, m30, m11), c23 = det2(m10, m21, m20, m11);
    // inv column k, row r (sign: column 0 and 2 use (+,-,+,-), column 1 and 3 use (-,+,-,+))
    const float i00 = cof3(m11, c00, m12, c04, m13, c08), i01 = -cof3(m01, c00, m02, c04, m03, c08);
    const float i02 = cof3(m01, c02, m02, c06, m03, c10), i03 = -cof3(m01, c03, m02, c07, m03, c11);
    const float i10 = -cof3(m10, c00, m12, c12, m13, c16), i11 = cof3(m00, c00, m02, c12, m03, c16);
    const float i12 = -cof3(m00, c02, m02, c14, m03, c18);
    const float i20 = cof3(m10, c04, m11, c12, m13, c20), i21 = -cof3(m00, c04, m01, c12, m03, c20);
    const float i22 = cof3(m00, c06, m01, c14, m03, c22);
    const float i30 = -cof3(m10, c08, m11, c16, m12, c20);
    const float det = add(add(mul(m00, i00), mul(m02, i20)), add(mul(m01, i10), mul(m03, i30)));   // (x+z)+(y+w)
    const float rcp = fdiv(1.0f, det);
    (void)i03;
    // n.col[c][r] = inv.col[r][c] * rcp
    n[0] = mul(i00, rcp); n[1] = mul(i10, rcp); n[2] = mul(i20, rcp); n[3] = 0.0f;
    n[4] = mul(i01, rcp); n[5] = mul(i11, rcp); n[6] = mul(i21, rcp); n[7] = 0.0f;
    n[8] = mul(i02, rcp); n[9] = mul(i12, rcp); n[10] = mul(i22, rcp); n[11] = 0.0f;
    n[12] = 0.0f; n[13] = 0.0f; n[14] = 0.0f; n[15] = 1.0f;
}

__global__ void __launch_bounds__(kSuThreads) scene_update_emit_kernel(const __grid_constant__ SceneUpdateParams p) {
    __shared__ float4 s_stage[kSuWarps][32 * kRowF4];
    __shared__ uint32_t s_cnt[kSuWarps], s_wrd[kSuWarps];
    __shared__ unsigned long long s_base[kSuWarps];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile = blockIdx.x, gid = tile * kSuThreads + tid;
    // inputs first (independent of the prefix): the loads are in flight while the lower tiles are summed
    const EntityNeed e = entity_need(p, gid);
    float4 t_pos = make_float4(0.f, 0.f, 0.f, 0.f), t_q = t_pos, t_scl = t_pos;
    if (e.has) {
        const float4* t = reinterpret_cast<const float4*>(p.transforms) + (size_t)gid * 3u;
        t_pos = __ldg(t); t_q = __ldg(t + 1); t_scl = __ldg(t + 2);
    }
    // ---- sums of the tiles below this one (count in the low half, words in the high half; neither can carry into
    //      the other: both totals are below 2^32 by construction of the buffers they index)
    unsigned long long below = 0ull;
    for (uint32_t i = tid; i < tile; i += kSuThreads) below += __ldcg(p.tile_sums + i);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) below += __shfl_xor_sync(0xFFFFFFFFu, below, d);
    // ---- CTA-level exclusive scan of (has, words)
    const uint32_t has_mask = __ballot_sync(0xFFFFFFFFu, e.has);
    uint32_t w_incl = e.words;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w_incl, d);
        if (lane >= (uint32_t)d) w_incl += t;
    }
    if (lane == 31u) { s_cnt[warp] = __popc(has_mask); s_wrd[warp] = w_incl; }
    if (lane == 0u) s_base[warp] = below;
    __syncthreads();
    unsigned long long base = 0ull;
#pragma unroll
    for (int i = 0; i < kSuWarps; ++i) base += s_base[i];
    uint32_t cnt_before = 0u, wrd_before = 0u, cnt_cta = 0u, wrd_cta = 0u;
#pragma unroll
    for (int i = 0; i < kSuWarps; ++i) {
        if ((uint32_t)i < warp) { cnt_before += s_cnt[i]; wrd_before += s_wrd[i]; }
        cnt_cta += s_cnt[i]; wrd_cta += s_wrd[i];
    }
    const uint32_t cursor0 = __ldcg(p.cursor_snapshot);
    const uint32_t warp_first = (uint32_t)(base & 0xFFFFFFFFull) + cnt_before;          // instance index of the warp's first meshed entity
    const uint32_t rank = __popc(has_mask & ((1u << lane) - 1u));
    const uint32_t instance = warp_first + rank;
    const unsigned long long words_before = (base >> 32) + wrd_before + (w_incl - e.words);
    uint32_t vo = e.vo;
    if (e.need) {
        vo = cursor0 + (uint32_t)words_before;
        p.visibility_offsets[gid] = vo;
    }
    if (e.has) {
        float m[16], n[16];
        model_from_srt(t_pos, t_q, t_scl, m);
        normal_from_model(m, n);
        float4* row = &s_stage[warp][rank * kRowF4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            row[k] = make_float4(m[4 * k], m[4 * k + 1], m[4 * k + 2], m[4 * k + 3]);
            row[4 + k] = make_float4(n[4 * k], n[4 * k + 1], n[4 * k + 2], n[4 * k + 3]);
        }
        uint32_t* d = p.entity_draw_words + 1u + 3u * (size_t)instance;
        d[0] = instance; d[1] = e.slot; d[2] = vo;
    }
    __syncwarp();
    {   // the warp's rows are consecutive GpuEntityData entries: coalesced 16-byte stores
        const uint32_t n_f4 = __popc(has_mask) * 8u;
        float4* out = p.entity_data + (size_t)warp_first * 8u;
        for (uint32_t i = lane; i < n_f4; i += 32u) out[i] = s_stage[warp][(i >> 3) * kRowF4 + (i & 7u)];
    }
    if (tile == gridDim.x - 1u && tid == 0u) {
        const unsigned long long total_cnt = (base & 0xFFFFFFFFull) + cnt_cta;
        const unsigned long long end = (unsigned long long)cursor0 + (base >> 32) + wrd_cta;
        p.entity_draw_words[0] = (uint32_t)total_cnt;
        *p.visibility_cursor = (uint32_t)end;
        if (end > (unsigned long long)p.visibility_capacity_words) *p.overflow_flag = 1u;
    }
}

cudaError_t launch_scene_update(const SceneUpdateParams& p, cudaStream_t stream) {
    const uint32_t tiles = (p.n_entities + kSuThreads - 1) / kSuThreads;
    // plain stream-ordered launches (no programmatic dependent launch: launch 2 needs all of launch 1)
    scene_update_sums_kernel<<<tiles, kSuThreads, 0, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    scene_update_emit_kernel<<<tiles, kSuThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace orbit
