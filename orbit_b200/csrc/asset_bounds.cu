// asset_bounds.cu — asset-side producers of the culling path's inputs, sm_100a.
//
// Stands in for the bounds part of compute_meshlets (src/assets/mesh.rs:292-338): per meshlet,
// meshopt::compute_meshlet_bounds -> GpuMeshlet{bounding_sphere, cone_axis (snorm8 x 3), cone_cutoff (snorm8)}, and for
// MeshData::compute_bounds (src/assets/mesh.rs:192-215): per mesh, AABB + sphere (centre = AABB middle, radius = largest
// vertex distance) -> GpuMeshInfo{bounding_sphere, aabb}.
//
// meshopt 0.2.0 is an un-vendored dependency (Cargo.toml:40): the algorithm restated here is meshoptimizer's published
// meshopt_computeMeshletBounds / computeBoundingSphere (clusterizer.cpp): triangle normals and corners of the
// non-degenerate triangles, Ritter-style bounding sphere of the corners (extreme points along the three axes, the
// longest of the three spans as the first diameter, then one pass growing the sphere over the points IN ORDER), the same
// sphere over the normals (its centre is the cone axis), mindp = smallest dot(normal, axis) -> cone wider than ~168
// degrees: cutoff 127 and a zero axis; else cutoff = sqrt(1 - mindp^2), axis and cutoff quantised to snorm8 with the cutoff
// rounded up by the axis' quantisation error. Arithmetic: binary32, every product and sum individually rounded, left to
// right as the source writes them (explicit intrinsics: no contraction) — the oracle restates the same order.
// PARITY UNPINNED against the reference itself (no executable meshopt here): pinned oracle <-> CUDA only.
//
// B200 design: one warp per meshlet. The lanes gather and classify the triangles (up to 128), compacting the
// non-degenerate ones IN ORDER into shared memory (ballot ranks); the extreme-point search is a lane-strided scan + warp
// arg-min / arg-max that breaks ties towards the smaller index exactly like the sequential loop; the order-dependent
// sphere growth runs redundantly on all lanes over broadcast shared-memory reads; min / max reductions close the cone.
#include "params.cuh"

namespace orbit {

constexpr int kAbWarps = 4;
constexpr uint32_t kAbMaxTris = 128u;

struct AbSmem {
    float corners[kAbMaxTris * 3][3];
    float normals[kAbMaxTris][3];
};

// first index attaining the minimum (kMax: maximum) of points[i][axis], i < count — ties towards the smaller index, as the
// sequential `pmin = (p < points[pmin]) ? i : pmin` loop leaves it
template <bool kMax>
__device__ __forceinline__ uint32_t arg_extreme(const float (*points)[3], uint32_t count, int axis, uint32_t lane) {
    float best = points[0][axis];
    uint32_t bi = 0u;
    for (uint32_t i = lane; i < count; i += 32u) {
        const float v = points[i][axis];
        if (kMax ? (v > best) : (v < best)) { best = v; bi = i; }
        else if (v == best && i < bi) bi = i;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const float ov = __shfl_xor_sync(0xFFFFFFFFu, best, d);
        const uint32_t oi = __shfl_xor_sync(0xFFFFFFFFu, bi, d);
        if ((kMax ? (ov > best) : (ov < best)) || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    return bi;
}

// computeBoundingSphere: result = (centre xyz, radius); all lanes return the same value
__device__ __forceinline__ float4 bounding_sphere(const float (*points)[3], uint32_t count, uint32_t lane) {
    uint32_t pmin[3], pmax[3];
#pragma unroll
    for (int axis = 0; axis < 3; ++axis) {
        pmin[axis] = arg_extreme<false>(points, count, axis, lane);
        pmax[axis] = arg_extreme<true>(points, count, axis, lane);
    }
    float paxisd2 = 0.0f;
    int paxis = 0;
#pragma unroll
    for (int axis = 0; axis < 3; ++axis) {
        const float* p1 = points[pmin[axis]];
        const float* p2 = points[pmax[axis]];
        const float dx = sub(p2[0], p1[0]), dy = sub(p2[1], p1[1]), dz = sub(p2[2], p1[2]);
        const float d2 = add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz));
        if (d2 > paxisd2) { paxisd2 = d2; paxis = axis; }
    }
    const float* p1 = points[pmin[paxis]];
    const float* p2 = points[pmax[paxis]];
    float cx = fdiv(add(p1[0], p2[0]), 2.0f), cy = fdiv(add(p1[1], p2[1]), 2.0f), cz = fdiv(add(p1[2], p2[2]), 2.0f);
    float radius = fdiv(fsqrt(paxisd2), 2.0f);
    // grow the sphere over the points in order (order-dependent: every lane walks the same sequence)
    for (uint32_t i = 0; i < count; ++i) {
        const float px = points[i][0], py = points[i][1], pz = points[i][2];
        const float dx = sub(px, cx), dy = sub(py, cy), dz = sub(pz, cz);
        const float d2 = add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz));
        if (d2 > mul(radius, radius)) {
            const float d = fsqrt(d2);
            const float k = add(0.5f, fdiv(fdiv(radius, d), 2.0f));
            const float k1 = sub(1.0f, k);
            cx = add(mul(cx, k), mul(px, k1));
            cy = add(mul(cy, k), mul(py, k1));
            cz = add(mul(cz, k), mul(pz, k1));
            radius = fdiv(add(radius, d), 2.0f);
        }
    }
    return make_float4(cx, cy, cz, radius);
}

// meshopt_quantizeSnorm(v, 8)
__device__ __forceinline__ int quantize_snorm8(float v) {
    const float round = (v >= 0.0f) ? 0.5f : -0.5f;
    v = (v >= -1.0f) ? v : -1.0f;
    v = (v <= 1.0f) ? v : 1.0f;
    return (int)add(mul(v, 127.0f), round);
}

__global__ void __launch_bounds__(kAbWarps * 32) meshlet_bounds_kernel(const __grid_constant__ MeshletBoundsParams p) {
    __shared__ AbSmem s_all[kAbWarps];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    AbSmem& sm = s_all[warp];
    for (uint32_t m = blockIdx.x * kAbWarps + warp; m < p.n_meshlets; m += gridDim.x * kAbWarps) {
        __syncwarp();
        const uint4 mb = __ldg(reinterpret_cast<const uint4*>(p.meshlets) + 2u * (size_t)m + 1u);   // cone, vertex_offset, data_offset, packed counts
        const uint32_t vertex_offset = mb.y, data_offset = mb.z;
        const uint32_t vcount = (mb.w >> 16) & 0xFFu, tcount = mb.w >> 24;
        const uint32_t* vidx = p.meshlet_data + data_offset;
        const uint8_t* tris = reinterpret_cast<const uint8_t*>(vidx + vcount);
        if (tcount > kAbMaxTris) {          // beyond what one warp stages (the reference builds 64 / 64 meshlets)
            if (lane == 0u) *p.error_flag = 1u;
            continue;
        }
        // ---- triangle normals and corners of the non-degenerate triangles, in order
        uint32_t triangles = 0u;
        for (uint32_t t0 = 0; t0 < tcount; t0 += 32u) {
            const uint32_t t = t0 + lane;
            bool valid = false;
            float c[3][3], n[3] = {0.f, 0.f, 0.f};
            if (t < tcount) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint32_t local = tris[3u * t + (uint32_t)k];
                    const uint32_t gv = vertex_offset + __ldg(vidx + local);
                    const float* pos = reinterpret_cast<const float*>(p.vertices + (size_t)gv * p.vertex_stride);
                    c[k][0] = __ldg(pos); c[k][1] = __ldg(pos + 1); c[k][2] = __ldg(pos + 2);
                }
                const float ax = sub(c[1][0], c[0][0]), ay = sub(c[1][1], c[0][1]), az = sub(c[1][2], c[0][2]);
                const float bx = sub(c[2][0], c[0][0]), by = sub(c[2][1], c[0][1]), bz = sub(c[2][2], c[0][2]);
                const float nx = sub(mul(ay, bz), mul(az, by));
                const float ny = sub(mul(az, bx), mul(ax, bz));
                const float nz = sub(mul(ax, by), mul(ay, bx));
                const float area = fsqrt(add(add(mul(nx, nx), mul(ny, ny)), mul(nz, nz)));
                valid = !(area == 0.0f);        // degenerate triangles are invisible anyway
                if (valid) { n[0] = fdiv(nx, area); n[1] = fdiv(ny, area); n[2] = fdiv(nz, area); }
            }
            const uint32_t vm = __ballot_sync(0xFFFFFFFFu, valid);
            if (valid) {
                const uint32_t r = triangles + (uint32_t)__popc(vm & lt);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    sm.normals[r][k] = n[k];
#pragma unroll
                    for (int q = 0; q < 3; ++q) sm.corners[3u * r + (uint32_t)k][q] = c[k][q];
                }
            }
            triangles += (uint32_t)__popc(vm);
        }
        __syncwarp();
        uint32_t* out = reinterpret_cast<uint32_t*>(p.meshlets) + 8u * (size_t)m;
        if (triangles == 0u) {                  // degenerate cluster: zero bounds (trivial reject, cone data 0)
            if (lane < 5u) out[lane] = 0u;
            continue;
        }
        const float4 ps = bounding_sphere(sm.corners, triangles * 3u, lane);
        const float4 ns = bounding_sphere(sm.normals, triangles, lane);
        float ax = ns.x, ay = ns.y, az = ns.z;
        const float axislength = fsqrt(add(add(mul(ax, ax), mul(ay, ay)), mul(az, az)));
        const float inv = axislength == 0.0f ? 0.0f : fdiv(1.0f, axislength);
        ax = mul(ax, inv); ay = mul(ay, inv); az = mul(az, inv);
        // tight cone around all normals: mindp = cos(angle / 2)
        float mindp = 1.0f;
        for (uint32_t i = lane; i < triangles; i += 32u) {
            const float dp = add(add(mul(sm.normals[i][0], ax), mul(sm.normals[i][1], ay)), mul(sm.normals[i][2], az));
            mindp = (dp < mindp) ? dp : mindp;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) { const float o = __shfl_xor_sync(0xFFFFFFFFu, mindp, d); mindp = (o < mindp) ? o : mindp; }
        int8_t q[4];
        if (mindp <= 0.1f) {                    // cone wider than ~168 degrees: not useful -> trivial accept
            q[0] = q[1] = q[2] = 0; q[3] = 127;
        } else {
            // (the cone apex — the point on centre - t * axis behind every triangle — is not part of GpuMeshlet and is not computed)
            const float cutoff = fsqrt(sub(1.0f, mul(mindp, mindp)));
            const int q0 = quantize_snorm8(ax), q1 = quantize_snorm8(ay), q2 = quantize_snorm8(az);
            const float e0 = fabsf(sub(fdiv((float)(signed char)q0, 127.0f), ax));
            const float e1 = fabsf(sub(fdiv((float)(signed char)q1, 127.0f), ay));
            const float e2 = fabsf(sub(fdiv((float)(signed char)q2, 127.0f), az));
            const int qc = (int)add(mul(127.0f, add(add(add(cutoff, e0), e1), e2)), 1.0f);   // rounded UP: the 8-bit test must stay conservative
            q[0] = (int8_t)q0; q[1] = (int8_t)q1; q[2] = (int8_t)q2; q[3] = qc > 127 ? (int8_t)127 : (int8_t)qc;
        }
        if (lane == 0u) {
            out[0] = __float_as_uint(ps.x); out[1] = __float_as_uint(ps.y); out[2] = __float_as_uint(ps.z); out[3] = __float_as_uint(ps.w);
            out[4] = (uint32_t)(uint8_t)q[0] | ((uint32_t)(uint8_t)q[1] << 8) | ((uint32_t)(uint8_t)q[2] << 16) | ((uint32_t)(uint8_t)q[3] << 24);
        }
    }
}

// MeshData::compute_bounds: one CTA per mesh; min / max / max-distance are order-independent, so plain reductions are exact
__global__ void __launch_bounds__(256) mesh_bounds_kernel(const __grid_constant__ MeshBoundsParams p) {
    __shared__ float s_red[8][6];
    __shared__ float s_centre[3];
    __shared__ float s_r2[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint32_t mesh = blockIdx.x; mesh < p.n_meshes; mesh += gridDim.x) {
        const uint32_t first = __ldg(p.vertex_ranges + 2u * mesh), count = __ldg(p.vertex_ranges + 2u * mesh + 1u);
        if (count == 0u) continue;             // compute_bounds leaves the defaults
        const float inf = __uint_as_float(0x7F800000u);
        float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
        for (uint32_t i = tid; i < count; i += 256u) {
            const float* pos = reinterpret_cast<const float*>(p.vertices + (size_t)(first + i) * p.vertex_stride);
#pragma unroll
            for (int k = 0; k < 3; ++k) { const float v = __ldg(pos + k); lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v); }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], d));
                hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], d));
            }
        }
        __syncthreads();
        if (lane == 0u) { for (int k = 0; k < 3; ++k) { s_red[warp][k] = lo[k]; s_red[warp][3 + k] = hi[k]; } }
        __syncthreads();
        if (tid < 3u) {
            float a = s_red[0][tid], b = s_red[0][3 + tid];
            for (int w = 1; w < 8; ++w) { a = fminf(a, s_red[w][tid]); b = fmaxf(b, s_red[w][3 + tid]); }
            s_red[0][tid] = a; s_red[0][3 + tid] = b;
            s_centre[tid] = mul(add(b, a), 0.5f);          // (max + min) * 0.5
        }
        __syncthreads();
        const float cx = s_centre[0], cy = s_centre[1], cz = s_centre[2];
        float r2 = 0.0f;
        for (uint32_t i = tid; i < count; i += 256u) {
            const float* pos = reinterpret_cast<const float*>(p.vertices + (size_t)(first + i) * p.vertex_stride);
            const float dx = sub(__ldg(pos), cx), dy = sub(__ldg(pos + 1), cy), dz = sub(__ldg(pos + 2), cz);
            r2 = fmaxf(r2, add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz)));       // distance_squared, glam dot order
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xFFFFFFFFu, r2, d));
        if (lane == 0u) s_r2[warp] = r2;
        __syncthreads();
        if (tid == 0u) {
            float r = s_r2[0];
            for (int w = 1; w < 8; ++w) r = fmaxf(r, s_r2[w]);
            float* mi = reinterpret_cast<float*>(p.mesh_infos + (size_t)mesh * 128u);
            mi[0] = cx; mi[1] = cy; mi[2] = cz; mi[3] = fsqrt(r);                   // bounding_sphere @0
            mi[4] = s_red[0][0]; mi[5] = s_red[0][1]; mi[6] = s_red[0][2]; mi[7] = 0.0f;      // aabb.min @16 (vec3 + pad)
            mi[8] = s_red[0][3]; mi[9] = s_red[0][4]; mi[10] = s_red[0][5]; mi[11] = 0.0f;    // aabb.max @32
        }
        __syncthreads();
    }
}

cudaError_t launch_meshlet_bounds(const MeshletBoundsParams& p, int grid, cudaStream_t s) {
    meshlet_bounds_kernel<<<grid, kAbWarps * 32, 0, s>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_mesh_bounds(const MeshBoundsParams& p, int grid, cudaStream_t s) {
    mesh_bounds_kernel<<<grid, 256, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace orbit
