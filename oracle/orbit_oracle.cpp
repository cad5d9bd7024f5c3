// orbit_oracle.cpp — CPU restatement of the reference's visibility pipeline.
//
// *** TEST INFRASTRUCTURE.  Not part of the product. ***
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library, and only as the checker / CPU baseline.  liborbit_b200.so never links, loads or calls it.
//
// PARITY STATUS.  Thefefe/orbit has no test, golden vector or fixture for this path, and neither a
// Vulkan implementation nor a Rust toolchain exists in the build image, so the reference BINARY cannot be run here.
// What can be run are its shipped GPU programs: oracle/spirv_vm interprets shaders/{depth_reduce,entity_cull,
// meshlet_cull}.comp.spv and shaders/light_cluster/*.comp.spv, dispatched as draw_gen.rs / cluster.rs dispatch them, and
// tests/golden/spirv_reference.json holds their outputs (tests/golden/make_spirv_golden.py). This oracle is PINNED to
// those fixtures (tests/test_spirv_golden.py: Hi-Z texels and visibility words byte for byte, records / commands /
// light lists as sorted sets). It stays UNPINNED with respect to a real Vulkan driver in the three places SPIR-V leaves
// to the implementation (summation order of OpDot / OpMatrixTimes*, Log2, the sampler's texel footprint), where the
// interpreter makes the same choices as the contract below, and for the scene update (glam), which has no executable
// reference here. (The three shipped task shaders are interpreted as well and pin the payload output.) tests/golden/c1_and_clusters.json are outputs of THIS oracle.
//
// What each function follows (paths relative to the reference tree):
//   hiz_build           shaders/depth_reduce.comp:14-19, loop src/passes/draw_gen.rs:538-564,
//                       sampler ReduceMin src/graphics/device.rs:1404-1420
//   entity_cull         shaders/entity_cull.comp:104-245
//   meshlet_cull        shaders/meshlet_cull.comp:108-255
//   task payload        shaders/forward/forward_depth_prepass.task:224-256
//   mark_active         shaders/light_cluster/mark_active.comp:27-57, cluster_common.glsl:18-20
//   compact             shaders/light_cluster/active_cluster_compaction.comp:17-44
//   light_culling       shaders/light_cluster/light_culling.comp:34-151
//
// Arithmetic contract.  binary32 everywhere; every * + - individually rounded (built with
// -ffp-contract=off); fused multiply-add ONLY where the shipped SPIR-V has GLSL.std.450 Fma (read with
// oracle/tools/spv_dataflow.py); IEEE-correct / and sqrt; opaque SPIR-V ops evaluated in a fixed order:
//   dot(a,b)   = ((a0*b0 + a1*b1) + a2*b2) [+ a3*b3]
//   (M*v)[i]   = ((M[0][i]*v0 + M[1][i]*v1) + M[2][i]*v2) + M[3][i]*v3          (column-major M[col][row])
//   M*N        : column k = M * N[k]
//   length(v)  = sqrt(dot(v,v)); distance(a,b) = length(a-b)
//   FMax/FMin  = fmaxf/fminf (a NaN operand yields the other operand); FClamp(x,lo,hi) = fminf(fmaxf(x,lo),hi)
//   ConvertFToU: NaN -> 0, negative -> 0, >= 2^32 -> 0xFFFFFFFF, else truncate
//   1u << s    : 0 when s >= 32
//   log2       : orbit_log2f below (one deterministic implementation shared by oracle and kernels, restated
//                independently on each side) — used for the LOD index and the cluster depth slice
//   Hi-Z level : nearest-mip selection of log2(x) done exactly on exponent/mantissa (no log2 evaluation)
//   ReduceMin sample at a level of size (w,h): fx = u*w - 0.5; i0 = floor(fx); i1 = i0 + 1; both clamped to
//                [0,w-1]; same in y; result = min of the four texels.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/orbit_cuda.h"  // struct layouts only (OrbitCullInfo, OrbitSceneBuffers with HOST pointers, ...)

namespace {

// ---------------------------------------------------------------------------------------------------------
// leaf arithmetic
// ---------------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct M4 { float c[4][4]; };  // c[col][row]

inline float bits_f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t f_bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot2(float ax, float ay, float bx, float by) { return ax * bx + ay * by; }
inline float length3(V3 v) { return std::sqrt(dot3(v, v)); }

inline V4 mat_vec(const M4& m, V4 v) {
    V4 r;
    r.x = ((m.c[0][0] * v.x + m.c[1][0] * v.y) + m.c[2][0] * v.z) + m.c[3][0] * v.w;
    r.y = ((m.c[0][1] * v.x + m.c[1][1] * v.y) + m.c[2][1] * v.z) + m.c[3][1] * v.w;
    r.z = ((m.c[0][2] * v.x + m.c[1][2] * v.y) + m.c[2][2] * v.z) + m.c[3][2] * v.w;
    r.w = ((m.c[0][3] * v.x + m.c[1][3] * v.y) + m.c[2][3] * v.z) + m.c[3][3] * v.w;
    return r;
}

inline M4 mat_mat(const M4& a, const M4& b) {
    M4 r;
    for (int k = 0; k < 4; ++k) {
        V4 col = mat_vec(a, V4{b.c[k][0], b.c[k][1], b.c[k][2], b.c[k][3]});
        r.c[k][0] = col.x; r.c[k][1] = col.y; r.c[k][2] = col.z; r.c[k][3] = col.w;
    }
    return r;
}

inline uint32_t f2u(float f) {  // ConvertFToU, pinned
    if (!(f > 0.0f)) return 0u;  // NaN, negative, zero
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;
}
inline uint32_t shl1(uint32_t s) { return s >= 32u ? 0u : (1u << s); }

// Deterministic log2 (contract function). x = m*2^e with m in [sqrt(1/2), sqrt(2)); t = (m-1)/(m+1);
// log2(m) = t * P(t^2), P = (2/ln2)(1 + s/3 + s^2/5 + s^3/7 + s^4/9), evaluated by Horner with explicit fma;
// result = float(e) + t*P.
inline float orbit_log2f(float x) {
    uint32_t u = f_bits(x);
    if ((u << 1) == 0u) return -INFINITY;           // +-0
    if (u >> 31) return NAN;                        // negative (incl. -inf, -NaN)
    if (u >= 0x7F800000u) return x;                 // +inf, NaN
    int e = 0;
    if (u < 0x00800000u) { x = x * 8388608.0f; u = f_bits(x); e = -23; }  // subnormal: scale by 2^23 (exact)
    e += (int)(u >> 23) - 127;
    uint32_t mant = (u & 0x007FFFFFu) | 0x3F800000u;
    float m = bits_f(mant);                         // [1,2)
    if (m > 1.41421354f) { m = m * 0.5f; e += 1; }  // [~0.7071, ~1.4142]
    float t = (m - 1.0f) / (m + 1.0f);
    float s = t * t;
    float p = 0.3205986261f;                        // (2/ln2)/9
    p = std::fmaf(s, p, 0.4121982336f);             // (2/ln2)/7
    p = std::fmaf(s, p, 0.5770775080f);             // (2/ln2)/5
    p = std::fmaf(s, p, 0.9617958665f);             // (2/ln2)/3
    p = std::fmaf(s, p, 2.8853900433f);             // (2/ln2)
    return (float)e + t * p;
}

// Nearest-mip level for lod = log2(x), clamped to [0, levels-1]; exact (no log2 evaluation).
inline uint32_t hiz_level(float x, uint32_t levels, uint64_t* near_ties) {
    if (!(x > 0.0f)) return 0u;                      // log2(<=0) = -inf / NaN -> clamps to level 0
    uint32_t u = f_bits(x);
    if (u >= 0x7F800000u) return levels - 1u;        // +inf
    int e;
    if (u < 0x00800000u) return 0u;                  // subnormal: far below 2^-0.5
    e = (int)(u >> 23) - 127;
    float m = bits_f((u & 0x007FFFFFu) | 0x3F800000u);
    if (near_ties && std::fabs(m - 1.41421354f) <= 1e-5f * 1.41421354f) ++*near_ties;
    int k = e + (m > 1.41421354f ? 1 : 0);           // k - 1/2 <= log2 x < k + 1/2
    if (k < 0) k = 0;
    if (k > (int)levels - 1) k = (int)levels - 1;
    return (uint32_t)k;
}

struct HizView {
    const float* texels;
    OrbitHizInfo g;
};

inline void footprint(float u, uint32_t w, int& i0, int& i1) {
    float fx = u * (float)w - 0.5f;
    float f = std::floor(fx);
    int a;
    if (!(f >= 0.0f)) a = -1; else if (f >= (float)w) a = (int)w; else a = (int)f;
    int b = a + 1;
    int hi = (int)w - 1;
    i0 = a < 0 ? 0 : (a > hi ? hi : a);
    i1 = b < 0 ? 0 : (b > hi ? hi : b);
}

inline float sample_reduce_min(const float* level, uint32_t w, uint32_t h, float u, float v) {
    int x0, x1, y0, y1;
    footprint(u, w, x0, x1);
    footprint(v, h, y0, y1);
    float a = level[(size_t)y0 * w + x0], b = level[(size_t)y0 * w + x1];
    float c = level[(size_t)y1 * w + x0], d = level[(size_t)y1 * w + x1];
    return std::fmin(std::fmin(a, b), std::fmin(c, d));
}

void hiz_geometry(uint32_t dw, uint32_t dh, OrbitHizInfo* g) {
    auto npot = [](uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; };  // u32::next_power_of_two
    std::memset(g, 0, sizeof(*g));
    g->width = npot(dw) / 2; g->height = npot(dh) / 2;         // draw_gen.rs:458
    uint32_t mx = std::max(g->width, g->height);
    uint32_t levels = 1; { uint32_t t = mx; levels = 0; while (t) { ++levels; t >>= 1; } if (!levels) levels = 1; }
    g->levels = levels;                                        // math.rs:18-20: floor(log2(max))+1
    uint32_t off = 0;
    for (uint32_t l = 0; l < levels && l < ORBIT_HIZ_MAX_LEVELS; ++l) {
        g->level_offset[l] = off;
        off += std::max(g->width >> l, 1u) * std::max(g->height >> l, 1u);   // image.rs:531
    }
    g->total_texels = off;
}

struct Margins {  // predicates whose operands are within 1e-5 relative of each other (north-star epsilon rule)
    uint64_t plane = 0, cone = 0, cullable = 0, depth = 0, hiz_level = 0, lod = 0;
};
inline bool near(float a, float b) {
    float m = std::fmax(std::fabs(a), std::fabs(b));
    return std::fabs(a - b) <= 1e-5f * m;
}

struct Sphere { V3 c; float r; float r_model; float s; };

inline Sphere transform_sphere(const M4& m, const float* sph) {
    V4 p = mat_vec(m, V4{sph[0], sph[1], sph[2], 1.0f});
    Sphere o;
    o.c = V3{p.x / p.w, p.y / p.w, p.z / p.w};
    V3 X{m.c[0][0], m.c[0][1], m.c[0][2]}, Y{m.c[1][0], m.c[1][1], m.c[1][2]}, Z{m.c[2][0], m.c[2][1], m.c[2][2]};
    float s2 = std::fmax(dot3(X, X), std::fmax(dot3(Y, Y), dot3(Z, Z)));
    o.s = std::sqrt(s2);
    o.r_model = sph[3];
    o.r = o.r_model * o.s;
    return o;
}

inline bool frustum_test(const OrbitCullInfo& ci, const Sphere& s, Margins& mg) {
    bool visible = true;
    for (uint32_t i = 0; i < ci.cull_plane_count; ++i) {
        const float* pl = ci.cull_planes[i];
        float d = dot3(V3{pl[0], pl[1], pl[2]}, s.c) + pl[3];
        float nr = -s.r;
        if (near(d, nr)) ++mg.plane;
        visible = visible && (d > nr);
    }
    return visible;
}

// Occlusion block shared by both cull shaders (entity_cull.comp:146-191, meshlet_cull.comp:160-205).
// NOTE: in the perspective case c.z is negated in place and stays negated for the caller.
inline bool occlusion_test(const OrbitCullInfo& ci, Sphere& s, const HizView& hz, Margins& mg) {
    bool cullable = true;
    float ax, ay, az, aw, depth;
    if (ci.projection_type == 0u) {
        float zp = -s.c.z;
        s.c.z = zp;
        float thr = std::fmaf(s.r_model, s.s, ci.z_near);
        if (near(zp, thr)) ++mg.cullable;
        cullable = zp >= thr;
        float P00 = ci.p00_or_width_recip_x2, P11 = ci.p11_or_height_recip_x2;
        float nr = -s.r;
        float cx0 = -s.c.x, cx1 = -zp;
        float sx = std::sqrt(std::fmaf(nr, s.r, dot2(cx0, cx1, cx0, cx1)));
        float minx0 = sx * cx0 + nr * cx1, minx1 = s.r * cx0 + sx * cx1;
        float maxx0 = sx * cx0 + s.r * cx1, maxx1 = nr * cx0 + sx * cx1;
        float cy0 = -s.c.y, cy1 = -zp;
        float sy = std::sqrt(std::fmaf(nr, s.r, dot2(cy0, cy1, cy0, cy1)));
        float miny0 = sy * cy0 + nr * cy1, miny1 = s.r * cy0 + sy * cy1;
        float maxy0 = sy * cy0 + s.r * cy1, maxy1 = nr * cy0 + sy * cy1;
        float a0 = minx0 / minx1 * P00, a1 = miny0 / miny1 * P11, a2 = maxx0 / maxx1 * P00, a3 = maxy0 / maxy1 * P11;
        // aabb = a.xwzy * (.5,-.5,.5,-.5) + .5
        ax = std::fmaf(a0, 0.5f, 0.5f); ay = std::fmaf(a3, -0.5f, 0.5f);
        az = std::fmaf(a2, 0.5f, 0.5f); aw = std::fmaf(a1, -0.5f, 0.5f);
        depth = ci.z_near / std::fmaf(-s.r_model, s.s, zp);
    } else if (ci.projection_type == 1u) {
        float sr = ci.p00_or_width_recip_x2;  // both components from p00 (entity_cull.comp:166)
        float ctrx = s.c.x * sr, ctry = s.c.y * sr;
        float box = sr * s.r;
        float b0 = std::fmaf(box, -1.0f, ctrx), b1 = std::fmaf(box, -1.0f, ctry);
        float b2 = std::fmaf(box, 1.0f, ctrx), b3 = std::fmaf(box, 1.0f, ctry);
        auto cl = [](float v) { return std::fmin(std::fmax(v, -1.0f), 1.0f); };
        ax = std::fmaf(cl(b0), 0.5f, 0.5f); ay = std::fmaf(cl(b1), -0.5f, 0.5f);
        az = std::fmaf(cl(b2), 0.5f, 0.5f); aw = std::fmaf(cl(b3), -0.5f, 0.5f);
        float k = 1.0f / (ci.z_far - ci.z_near);
        depth = k * (std::fmaf(s.r_model, s.s, s.c.z) + ci.z_far);
    } else {
        return true;  // switch without matching case: aabb undefined in the reference; never produced by callers
    }
    if (!cullable) return true;
    float W = (az - ax) * (float)hz.g.width;
    float H = (aw - ay) * (float)hz.g.height;
    float u = (ax + az) * 0.5f, v = (ay + aw) * 0.5f;
    uint32_t lvl = hiz_level(std::fmax(W, H), hz.g.levels, &mg.hiz_level);
    uint32_t lw = std::max(hz.g.width >> lvl, 1u), lh = std::max(hz.g.height >> lvl, 1u);
    float sampled = sample_reduce_min(hz.texels + hz.g.level_offset[lvl], lw, lh, u, v);
    if (near(depth, sampled)) ++mg.depth;
    return depth >= sampled;
}

inline M4 load_m4(const void* p) { M4 m; std::memcpy(&m, p, 64); return m; }

}  // namespace

extern "C" {

typedef struct OracleStats {
    uint64_t near_plane, near_cone, near_cullable, near_depth, near_hiz_level, near_lod;
    uint64_t lanes;        // active lanes processed by the meshlet stage
    uint64_t records;      // dispatch records written / read
    uint64_t survivors;    // draw commands written
    uint64_t visible;      // lanes with visible == true
} OracleStats;

int oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Number of host threads the parallel loops use from now on (bench.py reports a single-thread figure beside the
// all-cores one). Returns the previous value.
int oracle_set_threads(int n) {
#ifdef _OPENMP
    const int prev = omp_get_max_threads();
    if (n > 0) omp_set_num_threads(n);
    return prev;
#else
    (void)n;
    return 1;
#endif
}

void oracle_hiz_geometry(uint32_t dw, uint32_t dh, OrbitHizInfo* out) { hiz_geometry(dw, dh, out); }

float oracle_log2f(float x) { return orbit_log2f(x); }
uint32_t oracle_hiz_level(float x, uint32_t levels) { return hiz_level(x, levels, nullptr); }

// Leaf predicates exposed for unit tests (tests/test_oracle_leaf.py).
// project_sphere + sphere_closest_depth (entity_cull.comp:83-102,152-163). c = view-space centre as the
// shader sees it BEFORE the in-place negation of z. out = {aabb.x, aabb.y, aabb.z, aabb.w, depth, cullable}.
void oracle_project_sphere(float cx, float cy, float cz, float r_model, float scale, float z_near, float p00, float p11, float* out) {
    float zp = -cz;
    float r = r_model * scale, nr = -r;
    out[5] = (zp >= std::fmaf(r_model, scale, z_near)) ? 1.0f : 0.0f;
    float c0 = -cx, c1 = -zp;
    float sx = std::sqrt(std::fmaf(nr, r, dot2(c0, c1, c0, c1)));
    float minx0 = sx * c0 + nr * c1, minx1 = r * c0 + sx * c1, maxx0 = sx * c0 + r * c1, maxx1 = nr * c0 + sx * c1;
    float d0 = -cy;
    float sy = std::sqrt(std::fmaf(nr, r, dot2(d0, c1, d0, c1)));
    float miny0 = sy * d0 + nr * c1, miny1 = r * d0 + sy * c1, maxy0 = sy * d0 + r * c1, maxy1 = nr * d0 + sy * c1;
    float a0 = minx0 / minx1 * p00, a1 = miny0 / miny1 * p11, a2 = maxx0 / maxx1 * p00, a3 = maxy0 / maxy1 * p11;
    out[0] = std::fmaf(a0, 0.5f, 0.5f); out[1] = std::fmaf(a3, -0.5f, 0.5f);
    out[2] = std::fmaf(a2, 0.5f, 0.5f); out[3] = std::fmaf(a1, -0.5f, 0.5f);
    out[4] = z_near / std::fmaf(-r_model, scale, zp);
}

// coneCull (meshlet_cull.comp:104-106) in the perspective case: returns 1 when the meshlet is back-facing.
int oracle_cone_cull(float cx, float cy, float cz, float r, float ax, float ay, float az, float cutoff) {
    V3 c{cx, cy, cz}, a{ax, ay, az};
    return dot3(c, a) >= std::fmaf(cutoff, length3(c), r) ? 1 : 0;
}

// depth_reduce.comp applied level by level (draw_gen.rs:538-564).
void oracle_hiz_build(const float* depth, uint32_t dw, uint32_t dh, float* texels) {
    OrbitHizInfo g; hiz_geometry(dw, dh, &g);
    for (uint32_t l = 0; l < g.levels; ++l) {
        const float* src = l == 0 ? depth : texels + g.level_offset[l - 1];
        uint32_t sw = l == 0 ? dw : std::max(g.width >> (l - 1), 1u);
        uint32_t sh = l == 0 ? dh : std::max(g.height >> (l - 1), 1u);
        uint32_t w = std::max(g.width >> l, 1u), h = std::max(g.height >> l, 1u);
        float* dst = texels + g.level_offset[l];
#pragma omp parallel for schedule(static) if (w * h > 4096)
        for (int64_t y = 0; y < (int64_t)h; ++y)
            for (uint32_t x = 0; x < w; ++x) {
                float u = ((float)x + 0.5f) / (float)w;
                float v = ((float)y + 0.5f) / (float)h;
                dst[(size_t)y * w + x] = sample_reduce_min(src, sw, sh, u, v);
            }
    }
}

// Sample helper exposed for unit tests of the ReduceMin rule.
float oracle_hiz_sample(const float* texels, uint32_t dw, uint32_t dh, float u, float v, float lod_arg) {
    OrbitHizInfo g; hiz_geometry(dw, dh, &g);
    uint32_t lvl = hiz_level(lod_arg, g.levels, nullptr);
    return sample_reduce_min(texels + g.level_offset[lvl], std::max(g.width >> lvl, 1u), std::max(g.height >> lvl, 1u), u, v);
}

// entity_cull.comp main(). Canonical output order: ascending entity-draw index, chunks ascending.
// Returns the number of dispatch records produced (also stored in the header even when > capacity).
uint64_t oracle_entity_cull(const OrbitCullInfo* cull, const OrbitSceneBuffers* scene, const float* hiz_texels,
                            uint32_t depth_w, uint32_t depth_h, void* dispatch_buffer, uint64_t capacity_records,
                            OracleStats* stats) {
    const OrbitCullInfo& ci = *cull;
    HizView hz{hiz_texels, {}};
    if (hiz_texels) hiz_geometry(depth_w, depth_h, &hz.g);
    const uint8_t* draws_base = (const uint8_t*)scene->entity_draws;
    uint32_t count; std::memcpy(&count, draws_base, 4);
    const OrbitEntityDraw* draws = (const OrbitEntityDraw*)(draws_base + ORBIT_ENTITY_DRAW_HEADER_BYTES);
    const OrbitMeshInfo* mesh_infos = (const OrbitMeshInfo*)scene->mesh_infos;
    const OrbitEntityData* entities = (const OrbitEntityData*)scene->entities;
    uint32_t begin = scene->draw_begin, end = scene->draw_end;
    if (begin == 0 && end == 0) end = count;
    end = std::min(end, count);
    const uint32_t pass = ci.occlusion_pass;
    const bool mocc = ci.meshlet_visibility_buffer != ORBIT_NO_BUFFER;
    const M4 view = load_m4(&ci.view_matrix);

    struct Emit { uint32_t entity_index, lod_offset, lod_count, vis_offset; };
    const uint32_t n = end > begin ? end - begin : 0;
    std::vector<Emit> emits(n);
    std::vector<uint8_t> vis_flags(n), draw_flags(n);
    Margins total;
#pragma omp parallel
    {
        Margins mg;
#pragma omp for schedule(static)
        for (int64_t ii = 0; ii < (int64_t)n; ++ii) {
            uint32_t gid = begin + (uint32_t)ii;
            OrbitEntityDraw d = draws[gid];
            const OrbitMeshInfo& mi = mesh_infos[d.mesh_index];
            bool visible = true, vib = true;
            if (pass == 1u || pass == 2u) vib = (scene->entity_visibility[gid / 32u] & (1u << (gid % 32u))) != 0u;
            if (pass == 1u) visible = vib;
            M4 mv = mat_mat(view, load_m4(&entities[d.entity_index].model_matrix));
            Sphere s = transform_sphere(mv, mi.bounding_sphere);
            if (visible) visible = frustum_test(ci, s, mg);
            if (pass == 2u && visible) visible = occlusion_test(ci, s, hz, mg);
            bool should_draw = visible;
            if (pass == 2u) should_draw = visible && (!vib || mocc);
            vis_flags[ii] = visible; draw_flags[ii] = should_draw;
            if (should_draw) {
                V3 t{ci.lod_target_pos_view_space[0], ci.lod_target_pos_view_space[1], ci.lod_target_pos_view_space[2]};
                V3 dv{t.x - s.c.x, t.y - s.c.y, t.z - s.c.z};
                float lod_distance = length3(dv) - s.r;
                float f = orbit_log2f(std::fmax(lod_distance, 0.0f) / ci.lod_base) / orbit_log2f(ci.lod_step);
                float g = std::fmax(f + 1.0f, 0.0f);
                if (g > 0.0f && near(g, std::floor(g + 0.5f)) && g < 64.0f) ++mg.lod;
                uint32_t lod = f2u(g);
                lod = std::min(std::max(lod, ci.min_mesh_lod), ci.max_mesh_lod);  // UClamp
                OrbitMeshLod L = mi.mesh_lods[std::min(lod, mi.lod_count - 1u) & 7u];
                emits[ii] = Emit{d.entity_index, L.meshlet_offset, L.meshlet_count, d.visibility_offset};
            }
        }
#pragma omp critical
        { total.plane += mg.plane; total.cone += mg.cone; total.cullable += mg.cullable; total.depth += mg.depth;
          total.hiz_level += mg.hiz_level; total.lod += mg.lod; }
    }
    // emission in canonical order + visibility writeback (ballot over 32 consecutive draws)
    uint8_t* out = (uint8_t*)dispatch_buffer;
    OrbitMeshletDispatch* recs = (OrbitMeshletDispatch*)(out + ORBIT_DISPATCH_HEADER_BYTES);
    uint64_t nrec = 0;
    for (uint32_t ii = 0; ii < n; ++ii) {
        if (!draw_flags[ii]) continue;
        const Emit& e = emits[ii];
        uint32_t chunks = (e.lod_count + 31u) / 32u;
        uint32_t vo = e.vis_offset;
        for (uint32_t k = 0; k < chunks; ++k) {
            OrbitMeshletDispatch r{e.entity_index, e.lod_offset + 32u * k, std::min(e.lod_count - 32u * k, 32u), vo};
            if (nrec < capacity_records) std::memcpy(&recs[nrec], &r, 16);
            ++nrec;
            vo += r.meshlet_count / 32u;
        }
    }
    uint32_t hdr[3] = {(uint32_t)nrec, 1u, 1u};   // fill_buffer {0,1,1} + atomicAdd (draw_gen.rs:356-363)
    std::memcpy(out, hdr, 12);
    if (pass == 2u) {
        for (uint32_t w0 = 0; w0 < n; w0 += 32u) {
            uint32_t word = 0;
            for (uint32_t b = 0; b < 32u && w0 + b < n; ++b) word |= (uint32_t)vis_flags[w0 + b] << b;
            scene->entity_visibility[(begin + w0) / 32u] = word;
        }
    }
    if (stats) {
        stats->near_plane += total.plane; stats->near_cullable += total.cullable; stats->near_depth += total.depth;
        stats->near_hiz_level += total.hiz_level; stats->near_lod += total.lod; stats->records += nrec;
    }
    return nrec;
}

// meshlet_cull.comp main(), one 32-lane group per dispatch record. Canonical output order: (record, lane).
// task_payloads (nullable): per record {u32 task_count; MeshTaskPayload} as the task-shader twins would emit.
uint64_t oracle_meshlet_cull(const OrbitCullInfo* cull, const OrbitSceneBuffers* scene, const float* hiz_texels,
                             uint32_t depth_w, uint32_t depth_h, const void* dispatch_buffer, void* draw_buffer,
                             uint64_t capacity_draws, void* task_payloads, OracleStats* stats) {
    const OrbitCullInfo& ci = *cull;
    HizView hz{hiz_texels, {}};
    if (hiz_texels) hiz_geometry(depth_w, depth_h, &hz.g);
    const uint8_t* in = (const uint8_t*)dispatch_buffer;
    uint32_t nrec; std::memcpy(&nrec, in, 4);
    const OrbitMeshletDispatch* recs = (const OrbitMeshletDispatch*)(in + ORBIT_DISPATCH_HEADER_BYTES);
    const OrbitMeshlet* meshlets = (const OrbitMeshlet*)scene->meshlets;
    const OrbitEntityData* entities = (const OrbitEntityData*)scene->entities;
    const uint8_t* materials = (const uint8_t*)scene->materials;
    const uint32_t pass = ci.occlusion_pass;
    const bool mocc = ci.meshlet_visibility_buffer != ORBIT_NO_BUFFER;
    const bool use_vis = (pass == 1u || pass == 2u) && mocc;
    const M4 view = load_m4(&ci.view_matrix);
    const float K = 0.007874015718698502f;  // SPIR-V constant 0x3c010204: x/127.0 became x*fl(1/127)

    std::vector<uint32_t> draw_mask(nrec), vis_mask(nrec);
    Margins total;
    uint64_t lanes_total = 0, visible_total = 0;
#pragma omp parallel
    {
        Margins mg; uint64_t lanes = 0, vis_n = 0;
#pragma omp for schedule(static)
        for (int64_t ri = 0; ri < (int64_t)nrec; ++ri) {
            OrbitMeshletDispatch rec = recs[ri];
            M4 mv = mat_mat(view, load_m4(&entities[rec.entity_index].model_matrix));
            uint32_t word = use_vis ? scene->meshlet_visibility[rec.visibility_offset] : 0u;
            uint32_t dm = 0, vm = 0;
            for (uint32_t lane = 0; lane < 32u && lane < rec.meshlet_count; ++lane) {
                ++lanes;
                const OrbitMeshlet& m = meshlets[rec.meshlet_offset + lane];
                Sphere s = transform_sphere(mv, m.bounding_sphere);
                V4 a4 = mat_vec(mv, V4{(float)m.cone_axis[0] * K, (float)m.cone_axis[1] * K, (float)m.cone_axis[2] * K, 0.0f});
                V3 axis{a4.x, a4.y, a4.z};
                float cutoff = (float)m.cone_cutoff * K;
                uint32_t alpha; std::memcpy(&alpha, materials + (size_t)m.material_index * ORBIT_MATERIAL_STRIDE_BYTES + ORBIT_MATERIAL_ALPHA_MODE_OFFSET, 4);
                bool visible = true, vib = true;
                if (use_vis) vib = (word & (1u << lane)) != 0u;   // lane/32 == 0, lane%32 == lane
                if (pass == 1u) visible = vib;
                if (visible) visible = frustum_test(ci, s, mg);
                if (visible) {
                    if (ci.projection_type == 0u) {
                        float lhs = dot3(s.c, axis), rhs = std::fmaf(cutoff, length3(s.c), s.r);
                        if (near(lhs, rhs)) ++mg.cone;
                        visible = !(lhs >= rhs);
                    } else if (ci.projection_type == 1u) {
                        V3 cam{s.c.x - 0.0f, s.c.y - 0.0f, s.c.z - (-1.0f)};
                        V3 q{s.c.x - cam.x, s.c.y - cam.y, s.c.z - cam.z};
                        float lhs = dot3(q, axis), rhs = std::fmaf(cutoff, length3(q), s.r);
                        if (near(lhs, rhs)) ++mg.cone;
                        visible = !(lhs >= rhs);
                    }
                }
                if (mocc && pass == 2u && visible) visible = occlusion_test(ci, s, hz, mg);
                bool should_draw = visible && (shl1(alpha) & ci.alpha_mode_flags) != 0u;
                if (pass == 2u && mocc && (shl1(alpha) & ci.noskip_alpha_mode) == 0u) should_draw = visible && !vib;
                dm |= (uint32_t)should_draw << lane;
                vm |= (uint32_t)visible << lane;
                vis_n += visible;
            }
            draw_mask[ri] = dm; vis_mask[ri] = vm;
        }
#pragma omp critical
        { total.plane += mg.plane; total.cone += mg.cone; total.cullable += mg.cullable; total.depth += mg.depth;
          total.hiz_level += mg.hiz_level; lanes_total += lanes; visible_total += vis_n; }
    }
    // canonical emission
    uint8_t* out = (uint8_t*)draw_buffer;
    uint64_t ndraw = 0;
    for (uint32_t ri = 0; ri < nrec; ++ri) {
        OrbitMeshletDispatch rec = recs[ri];
        uint32_t dm = draw_mask[ri];
        uint32_t tcount = 0;
        uint8_t* tp = task_payloads ? (uint8_t*)task_payloads + (size_t)ri * 44u : nullptr;
        if (tp) { std::memset(tp, 0, 44); std::memcpy(tp + 4, &rec.entity_index, 4); std::memcpy(tp + 8, &rec.meshlet_offset, 4); }
        while (dm) {
            uint32_t lane = (uint32_t)__builtin_ctz(dm); dm &= dm - 1u;
            const OrbitMeshlet& m = meshlets[rec.meshlet_offset + lane];
            OrbitMeshletDrawCommand c;
            c.cmd_index_count = (uint32_t)m.triangle_count * 3u;
            c.cmd_instance_count = 1u;
            c.cmd_first_index = (m.data_offset + (uint32_t)m.vertex_count) * 4u;
            c.cmd_vertex_offset = (int32_t)m.data_offset;
            c.cmd_first_instance = rec.entity_index;
            c.meshlet_vertex_offset = m.vertex_offset;
            c.meshlet_index = rec.meshlet_offset + lane;
            if (ndraw < capacity_draws) std::memcpy(out + ORBIT_DRAW_HEADER_BYTES + ndraw * 28u, &c, 28);
            ++ndraw;
            if (tp) tp[12 + tcount] = (uint8_t)lane;
            ++tcount;
        }
        if (tp) std::memcpy(tp, &tcount, 4);
    }
    uint32_t cnt = (uint32_t)ndraw;
    std::memcpy(out, &cnt, 4);
    if (pass == 2u && mocc)
        for (uint32_t ri = 0; ri < nrec; ++ri) scene->meshlet_visibility[recs[ri].visibility_offset] = vis_mask[ri];
    if (stats) {
        stats->near_plane += total.plane; stats->near_cone += total.cone; stats->near_cullable += total.cullable;
        stats->near_depth += total.depth; stats->near_hiz_level += total.hiz_level;
        stats->lanes += lanes_total; stats->records += nrec; stats->survivors += ndraw; stats->visible += visible_total;
    }
    return ndraw;
}

// mark_active.comp (sample_count == 1 path) over a W x H depth buffer.
void oracle_mark_active(const OrbitClusterParams* p, const float* depth, uint32_t* tile_masks, OrbitClusterDepthBounds* bounds) {
    const OrbitClusterCullInfo& ci = p->info;
    const uint32_t cx = ci.cluster_count[0], cy = ci.cluster_count[1], cz = ci.cluster_count[2];
    std::memset(tile_masks, 0, sizeof(uint32_t) * (size_t)cx * cy);                 // cluster.rs:447-448 fill 0
    std::memset(bounds, 0, sizeof(OrbitClusterDepthBounds) * (size_t)cx * cy * cz); // cluster.rs:449-450
    const uint32_t W = ci.screen_size[0], H = ci.screen_size[1];
    for (uint32_t y = 0; y < H; ++y)
        for (uint32_t x = 0; x < W; ++x) {
            float d = depth[(size_t)y * W + x];
            uint32_t tx = x / ci.tile_size_px, ty = y / ci.tile_size_px;
            float z = ci.z_near / d;
            uint32_t slice = f2u(std::fmaf(orbit_log2f(z), p->z_scale, p->z_bias));
            uint32_t mask = shl1(slice);
            if (slice < cz) {
                size_t idx = (size_t)tx + (size_t)ty * cx + (size_t)slice * cx * cy;
                bounds[idx].min_depth = std::max(bounds[idx].min_depth, f_bits(1.0f - d));
                bounds[idx].max_depth = std::max(bounds[idx].max_depth, f_bits(d));
            }
            if (mask > 0u) tile_masks[tx + ty * cx] |= mask;
        }
}

// active_cluster_compaction.comp; canonical order = ascending cluster index.
uint32_t oracle_compact_clusters(const OrbitClusterParams* p, const uint32_t* tile_masks, void* unique_clusters) {
    const OrbitClusterCullInfo& ci = p->info;
    const uint32_t cx = ci.cluster_count[0], cy = ci.cluster_count[1], cz = ci.cluster_count[2];
    uint32_t* out = (uint32_t*)unique_clusters;
    uint32_t n = 0;
    for (uint32_t z = 0; z < cz; ++z)
        for (uint32_t y = 0; y < cy; ++y)
            for (uint32_t x = 0; x < cx; ++x)
                if (tile_masks[x + y * cx] & shl1(z)) out[4 + n++] = x + y * cx + z * cx * cy;
    out[0] = (n + 255u) / 256u; out[1] = 1u; out[2] = 1u; out[3] = n;
    return n;
}

// light_culling.comp; canonical packing = clusters in compacted-list order, each list ascending.
uint64_t oracle_light_culling(const OrbitClusterParams* p, const void* lights_, const OrbitClusterDepthBounds* bounds,
                              const void* unique_clusters, uint32_t* offset_count_image, void* light_index_list,
                              uint64_t capacity_indices) {
    const OrbitClusterCullInfo& ci = p->info;
    const OrbitLightData* lights = (const OrbitLightData*)lights_;
    const uint32_t cx = ci.cluster_count[0], cy = ci.cluster_count[1], cz = ci.cluster_count[2];
    const uint32_t* uc = (const uint32_t*)unique_clusters;
    const uint32_t nactive = uc[3];
    const M4 s2v = load_m4(&ci.screen_to_view_matrix), w2v = load_m4(&ci.world_to_view_matrix);
    const uint32_t L = ci.global_light_count;
    std::memset(offset_count_image, 0, 8 * (size_t)cx * cy * cz);
    // light positions in view space (recomputed per (cluster,light) in the shader; same value every time)
    std::vector<V4> lv(L);
    for (uint32_t j = 0; j < L; ++j) {
        V4 c = mat_vec(w2v, V4{lights[j].position[0], lights[j].position[1], lights[j].position[2], 1.0f});
        lv[j] = V4{c.x, c.y, c.z, lights[j].outer_radius};
    }
    std::vector<std::vector<uint32_t>> lists(nactive);
    const float sx = (float)ci.screen_size[0], sy = (float)ci.screen_size[1];
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t a = 0; a < (int64_t)nactive; ++a) {
        uint32_t idx = uc[4 + a];
        uint32_t z = idx / (cx * cy); uint32_t rem = idx - z * cx * cy; uint32_t y = rem / cx; uint32_t x = rem - y * cx;
        float minx = (float)(x * ci.tile_size_px), miny = (float)(y * ci.tile_size_px);
        float maxx = std::fmin(minx + (float)ci.tile_size_px, sx), maxy = std::fmin(miny + (float)ci.tile_size_px, sy);
        auto unproject = [&](float px, float py) {
            float tx = px / sx, ty = py / sy;
            float clx = tx * 2.0f - 1.0f, cly = (1.0f - ty) * 2.0f - 1.0f;
            V4 v = mat_vec(s2v, V4{clx, cly, 1.0f, 1.0f});
            return V3{v.x / v.w, v.y / v.w, v.z / v.w};
        };
        V3 vmin = unproject(minx, miny), vmax = unproject(maxx, maxy);
        float min_d = 1.0f - bits_f(bounds[idx].min_depth);
        float max_d = bits_f(bounds[idx].max_depth);
        float cnear = ci.z_near / max_d, cfar = ci.z_near / min_d;
        auto pt = [](V3 v, float zd) {
            float dn = (0.0f * v.x + 0.0f * v.y) + (-1.0f) * v.z;   // dot((0,0,-1), v)
            float t = zd / dn;
            return V3{v.x * t, v.y * t, v.z * t};
        };
        V3 p0 = pt(vmin, cnear), p1 = pt(vmin, cfar), p2 = pt(vmax, cnear), p3 = pt(vmax, cfar);
        auto mn = [](float a, float b, float c, float d) { return std::fmin(std::fmin(a, b), std::fmin(c, d)); };
        auto mx = [](float a, float b, float c, float d) { return std::fmax(std::fmax(a, b), std::fmax(c, d)); };
        float lo[3] = {mn(p0.x, p1.x, p2.x, p3.x), mn(p0.y, p1.y, p2.y, p3.y), mn(p0.z, p1.z, p2.z, p3.z)};
        float hi[3] = {mx(p0.x, p1.x, p2.x, p3.x), mx(p0.y, p1.y, p2.y, p3.y), mx(p0.z, p1.z, p2.z, p3.z)};
        std::vector<uint32_t>& out = lists[a];
        for (uint32_t j = 0; j < L && out.size() < ORBIT_MAX_LIGHTS_PER_CLUSTER; ++j) {
            bool hit = true;
            if (lights[j].light_type == ORBIT_LIGHT_POINT) {
                float c[3] = {lv[j].x, lv[j].y, lv[j].z};
                float acc = 0.0f;
                for (int k = 0; k < 3; ++k) {
                    float v = c[k];
                    if (v < lo[k]) acc = std::fmaf(lo[k] - v, lo[k] - v, acc);
                    if (v > hi[k]) acc = std::fmaf(v - hi[k], v - hi[k], acc);
                }
                hit = acc <= lv[j].w * lv[j].w;
            }
            if (hit) out.push_back(j);
        }
    }
    uint8_t* base = (uint8_t*)light_index_list;
    uint32_t* indices = (uint32_t*)(base + ORBIT_LIGHT_INDEX_HEADER_BYTES);
    uint64_t total = 0;
    for (uint32_t a = 0; a < nactive; ++a) {
        uint32_t idx = uc[4 + a];
        offset_count_image[2 * (size_t)idx + 0] = (uint32_t)total;
        offset_count_image[2 * (size_t)idx + 1] = (uint32_t)lists[a].size();
        for (uint32_t v : lists[a]) { if (total < capacity_indices) indices[total] = v; ++total; }
    }
    uint32_t t32 = (uint32_t)total;
    std::memcpy(base, &t32, 4);
    return total;
}

// ---- SceneData::update_scene (src/scene.rs:404-492) --------------------------------------------------------
// glam 0.24 (Cargo.toml:24; NOT vendored in /root/reference — restated from its published source, x86_64/SSE2
// build, no FMA contraction):
//   Mat4::from_scale_rotation_translation = quat_to_axes (x2=x+x .. wz=w*z2; x_axis = (1-(yy+zz), xy+wz, xz-wy, 0) ..)
//     with each axis multiplied by the matching scale component, w_axis = (translation, 1);
//   Mat4::inverse = the GLM cofactor scheme: 18 2x2 sub-determinants a*b - c*d, inv_k = (v*f - v*f) + v*f with the
//     checkerboard sign, determinant = dot(x_axis, first row of the adjugate) summed as (x+z)+(y+w) (SSE2 dot4),
//     rcp = 1/det, every element multiplied by rcp.
// PARITY UNPINNED: glam itself cannot be run here; the model matrix is plain enough to be safe, the rounding of
// the normal matrix (never read by the culling path) depends on the summation order assumed above.
static void glam_from_srt(const float* pos, const float* q, const float* scl, float m[16]) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2;
    const float wx = w * x2, wy = w * y2, wz = w * z2;
    const float ax[3] = {1.0f - (yy + zz), xy + wz, xz - wy};
    const float ay[3] = {xy - wz, 1.0f - (xx + zz), yz + wx};
    const float az[3] = {xz + wy, yz - wx, 1.0f - (xx + yy)};
    for (int i = 0; i < 3; ++i) { m[i] = ax[i] * scl[0]; m[4 + i] = ay[i] * scl[1]; m[8 + i] = az[i] * scl[2]; m[12 + i] = pos[i]; }
    m[3] = 0.0f * scl[0]; m[7] = 0.0f * scl[1]; m[11] = 0.0f * scl[2]; m[15] = 1.0f;   // Vec4 * f32 multiplies the 0 lane too
}

static void glam_inverse(const float m[16], float out[16]) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m03 = m[3], m10 = m[4], m11 = m[5], m12 = m[6], m13 = m[7];
    const float m20 = m[8], m21 = m[9], m22 = m[10], m23 = m[11], m30 = m[12], m31 = m[13], m32 = m[14], m33 = m[15];
    const float c00 = m22 * m33 - m32 * m23, c02 = m12 * m33 - m32 * m13, c03 = m12 * m23 - m22 * m13;
    const float c04 = m21 * m33 - m31 * m23, c06 = m11 * m33 - m31 * m13, c07 = m11 * m23 - m21 * m13;
    const float c08 = m21 * m32 - m31 * m22, c10 = m11 * m32 - m31 * m12, c11 = m11 * m22 - m21 * m12;
    const float c12 = m20 * m33 - m30 * m23, c14 = m10 * m33 - m30 * m13, c15 = m10 * m23 - m20 * m13;
    const float c16 = m20 * m32 - m30 * m22, c18 = m10 * m32 - m30 * m12, c19 = m10 * m22 - m20 * m12;
    const float c20 = m20 * m31 - m30 * m21, c22 = m10 * m31 - m30 * m11, c23 = m10 * m21 - m20 * m11;
    const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    const float v0[4] = {m10, m00, m00, m00}, v1[4] = {m11, m01, m01, m01}, v2[4] = {m12, m02, m02, m02}, v3[4] = {m13, m03, m03, m03};
    float inv[16];
    for (int k = 0; k < 4; ++k) {
        const float sa = (k & 1) ? -1.0f : 1.0f, sb = -sa;
        inv[k]      = ((v1[k] * f0[k] - v2[k] * f1[k]) + v3[k] * f2[k]) * sa;
        inv[4 + k]  = ((v0[k] * f0[k] - v2[k] * f3[k]) + v3[k] * f4[k]) * sb;
        inv[8 + k]  = ((v0[k] * f1[k] - v1[k] * f3[k]) + v3[k] * f5[k]) * sa;
        inv[12 + k] = ((v0[k] * f2[k] - v1[k] * f4[k]) + v2[k] * f5[k]) * sb;
    }
    const float d0 = m00 * inv[0], d1 = m01 * inv[4], d2 = m02 * inv[8], d3 = m03 * inv[12];
    const float det = (d0 + d2) + (d1 + d3);
    const float rcp = 1.0f / det;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * rcp;
}

// transforms: 12 floats per entity (OrbitTransform). Returns 1 when the visibility buffer overflowed.
int oracle_scene_update(const float* transforms, const uint32_t* mesh_slots, uint32_t* visibility_offsets, const void* mesh_infos_,
                        uint32_t* visibility_cursor, uint32_t n_entities, uint32_t visibility_capacity_words,
                        void* entity_data_, void* entity_draws_) {
    const OrbitMeshInfo* mesh_infos = (const OrbitMeshInfo*)mesh_infos_;
    float* entity_data = (float*)entity_data_;
    uint32_t* draw_words = (uint32_t*)entity_draws_;
    uint32_t count = 0;
    uint64_t cursor = *visibility_cursor;
    int overflow = 0;
    for (uint32_t e = 0; e < n_entities; ++e) {              // scene.rs:419
        const uint32_t slot = mesh_slots[e];
        if (slot == ORBIT_NO_MESH) continue;
        uint32_t vo = visibility_offsets[e];
        if (vo == ORBIT_NO_VISIBILITY_RANGE) {               // scene.rs:424-431; allocator without frees = bump pointer
            const uint32_t mc = mesh_infos[slot].mesh_lods[0].meshlet_count;
            const uint32_t words = (mc >> 5) + ((mc & 31u) ? 1u : 0u);
            vo = (uint32_t)cursor;
            cursor += words;
            if (cursor > visibility_capacity_words) overflow = 1;
            visibility_offsets[e] = vo;
        }
        const float* t = transforms + (size_t)e * 12;
        float model[16], inv[16];
        glam_from_srt(t, t + 4, t + 8, model);
        glam_inverse(model, inv);
        float* out = entity_data + (size_t)count * 32;
        std::memcpy(out, model, 64);
        float* nm = out + 16;                                // from_mat3(from_mat4(inverse.transpose()))
        for (int c = 0; c < 3; ++c) { for (int r = 0; r < 3; ++r) nm[c * 4 + r] = inv[r * 4 + c]; nm[c * 4 + 3] = 0.0f; }
        nm[12] = 0.0f; nm[13] = 0.0f; nm[14] = 0.0f; nm[15] = 1.0f;
        draw_words[1 + 3 * (size_t)count + 0] = count;       // instance_index
        draw_words[1 + 3 * (size_t)count + 1] = slot;
        draw_words[1 + 3 * (size_t)count + 2] = vo;
        ++count;
    }
    draw_words[0] = count;
    *visibility_cursor = (uint32_t)cursor;
    return overflow;
}

// ---- asset-side producers --------------------------------------------------------------------------
// meshlet bounds: the bounds part of compute_meshlets (src/assets/mesh.rs:292-338) = meshopt::compute_meshlet_bounds.
// meshopt 0.2.0 (Cargo.toml:40) is NOT vendored in the reference tree; this restates meshoptimizer's published
// meshopt_computeMeshletBounds / computeBoundingSphere (clusterizer.cpp). PARITY UNPINNED against the reference (no
// executable meshopt here); the CUDA kernel (csrc/asset_bounds.cu) is held to this restatement bit for bit.
static void ref_bounding_sphere(float result[4], const float (*points)[3], size_t count) {
    size_t pmin[3] = {0, 0, 0}, pmax[3] = {0, 0, 0};
    for (size_t i = 0; i < count; ++i) {
        const float* p = points[i];
        for (int axis = 0; axis < 3; ++axis) {
            pmin[axis] = (p[axis] < points[pmin[axis]][axis]) ? i : pmin[axis];
            pmax[axis] = (p[axis] > points[pmax[axis]][axis]) ? i : pmax[axis];
        }
    }
    float paxisd2 = 0;
    int paxis = 0;
    for (int axis = 0; axis < 3; ++axis) {
        const float* p1 = points[pmin[axis]];
        const float* p2 = points[pmax[axis]];
        const float dx = p2[0] - p1[0], dy = p2[1] - p1[1], dz = p2[2] - p1[2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 > paxisd2) { paxisd2 = d2; paxis = axis; }
    }
    const float* p1 = points[pmin[paxis]];
    const float* p2 = points[pmax[paxis]];
    float center[3] = {(p1[0] + p2[0]) / 2, (p1[1] + p2[1]) / 2, (p1[2] + p2[2]) / 2};
    float radius = std::sqrt(paxisd2) / 2;
    for (size_t i = 0; i < count; ++i) {
        const float* p = points[i];
        const float dx = p[0] - center[0], dy = p[1] - center[1], dz = p[2] - center[2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 > radius * radius) {
            const float d = std::sqrt(d2);
            const float k = 0.5f + (radius / d) / 2;
            const float k1 = 1 - k;
            center[0] = center[0] * k + p[0] * k1;
            center[1] = center[1] * k + p[1] * k1;
            center[2] = center[2] * k + p[2] * k1;
            radius = (radius + d) / 2;
        }
    }
    result[0] = center[0]; result[1] = center[1]; result[2] = center[2]; result[3] = radius;
}

static int ref_quantize_snorm8(float v) {
    const float round = (v >= 0 ? 0.5f : -0.5f);
    v = (v >= -1) ? v : -1;
    v = (v <= +1) ? v : +1;
    return int(v * 127.0f + round);
}

// meshlets: OrbitMeshlet[n] in / out. Returns the number of meshlets skipped because they hold more than 128 triangles.
int oracle_meshlet_bounds(const void* vertices_, uint32_t vertex_stride, const uint32_t* meshlet_data, void* meshlets_, uint32_t n_meshlets) {
    const uint8_t* vertices = (const uint8_t*)vertices_;
    OrbitMeshlet* meshlets = (OrbitMeshlet*)meshlets_;
    int skipped = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : skipped)
    for (int64_t m = 0; m < (int64_t)n_meshlets; ++m) {
        OrbitMeshlet& ml = meshlets[m];
        const uint32_t tcount = ml.triangle_count;
        if (tcount > 128u) { ++skipped; continue; }
        const uint32_t* vidx = meshlet_data + ml.data_offset;
        const uint8_t* tris = (const uint8_t*)(vidx + ml.vertex_count);
        float normals[128][3];
        float corners[128 * 3][3];
        size_t triangles = 0;
        for (uint32_t t = 0; t < tcount; ++t) {
            const float* pc[3];
            for (int k = 0; k < 3; ++k) pc[k] = (const float*)(vertices + (size_t)(ml.vertex_offset + vidx[tris[3 * t + k]]) * vertex_stride);
            const float p10[3] = {pc[1][0] - pc[0][0], pc[1][1] - pc[0][1], pc[1][2] - pc[0][2]};
            const float p20[3] = {pc[2][0] - pc[0][0], pc[2][1] - pc[0][1], pc[2][2] - pc[0][2]};
            const float nx = p10[1] * p20[2] - p10[2] * p20[1];
            const float ny = p10[2] * p20[0] - p10[0] * p20[2];
            const float nz = p10[0] * p20[1] - p10[1] * p20[0];
            const float area = std::sqrt(nx * nx + ny * ny + nz * nz);
            if (area == 0.f) continue;                       // degenerate triangles are invisible anyway
            normals[triangles][0] = nx / area; normals[triangles][1] = ny / area; normals[triangles][2] = nz / area;
            for (int k = 0; k < 3; ++k) std::memcpy(corners[3 * triangles + k], pc[k], 12);
            ++triangles;
        }
        if (triangles == 0) {                                // degenerate cluster: zeroed bounds
            ml.bounding_sphere[0] = ml.bounding_sphere[1] = ml.bounding_sphere[2] = ml.bounding_sphere[3] = 0.f;
            ml.cone_axis[0] = ml.cone_axis[1] = ml.cone_axis[2] = 0; ml.cone_cutoff = 0;
            continue;
        }
        float psphere[4], nsphere[4];
        ref_bounding_sphere(psphere, corners, triangles * 3);
        ref_bounding_sphere(nsphere, normals, triangles);
        float axis[3] = {nsphere[0], nsphere[1], nsphere[2]};
        const float axislength = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
        const float invaxislength = axislength == 0.f ? 0.f : 1.f / axislength;
        axis[0] *= invaxislength; axis[1] *= invaxislength; axis[2] *= invaxislength;
        float mindp = 1.f;
        for (size_t i = 0; i < triangles; ++i) {
            const float dp = normals[i][0] * axis[0] + normals[i][1] * axis[1] + normals[i][2] * axis[2];
            mindp = (dp < mindp) ? dp : mindp;
        }
        for (int k = 0; k < 4; ++k) ml.bounding_sphere[k] = psphere[k];
        if (mindp <= 0.1f) {                                 // normal cone ~168 degrees or wider: trivial accept
            ml.cone_axis[0] = ml.cone_axis[1] = ml.cone_axis[2] = 0; ml.cone_cutoff = 127;
            continue;
        }
        const float cutoff = std::sqrt(1 - mindp * mindp);
        const int q0 = ref_quantize_snorm8(axis[0]), q1 = ref_quantize_snorm8(axis[1]), q2 = ref_quantize_snorm8(axis[2]);
        const float e0 = std::fabs((signed char)q0 / 127.f - axis[0]);
        const float e1 = std::fabs((signed char)q1 / 127.f - axis[1]);
        const float e2 = std::fabs((signed char)q2 / 127.f - axis[2]);
        const int qc = int(127 * (cutoff + e0 + e1 + e2) + 1);
        ml.cone_axis[0] = (int8_t)q0; ml.cone_axis[1] = (int8_t)q1; ml.cone_axis[2] = (int8_t)q2;
        ml.cone_cutoff = (qc > 127) ? (int8_t)127 : (int8_t)qc;
    }
    return skipped;
}

// MeshData::compute_bounds (src/assets/mesh.rs:192-215): vertex_ranges[2m] = first vertex, [2m+1] = count.
void oracle_mesh_bounds(const void* vertices_, uint32_t vertex_stride, const uint32_t* vertex_ranges, void* mesh_infos_, uint32_t n_meshes) {
    const uint8_t* vertices = (const uint8_t*)vertices_;
    uint8_t* mesh_infos = (uint8_t*)mesh_infos_;
    for (uint32_t m = 0; m < n_meshes; ++m) {
        const uint32_t first = vertex_ranges[2 * m], count = vertex_ranges[2 * m + 1];
        if (count == 0) continue;
        auto pos = [&](uint32_t i) { return (const float*)(vertices + (size_t)(first + i) * vertex_stride); };
        float lo[3] = {pos(0)[0], pos(0)[1], pos(0)[2]}, hi[3] = {lo[0], lo[1], lo[2]};
        for (uint32_t i = 0; i < count; ++i)
            for (int k = 0; k < 3; ++k) { lo[k] = std::fmin(lo[k], pos(i)[k]); hi[k] = std::fmax(hi[k], pos(i)[k]); }
        const float c[3] = {(hi[0] + lo[0]) * 0.5f, (hi[1] + lo[1]) * 0.5f, (hi[2] + lo[2]) * 0.5f};
        float r2 = 0.0f;
        for (uint32_t i = 0; i < count; ++i) {
            const float dx = pos(i)[0] - c[0], dy = pos(i)[1] - c[1], dz = pos(i)[2] - c[2];
            r2 = std::fmax(r2, dx * dx + dy * dy + dz * dz);
        }
        float* mi = (float*)(mesh_infos + (size_t)m * 128);
        mi[0] = c[0]; mi[1] = c[1]; mi[2] = c[2]; mi[3] = std::sqrt(r2);
        mi[4] = lo[0]; mi[5] = lo[1]; mi[6] = lo[2]; mi[7] = 0.0f;
        mi[8] = hi[0]; mi[9] = hi[1]; mi[10] = hi[2]; mi[11] = 0.0f;
    }
}

}  // extern "C"
