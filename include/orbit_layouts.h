/*
 * orbit_layouts.h — byte layouts of every buffer that crosses the drop-in boundary.
 *
 * These are the reference's own GPU structs (std430 on the GLSL side, #[repr(C)] + bytemuck::Pod on
 * the Rust side), restated as plain C so that C, C++, CUDA and (via the mirrored rust/ crate) Rust all
 * agree on them.  Nothing here is ML vocabulary: entities, meshes, meshlets, dispatch records, draw
 * commands, visibility words, depth pyramid, clusters, lights.
 *
 * Citations are relative to the reference tree (Thefefe/orbit):
 *   GpuCullInfo                 src/passes/draw_gen.rs:208-237      shaders/include/types.glsl:202-228
 *   GpuEntityData               src/scene.rs:120-125                types.glsl:75-78
 *   GpuEntityDraw (+ buffer)    src/scene.rs:127-133,475-485        types.glsl:112-121
 *   GpuMeshInfo / GpuMeshLod    src/assets/mod.rs:18-43             types.glsl:128-141
 *   GpuMeshlet                  src/assets/mod.rs:111-122           types.glsl:143-152
 *   GpuMaterialData             src/assets/mod.rs:171-191           types.glsl:92-110
 *   MeshletDispatch (+ buffer)  types.glsl:166-178
 *   GpuMeshletDrawCommand       src/assets/mod.rs:98-109            types.glsl:180-194
 *   MeshTaskPayload             types.glsl:196-200
 *   GpuLightData                src/scene.rs:278-291                types.glsl:16-27
 *   ClusterCullInfo             src/passes/cluster.rs:186-207       shaders/light_cluster/light_culling.comp:8-26
 *   ClusterDepthBounds          types.glsl:251-255
 *   CompactedClusterIndexList   types.glsl:270-276
 *   ClusterLightIndices         types.glsl:246-249
 */
#ifndef ORBIT_LAYOUTS_H
#define ORBIT_LAYOUTS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define ORBIT_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define ORBIT_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

#define ORBIT_MAX_CULL_PLANES 12u        /* draw_gen.rs:205 */
#define ORBIT_MESHLET_DISPATCH_SIZE 32u  /* spec constant 0 := task_shader_workgroup_size() = 32, device.rs:369-372 */
#define ORBIT_MAX_MESH_LODS 8u           /* assets/mod.rs:16 */
#define ORBIT_NO_BUFFER 0xFFFFFFFFu      /* draw_gen.rs:140-142: descriptor index when a buffer is absent */
#define ORBIT_MAX_LIGHTS_PER_CLUSTER 256u /* light_culling.comp:132 */

/* occlusion_pass values — draw_gen.rs:96-102 */
#define ORBIT_PASS_NONE 0u
#define ORBIT_PASS_VISIBILITY_READ 1u
#define ORBIT_PASS_VISIBILITY_WRITE 2u

/* projection_type — draw_gen.rs:165-168 */
#define ORBIT_PROJ_PERSPECTIVE 0u
#define ORBIT_PROJ_ORTHOGRAPHIC 1u

/* alpha modes (assets/mod.rs:124-130) and AlphaModeFlags bits (draw_gen.rs:630-641) */
#define ORBIT_ALPHA_OPAQUE 0u
#define ORBIT_ALPHA_MASKED 1u
#define ORBIT_ALPHA_TRANSPARENT 2u

/* light types — types.glsl:12-14 */
#define ORBIT_LIGHT_SKY 0u
#define ORBIT_LIGHT_DIRECTIONAL 1u
#define ORBIT_LIGHT_POINT 2u

/* Column-major 4x4, glam::Mat4 memory order == GLSL mat4: m[col][row]. */
typedef struct OrbitMat4 { float m[4][4]; } OrbitMat4;

typedef struct OrbitCullInfo {
    OrbitMat4 view_matrix;                 /*   0 */
    OrbitMat4 reprojection_matrix;         /*  64  always zero in the reference */
    float     cull_planes[ORBIT_MAX_CULL_PLANES][4]; /* 128  view space (nx,ny,nz,d), normalised */
    uint32_t  cull_plane_count;            /* 320 */
    uint32_t  alpha_mode_flags;            /* 324 */
    uint32_t  noskip_alpha_mode;           /* 328 */
    uint32_t  occlusion_pass;              /* 332 */
    uint32_t  visibility_buffer;           /* 336  descriptor index in the reference; ignored here */
    uint32_t  meshlet_visibility_buffer;   /* 340  ORBIT_NO_BUFFER disables meshlet occlusion culling */
    uint32_t  depth_pyramid;               /* 344  descriptor index; ignored here */
    uint32_t  secondary_depth_pyramid;     /* 348 */
    uint32_t  projection_type;             /* 352 */
    float     p00_or_width_recip_x2;       /* 356 */
    float     p11_or_height_recip_x2;      /* 360 */
    float     z_near;                      /* 364 */
    float     z_far;                       /* 368 */
    float     lod_base;                    /* 372 */
    float     lod_step;                    /* 376 */
    uint32_t  min_mesh_lod;                /* 380 */
    float     lod_target_pos_view_space[3];/* 384 */
    uint32_t  max_mesh_lod;                /* 396 */
} OrbitCullInfo;
ORBIT_STATIC_ASSERT(sizeof(OrbitCullInfo) == 400, "GpuCullInfo is 400 bytes");
ORBIT_STATIC_ASSERT(offsetof(OrbitCullInfo, cull_planes) == 128, "cull_planes@128");
ORBIT_STATIC_ASSERT(offsetof(OrbitCullInfo, cull_plane_count) == 320, "cull_plane_count@320");
ORBIT_STATIC_ASSERT(offsetof(OrbitCullInfo, occlusion_pass) == 332, "occlusion_pass@332");
ORBIT_STATIC_ASSERT(offsetof(OrbitCullInfo, meshlet_visibility_buffer) == 340, "meshlet_visibility_buffer@340");
ORBIT_STATIC_ASSERT(offsetof(OrbitCullInfo, projection_type) == 352, "projection_type@352");
ORBIT_STATIC_ASSERT(offsetof(OrbitCullInfo, z_near) == 364, "z_near@364");
ORBIT_STATIC_ASSERT(offsetof(OrbitCullInfo, lod_target_pos_view_space) == 384, "lod_target@384");
ORBIT_STATIC_ASSERT(offsetof(OrbitCullInfo, max_mesh_lod) == 396, "max_mesh_lod@396");

typedef struct OrbitEntityData {
    OrbitMat4 model_matrix;   /*  0 */
    OrbitMat4 normal_matrix;  /* 64  not read by the culling path */
} OrbitEntityData;
ORBIT_STATIC_ASSERT(sizeof(OrbitEntityData) == 128, "GpuEntityData is 128 bytes");

/* Transform (src/scene.rs:18-23: position Vec3, orientation Quat, scale Vec3). The reference's struct is a plain
 * Rust struct without a defined byte layout; this is the layout the scene-update entry point reads: three 16-byte
 * rows so that one entity is three aligned vector loads. */
typedef struct OrbitTransform {
    float position[3];    /*  0 */
    float _pad0;
    float orientation[4]; /* 16  x, y, z, w (glam Quat memory order) */
    float scale[3];       /* 32 */
    float _pad1;
} OrbitTransform;
ORBIT_STATIC_ASSERT(sizeof(OrbitTransform) == 48, "OrbitTransform is 48 bytes");
#define ORBIT_NO_MESH 0xFFFFFFFFu        /* EntityData::mesh == None (scene.rs:63) */
#define ORBIT_NO_VISIBILITY_RANGE 0xFFFFFFFFu /* EntityData::visibility_buffer_range == None (scene.rs:65) */

typedef struct OrbitEntityDraw {
    uint32_t entity_index;
    uint32_t mesh_index;
    uint32_t visibility_offset;  /* first meshlet-visibility word of this draw, scene.rs:422-431 */
} OrbitEntityDraw;
ORBIT_STATIC_ASSERT(sizeof(OrbitEntityDraw) == 12, "GpuEntityDraw is 12 bytes");
/* EntityDrawBuffer: uint32 count @0, OrbitEntityDraw draws[] @4 */
#define ORBIT_ENTITY_DRAW_HEADER_BYTES 4u

typedef struct OrbitMeshLod {
    uint32_t meshlet_offset;  /* absolute index into the meshlet array, assets/mod.rs:82-85 */
    uint32_t meshlet_count;
} OrbitMeshLod;

typedef struct OrbitMeshInfo {
    float        bounding_sphere[4];  /*  0 */
    float        aabb_min[4];         /* 16 */
    float        aabb_max[4];         /* 32 */
    uint32_t     vertex_offset;       /* 48 */
    uint32_t     meshlet_data_offset; /* 52 */
    uint32_t     lod_count;           /* 56 */
    uint32_t     _padding;            /* 60 */
    OrbitMeshLod mesh_lods[ORBIT_MAX_MESH_LODS]; /* 64 */
} OrbitMeshInfo;
ORBIT_STATIC_ASSERT(sizeof(OrbitMeshInfo) == 128, "GpuMeshInfo is 128 bytes");
ORBIT_STATIC_ASSERT(offsetof(OrbitMeshInfo, lod_count) == 56, "lod_count@56");
ORBIT_STATIC_ASSERT(offsetof(OrbitMeshInfo, mesh_lods) == 64, "mesh_lods@64");

typedef struct OrbitMeshlet {
    float    bounding_sphere[4]; /*  0  model space xyz, r */
    int8_t   cone_axis[3];       /* 16  snorm8 */
    int8_t   cone_cutoff;        /* 19  snorm8 */
    uint32_t vertex_offset;      /* 20 */
    uint32_t data_offset;        /* 24 */
    uint16_t material_index;     /* 28 */
    uint8_t  vertex_count;       /* 30 */
    uint8_t  triangle_count;     /* 31 */
} OrbitMeshlet;
ORBIT_STATIC_ASSERT(sizeof(OrbitMeshlet) == 32, "GpuMeshlet is 32 bytes");
ORBIT_STATIC_ASSERT(offsetof(OrbitMeshlet, vertex_offset) == 20, "vertex_offset@20");
ORBIT_STATIC_ASSERT(offsetof(OrbitMeshlet, material_index) == 28, "material_index@28");

/* GpuMaterialData is 80 bytes; the culling path reads only alpha_mode. */
#define ORBIT_MATERIAL_STRIDE_BYTES 80u
#define ORBIT_MATERIAL_ALPHA_MODE_OFFSET 64u

typedef struct OrbitMeshletDispatch {
    uint32_t entity_index;
    uint32_t meshlet_offset;
    uint32_t meshlet_count;      /* 1..32 */
    uint32_t visibility_offset;
} OrbitMeshletDispatch;
ORBIT_STATIC_ASSERT(sizeof(OrbitMeshletDispatch) == 16, "MeshletDispatch is 16 bytes");
/* MeshletDispatchBuffer: uint32 workgroup_count_x @0, _y @4 (=1), _z @8 (=1), records @12 */
#define ORBIT_DISPATCH_HEADER_BYTES 12u

typedef struct OrbitMeshletDrawCommand {
    uint32_t cmd_index_count;        /*  0  triangle_count * 3 */
    uint32_t cmd_instance_count;     /*  4  1 */
    uint32_t cmd_first_index;        /*  8  (data_offset + vertex_count) * 4 */
    int32_t  cmd_vertex_offset;      /* 12  int(data_offset) */
    uint32_t cmd_first_instance;     /* 16  entity_index */
    uint32_t meshlet_vertex_offset;  /* 20 */
    uint32_t meshlet_index;          /* 24 */
} OrbitMeshletDrawCommand;
ORBIT_STATIC_ASSERT(sizeof(OrbitMeshletDrawCommand) == 28, "GpuMeshletDrawCommand is 28 bytes");
/* MeshletDrawCommandBuffer: uint32 count @0, commands @4 */
#define ORBIT_DRAW_HEADER_BYTES 4u

typedef struct OrbitMeshTaskPayload {
    uint32_t entity_index;
    uint32_t meshlet_offset;
    uint8_t  meshlet_indices[32];
} OrbitMeshTaskPayload;
ORBIT_STATIC_ASSERT(sizeof(OrbitMeshTaskPayload) == 40, "MeshTaskPayload is 40 bytes");

typedef struct OrbitLightData {
    uint32_t light_type;         /*  0 */
    uint32_t shadow_data_index;  /*  4 */
    uint32_t irradiance_map;     /*  8 */
    uint32_t prefiltered_map;    /* 12 */
    float    color[3];           /* 16 */
    float    intensity;          /* 28 */
    float    position[3];        /* 32 */
    float    inner_radius;       /* 44 */
    float    direction[3];       /* 48 */
    float    outer_radius;       /* 60 */
} OrbitLightData;
ORBIT_STATIC_ASSERT(sizeof(OrbitLightData) == 64, "GpuLightData is 64 bytes");
ORBIT_STATIC_ASSERT(offsetof(OrbitLightData, position) == 32, "position@32");
ORBIT_STATIC_ASSERT(offsetof(OrbitLightData, outer_radius) == 60, "outer_radius@60");

typedef struct OrbitClusterCullInfo {
    OrbitMat4 world_to_view_matrix;   /*   0 */
    OrbitMat4 screen_to_view_matrix;  /*  64  inverse projection */
    uint32_t  cluster_count[3];       /* 128 */
    uint32_t  tile_size_px;           /* 140 */
    uint32_t  screen_size[2];         /* 144 */
    float     z_near;                 /* 152 */
    float     z_far;                  /* 156 */
    uint32_t  unique_cluster_buffer;  /* 160  descriptor ids in the reference; ignored here */
    uint32_t  cluster_offset_image;   /* 164 */
    uint32_t  light_index_buffer;     /* 168 */
    uint32_t  depth_bounds_buffer;    /* 172 */
    uint32_t  global_light_count;     /* 176 */
    uint32_t  global_light_list;      /* 180 */
    uint32_t  _padding[2];            /* 184 */
} OrbitClusterCullInfo;
ORBIT_STATIC_ASSERT(sizeof(OrbitClusterCullInfo) == 192, "ClusterCullInfo is 192 bytes");
ORBIT_STATIC_ASSERT(offsetof(OrbitClusterCullInfo, cluster_count) == 128, "cluster_count@128");
ORBIT_STATIC_ASSERT(offsetof(OrbitClusterCullInfo, global_light_count) == 176, "global_light_count@176");

typedef struct OrbitClusterDepthBounds {
    uint32_t min_depth;  /* max over pixels of float bits of (1 - d) */
    uint32_t max_depth;  /* max over pixels of float bits of d */
} OrbitClusterDepthBounds;
ORBIT_STATIC_ASSERT(sizeof(OrbitClusterDepthBounds) == 8, "ClusterDepthBounds is 8 bytes");

/* CompactedClusterIndexList: wg_x @0, wg_y @4, wg_z @8, cluster_count @12, indices @16 */
#define ORBIT_COMPACT_CLUSTER_HEADER_BYTES 16u
/* ClusterLightIndices: light_count @0 (running total), indices @4 */
#define ORBIT_LIGHT_INDEX_HEADER_BYTES 4u
/* cluster offset image: 3-D R32G32_UINT, texel [x,y,z] at ((z*cy + y)*cx + x)*8 = (light_offset, light_count) */

#endif /* ORBIT_LAYOUTS_H */
