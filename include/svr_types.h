/*
 * svr_types.h -- plain-C mirror of the scene structs that cross the render boundary.
 *
 * The reference passes C++ PODs built on packed GLM vectors through seven `extern "C"`
 * functions (pathtracer.h:17-24, raycasting.h:8).  The types below have exactly the same size,
 * alignment and field offsets (probed with nvcc 12.9 / x86-64), so a
 * `cudaVolume&` on the reference side and a `svr_volume*` on this side are the same bits on
 * the wire.  Nothing here depends on GLM, Qt, VTK or torch.
 *
 *   reference type             file:line                               size/align
 *   glm::vec3                  (GLM, packed)                           12 / 4
 *   glm::u8vec4                (GLM, packed)                            4 / 1
 *   cudaBBox                   core/geometry/cuda_bbox.h:66-69         36 / 4
 *   cudaVolume                 core/cuda_volume.h:111-121             112 / 8
 *   cudaTransferFunction       core/cuda_transfer_function.h:57-59     16 / 8
 *   cudaCamera                 core/cuda_camera.h:98-106               76 / 4
 *   cudaDisk                   core/geometry/cuda_disk.h:58-61         28 / 4
 *   cudaAreaLight              core/lights/cuda_arealight.h:68-71      44 / 4
 *   cudaEnvironmentLight       core/lights/cuda_environment_light.h:74-78  32 / 8
 *   RenderParams               core/render_parameters.h:34-37          16 / 8
 */
#ifndef SVR_TYPES_H
#define SVR_TYPES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svr_vec2 { float x, y; } svr_vec2;
typedef struct svr_vec3 { float x, y, z; } svr_vec3;
typedef struct svr_vec4 { float x, y, z, w; } svr_vec4;
typedef struct svr_u8vec4 { uint8_t x, y, z, w; } svr_u8vec4;

/* cudaTextureObject_t is `unsigned long long` (driver_types.h). */
typedef unsigned long long svr_texture_t;

/* cudaBBox: world-space box + reciprocal size (cuda_bbox.h:25-31). */
typedef struct svr_bbox {
    svr_vec3 vmin;
    svr_vec3 vmax;
    svr_vec3 invSize;
} svr_bbox;

/* cudaVolume: members are private in the reference (cuda_volume.h:111-121); offsets in comments. */
typedef struct svr_volume {
    svr_bbox bbox;            /*   0 */
    uint32_t _pad0;           /*  36 */
    svr_texture_t tex;        /*  40  3-D, 1 channel, linear, border, normalised coords */
    float densityScale;       /*  48 */
    float invMaxMagnitude;    /*  52 */
    float gradientFactor;     /*  56 */
    svr_vec3 spacing;         /*  60 */
    svr_vec3 invSpacing;      /*  72 */
    svr_vec2 x_clip;          /*  84 */
    svr_vec2 y_clip;          /*  92 */
    svr_vec2 z_clip;          /* 100 */
    uint32_t _pad1;           /* 108 */
} svr_volume;

/* cudaTransferFunction (cuda_transfer_function.h:57-59). */
typedef struct svr_transfer_function {
    svr_texture_t tex;        /* 0  1-D float4 x 1024, linear, clamp, normalised coords */
    float maxOpacity;         /* 8  global majorant sigma_max */
    uint32_t _pad0;           /* 12 */
} svr_transfer_function;

/* cudaCamera (cuda_camera.h:98-106; `apeture` spelled as in the reference). */
typedef struct svr_camera {
    uint32_t imageW, imageH;  /* 0, 4 */
    float exposure;           /* 8 */
    float apeture;            /* 12 */
    float focalLength;        /* 16 */
    float aspectRatio;        /* 20 */
    float tanFovxOverTwo;     /* 24 */
    svr_vec3 pos;             /* 28 */
    svr_vec3 u, v, w;         /* 40, 52, 64 */
} svr_camera;

/* cudaDisk (cuda_disk.h:58-61). */
typedef struct svr_disk {
    float radius;
    svr_vec3 center;
    svr_vec3 normal;
} svr_disk;

/* cudaAreaLight (cuda_arealight.h:68-71). */
typedef struct svr_area_light {
    svr_disk disk;
    svr_vec3 color;
    float intensity;
} svr_area_light;

/* cudaEnvironmentLight (cuda_environment_light.h:74-78). tex == 0 => constant defaultRadiance. */
typedef struct svr_env_light {
    svr_texture_t tex;        /* 0  2-D float4 lat-long, linear, wrap, normalised */
    svr_vec3 defaultRadiance; /* 8 */
    float intensity;          /* 20 */
    svr_vec2 offset;          /* 24 */
} svr_env_light;

/* RenderParams (render_parameters.h:34-37). hdrBuffer: device, imageW*imageH packed vec3. */
typedef struct svr_render_params {
    uint32_t traceDepth;      /* 0 */
    uint32_t frameNo;         /* 4 */
    svr_vec3* hdrBuffer;      /* 8 */
} svr_render_params;

#define SVR_MAX_LIGHT_SOURCES 8 /* common.h:11 */
#define SVR_TF_TABLE_SIZE 1024  /* gui/transferfunction.h:29 */

#if defined(__cplusplus) || (defined(__STDC_VERSION__) && __STDC_VERSION__ >= 201112L)
#ifdef __cplusplus
#define SVR_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define SVR_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif
SVR_STATIC_ASSERT(sizeof(svr_vec3) == 12, "glm::vec3 is 12 bytes");
SVR_STATIC_ASSERT(sizeof(svr_u8vec4) == 4, "glm::u8vec4 is 4 bytes");
SVR_STATIC_ASSERT(sizeof(svr_bbox) == 36, "cudaBBox is 36 bytes");
SVR_STATIC_ASSERT(sizeof(svr_volume) == 112, "cudaVolume is 112 bytes");
SVR_STATIC_ASSERT(offsetof(svr_volume, tex) == 40, "cudaVolume::tex @40");
SVR_STATIC_ASSERT(offsetof(svr_volume, densityScale) == 48, "cudaVolume::densityScale @48");
SVR_STATIC_ASSERT(offsetof(svr_volume, spacing) == 60, "cudaVolume::spacing @60");
SVR_STATIC_ASSERT(offsetof(svr_volume, invSpacing) == 72, "cudaVolume::invSpacing @72");
SVR_STATIC_ASSERT(offsetof(svr_volume, x_clip) == 84, "cudaVolume::x_clip @84");
SVR_STATIC_ASSERT(offsetof(svr_volume, z_clip) == 100, "cudaVolume::z_clip @100");
SVR_STATIC_ASSERT(sizeof(svr_transfer_function) == 16, "cudaTransferFunction is 16 bytes");
SVR_STATIC_ASSERT(offsetof(svr_transfer_function, maxOpacity) == 8, "maxOpacity @8");
SVR_STATIC_ASSERT(sizeof(svr_camera) == 76, "cudaCamera is 76 bytes");
SVR_STATIC_ASSERT(offsetof(svr_camera, pos) == 28, "cudaCamera::pos @28");
SVR_STATIC_ASSERT(offsetof(svr_camera, w) == 64, "cudaCamera::w @64");
SVR_STATIC_ASSERT(sizeof(svr_disk) == 28, "cudaDisk is 28 bytes");
SVR_STATIC_ASSERT(sizeof(svr_area_light) == 44, "cudaAreaLight is 44 bytes");
SVR_STATIC_ASSERT(sizeof(svr_env_light) == 32, "cudaEnvironmentLight is 32 bytes");
SVR_STATIC_ASSERT(offsetof(svr_env_light, intensity) == 20, "env intensity @20");
SVR_STATIC_ASSERT(sizeof(svr_render_params) == 16, "RenderParams is 16 bytes");
SVR_STATIC_ASSERT(offsetof(svr_render_params, hdrBuffer) == 8, "RenderParams::hdrBuffer @8");
#endif

#ifdef __cplusplus
}
#endif

#endif /* SVR_TYPES_H */
