/*
 * TEST INFRASTRUCTURE -- not product code.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * C ABI of the CPU oracle: a host restatement of the reference's render hot path
 * (pathtracer.cu, raycasting.cu, core/ headers) with a software texture sampler and a host XORWOW.
 */
#ifndef SVR_ORACLE_H
#define SVR_ORACLE_H

#include "../include/svr_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svr_oracle_scene {
    const void* voxels;           /* host, x-fastest */
    int32_t format;               /* svr_voxel_format: 0 u8, 1 u16, 2 f16, 3 f32 */
    uint32_t nx, ny, nz;
    svr_volume volume;            /* .tex ignored; everything else as the GPU side sees it */
    const float* tfTable;         /* host, tfSize x (r,g,b,opacity) */
    uint32_t tfSize;
    svr_transfer_function tf;     /* .tex ignored; maxOpacity used as the global majorant */
    svr_camera camera;
    svr_area_light lights[SVR_MAX_LIGHT_SOURCES];
    uint32_t numLights;
    svr_env_light env;            /* constant radiance only (tex ignored) */
    int32_t envEnabled;           /* 0 = as shipped (pathtracer.cu:233 commented out) */
    int32_t filterMode;           /* 0 = the texture unit's filter as measured on a B200 (integer 1/256 weights
                                     built in two rounded stages, 16-bit results for u8/u16 reads: see
                                     svr_oracle.cpp); 2 = plain fp32 trilinear weights */
} svr_oracle_scene;

enum { SVR_ORACLE_CNT_TRACK_TAPS = 0, SVR_ORACLE_CNT_SHADOW_TAPS = 1, SVR_ORACLE_CNT_SHADE_TAPS = 2,
       SVR_ORACLE_CNT_TF_LOOKUPS = 3, SVR_ORACLE_CNT_SCATTERS = 4, SVR_ORACLE_CNT_PATHS = 5,
       SVR_ORACLE_CNT_STEPS = 7, SVR_ORACLE_CNT_COUNT = 16 };

int svr_oracle_threads(void);
void svr_oracle_set_threads(int n);

/* raycasting.cu:15-67 over rows [y0,y1).  outRGBA (float4/pixel, pre-quantisation L) and outU8 may
 * each be NULL.  strideW = the reference's compile-time WIDTH (row stride). */
void svr_oracle_raycast(const svr_oracle_scene* scene, float stepSize, uint32_t strideW, uint32_t y0, uint32_t y1,
                        float* outRGBA, uint8_t* outU8, uint64_t* counters);

/* pathtracer.cu:200-280 for frames frameNo0 .. frameNo0+nFrames-1 over rows [y0,y1): hdr is the packed
 * vec3 running mean (pathtracer.cu:81-84), cleared when a frame number is 0 (pathtracer.cu:297-300). */
void svr_oracle_pathtrace(const svr_oracle_scene* scene, uint32_t traceDepth, uint32_t frameNo0, uint32_t nFrames,
                          uint32_t strideW, uint32_t y0, uint32_t y1, float* hdr, uint64_t* counters);

/* the same over rows y0, y0+yStep, ... < y1 only (bounded image-wide sample for the CPU baseline) */
void svr_oracle_pathtrace_strided(const svr_oracle_scene* scene, uint32_t traceDepth, uint32_t frameNo0, uint32_t nFrames,
                                  uint32_t strideW, uint32_t y0, uint32_t y1, uint32_t yStep, float* hdr, uint64_t* counters);

/* pathtracer.cu:282-290 + tonemapping.h:13-27 */
void svr_oracle_tonemap(const float* hdr, float exposure, uint64_t npix, uint8_t* outU8);

/* building blocks exposed for unit tests */
float svr_oracle_tex3d(const svr_oracle_scene* scene, float u, float v, float w);
void svr_oracle_tf(const svr_oracle_scene* scene, float intensity, float* rgba);
uint32_t svr_oracle_wang_hash(uint32_t a);
void svr_oracle_xorwow_uniforms(uint64_t seed, uint32_t n, float* out);

#ifdef __cplusplus
}
#endif
#endif
