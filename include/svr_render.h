/*
 * svr_render.h -- C ABI of libsvr_b200.so, the B200-native drop-in for the render hot path of
 * SunVolumeRender (pathtracer.cu + raycasting.cu).
 *
 * Part 1 is the reference's own boundary: the seven unmangled symbols a host (gui/canvas.cpp or
 * tools/svr_headless.cpp) links against.  The reference declares them `extern "C"` with C++
 * reference parameters; a reference and a pointer are the same thing at the ABI level, so the
 * prototypes below are binary-identical to pathtracer.h:17-24 and raycasting.h:8 and a host
 * compiled against the reference headers links to this library unchanged.
 *
 * Part 2 is the headless extension the reference lacks: batched samples-per-pixel, float
 * outputs for parity checks, sum-buffers for the multi-GPU split, resource builders that mirror
 * the reference's loaders, synthetic volume generators and tap counters.
 *
 * All pointers named `img`, `hdrBuffer`, `sum`, `out`, `dev*` are DEVICE pointers unless the name
 * says host.  No torch types, no C++ types.
 */
#ifndef SVR_RENDER_H
#define SVR_RENDER_H

#include "svr_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Part 1 -- the reference boundary (error convention = utils/helper_cuda.h:967-981: print
 * "CUDA error at file:line ...", cudaDeviceReset(), exit(EXIT_FAILURE); no return codes).
 * ------------------------------------------------------------------------------------------ */

/* pathtracer.h:17 / pathtracer.cu:292-304.  One call = one sample per pixel with seed stream
 * `frameNo`; frameNo == 0 clears hdrBuffer; hdrBuffer holds the running mean afterwards and img
 * the tone-mapped image.  Asynchronous on the library stream (svr_set_stream). */
void render_pathtracer(svr_u8vec4* img, const svr_render_params* renderParams);

/* pathtracer.h:20-24 / pathtracer.cu:34-68.  Copy the POD into library-owned device state.
 * setup_volume / setup_transferfunction also (re)build the macrocell majorant grid. */
void setup_volume(const svr_volume* vol);
void setup_transferfunction(const svr_transfer_function* tf);
void setup_camera(const svr_camera* cam);
void setup_env_lights(const svr_env_light* light);
void setup_area_lights(svr_area_light* lights, uint32_t n); /* n is clamped to 8 */

/* raycasting.h:8 / raycasting.cu:69-75.  Deterministic front-to-back compositing; takes its
 * scene by argument, not from setup_*.  stepSize = 0.5*|spacing| in the reference caller
 * (core/VolumeReader.cpp:198-201, gui/canvas.cpp:92); the march step is stepSize*0.5. */
void render_raycasting(svr_u8vec4* img, svr_volume* volume, svr_transfer_function* transferFunction,
                       svr_camera* camera, float stepSize);

/* ------------------------------------------------------------------------------------------
 * Part 2 -- headless extension.  Functions returning int: 0 = ok, non-zero = error
 * (svr_last_error() gives the text).  They never call exit().
 * ------------------------------------------------------------------------------------------ */

#define SVR_VERSION 100

int svr_version(void);
const char* svr_last_error(void);

/* Stream all launches and copies are issued on (a cudaStream_t; NULL = legacy default stream). */
int svr_set_stream(void* cuda_stream);
/* Bind the calling thread to a device and (re)initialise library state there. */
int svr_set_device(int device);

enum svr_option {
    /* path tracer estimator:
     *   0 = reference-compatible: global majorant, XORWOW stream seeded wangHash(frameNo)+pixel
     *       in the reference's draw order (path-exact twin of kernel_pathtracer);
     *   1 = global majorant, Philox4x32-10 counter RNG keyed (seed, pixel, sample);
     *   2 = local majorants (macrocell DDA) + Philox (default, fastest; same expectation). */
    SVR_OPT_PT_MODE = 0,
    /* shadow-ray estimator: 0 = binary delta tracking (transmittance.h:10-17), 1 = ratio tracking */
    SVR_OPT_SHADOW_ESTIMATOR = 1,
    /* 0 = escaped paths add nothing (pathtracer.cu:233 is commented out); 1 = add envLight */
    SVR_OPT_ENV_ENABLED = 2,
    /* macrocell edge in voxels: 0 (default) = chosen from the scene, about twice the mean free path inside
     * the medium, re-evaluated when volume, transfer function or density scale change; or a power of two in 2..64 */
    SVR_OPT_MACROCELL_SIZE = 3,
    /* ray caster empty-space skipping through the macrocell grid: 0 off, 1 on (default) */
    SVR_OPT_RC_SKIP = 4,
    /* Philox key */
    SVR_OPT_SEED = 5,
    /* 1 = kernels also count taps / lookups / steps (slower; used for bytes_algo) */
    SVR_OPT_COUNTERS = 6,
    /* threads per block of the path tracer / ray caster (tuning) */
    SVR_OPT_PT_BLOCK = 7,
    SVR_OPT_RC_BLOCK = 8,
    /* path-tracer kernel shape: 2 = sample-parallel warp (the lanes of a warp take different samples of
     * the same pixel), 1 = megakernel (one pixel per lane, samples one after the other),
     * 0 = phase-scheduled warp (generate / march / collide / event / bounce phases, the warp
     * votes each round and runs the phase most lanes wait in), 3 = sample-parallel warp with the scatter
     * queue at every depth (see SVR_OPT_PT_QUEUE_MIN_DEPTH), 4 = majorant-profile kernel (see SVR_OPT_PT_PROFILE),
     * 5 = ray pool + event queue per warp: lanes take rays and scatter events from per-warp queues instead of owning a path */
    SVR_OPT_PT_KERNEL = 9,
    /* phase-scheduled kernel: macrocell visits per MARCH round (0 = default 4) */
    SVR_OPT_PT_ROUNDS = 10,
    /* 1 (default) = empty macrocells record how far the empty space around them extends and rays
     * leap over it; 0 = one cell at a time.  Images are bit-identical either way. */
    SVR_OPT_LEAP = 11,
    /* 1 (default) = each pixel finds once, along its centre ray, how far the camera rays of all its
     * samples can be advanced through empty macrocells; 0 = every sample walks from the volume face.
     * Images are bit-identical either way (only empty cells are skipped). */
    SVR_OPT_PT_ENTRY_CACHE = 12,
    /* sample-parallel kernel (SVR_OPT_PT_KERNEL = 2): pixels a warp renders one after the other (1..64; 0, the
     * default, lets the library choose by launch length: 1 from 128 samples per launch on, 2 below), and the
     * smallest batch (samples per pixel per launch) the kernel is used for; smaller batches run the megakernel.
     * Images differ from the other shapes only in float summation order. */
    SVR_OPT_PT_WARP_PIXELS = 13,
    SVR_OPT_PT_WARP_MIN_SPP = 14,
    /* SVR_OPT_PT_KERNEL = 2 with local majorants (SVR_OPT_PT_MODE = 2): from this traceDepth on the
     * sample-parallel kernel runs with a per-warp queue of scatter events in shared memory (kernel shape 3:
     * camera-ray rounds and scatter-event rounds that each start with all 32 lanes busy, however unequal the
     * path lengths).  Default 8; 0 = never.  Images differ from shape 2 only in float summation order. */
    SVR_OPT_PT_QUEUE_MIN_DEPTH = 15,
    /* 1 (default) = setup_volume / _transferfunction / _camera / _env_lights return after cudaDeviceSynchronize,
     * as the reference's do (pathtracer.cu:34-55).  0 = they only store the PODs (this library passes the scene
     * as kernel parameters, there is no device copy to wait for): a streaming host whose copy streams must keep
     * running across the setup calls of the next frame turns the synchronisation off. */
    SVR_OPT_SETUP_SYNC = 16,
    /* 1 (default) = a pixel whose camera rays provably pass every area-light disk skips get_nearest_light_sample
     * (pathtracer.cu:214-215); 0 = every camera ray tests every disk, as the reference does.  Images are
     * bit-identical either way (the cull is conservative). */
    SVR_OPT_PT_LIGHT_CULL = 17,
    /* 1 = with local majorants and a pinhole camera the sample-parallel shapes track camera rays against a per-pixel
     * majorant PROFILE the warp builds once per pixel (kernel shape 4: no lane walks a camera ray, every tentative
     * collision of a camera ray runs with the lanes packed; scatter queue as shape 3).  Same estimator, other random
     * walks: images agree with shapes 1-3 statistically.  0 (default) = shapes 2 / 3 as selected: the profile kernel
     * executes fewer cell visits but measured slower on every BASELINE configuration. */
    SVR_OPT_PT_PROFILE = 18,
    /* shape 4: idle lanes take the pixel's next camera samples together once this many lanes wait (1..32, 0 = default 8) */
    SVR_OPT_PT_REFILL = 19,
    /* kernel shape 5 (ray pool + event queue per warp, SVR_OPT_PT_KERNEL = 5): pixels in a warp's run (1..16, 0 = default 16);
     * rays of all of them share the warp's pool */
    SVR_OPT_PT_POOL_PIXELS = 20,
    /* 1 = the environment light (when SVR_OPT_ENV_ENABLED) is a NEXT-EVENT target beside the area lights: at every scatter
     * event one of (area lights + environment) is picked uniformly; the environment is sampled by importance from its
     * luminance (a 512 x 256 direction grid built from the map, its offset and intensity; rebuilt when they change) and
     * weighted with both lobes of the shading model; bounce rays that leave the medium then add nothing.  Same
     * expectation as collecting the sky on escape (0, default: what the reference's commented-out line would do), far
     * lower variance for maps with small bright features.  Ignored by the reference-twin mode and kernel shape 5. */
    SVR_OPT_ENV_NEE = 21,
    /* sample-parallel kernel (shape 2): 1 = the warps of a block share one row of pixels and split every pixel's samples
     * between them (batches of at least 32 samples per warp); 0 (default) = one row per warp.  A pixel then occupies a warp
     * for a quarter of the time: shorter blocks, a shorter end of the launch -- for short frames (one frame divided over
     * several GPUs).  The image differs from 0 only in the order of four float additions per pixel. */
    SVR_OPT_PT_BLOCK_SPLIT = 22,
    /* Every pixel's classification -- how far its camera rays can skip empty space, whether they can hit a light, whether the
     * pixel is all sky -- is made for the whole image by a kernel of its own (one lane per pixel) and read by the render kernels
     * (shapes 1-3).  1 (default) = it is kept across launches -- the frames of a progressive render, the batches of an
     * accumulation -- and made again only after something a pixel can see has changed (any setup_*, upload or option);
     * 0 = made again for every launch.  Images are bit-identical either way (only empty space is skipped). */
    SVR_OPT_PT_PIXEL_CACHE = 23,
    /* 1 (default) = svr_volume_upload from a DEVICE buffer into an array made by svr_volume_create, once the macrocell grid
     * exists for that array, runs as one kernel that stores the voxels into the array and reduces the grid's value ranges from
     * the same pass (streamed time series: a new volume per frame).  0 = copy, then rebuild the ranges from the array at the
     * next render.  Arrays, ranges and images are bit-identical either way. */
    SVR_OPT_FUSED_UPLOAD = 24,
    /* render_pathtracer (one sample per call): once a progressive render has reached frame 16 with nothing changed, the next 32
     * samples of every pixel are computed in ONE launch of the sample-parallel kernel and kept; the following calls only fold the
     * kept sample of their frame into the running mean and tone-map (the same arithmetic on the same sample values: hdrBuffer and
     * image are bit-identical, frame for frame, to computing one sample per call).  Any setup_*, upload or option, a frameNo other
     * than the next one, another hdrBuffer, size or traceDepth discards what was kept.  Value = the batch (default 32, at most
     * 32); > 0: the library times one single-sample launch, the first batch and one fold per scene and stops batching for that
     * scene when it does not pay (views in which the volume fills the frame); < 0: always batch -value samples; 0, 1, -1 = off.
     * Not used with SVR_OPT_COUNTERS, nor where the scatter-queue kernel would run (traceDepth >= SVR_OPT_PT_QUEUE_MIN_DEPTH).
     * Memory: batch x pixels x 12 bytes (at most 1 GiB; larger canvases get smaller batches). */
    SVR_OPT_PT_LOOKAHEAD = 25,
    SVR_OPT_COUNT_
};
/* Defaults can also come from the environment, read once at first use, for hosts that only know the
 * seven reference entry points: SVR_PT_MODE, SVR_SHADOW_ESTIMATOR, SVR_ENV_ENABLED, SVR_RC_SKIP,
 * SVR_SEED, SVR_PT_KERNEL (same values as the options). */
int svr_set_option(int key, int value);
int svr_get_option(int key);

/* Equivalent to `spp` consecutive render_pathtracer calls with frameNo, frameNo+1, ... but in one
 * launch: samples accumulate in registers, hdrBuffer is read and written once, the tone map
 * runs once.  frameNo == 0 clears first.  img may be NULL (no tone map). */
int svr_render_pathtracer_spp(svr_u8vec4* img, const svr_render_params* renderParams, uint32_t spp);

/* Multi-GPU building blocks.  `sum` is imageW*imageH float4: rgb = sum of
 * radiance samples, w = number of samples.  Samples [firstSample, firstSample+nSamples) are
 * rendered; clear != 0 zeroes `sum` first.  After an NCCL sum-reduce of the per-GPU buffers the
 * root calls svr_pathtracer_resolve: hdr = rgb / w (optional packed-vec3 output) and the
 * tone-mapped image (tonemapping.h:13-27), fused in one pass. */
int svr_pathtracer_accumulate(svr_vec4* sum, uint32_t traceDepth, uint32_t firstSample, uint32_t nSamples, int clear);
/* The IMAGE split of the multi-GPU path tracer ("interleaved image tiles, no reduce needed, only a
 * gather"): as svr_pathtracer_accumulate, for the row bands phase, phase + stride, ... only (bands of *bandRows rows, the
 * chosen kernel's block height, counted from row 0); rank r of N passes (r, N).  Pixels outside the call's bands are not
 * touched, so N ranks writing into ONE buffer (peer mappings of rank 0's buffer, svr_stage_*, svr_peer_*) assemble the frame
 * without a reduce, and every pixel's samples are summed by one GPU in the single-GPU order: the frame is bit-identical to
 * a single-GPU render.  Preferable to the sample split when there are few samples per GPU (per-pixel set-up is divided
 * too). */
int svr_pathtracer_accumulate_bands(svr_vec4* sum, uint32_t traceDepth, uint32_t firstSample, uint32_t nSamples, int clear,
                                    uint32_t phase, uint32_t stride, uint32_t* bandRows);
int svr_pathtracer_resolve(svr_u8vec4* img, svr_vec3* hdrOut, const svr_vec4* sum);

/* Staging buffers for streaming volumes to replicated scenes (the volume is replicated
 * per GPU; the reference is single-device, main.cpp:7-33, and has no counterpart).  One process per GPU:
 * every process allocates linear staging buffers (svr_stage_alloc = cudaMalloc), exports them to the
 * process that uploads (svr_stage_export = cudaIpcGetMemHandle, 64 bytes to send over any channel), which
 * imports them from ITS device (svr_stage_import = cudaIpcOpenMemHandle with lazy peer access, so the
 * mapping is a peer mapping over NVLink) and fills them with svr_stage_copy (cudaMemcpyAsync, kind
 * default: pinned host, local or imported peer memory -- copy engines only, no SMs, so the transfer runs
 * beside a render kernel that occupies every SM).  svr_volume_upload(vol, stage, 1) then moves the staged
 * voxels into the volume's cudaArray.  `stream` is a cudaStream_t. */
int svr_stage_alloc(void** dev_ptr, uint64_t bytes);
int svr_stage_free(void* dev_ptr);
int svr_stage_export(const void* dev_ptr, unsigned char handle_out[64]);
int svr_stage_import(const unsigned char handle[64], void** peer_ptr);
int svr_stage_release(void* peer_ptr);
int svr_stage_copy(void* dst, const void* src, uint64_t bytes, void* stream);

/* Completion flags between the GPUs of one box (ray casting on N GPUs without a collective): every rank renders its
 * bands straight into rank 0's image through a peer mapping of that buffer (svr_stage_alloc / _export / _import; pass
 * the mapped pointer as `img`), then raises rank 0's flag with svr_peer_signal (a peer mapping of an 8-byte,
 * zero-initialised svr_stage_alloc buffer: word 0 counts signals, word 1 is set when a wait timed out); rank 0's stream
 * waits with svr_peer_wait until word 0 has reached `expected` (wrap-safe; gives up after timeout_ms).  Both are
 * stream-ordered launches on the library's stream: no host synchronisation, no NCCL. */
int svr_peer_signal(void* peer_flag);
int svr_peer_wait(void* flag, uint32_t expected, uint32_t timeout_ms);

/* Ray caster variants: float RGBA before quantisation (parity is checked on these), and a row
 * range [y0, y1) for the image-tile split across GPUs.  out/img are full-frame buffers. */
int svr_render_raycasting_f32(svr_vec4* out, const svr_volume* volume, const svr_transfer_function* tf,
                              const svr_camera* camera, float stepSize);
int svr_render_raycasting_rows(svr_u8vec4* img, svr_vec4* outOrNull, const svr_volume* volume,
                               const svr_transfer_function* tf, const svr_camera* camera, float stepSize,
                               uint32_t y0, uint32_t y1);

/* The balanced split: the image is cut into bands of *bandRows rows (the kernel's block height), numbered from
 * the middle of the image outwards (0 = the middle band, 1 = the one below it, 2 = the one above, ...), and this
 * call renders bands phase, phase + stride, phase + 2 stride, ... -- rank r of N passes (r, N).  Contiguous row blocks
 * give the ranks that see the body several times the work of the ranks that see its margins; interleaved bands
 * do not.  Rows outside the call's bands are left untouched (zero the image first and sum-reduce the u8 images:
 * disjoint bands make the sum a gather; or let every rank write into one image through peer mappings, svr_peer_*).
 * The headless ray-cast entry points (_f32, _rows, _bands) rebuild the empty-space majorants when the transfer function's
 * HANDLE changes or an edit was announced (setup_transferfunction, svr_tf_upload); only render_raycasting itself also
 * checks the table's CONTENT on every call (the reference's host may change it behind an unchanged handle). */
int svr_render_raycasting_bands(svr_u8vec4* img, svr_vec4* outOrNull, const svr_volume* volume,
                                const svr_transfer_function* tf, const svr_camera* camera, float stepSize,
                                uint32_t phase, uint32_t stride, uint32_t* bandRows);

/* ---- resource builders (the input contract of the path) ---- */
enum svr_voxel_format { SVR_VOXEL_U8 = 0, SVR_VOXEL_U16 = 1, SVR_VOXEL_F16 = 2, SVR_VOXEL_F32 = 3 };

/* Mirrors VolumeReader::CreateTextures + CreateDeviceVolume (core/VolumeReader.cpp:138-185):
 * 3-D cudaArray, border addressing, linear filter, normalised-float reads (element reads for
 * f16/f32), normalised coordinates; bbox = +-0.5*dim*spacing; invMaxMagnitude = 1/maxGradMag
 * (maxGradMag <= 0: computed on the device by central differences on the raw values, f16/f32
 * scaled by 65535).  Clip planes (-1,1), densityScale 1, gradientFactor 0.5 (gui/canvas.cpp:19,31-32).
 * `data` is x-fastest; data_on_device selects the memcpy kind. */
int svr_volume_create(svr_volume* out, const void* data, int data_on_device, int format,
                      uint32_t nx, uint32_t ny, uint32_t nz, float sx, float sy, float sz, float maxGradMag);
int svr_volume_destroy(svr_volume* vol);
/* Re-uploads voxels (same dims and format, x-fastest) into the array behind vol->tex -- the second
 * half of VolumeReader::CreateTextures (core/VolumeReader.cpp:146-161) for a volume that is already
 * bound -- and drops the macrocell cache built from the old contents.  Asynchronous on the library
 * stream when `data` is device or pinned host memory. */
int svr_volume_upload(const svr_volume* vol, const void* data, int data_on_device);
/* Drop cached macrocell data (keyed on the cudaArray handle); call after re-uploading voxels
 * into an existing array. */
int svr_volume_invalidate_cache(void);

/* Mirrors TransferFunction::TransferFunction (gui/transferfunction.cpp:17-44): 1024 float4
 * (r,g,b,opacity) -> 1-D cudaArray texture, linear, clamp, normalised; maxOpacity = max opacity. */
int svr_tf_create(svr_transfer_function* out, const float* host_rgba, uint32_t n);
int svr_tf_destroy(svr_transfer_function* tf);
/* New table contents (same size) into the array behind tf->tex; updates tf->maxOpacity.  The caller
 * then publishes the struct with setup_transferfunction, as after any edit. */
int svr_tf_upload(svr_transfer_function* tf, const float* host_rgba, uint32_t n);

/* Mirrors Lights::SetEnvironmentLight (core/lights/lights.cpp:31-75): w x h float4 lat-long map. */
int svr_env_create(svr_env_light* out, const float* host_rgba, uint32_t w, uint32_t h);
int svr_env_destroy(svr_env_light* env);

/* ---- synthetic volumes, written x-fastest into device memory ---- */
enum svr_volume_kind { SVR_GEN_SPHERE = 0, SVR_GEN_CT = 1, SVR_GEN_CLOUD = 2 };
int svr_generate_volume(void* dev_out, int kind, int format, uint32_t n, uint32_t seed);
/* max over voxels of the central-difference gradient magnitude (VolumeReader.cpp:70-76 semantics) */
int svr_max_gradient_magnitude(const void* dev_data, int format, uint32_t nx, uint32_t ny, uint32_t nz,
                               float sx, float sy, float sz, float* host_out);

/* ---- counters (valid while SVR_OPT_COUNTERS = 1) ---- */
enum svr_counter {
    SVR_CNT_TRACK_TAPS = 0,   /* volume taps in primary/secondary tracking loops */
    SVR_CNT_SHADOW_TAPS = 1,  /* volume taps in shadow-ray tracking */
    SVR_CNT_SHADE_TAPS = 2,   /* volume taps at scatter events / ray-cast steps (1 + 6 gradient) */
    SVR_CNT_TF_LOOKUPS = 3,   /* transfer-function lookups */
    SVR_CNT_SCATTERS = 4,     /* scatter events */
    SVR_CNT_PATHS = 5,        /* paths (pixel samples) or rays */
    SVR_CNT_CELLS = 6,        /* macrocells visited */
    SVR_CNT_STEPS = 7,        /* ray-cast steps executed (incl. skipped: see SKIPPED) */
    SVR_CNT_SKIPPED = 8,      /* ray-cast steps skipped as provably empty */
    SVR_CNT_COUNT_ = 16
};
int svr_counters_reset(void);
int svr_counters_read(uint64_t* host_out, uint32_t n);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
uint64_t svr_launch_count(void);
/* How many svr_volume_upload calls took the one-pass path (SVR_OPT_FUSED_UPLOAD) since the library was loaded. */
uint64_t svr_fused_upload_count(void);
/* How many look-ahead batches (SVR_OPT_PT_LOOKAHEAD) render_pathtracer has launched since the library was loaded. */
uint64_t svr_lookahead_batch_count(void);

/* Gather-roofline microbenchmarks: `taps_per_thread` dependent-free tex3D taps per thread over
 * the bound volume, coherent (ray-like) or random.  Returns taps issued via *host_taps. */
int svr_microbench_taps(const svr_volume* vol, int random, uint32_t threads, uint32_t taps_per_thread,
                        float* dev_sink, uint64_t* host_taps);

/* Layout study behind the choice of voxel storage: the same taps as svr_microbench_taps done in
 * software -- eight loads + the trilinear filter in the SM -- from a linear copy of n^3 16-bit voxels (layout 1, x fastest) or
 * from a bricked copy (layout 2: 8^3-voxel bricks, each contiguous, bricks in Morton order; svr_layout_brick makes it, n a
 * multiple of 8).  Not used by any render path. */
int svr_layout_brick(const void* dev_linear, void* dev_bricked, uint32_t n);
int svr_microbench_soft_taps(const void* dev_voxels16, uint32_t n, int layout, int random, uint32_t threads,
                             uint32_t taps_per_thread, float* dev_sink, uint64_t* host_taps);

/* ---- inspection hooks (used by the parity tests; cheap, never on the render path) ---- */
/* Raw fetches through the caller's texture objects, exactly as the kernels issue them:
 * out[i] = tex3D<float>(vol->tex, uvw[3i], uvw[3i+1], uvw[3i+2]) (normalised coordinates, before
 * densityScale); rgba[4i..] = tex1D<float4>(tf->tex, x[i]).  All pointers are device pointers. */
int svr_debug_sample_volume(const svr_volume* vol, const float* dev_uvw, uint32_t n, float* dev_out);
int svr_debug_sample_tf(const svr_transfer_function* tf, const float* dev_x, uint32_t n, float* dev_rgba);
/* Macrocell grid built for the scene of the last render call: dims[3] = cells per axis, *cell = edge
 * in voxels.  svr_grid_copy copies majorants (dims[0]*dims[1]*dims[2] floats, x fastest) and/or
 * the (min,max) intensity ranges (2 floats per cell) to HOST memory; either may be NULL. */
int svr_grid_info(int32_t* dims, int32_t* cell);
int svr_grid_copy(float* host_majorant, float* host_range);

#ifdef __cplusplus
}
#endif

#endif /* SVR_RENDER_H */
