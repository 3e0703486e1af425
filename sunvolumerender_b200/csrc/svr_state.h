// svr_state.h -- host-side state shared by the translation units of libsvr_b200.so.
//
// Like the reference (one set of __constant__ scene PODs per process, pathtracer.cu:34-68) the
// library holds ONE scene per process and is not re-entrant; all work is issued on one stream.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/svr_render.h"
#include "svr_scene.cuh"

namespace svr {

struct HostState {
    cudaStream_t stream = nullptr;
    DevScene scene;  // what setup_* stored (zero-initialised in state())
    int options[SVR_OPT_COUNT_];

    // macrocell cache -- callee-owned derived data, keyed on the caller's cudaArray handle
    cudaArray_t gridArray = nullptr;      // volume array the range grid was built from
    int gridCell = 0;
    int3 gridDims = {0, 0, 0};
    int3 volDims = {0, 0, 0};
    float2* dRange = nullptr;
    float* dMajorant = nullptr;              // padded: (gx+2)(gy+2)(gz+2), one empty cell around the grid
    uint8_t* dDist[2] = {nullptr, nullptr};  // ping-pong Chebyshev distance to the nearest non-empty cell
    int* dOcc = nullptr;                     // bounding box of the non-empty cells: lo xyz, hi xyz
    int majorantLeap = -1;
    float* dTfSparse = nullptr;           // range-max sparse table over the TF opacity
    float4* dTfTable = nullptr;           // linear copy of the TF array
    int tfEntries = 0;
    cudaTextureObject_t volPointTex = 0;  // point-sampled view of gridArray
    cudaSurfaceObject_t uploadSurf = 0;   // store view of uploadSurfArray (svr_volume_upload from a device buffer, svr_macrocell.cu)
    cudaArray_t uploadSurfArray = nullptr;
    unsigned long long fusedUploads = 0;
    void* hMailbox = nullptr;             // 64 bytes of mapped pinned memory: small results reach the host without a copy engine
    void* dMailbox = nullptr;
    bool rangeValid = false;              // false until the range grid reflects the array's current voxels
    bool fingerprintDue = false;          // setup_volume was called since the voxels were last looked at
    unsigned long long* dFingerprint = nullptr;  // [0] sampled hash of the voxels the ranges were built from, [1] scratch
    bool majorantValid = false;           // false after setup_volume/setup_transferfunction
    float majorantDensityScale = 0.f;
    cudaArray_t majorantTfArray = nullptr;
    unsigned long long* dTfHash = nullptr;  // [0] hash of the TF table the majorants were built from, [1] of the live table, [2] stale flag
    cudaEvent_t tfCheckEvent = nullptr;     // behind the ray caster's per-call hash of the live table (svr_macrocell.cu: tf_check_*)
    bool tfCheckPending = false;

    // automatic macrocell size (SVR_OPT_MACROCELL_SIZE = 0): the choice and the scene it was made for
    int autoCell = 0;
    cudaArray_t autoArray = nullptr, autoTfArray = nullptr;
    float autoDensityScale = 0.f;
    unsigned autoEpoch = 0, uploadEpoch = 0;  // uploadEpoch counts svr_volume_upload / svr_tf_upload calls
    void* dStats = nullptr;

    // importance sampler of the environment light (svr_env_io.cu), keyed on what it was built from
    float* dEnvMarg = nullptr;
    float* dEnvCond = nullptr;
    svr_env_light envSamplerKey = {};
    bool envSamplerValid = false;

    // Per-pixel classification (classify_pixel: entry skip, light cull, all-sky flag) kept across the 1-sample-per-call frames
    // of a progressive render: valid while sceneEpoch has not moved since it was stored
    float2* dPixelCache = nullptr;
    size_t pixelCacheCap = 0;
    unsigned long long sceneEpoch = 1, pixelCacheEpoch = 0;  // sceneEpoch: bumped by everything that can change what a pixel sees
    uint32_t pixelCacheW = 0, pixelCacheH = 0;
    int pixelCacheMode = -1;

    // Sample look-ahead of the 1-sample-per-call protocol (SVR_OPT_PT_LOOKAHEAD, svr_pathtrace.cu): the radiance of the samples
    // [aheadFirst, aheadFirst + aheadCount) of every pixel, computed in one sample-parallel launch; aheadNext is the frame the
    // next call must ask for
    float* dAhead = nullptr;
    size_t aheadCapFloats = 0;
    uint32_t aheadFirst = 0, aheadCount = 0, aheadNext = 0, aheadW = 0, aheadH = 0, aheadDepth = 0;
    unsigned long long aheadEpoch = 0, aheadBatches = 0;
    const void* aheadHdr = nullptr;
    // does it pay for this scene?  timed once per scene epoch: a single-sample launch [0,1], the first batch [2,3], one fold [4,5]
    cudaEvent_t aheadEv[6] = {};
    bool aheadEvReady = false, aheadTimeFold = false;
    unsigned long long aheadSingleEpoch = 0, aheadTimedEpoch = 0, aheadOffEpoch = 0;
    uint32_t aheadTimedBatch = 0;

    Counters* dCounters = nullptr;
    unsigned long long launches = 0;
    std::string lastError;
};

HostState& state();

// largest empty-space leap, in cells, a macrocell can record (svr_macrocell.cu stage 3)
#define SVR_LEAP_CAP 15

// drops the macrocell cache (range grid, majorants, distances, point-sampled view)
void release_grid(HostState& st);

// Part-1 error convention, utils/helper_cuda.h:967-981: print, cudaDeviceReset, exit(EXIT_FAILURE)
inline void check_fatal(cudaError_t e, const char* what, const char* file, int line)
{
    if (e != cudaSuccess) {
        fprintf(stderr, "CUDA error at %s:%d code=%d(%s) \"%s\" \n", file, line, (int)e, cudaGetErrorName(e), what);
        cudaDeviceReset();
        exit(EXIT_FAILURE);
    }
}
#define SVR_FATAL(x) ::svr::check_fatal((x), #x, __FILE__, __LINE__)

// Part-2 error convention: record the text, return non-zero
int fail(const char* where, cudaError_t e);
int fail_msg(const char* msg);
#define SVR_TRY(x)                                             \
    do {                                                       \
        cudaError_t e_ = (x);                                  \
        if (e_ != cudaSuccess) return ::svr::fail(#x, e_);     \
    } while (0)

// Builds / refreshes the macrocell grid for (vol, tf) and fills scene->grid, scene->volDim.
// `force` rebuilds the majorants even when the cache key matches (ray caster: the TF content can
// change behind an unchanged handle, gui/transferfunction.cpp:128-151).
// `maxAutoCell` caps the automatic cell size: delta tracking through a thin medium wants large cells, the
// ray caster (which only skips empty space with the grid) never gains from cells above 8 voxels.
int ensure_grid(DevScene* scene, bool force, int maxAutoCell = 32);

// true when the transfer-function table behind `tf` differs from the one the majorants were built from (one small
// launch + an 16-byte read-back; the ray caster's drop-in entry point, whose host may edit the table behind an unchanged
// handle, gui/transferfunction.cpp:128-151).  Also true when no majorants exist yet.
int tf_check_collect(bool* changed);
int tf_check_launch(const svr_transfer_function& tf, const unsigned int** staleFlag);
int upload_with_ranges(cudaArray_t arr, const cudaChannelFormatDesc& ch, const cudaExtent& ext, unsigned int flags, const void* devData, bool* done);

// Builds / refreshes the environment light's importance sampler for scene->env and fills scene->envS.
int ensure_env_sampler(DevScene* scene);

inline void count_launch(int n = 1) { state().launches += (unsigned long long)n; }

}  // namespace svr
