// TEST INFRASTRUCTURE -- scene builder of bench.py's reference arm (and nothing else).
//
// The reference arm times the reference's own unmodified kernels (oracle/_ref/libsvr_ref_<W>x<H>.so).  They need what
// the reference's host hands them: a 3-D cudaArray + texture object made as VolumeReader::CreateTextures makes them
// (core/VolumeReader.cpp:138-172), a 1-D float4 array + texture object as TransferFunction makes them
// (gui/transferfunction.cpp:30-44), and the cudaVolume fields VolumeReader::CreateDeviceVolume + Canvas::LoadVolume set
// (core/VolumeReader.cpp:174-185, gui/canvas.cpp:31-32).  This file builds exactly that with plain CUDA runtime calls,
// so that the arm's process never maps the product library: oracle/_ref/libsvr_refscene.so.
// The synthetic voxels come from the same device code the product's generator uses (svr_generate.cuh), so both arms
// of the bench render bit-identical volumes.  Mirrors tests/dropin/dropin_host.cu:62-107.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../include/svr_render.h"
#include "../sunvolumerender_b200/csrc/svr_generate.cuh"

using namespace svr_gen;

#define CK(x)                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) {                                                                      \
            fprintf(stderr, "ref_scene: %s: %s\n", #x, cudaGetErrorString(e_));                       \
            return (int)e_;                                                                           \
        }                                                                                             \
    } while (0)

static size_t bytes_of(int format) { return format == SVR_VOXEL_U8 ? 1 : (format == SVR_VOXEL_F32 ? 4 : 2); }

// Fills *vol and *tf.  `kind`, `format`, `seed` as svr_generate_volume; `tf_rgba` = n_tf x float4 on the host.
extern "C" int ref_scene_create(int kind, int format, uint32_t n, uint32_t seed, const float* tf_rgba, uint32_t n_tf, svr_volume* vol,
                                svr_transfer_function* tf)
{
    const size_t bpe = bytes_of(format), count = (size_t)n * n * n;
    void* lin = nullptr;
    CK(cudaMalloc(&lin, count * bpe));
    const int blocks = 148 * 8, threads = 256;
    unsigned int* dBits = nullptr;
    CK(cudaMalloc(&dBits, sizeof(unsigned int)));
    CK(cudaMemset(dBits, 0, sizeof(unsigned int)));
    cudaChannelFormatDesc ch;
    switch (format) {
        case SVR_VOXEL_U8:
            gen_kernel<uint8_t><<<blocks, threads>>>((uint8_t*)lin, kind, (int)n, seed);
            gradmax_kernel<uint8_t><<<blocks, threads>>>((const uint8_t*)lin, n, n, n, 0.5f, 0.5f, 0.5f, dBits);
            ch = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
            break;
        case SVR_VOXEL_U16:
            gen_kernel<uint16_t><<<blocks, threads>>>((uint16_t*)lin, kind, (int)n, seed);
            gradmax_kernel<uint16_t><<<blocks, threads>>>((const uint16_t*)lin, n, n, n, 0.5f, 0.5f, 0.5f, dBits);
            ch = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
            break;
        case SVR_VOXEL_F16:
            gen_kernel<__half><<<blocks, threads>>>((__half*)lin, kind, (int)n, seed);
            gradmax_kernel<__half><<<blocks, threads>>>((const __half*)lin, n, n, n, 0.5f, 0.5f, 0.5f, dBits);
            ch = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindFloat);
            break;
        default:
            gen_kernel<float><<<blocks, threads>>>((float*)lin, kind, (int)n, seed);
            gradmax_kernel<float><<<blocks, threads>>>((const float*)lin, n, n, n, 0.5f, 0.5f, 0.5f, dBits);
            ch = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
            break;
    }
    CK(cudaGetLastError());
    unsigned int bits = 0;
    CK(cudaMemcpy(&bits, dBits, sizeof(bits), cudaMemcpyDeviceToHost));
    cudaFree(dBits);
    float maxGrad;
    memcpy(&maxGrad, &bits, sizeof(float));
    if (!(maxGrad > 0.f)) maxGrad = 1.f;

    // ---- VolumeReader::CreateTextures (core/VolumeReader.cpp:138-172)
    cudaArray_t arr = nullptr;
    cudaExtent extent = make_cudaExtent(n, n, n);
    CK(cudaMalloc3DArray(&arr, &ch, extent));
    cudaMemcpy3DParms cp;
    memset(&cp, 0, sizeof(cp));
    cp.srcPtr = make_cudaPitchedPtr(lin, n * bpe, n, n);
    cp.dstArray = arr;
    cp.extent = extent;
    cp.kind = cudaMemcpyDeviceToDevice;
    CK(cudaMemcpy3D(&cp));
    CK(cudaFree(lin));
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = (format == SVR_VOXEL_U8 || format == SVR_VOXEL_U16) ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
    td.normalizedCoords = 1;
    cudaTextureObject_t tex = 0;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));

    // ---- VolumeReader::CreateDeviceVolume (core/VolumeReader.cpp:174-185) + Canvas::LoadVolume (gui/canvas.cpp:31-32), spacing 1
    memset(vol, 0, sizeof(*vol));
    const float half = (float)n - (float)n * 0.5f;
    vol->bbox.vmin = {-half, -half, -half};
    vol->bbox.vmax = {half, half, half};
    vol->bbox.invSize = {1.f / (half + half), 1.f / (half + half), 1.f / (half + half)};
    vol->tex = tex;
    vol->densityScale = 1.f;
    vol->invMaxMagnitude = 1.f / maxGrad;
    vol->gradientFactor = 0.5f;
    vol->spacing = {1.f, 1.f, 1.f};
    vol->invSpacing = {1.f, 1.f, 1.f};
    vol->x_clip = vol->y_clip = vol->z_clip = {-1.f, 1.f};

    // ---- TransferFunction (gui/transferfunction.cpp:17-44)
    cudaChannelFormatDesc ch4 = cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindFloat);
    cudaArray_t tfArr = nullptr;
    CK(cudaMallocArray(&tfArr, &ch4, n_tf));
    CK(cudaMemcpy2DToArray(tfArr, 0, 0, tf_rgba, sizeof(float) * 4 * n_tf, sizeof(float) * 4 * n_tf, 1, cudaMemcpyHostToDevice));
    rd.res.array.array = tfArr;
    cudaTextureDesc td1;
    memset(&td1, 0, sizeof(td1));
    td1.addressMode[0] = cudaAddressModeClamp;
    td1.filterMode = cudaFilterModeLinear;
    td1.normalizedCoords = 1;
    td1.readMode = cudaReadModeElementType;
    cudaTextureObject_t tfTex = 0;
    CK(cudaCreateTextureObject(&tfTex, &rd, &td1, nullptr));
    float maxOpacity = 0.f;
    for (uint32_t i = 0; i < n_tf; ++i) maxOpacity = maxOpacity > tf_rgba[4 * i + 3] ? maxOpacity : tf_rgba[4 * i + 3];
    memset(tf, 0, sizeof(*tf));
    tf->tex = tfTex;
    tf->maxOpacity = maxOpacity;
    CK(cudaDeviceSynchronize());
    return 0;
}

extern "C" int ref_scene_destroy(svr_volume* vol, svr_transfer_function* tf)
{
    cudaDeviceSynchronize();
    cudaResourceDesc rd;
    if (vol && vol->tex && cudaGetTextureObjectResourceDesc(&rd, vol->tex) == cudaSuccess) {
        cudaDestroyTextureObject(vol->tex);
        cudaFreeArray(rd.res.array.array);
        vol->tex = 0;
    }
    if (tf && tf->tex && cudaGetTextureObjectResourceDesc(&rd, tf->tex) == cudaSuccess) {
        cudaDestroyTextureObject(tf->tex);
        cudaFreeArray(rd.res.array.array);
        tf->tex = 0;
    }
    return 0;
}

// Copies the voxels behind vol->tex to the host (tests: both arms' volumes are bit-identical).
extern "C" int ref_scene_download(const svr_volume* vol, void* host_out, uint64_t bytes)
{
    cudaResourceDesc rd;
    CK(cudaGetTextureObjectResourceDesc(&rd, vol->tex));
    cudaChannelFormatDesc ch;
    cudaExtent ext;
    unsigned int flags = 0;
    CK(cudaArrayGetInfo(&ch, &ext, &flags, rd.res.array.array));
    const size_t bpe = (size_t)(ch.x + ch.y + ch.z + ch.w) / 8;
    if (bytes != (uint64_t)ext.width * ext.height * ext.depth * bpe) return -1;
    cudaMemcpy3DParms cp;
    memset(&cp, 0, sizeof(cp));
    cp.srcArray = rd.res.array.array;
    cp.extent = ext;
    cp.kind = cudaMemcpyDeviceToHost;
    cp.dstPtr = make_cudaPitchedPtr(host_out, ext.width * bpe, ext.width, ext.height);
    CK(cudaMemcpy3D(&cp));
    return 0;
}
