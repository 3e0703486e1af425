// dropin_host.cu -- TEST INFRASTRUCTURE.  A host program written the way gui/canvas.cpp is: it
// includes the REFERENCE's own headers (pathtracer.h, raycasting.h, core/*.h -- from /root/reference
// at build time, never copied), builds the scene with the reference's own classes and setters, and
// calls the seven entry points through the reference's own C++ prototypes.  The same object code is
// linked twice (oracle/Makefile, target `dropin`):
//     dropin_host_ref   against oracle/_ref/libsvr_ref_64x64.so   (the reference's kernels)
//     dropin_host_b200  against sunvolumerender_b200/libsvr_b200.so (this repository)
// so switching implementations is a change of the link line and nothing else -- the drop-in claim
// of INTEGRATION.md, checked by tests/test_gpu_dropin.py on the images both binaries write.
//
//   dropin_host_<x> <out.bin>      writes: raycast u8vec4 image | hdrBuffer (vec3 floats) | path-traced u8vec4 image
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define GLM_FORCE_NO_CTOR_INIT
#include <glm/glm.hpp>

#include "core/cuda_volume.h"
#include "core/lights/cuda_arealight.h"
#include "core/lights/cuda_environment_light.h"
#include "pathtracer.h"
#include "raycasting.h"

#ifndef WIDTH
#error "build with -DWIDTH=.. -DHEIGHT=.. (the reference's canvas size is compile-time, common.h:8-9)"
#endif

static void die(cudaError_t e, int line)
{
    if (e != cudaSuccess) {
        fprintf(stderr, "dropin_host: CUDA error %s at line %d\n", cudaGetErrorString(e), line);
        exit(1);
    }
}
#define CK(x) die((x), __LINE__)

int main(int argc, char** argv)
{
    const char* outPath = argc > 1 ? argv[1] : "dropin_out.bin";
    const int N = 32, W = WIDTH, H = HEIGHT;

    // ---- a small CT-like u16 volume: soft body, dense shell, air outside
    std::vector<unsigned short> vox((size_t)N * N * N);
    double maxGrad = 0.0;
    auto at = [&](int x, int y, int z) -> double {
        double px = (x + 0.5 - N / 2.0) / (N / 2.0), py = (y + 0.5 - N / 2.0) / (N / 2.0), pz = (z + 0.5 - N / 2.0) / (N / 2.0);
        double e = sqrt(px * px / 0.64 + py * py / 0.45 + pz * pz / 0.72);
        if (e >= 1.0) return 0.0;
        double v = e > 0.8 ? 0.85 : 0.35 + 0.1 * sin(9.0 * px) * cos(7.0 * py + 3.0 * pz);
        return v;
    };
    for (int z = 0; z < N; ++z)
        for (int y = 0; y < N; ++y)
            for (int x = 0; x < N; ++x) vox[((size_t)z * N + y) * N + x] = (unsigned short)(at(x, y, z) * 65535.0 + 0.5);
    for (int z = 1; z < N - 1; ++z)
        for (int y = 1; y < N - 1; ++y)
            for (int x = 1; x < N - 1; ++x) {
                auto v = [&](int a, int b, int c) { return (double)vox[((size_t)c * N + b) * N + a]; };
                double gx = 0.5 * (v(x + 1, y, z) - v(x - 1, y, z)), gy = 0.5 * (v(x, y + 1, z) - v(x, y - 1, z)),
                       gz = 0.5 * (v(x, y, z + 1) - v(x, y, z - 1));
                maxGrad = fmax(maxGrad, sqrt(gx * gx + gy * gy + gz * gz));
            }

    // ---- textures with the descriptors of VolumeReader::CreateTextures (core/VolumeReader.cpp:138-172)
    cudaChannelFormatDesc ch = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
    cudaArray_t volArr;
    CK(cudaMalloc3DArray(&volArr, &ch, make_cudaExtent(N, N, N)));
    cudaMemcpy3DParms cp = {0};
    cp.srcPtr = make_cudaPitchedPtr(vox.data(), N * sizeof(unsigned short), N, N);
    cp.dstArray = volArr;
    cp.extent = make_cudaExtent(N, N, N);
    cp.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&cp));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = volArr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    cudaTextureObject_t volTex;
    CK(cudaCreateTextureObject(&volTex, &rd, &td, nullptr));

    // ---- transfer function as TransferFunction uploads it (gui/transferfunction.cpp:17-44)
    const int TN = 1024;
    std::vector<float> table(4 * TN);
    float maxOpacity = 0.f;
    for (int i = 0; i < TN; ++i) {
        float x = (float)i / (TN - 1);
        table[4 * i + 0] = 0.9f - 0.6f * x;
        table[4 * i + 1] = 0.3f + 0.5f * x;
        table[4 * i + 2] = 0.2f + 0.7f * x * x;
        table[4 * i + 3] = 0.5f * fminf(x / 0.1f, 1.f);
        maxOpacity = fmaxf(maxOpacity, table[4 * i + 3]);
    }
    cudaChannelFormatDesc ch4 = cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindFloat);
    cudaArray_t tfArr;
    CK(cudaMallocArray(&tfArr, &ch4, TN));
    CK(cudaMemcpy2DToArray(tfArr, 0, 0, table.data(), sizeof(float) * 4 * TN, sizeof(float) * 4 * TN, 1, cudaMemcpyHostToDevice));
    rd.res.array.array = tfArr;
    cudaTextureDesc td1 = {};
    td1.addressMode[0] = cudaAddressModeClamp;
    td1.filterMode = cudaFilterModeLinear;
    td1.normalizedCoords = 1;
    td1.readMode = cudaReadModeElementType;
    cudaTextureObject_t tfTex;
    CK(cudaCreateTextureObject(&tfTex, &rd, &td1, nullptr));

    // ---- the reference's own scene classes, set up as VolumeReader::CreateDeviceVolume + Canvas do
    glm::vec3 half(0.5f * N);
    cudaVolume volume;
    volume.Set(cudaBBox(-half, half), glm::vec3(1.f), volTex);
    volume.SetClipPlane(glm::vec2(-1.f, 1.f), glm::vec2(-1.f, 1.f), glm::vec2(-1.f, 1.f));
    volume.SetDensityScale(1.f);
    volume.SetInvMaxMagnitude((float)(1.0 / maxGrad));
    volume.SetGradientFactor(0.5f);
    cudaTransferFunction tf;
    tf.Set(tfTex, maxOpacity);
    float eyeDist = 1.5f * N / (2.f * tanf(45.f * 0.5f * (float)M_PI / 180.f));  // gui/canvas.cpp:191-197
    cudaCamera camera(glm::vec3(0.f, 0.f, eyeDist), glm::vec3(1, 0, 0), glm::vec3(0, 1, 0), glm::vec3(0, 0, 1), 45.f, 0.f, 1.f, 1.f, W, H);
    cudaAreaLight lights[2];
    float R = 0.5f * sqrtf(3.f) * N;
    lights[0].Set(cudaDisk(glm::vec3(0.f, 1.5f * R + 1.f, 0.f), glm::vec3(0.f, -1.f, 0.f), 2.5f), glm::vec3(1.f), 500.f);
    lights[1].Set(cudaDisk(glm::vec3(30.f, 10.f, 25.f), glm::normalize(glm::vec3(-30.f, -10.f, -25.f)), 3.f), glm::vec3(1.f, 0.6f, 0.3f), 200.f);
    cudaEnvironmentLight env;
    env.Set(glm::vec3(0.5f));

    setup_volume(volume);
    setup_transferfunction(tf);
    setup_camera(camera);
    setup_env_lights(env);
    setup_area_lights(lights, 2);

    glm::u8vec4* img;
    CK(cudaMalloc(&img, (size_t)W * H * 4));
    CK(cudaMemset(img, 0, (size_t)W * H * 4));
    std::vector<unsigned char> rc((size_t)W * H * 4), ptImg((size_t)W * H * 4);
    std::vector<float> hdr((size_t)W * H * 3);

    // ---- Canvas::paintGL, ray-casting branch (gui/canvas.cpp:92)
    render_raycasting(img, volume, tf, camera, 0.5f * sqrtf(3.f));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(rc.data(), img, rc.size(), cudaMemcpyDeviceToHost));

    // ---- Canvas::paintGL, path-tracing branch: one call per frame, frameNo++ (gui/canvas.cpp:96,116)
    RenderParams params;
    params.SetupHDRBuffer(W, H);
    params.traceDepth = 2;
    for (params.frameNo = 0; params.frameNo < 4; ++params.frameNo) {
        render_pathtracer(img, params);
        CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(ptImg.data(), img, ptImg.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hdr.data(), params.hdrBuffer, hdr.size() * sizeof(float), cudaMemcpyDeviceToHost));

    FILE* f = fopen(outPath, "wb");
    if (!f) return 2;
    fwrite(rc.data(), 1, rc.size(), f);
    fwrite(hdr.data(), sizeof(float), hdr.size(), f);
    fwrite(ptImg.data(), 1, ptImg.size(), f);
    fclose(f);
    double s = 0;
    for (float v : hdr) s += v;
    printf("dropin_host: %dx%d, mean radiance %.6f, wrote %s\n", W, H, s / hdr.size(), outPath);
    return 0;
}
