// svr_generate.cuh -- synthetic procedural volumes and the gradient-magnitude reduction of the
// loaders (VolumeReader.cpp:70-76), as device code shared by the library (svr_api.cu) and by the reference arm's scene
// builder (oracle/ref_scene.cu), so that both arms of bench.py render bit-identical voxels.  None of this is reference
// code: the reference ships no data and no generator.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/svr_render.h"  // svr_volume_kind, svr_voxel_format

namespace svr_gen {


__device__ __forceinline__ uint32_t hash3(int x, int y, int z, uint32_t seed)
{
    uint32_t h = seed * 0x9E3779B1u;
    h ^= (uint32_t)x * 0x85EBCA77u;
    h = (h << 13) | (h >> 19);
    h ^= (uint32_t)y * 0xC2B2AE3Du;
    h = (h << 13) | (h >> 19);
    h ^= (uint32_t)z * 0x27D4EB2Fu;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    h *= 0x297A2D39u;
    h ^= h >> 15;
    return h;
}

__device__ __forceinline__ float lattice(int x, int y, int z, uint32_t seed)
{
    return (float)(hash3(x, y, z, seed) >> 8) * (1.f / 16777216.f);
}

// trilinear value noise with smoothstep weights, in [0,1)
__device__ float value_noise(float x, float y, float z, uint32_t seed)
{
    float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    float a = x - fx, b = y - fy, c = z - fz;
    a = a * a * (3.f - 2.f * a);
    b = b * b * (3.f - 2.f * b);
    c = c * c * (3.f - 2.f * c);
    float v000 = lattice(ix, iy, iz, seed), v100 = lattice(ix + 1, iy, iz, seed);
    float v010 = lattice(ix, iy + 1, iz, seed), v110 = lattice(ix + 1, iy + 1, iz, seed);
    float v001 = lattice(ix, iy, iz + 1, seed), v101 = lattice(ix + 1, iy, iz + 1, seed);
    float v011 = lattice(ix, iy + 1, iz + 1, seed), v111 = lattice(ix + 1, iy + 1, iz + 1, seed);
    float x00 = v000 + a * (v100 - v000), x10 = v010 + a * (v110 - v010);
    float x01 = v001 + a * (v101 - v001), x11 = v011 + a * (v111 - v011);
    float y0 = x00 + b * (x10 - x00), y1 = x01 + b * (x11 - x01);
    return y0 + c * (y1 - y0);
}

__device__ float fbm(float x, float y, float z, int octaves, uint32_t seed)
{
    float sum = 0.f, amp = 0.5f, norm = 0.f;
    for (int o = 0; o < octaves; ++o) {
        sum += amp * value_noise(x, y, z, seed + (uint32_t)o * 101u);
        norm += amp;
        x *= 2.f;
        y *= 2.f;
        z *= 2.f;
        amp *= 0.5f;
    }
    return sum / norm;
}

__device__ float density_at(int kind, int n, int x, int y, int z, uint32_t seed)
{
    float c = 0.5f * (float)n;
    float px = (float)x + 0.5f - c, py = (float)y + 0.5f - c, pz = (float)z + 0.5f - c;
    if (kind == SVR_GEN_SPHERE) {
        // C1: rho = clamp(1 - r / (0.45 N), 0, 1)
        float r = sqrtf(px * px + py * py + pz * pz);
        return fminf(fmaxf(1.f - r / (0.45f * (float)n), 0.f), 1.f);
    }
    float qx = px / c, qy = py / c, qz = pz / c;  // [-1, 1]
    if (kind == SVR_GEN_CT) {
        // C2/C3/C5: nested ellipsoid shells -- skin 0.25, soft tissue 0.45, bone 0.85 -- plus three
        // octaves of value noise (amplitude 0.05) inside the body; air is exactly 0
        float e = sqrtf(qx * qx / (0.80f * 0.80f) + qy * qy / (0.62f * 0.62f) + qz * qz / (0.88f * 0.88f));
        if (e >= 1.f) return 0.f;
        float v = e > 0.93f ? 0.25f : 0.45f;
        float b = sqrtf(qx * qx / (0.46f * 0.46f) + qy * qy / (0.36f * 0.36f) + qz * qz / (0.60f * 0.60f));
        if (b < 1.f && b > 0.78f) v = 0.85f;
        // two small dense inclusions ("vertebrae")
        float dx = qx - 0.18f, dy = qy + 0.1f, dz = qz - 0.2f;
        if (dx * dx + dy * dy + dz * dz < 0.01f) v = 0.85f;
        dx = qx + 0.2f, dy = qy - 0.05f, dz = qz + 0.3f;
        if (dx * dx + dy * dy + dz * dz < 0.008f) v = 0.85f;
        float s = 8.f;
        float nz3 = fbm(qx * s + 17.f, qy * s + 5.f, qz * s + 11.f, 3, seed);
        v += 0.05f * (2.f * nz3 - 1.f);
        return fminf(fmaxf(v, 0.f), 1.f);
    }
    // C4 cloud: 5-octave fBm shaped by a sphere mask
    float r = sqrtf(qx * qx + qy * qy + qz * qz);
    float mask = fminf(fmaxf((0.85f - r) / 0.35f, 0.f), 1.f);
    float f = fbm(qx * 4.f + 3.f, qy * 4.f + 7.f, qz * 4.f + 13.f, 5, seed);
    float d = (f - 0.42f) * 3.2f * mask;
    return fminf(fmaxf(d, 0.f), 1.f);
}

template <typename T>
__device__ __forceinline__ T encode(float d);
template <>
__device__ __forceinline__ uint8_t encode<uint8_t>(float d) { return (uint8_t)(d * 255.f + 0.5f); }
template <>
__device__ __forceinline__ uint16_t encode<uint16_t>(float d) { return (uint16_t)(d * 65535.f + 0.5f); }
template <>
__device__ __forceinline__ __half encode<__half>(float d) { return __float2half_rn(d); }
template <>
__device__ __forceinline__ float encode<float>(float d) { return d; }

template <typename T>
__global__ void gen_kernel(T* out, int kind, int n, uint32_t seed)
{
    size_t total = (size_t)n * n * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % n), y = (int)((i / n) % n), z = (int)(i / ((size_t)n * n));
        out[i] = encode<T>(density_at(kind, n, x, y, z, seed));
    }
}

template <typename T>
__device__ __forceinline__ float raw_value(const T* d, size_t i);
template <>
__device__ __forceinline__ float raw_value<uint8_t>(const uint8_t* d, size_t i) { return (float)d[i] * 257.f; }  // as u16
template <>
__device__ __forceinline__ float raw_value<uint16_t>(const uint16_t* d, size_t i) { return (float)d[i]; }
template <>
__device__ __forceinline__ float raw_value<__half>(const __half* d, size_t i) { return __half2float(d[i]) * 65535.f; }
template <>
__device__ __forceinline__ float raw_value<float>(const float* d, size_t i) { return d[i] * 65535.f; }

// max |central-difference gradient| of the raw values, interior voxels, spacing-scaled
// (VolumeReader.cpp:70-76: vtkImageGradientMagnitude on the short data, then the maximum)
template <typename T>
__global__ void gradmax_kernel(const T* d, int nx, int ny, int nz, float hx, float hy, float hz, unsigned int* outBits)
{
    size_t total = (size_t)nx * ny * nz;
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((size_t)nx * ny));
        int x0 = max(x - 1, 0), x1 = min(x + 1, nx - 1);
        int y0 = max(y - 1, 0), y1 = min(y + 1, ny - 1);
        int z0 = max(z - 1, 0), z1 = min(z + 1, nz - 1);
        size_t row = ((size_t)z * ny + y) * nx, col = (size_t)z * ny * nx + x;
        float gx = (raw_value<T>(d, row + x1) - raw_value<T>(d, row + x0)) * hx;
        float gy = (raw_value<T>(d, col + (size_t)y1 * nx) - raw_value<T>(d, col + (size_t)y0 * nx)) * hy;
        float gz = (raw_value<T>(d, ((size_t)z1 * ny + y) * nx + x) - raw_value<T>(d, ((size_t)z0 * ny + y) * nx + x)) * hz;
        m = fmaxf(m, sqrtf(gx * gx + gy * gy + gz * gz));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(outBits, __float_as_uint(m));  // m >= 0: bit order == value order
}


}  // namespace svr_gen
