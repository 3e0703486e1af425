// svr_pathtrace.cu -- Monte Carlo volumetric path tracer behind render_pathtracer
// (pathtracer.h:17; kernel_pathtracer / hdr_to_ldr / clear_hdr_buffer, pathtracer.cu:86-94, 200-304).
//
// One launch renders a BATCH of samples per pixel: each lane owns a pixel, walks its samples one
// after the other without waiting for its neighbours, keeps the radiance sum in registers, and
// at the end merges once into the caller's accumulator and writes the tone-mapped pixel -- the
// reference's three launches per sample (clear, trace + 12-byte RMW, tone map; pathtracer.cu:
// 297-303) collapse into one launch per batch.
//
// Estimator modes (SVR_OPT_PT_MODE):
//   0  reference twin: global majorant tf.maxOpacity (woodcock_tracking.h:28-30), XORWOW stream
//      seeded wangHash(frameNo) + pixel with the reference's draw order -> path-for-path the same
//      walk as kernel_pathtracer up to fast-math / FMA-contraction rounding.
//   1  global majorant + Philox counter RNG (same estimator, different random stream).
//   2  local majorants: 3-D DDA over the macrocell grid (svr_macrocell.cu); within a cell the free
//      path is sampled against the cell's majorant, optical depth carries across cell faces, cells
//      with majorant 0 cost no fetch.  Delta tracking with any valid majorant samples the same
//      free-path distribution, so the expectation is unchanged.
// Shadow rays: binary delta-tracking estimate as the reference (transmittance.h:10-17), or ratio
// tracking (SVR_OPT_SHADOW_ESTIMATOR = 1).
#include "svr_rng.cuh"
#include "svr_state.h"

namespace svr {
Counters* device_counters();

namespace {

struct PtLaunch {
    uint32_t traceDepth;
    uint32_t firstSample;  // == frameNo of the first sample in the batch
    uint32_t nSamples;
    uint32_t y0, y1;       // row range
    float* hdr;            // packed vec3 running mean (reference ABI) or null
    uint32_t* img;         // tone-mapped u8vec4 or null
    float4* sum;           // rgb = sum of samples, w = sample count (multi-GPU partials) or null
    int32_t clearSum;
};

template <int MODE>
struct RngOf {
    typedef Philox type;
};
template <>
struct RngOf<0> {
    typedef XorwowCompat type;
};

// ---------------------------------------------------------------------------------------------
// free-path sampling
// ---------------------------------------------------------------------------------------------

// woodcock_tracking.h:20-51, global majorant.  Returns the collision distance or -FLT_MAX; tMin/tMax
// are the values the reference leaves in the (mutable) ray.
template <bool COUNT, class Rng>
SVR_DEV float track_global(const DevScene& s, const Ray& ray, Rng& rng, float* tMinOut, float* tMaxOut,
                           LocalCounters<COUNT>& lc, int slot)
{
    float tNear, tFar;
    if (!intersect_volume(s.vol, ray, &tNear, &tFar)) return -FLT_MAX;
    const float tMin = tNear < 0.f ? 1e-6f : tNear;
    const float tMax = tFar;
    *tMinOut = tMin;
    *tMaxOut = tMax;
    float t = tMin;
    const float sigmaMax = s.tf.maxOpacity;
    const float invSigmaMax = 1.f / sigmaMax;
    const float invSigmaMaxSampleInterval = 1.f / (sigmaMax * 1.f);  // BASE_SAMPLE_STEP_SIZE 1
    while (true) {
        t += -logf(rng.next_one_minus()) * invSigmaMaxSampleInterval;
        if (t > tMax) return -FLT_MAX;
        float intensity = intensity_at(s.vol, ray.orig + t * ray.dir);
        float sigma_t = tf_at(s.tf, intensity).w;
        lc.add(slot, 1);
        lc.add(SVR_CNT_TF_LOOKUPS, 1);
        if (rng.next() < sigma_t * invSigmaMax || t > tMax) break;
    }
    return t;
}

// Macrocell DDA state for a ray clipped to [tMin, tMax].
struct Dda {
    float tNx, tNy, tNz;  // ray parameter of the next cell face per axis
    float dtx, dty, dtz;  // parameter distance between faces
    int cx, cy, cz;
    int sx, sy, sz;
    SVR_DEV void init(const DevScene& s, const Ray& ray, float t)
    {
        const float3 toCell = f3(s.vol.bbox.invSize) * s.grid.scale;
        const float3 g0 = (ray.orig - f3(s.vol.bbox.vmin)) * toCell;
        const float3 dg = ray.dir * toCell;
        const float3 g = g0 + t * dg;
        cx = min(max((int)floorf(g.x), 0), s.grid.gx - 1);
        cy = min(max((int)floorf(g.y), 0), s.grid.gy - 1);
        cz = min(max((int)floorf(g.z), 0), s.grid.gz - 1);
        sx = dg.x > 0.f ? 1 : -1;
        sy = dg.y > 0.f ? 1 : -1;
        sz = dg.z > 0.f ? 1 : -1;
        const float ix = 1.f / dg.x, iy = 1.f / dg.y, iz = 1.f / dg.z;
        tNx = dg.x != 0.f ? ((float)(cx + (dg.x > 0.f ? 1 : 0)) - g0.x) * ix : FLT_MAX;
        tNy = dg.y != 0.f ? ((float)(cy + (dg.y > 0.f ? 1 : 0)) - g0.y) * iy : FLT_MAX;
        tNz = dg.z != 0.f ? ((float)(cz + (dg.z > 0.f ? 1 : 0)) - g0.z) * iz : FLT_MAX;
        dtx = dg.x != 0.f ? fabsf(ix) : FLT_MAX;
        dty = dg.y != 0.f ? fabsf(iy) : FLT_MAX;
        dtz = dg.z != 0.f ? fabsf(iz) : FLT_MAX;
    }
    SVR_DEV float exit_t() const { return fminf(fminf(tNx, tNy), tNz); }
    // step across the nearest face; false when the ray leaves the grid
    SVR_DEV bool step(const DevGrid& g)
    {
        if (tNx <= tNy && tNx <= tNz) {
            cx += sx;
            tNx += dtx;
            return (unsigned)cx < (unsigned)g.gx;
        }
        if (tNy <= tNz) {
            cy += sy;
            tNy += dty;
            return (unsigned)cy < (unsigned)g.gy;
        }
        cz += sz;
        tNz += dtz;
        return (unsigned)cz < (unsigned)g.gz;
    }
    SVR_DEV float majorant(const DevGrid& g) const
    {
        return __ldg(&g.majorant[((size_t)cz * g.gy + cy) * g.gx + cx]);
    }
};

// Delta tracking with per-macrocell majorants.  Same contract as track_global.
template <bool COUNT, class Rng>
SVR_DEV float track_local(const DevScene& s, const Ray& ray, Rng& rng, float* tMinOut, float* tMaxOut,
                          LocalCounters<COUNT>& lc, int slot)
{
    float tNear, tFar;
    if (!intersect_volume(s.vol, ray, &tNear, &tFar)) return -FLT_MAX;
    const float tMin = tNear < 0.f ? 1e-6f : tNear;
    const float tMax = tFar;
    *tMinOut = tMin;
    *tMaxOut = tMax;
    float t = tMin;
    Dda dda;
    dda.init(s, ray, t);
    float tau = -logf(rng.next_one_minus());  // optical depth still to travel
    while (true) {
        const float tExit = fminf(dda.exit_t(), tMax);
        const float sig = dda.majorant(s.grid);
        lc.add(SVR_CNT_CELLS, 1);
        const float d = fmaxf(tExit - t, 0.f) * sig;
        if (tau >= d) {
            // leaves the cell before colliding
            tau -= d;
            t = tExit;
            if (tExit >= tMax || !dda.step(s.grid)) return -FLT_MAX;
            continue;
        }
        t += tau / sig;
        float intensity = intensity_at(s.vol, ray.orig + t * ray.dir);
        float sigma_t = tf_at(s.tf, intensity).w;
        lc.add(slot, 1);
        lc.add(SVR_CNT_TF_LOOKUPS, 1);
        if (rng.next() * sig < sigma_t) return t;
        tau = -logf(rng.next_one_minus());
    }
}

template <int MODE, bool COUNT, class Rng>
SVR_DEV float sample_distance(const DevScene& s, const Ray& ray, Rng& rng, float* tMin, float* tMax,
                              LocalCounters<COUNT>& lc, int slot)
{
    if (MODE == 2) return track_local<COUNT>(s, ray, rng, tMin, tMax, lc, slot);
    return track_global<COUNT>(s, ray, rng, tMin, tMax, lc, slot);
}

// transmittance.h:10-17: 1 if a tracked flight from `start` toward `end` leaves the volume box
// (the segment is not clipped at the light), else 0.
template <int MODE, bool COUNT, class Rng>
SVR_DEV float transmittance_binary(const DevScene& s, float3 start, float3 end, Rng& rng, LocalCounters<COUNT>& lc)
{
    Ray ray;
    ray.orig = start;
    ray.dir = normalize(end - start);
    float tMin = 1e-6f, tMax = FLT_MAX;
    float t = sample_distance<MODE, COUNT>(s, ray, rng, &tMin, &tMax, lc, SVR_CNT_SHADOW_TAPS);
    bool flag = (t > tMin) && (t < tMax);
    return flag ? 0.f : 1.f;
}

// Ratio tracking over the same segment: T = prod(1 - sigma/majorant) at tentative collisions;
// same expectation as the binary estimator, lower variance, but walks the whole segment.
template <int MODE, bool COUNT, class Rng>
SVR_DEV float transmittance_ratio(const DevScene& s, float3 start, float3 end, Rng& rng, LocalCounters<COUNT>& lc)
{
    Ray ray;
    ray.orig = start;
    ray.dir = normalize(end - start);
    float tNear, tFar;
    if (!intersect_volume(s.vol, ray, &tNear, &tFar)) return 1.f;
    float t = tNear < 0.f ? 1e-6f : tNear;
    const float tMax = tFar;
    float T = 1.f;
    if (MODE == 2) {
        Dda dda;
        dda.init(s, ray, t);
        float tau = -logf(rng.next_one_minus());
        while (true) {
            const float tExit = fminf(dda.exit_t(), tMax);
            const float sig = dda.majorant(s.grid);
            lc.add(SVR_CNT_CELLS, 1);
            const float d = fmaxf(tExit - t, 0.f) * sig;
            if (tau >= d) {
                tau -= d;
                t = tExit;
                if (tExit >= tMax || !dda.step(s.grid)) return T;
                continue;
            }
            t += tau / sig;
            float sigma_t = tf_at(s.tf, intensity_at(s.vol, ray.orig + t * ray.dir)).w;
            lc.add(SVR_CNT_SHADOW_TAPS, 1);
            lc.add(SVR_CNT_TF_LOOKUPS, 1);
            T *= 1.f - sigma_t / sig;
            if (T < 0.02f) {  // Russian roulette on a nearly opaque segment
                if (rng.next() * 0.02f >= T) return 0.f;
                T = 0.02f;
            }
            tau = -logf(rng.next_one_minus());
        }
    } else {
        const float sigmaMax = s.tf.maxOpacity;
        const float inv = 1.f / sigmaMax;
        while (true) {
            t += -logf(rng.next_one_minus()) * inv;
            if (t > tMax) return T;
            float sigma_t = tf_at(s.tf, intensity_at(s.vol, ray.orig + t * ray.dir)).w;
            lc.add(SVR_CNT_SHADOW_TAPS, 1);
            lc.add(SVR_CNT_TF_LOOKUPS, 1);
            T *= 1.f - sigma_t * inv;
            if (T < 0.02f) {
                if (rng.next() * 0.02f >= T) return 0.f;
                T = 0.02f;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// shading (pathtracer.cu:96-198)
// ---------------------------------------------------------------------------------------------
struct VolumeSample {  // cuda_volume.h:124-132
    float3 ptInWorld, wo;
    float3 gradient;
    float4 color_opacity;
};

enum ShadingType { ISOTROPIC, BRDF };

// pathtracer.cu:106-131
SVR_DEV float3 bsdf(const VolumeSample& vs, float3 wi, ShadingType st)
{
    float3 diffuseColor = f3(vs.color_opacity.x, vs.color_opacity.y, vs.color_opacity.z);
    if (st == ISOTROPIC) return diffuseColor * hg_phase_f();
    float3 normal = normalize(vs.gradient);
    normal = dot(vs.wo, normal) < 0.f ? -normal : normal;
    float cosTerm = fmaxf(0.f, dot(wi, normal));
    float ks = schlick_fresnel(1.0f, SVR_IOR, cosTerm);
    float kd = 1.f - ks;
    float3 diffuse = diffuseColor * lambert_f();
    float3 specular = f3(1.f) * microfacet_f(wi, vs.wo, normal, SVR_IOR, SVR_ALPHA);
    return (kd * diffuse + ks * specular) * cosTerm;
}

// pathtracer.cu:133-169
template <bool EXACT_PI, class Rng>
SVR_DEV float3 sample_bsdf(const VolumeSample& vs, float3* wi, float* pdf, Rng& rng, ShadingType st)
{
    float3 color = f3(vs.color_opacity.x, vs.color_opacity.y, vs.color_opacity.z);
    if (st == ISOTROPIC) {
        hg_phase_sample<EXACT_PI>(vs.wo, wi, pdf, rng);
        return color * hg_phase_f();
    }
    float3 normal = normalize(vs.gradient);
    float cosTerm = dot(vs.wo, normal);
    if (cosTerm < 0.f) {
        cosTerm = -cosTerm;
        normal = -normal;
    }
    float ks = schlick_fresnel(1.f, SVR_IOR, cosTerm);
    float kd = 1.f - ks;
    float p = 0.25f + 0.5f * ks;
    if (rng.next() < p) {
        microfacet_sample(vs.wo, normal, SVR_ALPHA, wi, pdf, rng);
        float f = microfacet_f(*wi, vs.wo, normal, SVR_IOR, SVR_ALPHA);
        return f3(1.f) * f * ks / p;
    }
    lambert_sample<EXACT_PI>(normal, wi, pdf, rng);
    return color * lambert_f() * kd / (1.f - p);
}

// pathtracer.cu:171-198
template <int MODE, bool COUNT, class Rng>
SVR_DEV float3 estimate_direct_light(const DevScene& s, const VolumeSample& vs, Rng& rng, ShadingType st,
                                     LocalCounters<COUNT>& lc)
{
    constexpr bool EXACT_PI = MODE == 0;
    if (s.numLights == 0) return f3(0.f);
    int lightId = (int)((float)s.numLights * rng.next());
    lightId = lightId < (int)s.numLights ? lightId : (int)s.numLights - 1;
    const svr_area_light& light = s.lights[lightId];
    float3 lightPos, wi;
    float pdf;
    float3 Li = sample_light<EXACT_PI>(light, vs.ptInWorld, rng, &lightPos, &wi, &pdf);
    if (pdf > 0.f && max3(Li) > 0.f) {
        float Tr = s.shadowEstimator ? transmittance_ratio<MODE, COUNT>(s, vs.ptInWorld, lightPos, rng, lc)
                                     : transmittance_binary<MODE, COUNT>(s, vs.ptInWorld, lightPos, rng, lc);
        return (Tr * (float)s.numLights) * bsdf(vs, wi, st) * Li / pdf;
    }
    return f3(0.f);
}

// pathtracer.cu:96-103
template <class Rng>
SVR_DEV bool russian_roulette(float3* T, Rng& rng)
{
    float illum = 0.2126f * T->x + 0.7152f * T->y + 0.0722f * T->z;
    if (rng.next() > illum) return true;
    *T = *T / illum;
    return false;
}

// pathtracer.cu:200-278: one path; `offset` = idy * rowStride + idx seeds the reference stream
template <int MODE, bool COUNT>
SVR_DEV float3 trace_path(const DevScene& s, uint32_t idx, uint32_t idy, uint32_t offset, uint32_t sample,
                          uint32_t traceDepth, LocalCounters<COUNT>& lc)
{
    constexpr bool EXACT_PI = MODE == 0;
    typename RngOf<MODE>::type rng;
    rng.init(s.seedKey, offset, sample);

    float3 L = f3(0.f), T = f3(1.f);
    Ray ray = camera_ray_jittered<EXACT_PI>(s.cam, idx, idy, rng);
    lc.add(SVR_CNT_PATHS, 1);

    LightHit ls;
    const bool hitLight = nearest_light(s, ray, &ls);
    for (uint32_t k = 0; k < traceDepth; ++k) {
        float tMin, tMax;
        float t = sample_distance<MODE, COUNT>(s, ray, rng, &tMin, &tMax, lc, SVR_CNT_TRACK_TAPS);
        if ((k == 0) && hitLight) {
            t = t < 0.f ? FLT_MAX : t;
            if (ls.t < t) {
                float cosTerm = dot(ls.normal, -ray.dir);
                L += T * ls.radiance * (cosTerm <= 0.f ? 0.f : 1.f);
                break;
            }
        }
        if (t < 0.f) {
            if (s.envEnabled) L += T * env_radiance(s.env, ray.dir);  // the line commented out at pathtracer.cu:233
            break;
        }

        VolumeSample vs;
        vs.wo = -ray.dir;
        vs.ptInWorld = ray.orig + t * ray.dir;
        float intensity = intensity_at(s.vol, vs.ptInWorld);
        vs.color_opacity = tf_at(s.tf, intensity);
        vs.gradient = gradient_at(s.vol, vs.ptInWorld);
        float gradientMagnitude = sqrtf(dot(vs.gradient, vs.gradient));
        lc.add(SVR_CNT_SHADE_TAPS, 7);
        lc.add(SVR_CNT_TF_LOOKUPS, 1);
        lc.add(SVR_CNT_SCATTERS, 1);

        float3 wi = f3(0.f);
        float pdf = 0.f;
        const float gf = s.vol.gradientFactor;
        const float Pbrdf = vs.color_opacity.w *
                            (1.f - expf(-25.f * gf * gf * gf * gradientMagnitude * 65535.f * s.vol.invMaxMagnitude));
        const ShadingType st = (rng.next() < Pbrdf) ? BRDF : ISOTROPIC;

        L += T * estimate_direct_light<MODE, COUNT>(s, vs, rng, st, lc);

        float3 f = sample_bsdf<EXACT_PI>(vs, &wi, &pdf, rng, st);
        float cosTerm = fabsf(dot(normalize(vs.gradient), wi));
        if (max3(f) > 0.f && pdf > 0.f) {
            if (st == ISOTROPIC)
                T *= f / (pdf * (1.f - Pbrdf));
            else
                T *= f * cosTerm / (pdf * Pbrdf);
        }
        ray.orig = vs.ptInWorld;
        ray.dir = wi;
        if (k >= 3) {
            if (russian_roulette(&T, rng)) break;
        }
    }
    return L;
}

template <int MODE, bool COUNT>
__global__ void __launch_bounds__(256) pathtrace_kernel(const __grid_constant__ DevScene s, const PtLaunch a, Counters* cnt)
{
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t idx = blockIdx.x * 16u + (warp & 1u) * 8u + (lane & 7u);
    const uint32_t idy = a.y0 + blockIdx.y * (blockDim.x >> 4) + (warp >> 1) * 4u + (lane >> 3);
    const bool inside = idx < s.cam.imageW && idy < a.y1;
    LocalCounters<COUNT> lc;
    if (inside) {
        const uint32_t offset = idy * s.cam.imageW + idx;
        float3 sum = f3(0.f);
        for (uint32_t n = 0; n < a.nSamples; ++n)
            sum += trace_path<MODE, COUNT>(s, idx, idy, offset, a.firstSample + n, a.traceDepth, lc);

        if (a.sum) {
            float4 prev = a.clearSum ? make_float4(0.f, 0.f, 0.f, 0.f) : a.sum[offset];
            a.sum[offset] = make_float4(prev.x + sum.x, prev.y + sum.y, prev.z + sum.z, prev.w + (float)a.nSamples);
        }
        if (a.hdr) {
            // running_estimate (pathtracer.cu:81-84) for one sample; its closed form for a batch
            float* h = a.hdr + 3 * (size_t)offset;
            const float N0 = (float)a.firstSample;
            float3 acc = a.firstSample == 0 ? f3(0.f) : f3(h[0], h[1], h[2]);  // frameNo==0 clears (pathtracer.cu:297-300)
            if (a.nSamples == 1)
                acc = acc + (sum - acc) / (N0 + 1.f);
            else
                acc = (acc * N0 + sum) / (N0 + (float)a.nSamples);
            h[0] = acc.x;
            h[1] = acc.y;
            h[2] = acc.z;
            if (a.img) {
                float3 l = tone_map(acc, s.cam.exposure);  // hdr_to_ldr, pathtracer.cu:282-290
                a.img[offset] = pack_u8x4(l.x * 255.f, l.y * 255.f, l.z * 255.f, 255.f);
            }
        }
    }
    lc.flush(cnt);
}

// root-side resolve of summed partials: hdr = rgb / w, tone map, both in one pass
__global__ void resolve_kernel(const float4* __restrict__ sum, float* __restrict__ hdr, uint32_t* __restrict__ img,
                               uint32_t npix, float exposure)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 v = sum[i];
    float inv = v.w > 0.f ? 1.f / v.w : 0.f;
    float3 acc = f3(v.x * inv, v.y * inv, v.z * inv);
    if (hdr) {
        hdr[3 * (size_t)i + 0] = acc.x;
        hdr[3 * (size_t)i + 1] = acc.y;
        hdr[3 * (size_t)i + 2] = acc.z;
    }
    if (img) {
        float3 l = tone_map(acc, exposure);
        img[i] = pack_u8x4(l.x * 255.f, l.y * 255.f, l.z * 255.f, 255.f);
    }
}

template <int MODE>
void launch_mode(dim3 grid, int block, cudaStream_t stream, const DevScene& sc, const PtLaunch& a, Counters* cnt)
{
    if (cnt) pathtrace_kernel<MODE, true><<<grid, block, 0, stream>>>(sc, a, cnt);
    else pathtrace_kernel<MODE, false><<<grid, block, 0, stream>>>(sc, a, cnt);
}

int launch_pathtrace(PtLaunch a)
{
    HostState& st = state();
    DevScene sc = st.scene;
    if (!sc.vol.tex || !sc.tf.tex) return fail_msg("render_pathtracer: setup_volume / setup_transferfunction not called");
    if (sc.cam.imageW == 0 || sc.cam.imageH == 0) return fail_msg("render_pathtracer: setup_camera not called");
    const int mode = st.options[SVR_OPT_PT_MODE];
    sc.envEnabled = st.options[SVR_OPT_ENV_ENABLED];
    sc.shadowEstimator = st.options[SVR_OPT_SHADOW_ESTIMATOR];
    sc.seedKey = wang_hash((uint32_t)st.options[SVR_OPT_SEED]);
    if (mode == 2) {
        int rc = ensure_grid(&sc, false);
        if (rc) return rc;
    } else {
        memset(&sc.grid, 0, sizeof(sc.grid));
    }
    if (a.y1 > sc.cam.imageH) a.y1 = sc.cam.imageH;
    if (a.y0 >= a.y1 || a.nSamples == 0) return 0;
    Counters* cnt = nullptr;
    if (st.options[SVR_OPT_COUNTERS]) {
        cnt = device_counters();
        if (!cnt) return fail_msg("render_pathtracer: counter allocation failed");
    }
    const int block = st.options[SVR_OPT_PT_BLOCK];
    const uint32_t tileH = (uint32_t)block / 16u;
    dim3 grid((sc.cam.imageW + 15u) / 16u, ((a.y1 - a.y0) + tileH - 1u) / tileH);
    switch (mode) {
        case 0: launch_mode<0>(grid, block, st.stream, sc, a, cnt); break;
        case 1: launch_mode<1>(grid, block, st.stream, sc, a, cnt); break;
        default: launch_mode<2>(grid, block, st.stream, sc, a, cnt); break;
    }
    count_launch();
    SVR_TRY(cudaGetLastError());
    return 0;
}

}  // namespace
}  // namespace svr

using namespace svr;

// pathtracer.h:17 / pathtracer.cu:292-304.  One call = one sample per pixel; asynchronous.
extern "C" void render_pathtracer(svr_u8vec4* img, const svr_render_params* renderParams)
{
    PtLaunch a;
    memset(&a, 0, sizeof(a));
    a.traceDepth = renderParams->traceDepth;
    a.firstSample = renderParams->frameNo;
    a.nSamples = 1;
    a.y0 = 0;
    a.y1 = 0xffffffffu;
    a.hdr = (float*)renderParams->hdrBuffer;
    a.img = (uint32_t*)img;
    int rc = launch_pathtrace(a);
    if (rc) {
        fprintf(stderr, "CUDA error at %s:%d code=%d \"%s\" \n", __FILE__, __LINE__, rc, svr_last_error());
        cudaDeviceReset();
        exit(EXIT_FAILURE);
    }
}

extern "C" int svr_render_pathtracer_spp(svr_u8vec4* img, const svr_render_params* renderParams, uint32_t spp)
{
    if (!renderParams || !renderParams->hdrBuffer) return fail_msg("svr_render_pathtracer_spp: hdrBuffer is null");
    PtLaunch a;
    memset(&a, 0, sizeof(a));
    a.traceDepth = renderParams->traceDepth;
    a.firstSample = renderParams->frameNo;
    a.nSamples = spp;
    a.y0 = 0;
    a.y1 = 0xffffffffu;
    a.hdr = (float*)renderParams->hdrBuffer;
    a.img = (uint32_t*)img;
    return launch_pathtrace(a);
}

extern "C" int svr_pathtracer_accumulate(svr_vec4* sum, uint32_t traceDepth, uint32_t firstSample, uint32_t nSamples, int clear)
{
    if (!sum) return fail_msg("svr_pathtracer_accumulate: sum is null");
    PtLaunch a;
    memset(&a, 0, sizeof(a));
    a.traceDepth = traceDepth;
    a.firstSample = firstSample;
    a.nSamples = nSamples;
    a.y0 = 0;
    a.y1 = 0xffffffffu;
    a.sum = (float4*)sum;
    a.clearSum = clear;
    return launch_pathtrace(a);
}

extern "C" int svr_pathtracer_resolve(svr_u8vec4* img, svr_vec3* hdrOut, const svr_vec4* sum)
{
    HostState& st = state();
    if (!sum) return fail_msg("svr_pathtracer_resolve: sum is null");
    uint32_t npix = st.scene.cam.imageW * st.scene.cam.imageH;
    if (!npix) return fail_msg("svr_pathtracer_resolve: setup_camera not called");
    resolve_kernel<<<(npix + 255u) / 256u, 256, 0, st.stream>>>((const float4*)sum, (float*)hdrOut, (uint32_t*)img, npix,
                                                               st.scene.cam.exposure);
    count_launch();
    SVR_TRY(cudaGetLastError());
    return 0;
}
