// svr_pathtrace.cu -- Monte Carlo volumetric path tracer behind render_pathtracer
// (pathtracer.h:17; kernel_pathtracer / hdr_to_ldr / clear_hdr_buffer, pathtracer.cu:86-94, 200-304).
//
// One launch renders a BATCH of samples per pixel: each lane owns a pixel, walks its samples one
// after the other, keeps the radiance sum in registers, and at the end merges once into the
// caller's accumulator and writes the tone-mapped pixel -- the reference's three launches per
// sample (clear, trace + 12-byte RMW, tone map; pathtracer.cu:297-303) collapse into one launch
// per batch.
//
// Estimator modes (SVR_OPT_PT_MODE):
//   0  reference twin: global majorant tf.maxOpacity (woodcock_tracking.h:28-30), XORWOW stream
//      seeded wangHash(frameNo) + pixel with the reference's draw order -> path-for-path the same
//      walk as kernel_pathtracer up to fast-math / FMA-contraction rounding.
//   1  global majorant + Philox counter RNG (same estimator, different random stream).
//   2  local majorants: a walk over the macrocell grid (svr_macrocell.cu); within a cell the free
//      path is sampled against the cell's majorant, optical depth carries across cell faces, cells
//      with majorant 0 cost no fetch and are leapt over several at a time.  Delta tracking with any
//      valid majorant samples the same free-path distribution, so the expectation is unchanged.
// Shadow rays: binary delta-tracking estimate as the reference (transmittance.h:10-17), or ratio
// tracking (SVR_OPT_SHADOW_ESTIMATOR = 1).
//
// Kernel shapes (SVR_OPT_PT_KERNEL):
//   0  per-lane state machine ("wavefront in a warp"): the reference's three nested data-dependent
//      loops (bounces x tracking x shadow tracking) are flattened into
//      GENERATE -> TRACK -> EVENT phases; all lanes of a warp run the tracking phase together
//      whatever their ray is (camera, bounce or shadow ray), tentative collisions are evaluated
//      together, and the expensive shading code runs once per round for every lane that has an
//      event.  A lane whose path ends starts its next sample at once.
//   1  megakernel: the reference's loop nest as written, one path at a time per lane.
#include "svr_rng.cuh"
#include "svr_state.h"

// Resident 256-thread blocks per SM the path-tracing kernels are compiled for: sets the register
// budget (65536 / (256 * N)).  Chosen by measurement, see DESIGN.md.
#ifndef SVR_PT_MIN_BLOCKS
#define SVR_PT_MIN_BLOCKS 1
#endif

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
    int32_t trackRounds;   // state machine: tracking rounds between event phases (0 = until all lanes have an event)
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
// free-path sampling, split into march() (advance to the next tentative collision or out of the
// volume; no volume fetch) and collide() (fetch + accept/reject) so that a warp can run each half
// convergently.
// ---------------------------------------------------------------------------------------------
enum MarchResult { MARCH_COLLIDE = 0, MARCH_ESCAPED = 1 };

// woodcock_tracking.h:20-51, global majorant tf.GetMaxOpacity()
struct TrackGlobal {
    float t, tMin, tMax;

    template <class Rng>
    SVR_DEV bool begin(const DevScene& s, const Ray& ray, Rng&)
    {
        float tNear, tFar;
        if (!intersect_volume(s.vol, ray, &tNear, &tFar)) return false;
        tMin = tNear < 0.f ? 1e-6f : tNear;
        tMax = tFar;
        t = tMin;
        return true;
    }
    template <bool COUNT, class Rng>
    SVR_DEV MarchResult march(const DevScene& s, const Ray&, Rng& rng, LocalCounters<COUNT>&)
    {
        const float invSigmaMaxSampleInterval = 1.f / (s.tf.maxOpacity * 1.f);  // BASE_SAMPLE_STEP_SIZE 1
        t += -logf(rng.next_one_minus()) * invSigmaMaxSampleInterval;
        return t > tMax ? MARCH_ESCAPED : MARCH_COLLIDE;
    }
    // true = real collision at t
    template <bool COUNT, class Rng>
    SVR_DEV bool collide(const DevScene& s, const Ray& ray, Rng& rng, LocalCounters<COUNT>& lc, int slot, float* ratioT)
    {
        const float invSigmaMax = 1.f / s.tf.maxOpacity;
        float intensity = intensity_at(s.vol, ray.orig + t * ray.dir);
        float sigma_t = tf_at(s.tf, intensity).w;
        lc.add(slot, 1);
        lc.add(SVR_CNT_TF_LOOKUPS, 1);
        if (ratioT) {
            *ratioT *= 1.f - sigma_t * invSigmaMax;
            return false;
        }
        return rng.next() < sigma_t * invSigmaMax || t > tMax;
    }
};

// Delta tracking against per-macrocell majorants with empty-space leaping.
// Occupied cells are walked with a classic incremental 3-D DDA; an empty cell stores the radius of
// the empty cube around it, and the ray jumps to that cube's far face in one step, after which the
// DDA state is rebuilt from the new cell.
struct TrackLocal {
    float t, tMin, tMax;
    float tau;            // optical depth still to travel before the next tentative collision
    float sig;            // majorant of the cell the tentative collision lies in
    float3 invDg, kk;     // face crossing: t = bound * invDg - kk per axis (cell coordinates)
    float tNx, tNy, tNz;  // ray parameter at the next face of the current cell, per axis
    int cx, cy, cz;

    SVR_DEV void locate(const DevScene& s, const Ray& ray, float tt, int& ox, int& oy, int& oz) const
    {
        const float3 toCell = f3(s.vol.bbox.invSize) * s.grid.scale;
        const float3 g = (ray.orig + tt * ray.dir - f3(s.vol.bbox.vmin)) * toCell;
        ox = (int)floorf(g.x);
        oy = (int)floorf(g.y);
        oz = (int)floorf(g.z);
    }
    SVR_DEV void faces()
    {
        tNx = fmaf((float)(invDg.x > 0.f ? cx + 1 : cx), invDg.x, -kk.x);
        tNy = fmaf((float)(invDg.y > 0.f ? cy + 1 : cy), invDg.y, -kk.y);
        tNz = fmaf((float)(invDg.z > 0.f ? cz + 1 : cz), invDg.z, -kk.z);
    }
    template <class Rng>
    SVR_DEV bool begin(const DevScene& s, const Ray& ray, Rng& rng)
    {
        float tNear, tFar;
        if (!intersect_volume(s.vol, ray, &tNear, &tFar)) return false;
        tMin = tNear < 0.f ? 1e-6f : tNear;
        tMax = tFar;
        t = tMin;
        const float3 toCell = f3(s.vol.bbox.invSize) * s.grid.scale;
        const float3 g0 = (ray.orig - f3(s.vol.bbox.vmin)) * toCell;
        const float3 dg = ray.dir * toCell;
        // an axis the ray does not move along never produces a crossing: 0 * bound + FLT_MAX
        invDg.x = dg.x != 0.f ? 1.f / dg.x : 0.f;
        invDg.y = dg.y != 0.f ? 1.f / dg.y : 0.f;
        invDg.z = dg.z != 0.f ? 1.f / dg.z : 0.f;
        kk.x = dg.x != 0.f ? g0.x * invDg.x : -FLT_MAX;
        kk.y = dg.y != 0.f ? g0.y * invDg.y : -FLT_MAX;
        kk.z = dg.z != 0.f ? g0.z * invDg.z : -FLT_MAX;
        locate(s, ray, t, cx, cy, cz);
        cx = min(max(cx, 0), s.grid.gx - 1);
        cy = min(max(cy, 0), s.grid.gy - 1);
        cz = min(max(cz, 0), s.grid.gz - 1);
        faces();
        tau = -logf(rng.next_one_minus());
        return true;
    }
    template <bool COUNT, class Rng>
    SVR_DEV MarchResult march(const DevScene& s, const Ray& ray, Rng&, LocalCounters<COUNT>& lc)
    {
        const DevGrid& g = s.grid;
        while (true) {
            const float m = __ldg(&g.majorant[(cz * g.gy + cy) * g.gx + cx]);
            lc.add(SVR_CNT_CELLS, 1);
            if (m > 0.f) {
                // occupied cell: spend optical depth m * length, or collide inside it
                const float tN = fminf(fminf(tNx, tNy), tNz);
                const float dd = fmaxf(fminf(tN, tMax) - t, 0.f) * m;
                if (tau < dd) {
                    t += tau / m;
                    sig = m;
                    return MARCH_COLLIDE;
                }
                tau -= dd;
                t = tN;
                if (tN >= tMax) return MARCH_ESCAPED;
                if (tNx <= tNy && tNx <= tNz) {
                    cx += invDg.x > 0.f ? 1 : -1;
                    tNx += fabsf(invDg.x);
                    if ((unsigned)cx >= (unsigned)g.gx) return MARCH_ESCAPED;
                } else if (tNy <= tNz) {
                    cy += invDg.y > 0.f ? 1 : -1;
                    tNy += fabsf(invDg.y);
                    if ((unsigned)cy >= (unsigned)g.gy) return MARCH_ESCAPED;
                } else {
                    cz += invDg.z > 0.f ? 1 : -1;
                    tNz += fabsf(invDg.z);
                    if ((unsigned)cz >= (unsigned)g.gz) return MARCH_ESCAPED;
                }
            } else {
                // empty cell, and so is the cube of radius d-1 around it: leap to that cube's far face
                const int d = max((int)(-m), 1);
                const float lx = fmaf((float)(invDg.x > 0.f ? cx + d : cx - d + 1), invDg.x, -kk.x);
                const float ly = fmaf((float)(invDg.y > 0.f ? cy + d : cy - d + 1), invDg.y, -kk.y);
                const float lz = fmaf((float)(invDg.z > 0.f ? cz + d : cz - d + 1), invDg.z, -kk.z);
                const float tL = fminf(fminf(lx, ly), lz);
                if (tL >= tMax) return MARCH_ESCAPED;
                t = fmaxf(t, tL);
                // exact on the exit axis, re-located and clamped to the cube on the others
                int nx, ny, nz;
                locate(s, ray, t, nx, ny, nz);
                nx = min(max(nx, cx - d + 1), cx + d - 1);
                ny = min(max(ny, cy - d + 1), cy + d - 1);
                nz = min(max(nz, cz - d + 1), cz + d - 1);
                if (lx <= ly && lx <= lz) nx = invDg.x > 0.f ? cx + d : cx - d;
                else if (ly <= lz) ny = invDg.y > 0.f ? cy + d : cy - d;
                else nz = invDg.z > 0.f ? cz + d : cz - d;
                if ((unsigned)nx >= (unsigned)g.gx || (unsigned)ny >= (unsigned)g.gy || (unsigned)nz >= (unsigned)g.gz)
                    return MARCH_ESCAPED;
                cx = nx;
                cy = ny;
                cz = nz;
                faces();
            }
        }
    }
    template <bool COUNT>
    SVR_DEV bool collide(const DevScene& s, const Ray& ray, Philox& rng, LocalCounters<COUNT>& lc, int slot, float* ratioT)
    {
        float intensity = intensity_at(s.vol, ray.orig + t * ray.dir);
        float sigma_t = tf_at(s.tf, intensity).w;
        lc.add(slot, 1);
        lc.add(SVR_CNT_TF_LOOKUPS, 1);
        float ua, ub;
        rng.next2(ua, ub);  // one Philox block: the accept draw and the next free-flight draw
        tau = -logf(1.f - ub);
        if (ratioT) {
            *ratioT *= 1.f - sigma_t / sig;
            return false;
        }
        return ua * sig < sigma_t;
    }
};

template <int MODE>
struct TrackOf {
    typedef TrackGlobal type;
};
template <>
struct TrackOf<2> {
    typedef TrackLocal type;
};

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

// pathtracer.cu:96-103 (the 0.0722 term is a double product in the reference)
template <bool EXACT, class Rng>
SVR_DEV bool russian_roulette(float3* T, Rng& rng)
{
    float illum;
    if (EXACT) illum = (float)((double)(0.2126f * T->x + 0.7152f * T->y) + 0.0722 * (double)T->z);
    else illum = 0.2126f * T->x + 0.7152f * T->y + 0.0722f * T->z;
    if (rng.next() > illum) return true;
    *T = *T / illum;
    return false;
}

// Per-lane path state shared by both kernel shapes.
template <int MODE>
struct PathState {
    typename RngOf<MODE>::type rng;
    typename TrackOf<MODE>::type trk;
    Ray ray;          // the ray being tracked (camera / bounce ray, or the shadow ray)
    float3 L, T;
    uint32_t k;
    LightHit ls;      // nearest light along the camera ray (pathtracer.cu:214-215)
    bool hitLight;
    // stash across the shadow ray
    VolumeSample vs;
    float3 pending;   // numLights * bsdf * Li / pdf, to be multiplied by the transmittance
    float Pbrdf;
    float ratioT;     // ratio-tracking running transmittance
    ShadingType st;
    bool shadow;      // the tracked ray is a shadow ray
};

enum EventResult { EV_TRACK = 0, EV_PATH_DONE = 1 };

// pathtracer.cu:204-215: seed, camera ray, nearest light; then start tracking the camera ray.
// Returns false when the ray misses the volume (an immediate "escaped" event).
template <int MODE>
SVR_DEV bool path_begin(const DevScene& s, PathState<MODE>& ps, uint32_t idx, uint32_t idy, uint32_t offset, uint32_t sample)
{
    constexpr bool EXACT_PI = MODE == 0;
    ps.rng.init(s.seedKey, offset, sample);
    ps.L = f3(0.f);
    ps.T = f3(1.f);
    ps.k = 0;
    ps.shadow = false;
    ps.ray = camera_ray_jittered<EXACT_PI>(s.cam, idx, idy, ps.rng);
    ps.hitLight = nearest_light(s, ps.ray, &ps.ls);
    return ps.trk.begin(s, ps.ray, ps.rng);
}

// Everything between two tracked flights (pathtracer.cu:218-277 plus 171-198).  `t` is the
// collision distance of the flight that just ended, or -FLT_MAX when it left the volume.
template <int MODE, bool COUNT>
SVR_DEV EventResult path_event(const DevScene& s, PathState<MODE>& ps, float t, uint32_t traceDepth, LocalCounters<COUNT>& lc)
{
    constexpr bool EXACT_PI = MODE == 0;
    while (true) {
        if (!ps.shadow) {
            if ((ps.k == 0) && ps.hitLight) {
                float tt = t < 0.f ? FLT_MAX : t;
                if (ps.ls.t < tt) {
                    float cosTerm = dot(ps.ls.normal, -ps.ray.dir);
                    ps.L += ps.T * ps.ls.radiance * (cosTerm <= 0.f ? 0.f : 1.f);
                    return EV_PATH_DONE;
                }
            }
            if (t < 0.f) {
                if (s.envEnabled) ps.L += ps.T * env_radiance(s.env, ps.ray.dir);  // the line commented out at pathtracer.cu:233
                return EV_PATH_DONE;
            }
            VolumeSample& vs = ps.vs;
            vs.wo = -ps.ray.dir;
            vs.ptInWorld = ps.ray.orig + t * ps.ray.dir;
            float intensity = intensity_at(s.vol, vs.ptInWorld);
            vs.color_opacity = tf_at(s.tf, intensity);
            vs.gradient = gradient_at(s.vol, vs.ptInWorld);
            float gradientMagnitude = sqrtf(dot(vs.gradient, vs.gradient));
            lc.add(SVR_CNT_SHADE_TAPS, 7);
            lc.add(SVR_CNT_TF_LOOKUPS, 1);
            lc.add(SVR_CNT_SCATTERS, 1);
            const float gf = s.vol.gradientFactor;
            ps.Pbrdf = vs.color_opacity.w *
                       (1.f - expf(-25.f * gf * gf * gf * gradientMagnitude * 65535.f * s.vol.invMaxMagnitude));
            ps.st = (ps.rng.next() < ps.Pbrdf) ? BRDF : ISOTROPIC;

            // estimate_direct_light, pathtracer.cu:171-198
            bool needShadow = false;
            if (s.numLights != 0) {
                int lightId = (int)((float)s.numLights * ps.rng.next());
                lightId = lightId < (int)s.numLights ? lightId : (int)s.numLights - 1;
                float3 lightPos, wi;
                float pdf;
                float3 Li = sample_light<EXACT_PI>(s.lights[lightId], vs.ptInWorld, ps.rng, &lightPos, &wi, &pdf);
                if (pdf > 0.f && max3(Li) > 0.f) {
                    ps.pending = (float)s.numLights * bsdf(vs, wi, ps.st) * Li / pdf;
                    // transmittance.h:10-17: track from the sample toward the light through the whole box
                    ps.ray.orig = vs.ptInWorld;
                    ps.ray.dir = normalize(lightPos - vs.ptInWorld);
                    ps.shadow = true;
                    ps.ratioT = 1.f;
                    needShadow = true;
                }
            }
            if (needShadow) {
                if (ps.trk.begin(s, ps.ray, ps.rng)) return EV_TRACK;
                t = -FLT_MAX;  // shadow ray misses the box: unoccluded
                continue;
            }
            ps.shadow = true;  // no light sample: fall through to the bounce with nothing pending
            ps.pending = f3(0.f);
            ps.ratioT = 1.f;
            t = -FLT_MAX;
            continue;
        }

        // ---- the shadow flight ended: add the direct light, then bounce (pathtracer.cu:257-276)
        {
            float Tr;
            if (s.shadowEstimator) Tr = ps.ratioT;
            else Tr = ((t > ps.trk.tMin) && (t < ps.trk.tMax)) ? 0.f : 1.f;
            ps.L += ps.T * (Tr * ps.pending);
            ps.shadow = false;
            const VolumeSample& vs = ps.vs;
            // the last bounce's BSDF sample is never used: skip it unless reproducing the reference's draws
            if (MODE != 0 && ps.k + 1 >= traceDepth) return EV_PATH_DONE;
            float3 wi = f3(0.f);
            float pdf = 0.f;
            float3 f = sample_bsdf<EXACT_PI>(vs, &wi, &pdf, ps.rng, ps.st);
            float cosTerm = fabsf(dot(normalize(vs.gradient), wi));
            if (max3(f) > 0.f && pdf > 0.f) {
                if (ps.st == ISOTROPIC)
                    ps.T *= f / (pdf * (1.f - ps.Pbrdf));
                else
                    ps.T *= f * cosTerm / (pdf * ps.Pbrdf);
            }
            ps.ray.orig = vs.ptInWorld;
            ps.ray.dir = wi;
            if (ps.k >= 3) {
                if (russian_roulette<EXACT_PI>(&ps.T, ps.rng)) return EV_PATH_DONE;
            }
            ps.k += 1;
            if (ps.k >= traceDepth) return EV_PATH_DONE;
            if (ps.trk.begin(s, ps.ray, ps.rng)) return EV_TRACK;
            t = -FLT_MAX;  // bounce ray misses the (clipped) box
        }
    }
}

// Russian roulette for ratio tracking on a nearly opaque segment; true = terminate with T = 0
template <class Rng>
SVR_DEV bool ratio_roulette(float& T, Rng& rng)
{
    if (T >= 0.02f) return false;
    if (rng.next() * 0.02f >= T) {
        T = 0.f;
        return true;
    }
    T = 0.02f;
    return false;
}

SVR_DEV void write_pixel(const DevScene& s, const PtLaunch& a, uint32_t offset, float3 sum)
{
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

// ---------------------------------------------------------------------------------------------
// kernel shape 1: megakernel (the reference's loop nest)
// ---------------------------------------------------------------------------------------------
template <int MODE, bool COUNT>
__global__ void __launch_bounds__(256, SVR_PT_MIN_BLOCKS) pathtrace_mega_kernel(const __grid_constant__ DevScene s, const PtLaunch a, Counters* cnt)
{
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t idx = blockIdx.x * 16u + (warp & 1u) * 8u + (lane & 7u);
    const uint32_t idy = a.y0 + blockIdx.y * (blockDim.x >> 4) + (warp >> 1) * 4u + (lane >> 3);
    const bool inside = idx < s.cam.imageW && idy < a.y1;
    LocalCounters<COUNT> lc;
    if (inside) {
        const uint32_t offset = idy * s.cam.imageW + idx;
        float3 sum = f3(0.f);
        PathState<MODE> ps;
        for (uint32_t n = 0; n < a.nSamples; ++n) {
            lc.add(SVR_CNT_PATHS, 1);
            bool tracking = path_begin<MODE>(s, ps, idx, idy, offset, a.firstSample + n);
            float t = -FLT_MAX;
            while (true) {
                if (tracking) {
                    const int slot = ps.shadow ? SVR_CNT_SHADOW_TAPS : SVR_CNT_TRACK_TAPS;
                    float* ratio = (ps.shadow && s.shadowEstimator) ? &ps.ratioT : nullptr;
                    while (true) {
                        if (ps.trk.template march<COUNT>(s, ps.ray, ps.rng, lc) == MARCH_ESCAPED) {
                            t = -FLT_MAX;
                            break;
                        }
                        if (ps.trk.template collide<COUNT>(s, ps.ray, ps.rng, lc, slot, ratio)) {
                            t = ps.trk.t;
                            break;
                        }
                        if (ratio && ratio_roulette(ps.ratioT, ps.rng)) {
                            t = -FLT_MAX;
                            break;
                        }
                    }
                }
                if (path_event<MODE, COUNT>(s, ps, t, a.traceDepth, lc) == EV_PATH_DONE) break;
                tracking = true;
            }
            sum += ps.L;
        }
        write_pixel(s, a, offset, sum);
    }
    lc.flush(cnt);
}

// ---------------------------------------------------------------------------------------------
// kernel shape 0: per-lane state machine
// ---------------------------------------------------------------------------------------------
enum Phase { PH_GEN = 0, PH_TRACK = 1, PH_EVENT = 2, PH_DONE = 3 };

template <int MODE, bool COUNT>
__global__ void __launch_bounds__(256, SVR_PT_MIN_BLOCKS) pathtrace_sm_kernel(const __grid_constant__ DevScene s, const PtLaunch a, Counters* cnt)
{
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t idx = blockIdx.x * 16u + (warp & 1u) * 8u + (lane & 7u);
    const uint32_t idy = a.y0 + blockIdx.y * (blockDim.x >> 4) + (warp >> 1) * 4u + (lane >> 3);
    const bool inside = idx < s.cam.imageW && idy < a.y1;
    const uint32_t offset = idy * s.cam.imageW + idx;
    LocalCounters<COUNT> lc;

    PathState<MODE> ps;
    float3 sum = f3(0.f);
    uint32_t n = 0;
    int phase = inside ? PH_GEN : PH_DONE;
    float tEvent = -FLT_MAX;

    while (true) {
        // ---- GENERATE: lanes without a path start their next sample
        if (phase == PH_GEN) {
            if (n == a.nSamples) {
                phase = PH_DONE;
            } else {
                lc.add(SVR_CNT_PATHS, 1);
                const bool tracking = path_begin<MODE>(s, ps, idx, idy, offset, a.firstSample + n);
                ++n;
                tEvent = -FLT_MAX;
                phase = tracking ? PH_TRACK : PH_EVENT;
            }
        }
        if (__all_sync(0xffffffffu, phase == PH_DONE)) break;

        // ---- TRACK: every tracking lane marches to its next tentative collision, then all of them
        //      evaluate it together; repeat until the round budget is spent or nobody is tracking
        int rounds = a.trackRounds;
        while (__any_sync(0xffffffffu, phase == PH_TRACK)) {
            if (phase == PH_TRACK) {
                if (ps.trk.template march<COUNT>(s, ps.ray, ps.rng, lc) == MARCH_ESCAPED) {
                    tEvent = -FLT_MAX;
                    phase = PH_EVENT;
                }
            }
            if (phase == PH_TRACK) {
                const int slot = ps.shadow ? SVR_CNT_SHADOW_TAPS : SVR_CNT_TRACK_TAPS;
                float* ratio = (ps.shadow && s.shadowEstimator) ? &ps.ratioT : nullptr;
                if (ps.trk.template collide<COUNT>(s, ps.ray, ps.rng, lc, slot, ratio)) {
                    tEvent = ps.trk.t;
                    phase = PH_EVENT;
                } else if (ratio && ratio_roulette(ps.ratioT, ps.rng)) {
                    tEvent = -FLT_MAX;
                    phase = PH_EVENT;
                }
            }
            if (a.trackRounds > 0 && --rounds == 0) break;
        }

        // ---- EVENT: shade / add direct light / bounce / finish, for every lane whose flight ended
        if (phase == PH_EVENT) {
            if (path_event<MODE, COUNT>(s, ps, tEvent, a.traceDepth, lc) == EV_PATH_DONE) {
                sum += ps.L;
                phase = PH_GEN;
            } else {
                phase = PH_TRACK;
            }
        }
    }
    if (inside) write_pixel(s, a, offset, sum);
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
void launch_mode(int shape, dim3 grid, int block, cudaStream_t stream, const DevScene& sc, const PtLaunch& a, Counters* cnt)
{
    if (shape == 1) {
        if (cnt) pathtrace_mega_kernel<MODE, true><<<grid, block, 0, stream>>>(sc, a, cnt);
        else pathtrace_mega_kernel<MODE, false><<<grid, block, 0, stream>>>(sc, a, cnt);
    } else {
        if (cnt) pathtrace_sm_kernel<MODE, true><<<grid, block, 0, stream>>>(sc, a, cnt);
        else pathtrace_sm_kernel<MODE, false><<<grid, block, 0, stream>>>(sc, a, cnt);
    }
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
    a.trackRounds = st.options[SVR_OPT_PT_ROUNDS];
    Counters* cnt = nullptr;
    if (st.options[SVR_OPT_COUNTERS]) {
        cnt = device_counters();
        if (!cnt) return fail_msg("render_pathtracer: counter allocation failed");
    }
    const int block = st.options[SVR_OPT_PT_BLOCK];
    const int shape = st.options[SVR_OPT_PT_KERNEL];
    const uint32_t tileH = (uint32_t)block / 16u;
    dim3 grid((sc.cam.imageW + 15u) / 16u, ((a.y1 - a.y0) + tileH - 1u) / tileH);
    switch (mode) {
        case 0: launch_mode<0>(shape, grid, block, st.stream, sc, a, cnt); break;
        case 1: launch_mode<1>(shape, grid, block, st.stream, sc, a, cnt); break;
        default: launch_mode<2>(shape, grid, block, st.stream, sc, a, cnt); break;
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
