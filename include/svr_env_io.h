/*
 * svr_env_io.h -- C ABI of the environment-map input stage: what Lights::SetEnvironmentLight(filename)
 * does with stb_image (core/lights/lights.cpp:31-75): a Radiance RGBE (.hdr) picture -> w x h float4
 * lat-long texture (wrap, linear, normalised coordinates) -> cudaEnvironmentLight.
 * Functions return 0 on success; svr_last_error() has the text otherwise.
 */
#ifndef SVR_ENV_IO_H
#define SVR_ENV_IO_H

#include "svr_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Decodes a Radiance .hdr file (#?RADIANCE / #?RGBE, FORMAT=32-bit_rle_rgbe, "-Y h +X w", flat or
 * new-style run-length scanlines) to linear floats, channel = mantissa byte * 2^(exponent - 136) as
 * stbi_loadf does.  Call with rgb_out == NULL to get the size; then with room for 3 * w * h floats AND *w, *h
 * still holding that size (in-out: a file whose size changed in between is an error, not an overrun).
 * Pictures above 32768 per side or 2^27 pixels, and files too short for the picture their header declares, are
 * rejected before anything is allocated.  Host only: needs no GPU. */
int svr_hdr_read(const char* path, float* rgb_out, uint32_t* w, uint32_t* h);

/* Lights::SetEnvironmentLight(filename) (core/lights/lights.cpp:31-75): reads the file, expands RGB to
 * float4 (alpha 0), creates the 2-D array + texture object and fills *out as cudaEnvironmentLight::Set(tex)
 * does (intensity 1, offset 0).  Destroy with svr_env_destroy (svr_render.h). */
int svr_env_load_hdr(const char* path, svr_env_light* out);

/* Inspection hook: the importance sampler SVR_OPT_ENV_NEE uses for the environment light last passed to setup_env_lights --
 * cumulative row probabilities (*h + 1 floats) and per-row cumulative cell probabilities (*h rows of *w + 1 floats) over
 * the direction grid (u, v) = (phi / 2 pi, theta / pi).  Either pointer may be NULL. */
int svr_env_sampler_copy(float* host_marg, float* host_cond, uint32_t* w, uint32_t* h);

#ifdef __cplusplus
}
#endif
#endif /* SVR_ENV_IO_H */
