/* TEST INFRASTRUCTURE -- the reference's own environment-map decoder, compiled where it lies.
 *
 * Lights::SetEnvironmentLight reads .hdr files with stbi_loadf (core/lights/lights.cpp:34) from the vendored
 * utils/stb_image.h.  This translation unit is nothing but that header's implementation behind one exported function,
 * built by oracle/Makefile with -I$(REF) into oracle/_ref/libsvr_stbhdr.so (git-ignored; no reference source is copied).
 * tests/test_env_io.py holds svr_hdr_read (include/svr_env_io.h) to it bit for bit, and tests/golden/make_hdr_golden.py
 * freezes its outputs into tests/golden/hdr_stb.npz for boxes without /root/reference. */
#define STB_IMAGE_IMPLEMENTATION
#define STBI_ONLY_HDR
#include "utils/stb_image.h"

/* Returns the malloc'ed w * h * n floats of stbi_loadf(path, &w, &h, &n, 0) (NULL on failure); free with ref_stbi_free. */
float* ref_stbi_loadf(const char* path, int* w, int* h, int* n) { return stbi_loadf(path, w, h, n, 0); }
void ref_stbi_free(float* p) { stbi_image_free(p); }
const char* ref_stbi_failure(void) { return stbi_failure_reason(); }
