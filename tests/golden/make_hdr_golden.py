"""Generates tests/golden/hdr_stb.npz: Radiance .hdr files (as bytes) and what the REFERENCE's decoder makes of them --
stbi_loadf from /root/reference/utils/stb_image.h, compiled where it lies into oracle/_ref/libsvr_stbhdr.so (oracle/Makefile,
target stbhdr).  Run here (the container with /root/reference): python tests/golden/make_hdr_golden.py
The fixture lets tests/test_env_io.py hold svr_hdr_read to the reference's decoder on boxes without the reference."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hdr_oracle as H  # noqa: E402


def stb_load(lib, path):
    w, h, n = C.c_int(), C.c_int(), C.c_int()
    lib.ref_stbi_loadf.restype = C.POINTER(C.c_float)
    p = lib.ref_stbi_loadf(str(path).encode(), C.byref(w), C.byref(h), C.byref(n))
    if not p:
        return None
    out = np.ctypeslib.as_array(p, shape=(h.value, w.value, n.value)).copy()
    lib.ref_stbi_free(p)
    return out


def cases():
    rng = np.random.default_rng(2024)
    for name, w, h, rle, magic in (("rle_64x32", 64, 32, True, "#?RADIANCE"), ("flat_64x32", 64, 32, False, "#?RADIANCE"), ("rle_300x7", 300, 7, True, "#?RADIANCE"),
                                   ("narrow_6x9", 6, 9, True, "#?RADIANCE"), ("rle_1024x3", 1024, 3, True, "#?RADIANCE")):
        v, u = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
        img = np.stack([0.3 + 0.7 * u, 0.2 + 0.5 * v, 1.0 - 0.6 * v], axis=2) * rng.uniform(0.5, 1.5, (h, w, 3))
        img[h // 5: h // 5 + 2, w // 3: w // 3 + 3] = [900.0, 700.0, 350.0]
        img[-2:, :] = 0.0
        img[h // 2, :] = 0.5
        img[0, : w // 2] = 1e-6 * rng.uniform(0.5, 1.0, (w // 2, 3))   # tiny values: small exponents
        yield name, H.float_to_rgbe(img), rle, magic


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsvr_stbhdr.so"))
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, rgbe, rle, magic in cases():
            p = H.write_hdr(os.path.join(d, name + ".hdr"), rgbe, rle=rle, magic=magic)
            dec = stb_load(lib, p)
            assert dec is not None and dec.shape[2] == 3, name
            out[name + "_file"] = np.frombuffer(open(p, "rb").read(), np.uint8)
            out[name + "_stb"] = dec
        # what the reference's stb_image v2.12 does NOT read although svr_hdr_read does: the "#?RGBE" signature
        name, rgbe, rle, _ = next(cases())
        p = H.write_hdr(os.path.join(d, "rgbe_magic.hdr"), rgbe, rle=rle, magic="#?RGBE")
        out["stb_rejects_rgbe_signature"] = np.array([stb_load(lib, p) is None])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hdr_stb.npz"), **out)
    print("wrote", len(out) // 2, "cases")


if __name__ == "__main__":
    main()
