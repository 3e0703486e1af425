"""Radiance .hdr reader through the C ABI (include/svr_env_io.h; host code, no GPU) against the numpy
encoder/decoder in oracle/hdr_oracle.py."""
import ctypes as C

import numpy as np
import pytest

from oracle import hdr_oracle as H
from sunvolumerender_b200 import _lib as L


def _read(path):
    lib = L.load()
    w, h = C.c_uint32(), C.c_uint32()
    rc = lib.svr_hdr_read(str(path).encode(), None, C.byref(w), C.byref(h))
    if rc:
        return rc, None, lib.svr_last_error().decode()
    out = np.zeros((h.value, w.value, 3), np.float32)
    rc = lib.svr_hdr_read(str(path).encode(), C.c_void_p(out.ctypes.data), C.byref(w), C.byref(h))
    return rc, out, lib.svr_last_error().decode()


def _sky(h, w, rng):
    v, u = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    img = np.stack([0.3 + 0.7 * u, 0.2 + 0.5 * v, 1.0 - 0.6 * v], axis=2)
    img[h // 5: h // 5 + 3, w // 3: w // 3 + 4] = [900.0, 700.0, 350.0]   # a sun
    img[-4:, :] = 0.0                                                      # black ground rows (exponent 0)
    img[h // 2, :] = 0.5                                                   # a long run
    return img * rng.uniform(0.8, 1.2, img.shape)


@pytest.mark.parametrize("rle,w,h,magic", [(True, 64, 32, "#?RADIANCE"), (False, 64, 32, "#?RGBE"), (True, 300, 7, "#?RADIANCE"), (True, 6, 9, "#?RADIANCE")])
def test_decode_is_bit_exact(tmp_path, rle, w, h, magic):
    rgbe = H.float_to_rgbe(_sky(h, w, np.random.default_rng(w)))
    p = H.write_hdr(tmp_path / "sky.hdr", rgbe, rle=rle, magic=magic)
    rc, img, err = _read(p)
    assert rc == 0, err
    exp = H.rgbe_to_float(rgbe)
    assert img.shape == (h, w, 3) and np.array_equal(img.view(np.uint32), exp.view(np.uint32))
    assert img.max() > 500 and (img[-1] == 0).all()


def test_rejected_files(tmp_path):
    rgbe = H.float_to_rgbe(_sky(16, 32, np.random.default_rng(0)))
    good = H.write_hdr(tmp_path / "g.hdr", rgbe).read_bytes()
    cases = {
        "magic": good.replace(b"#?RADIANCE", b"#?RADIANCX"),
        "format": good.replace(b"32-bit_rle_rgbe", b"32-bit_rle_xyze"),
        "layout": good.replace(b"-Y 16 +X 32", b"+Y 16 +X 32"),
        "trunc": good[: len(good) - 40],
    }
    for name, data in cases.items():
        p = tmp_path / (name + ".hdr")
        p.write_bytes(data)
        rc, _, err = _read(p)
        assert rc != 0 and "svr_hdr_read" in err, name
    rc, _, err = _read(tmp_path / "missing.hdr")
    assert rc != 0 and "unable to load" in err


def _stb():
    """The reference's own decoder (stbi_loadf of utils/stb_image.h, compiled where it lies: oracle/Makefile, stbhdr)."""
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libsvr_stbhdr.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_stbi_loadf.restype = C.POINTER(C.c_float)
    return lib


def test_decode_is_bit_exact_against_the_references_stb_image_golden(tmp_path):
    """tests/golden/hdr_stb.npz: files and what stbi_loadf (the call at core/lights/lights.cpp:34) decodes from them,
    frozen by tests/golden/make_hdr_golden.py where /root/reference is mounted."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hdr_stb.npz"))
    names = sorted(k[:-5] for k in g.files if k.endswith("_file"))
    assert len(names) >= 5
    for name in names:
        p = tmp_path / (name + ".hdr")
        p.write_bytes(g[name + "_file"].tobytes())
        rc, img, err = _read(p)
        assert rc == 0, (name, err)
        ref = g[name + "_stb"]
        assert img.shape == ref.shape and np.array_equal(img.view(np.uint32), ref.view(np.uint32)), name
    assert bool(g["stb_rejects_rgbe_signature"][0])   # a superset: this library also reads the "#?RGBE" signature


def test_decode_is_bit_exact_against_stb_image_live(tmp_path):
    """The same against the decoder itself on fresh random pictures (skipped where oracle/_ref/libsvr_stbhdr.so is absent)."""
    lib = _stb()
    if lib is None:
        pytest.skip("oracle/_ref/libsvr_stbhdr.so not built (needs /root/reference)")
    rng = np.random.default_rng(11)
    for i in range(12):
        w, h = int(rng.integers(1, 200)), int(rng.integers(1, 40))
        img = rng.uniform(0, 1, (h, w, 3)) ** 4 * 10.0 ** rng.uniform(-6, 4)
        img[rng.uniform(0, 1, (h, w)) < 0.2] = 0.0
        p = H.write_hdr(tmp_path / f"r{i}.hdr", H.float_to_rgbe(img), rle=bool(i % 2))
        wi, hi, ni = C.c_int(), C.c_int(), C.c_int()
        ptr = lib.ref_stbi_loadf(str(p).encode(), C.byref(wi), C.byref(hi), C.byref(ni))
        assert ptr and (wi.value, hi.value, ni.value) == (w, h, 3)
        ref = np.ctypeslib.as_array(ptr, shape=(h, w, 3)).copy()
        lib.ref_stbi_free(ptr)
        rc, mine, err = _read(p)
        assert rc == 0, err
        assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32)), (i, w, h)
