"""TEST INFRASTRUCTURE -- not product code.  Radiance RGBE (.hdr) writer and decoder in numpy: the checker
for include/svr_env_io.h.  The reference reads environment maps with stbi_loadf (utils/stb_image.h,
called at core/lights/lights.cpp:34); the decode rule restated here is stb's: channel = mantissa byte *
2^(exponent - 136), exponent byte 0 = black, no gamma for HDR input."""
import numpy as np


def float_to_rgbe(rgb):
    """(h, w, 3) float -> (h, w, 4) uint8, the classic Radiance encoding."""
    rgb = np.asarray(rgb, np.float64)
    m = rgb.max(axis=2)
    out = np.zeros(rgb.shape[:2] + (4,), np.uint8)
    nz = m > 1e-32
    frac, exp = np.frexp(np.where(nz, m, 1.0))
    scale = np.where(nz, frac * 256.0 / np.where(nz, m, 1.0), 0.0)
    out[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(nz, exp + 128, 0).astype(np.uint8)
    return out


def rgbe_to_float(rgbe):
    rgbe = np.asarray(rgbe, np.uint8)
    e = rgbe[..., 3].astype(np.int32)
    f = np.where(e != 0, np.ldexp(np.float32(1.0), e - 136), np.float32(0.0)).astype(np.float32)
    return (rgbe[..., :3].astype(np.float32) * f[..., None]).astype(np.float32)


def _rle_channel(row):
    out = bytearray()
    i, n = 0, len(row)
    while i < n:
        run = 1
        while i + run < n and run < 127 and row[i + run] == row[i]:
            run += 1
        if run >= 4:
            out += bytes([128 + run, row[i]])
            i += run
        else:
            j = i
            while j < n and j - i < 128:
                r = 1
                while j + r < n and r < 4 and row[j + r] == row[j]:
                    r += 1
                if r >= 4:
                    break
                j += 1
            out += bytes([j - i]) + bytes(row[i:j])
            i = j
    return bytes(out)


def write_hdr(path, rgbe, rle=True, magic="#?RADIANCE", extra_header=("EXPOSURE=1.0",)):
    h, w = rgbe.shape[:2]
    with open(path, "wb") as f:
        f.write((magic + "\n" + "".join(l + "\n" for l in extra_header) + "FORMAT=32-bit_rle_rgbe\n\n" + f"-Y {h} +X {w}\n").encode())
        if not rle or w < 8 or w >= 32768:
            f.write(rgbe.tobytes())
        else:
            for y in range(h):
                f.write(bytes([2, 2, w >> 8, w & 255]))
                for k in range(4):
                    f.write(_rle_channel(rgbe[y, :, k].tolist()))
    return path
