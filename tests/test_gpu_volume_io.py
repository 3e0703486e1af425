"""GPU tests of the volume input stage (include/svr_volume_io.h; core/VolumeReader.cpp:13-94, 124-185)
against the numpy restatement oracle/metaimage_oracle.py: bit-exact for every integer result."""
import numpy as np
import pytest
import torch

from oracle import metaimage_oracle as M
from sunvolumerender_b200 import _lib as L
from sunvolumerender_b200 import scene as S

from _gpu_common import raycast_f32

pytestmark = pytest.mark.gpu


def _ct_like(n, dtype, rng, lo, hi):
    z, y, x = np.meshgrid(*[np.linspace(-1, 1, n)] * 3, indexing="ij")
    r = np.sqrt(x * x / 0.6 + y * y / 0.4 + z * z / 0.7)
    v = np.where(r < 1, np.where(r > 0.8, 0.9, 0.4 + 0.1 * np.sin(7 * x) * np.cos(5 * y)), 0.0)
    v = lo + v * (hi - lo) + rng.normal(0, 0.01 * (hi - lo), v.shape) * (r < 1)
    if np.dtype(dtype).kind == "f":
        return v.astype(dtype)
    info = np.iinfo(dtype)
    return np.clip(np.round(v), info.min, info.max).astype(dtype)


def _check(renderer, stats, hist, data, spacing):
    exp = M.preprocess(data, spacing)
    n = data.shape
    assert list(stats.dim) == [n[2], n[1], n[0]]
    assert list(stats.spacing) == [np.float32(s) for s in spacing]
    assert (stats.data_min, stats.data_max) == (exp["data_min"], exp["data_max"])
    assert stats.histogram_bins == max(exp["data_max"] - exp["data_min"], 0)
    assert stats.histogram_total == exp["histogram_total"]
    assert np.array_equal(hist[: len(exp["histogram"])], exp["histogram"][: len(hist)])
    assert stats.max_gradient_magnitude == exp["max_gradient_magnitude"]
    got = renderer.download_volume((n[2], n[1], n[0]))
    assert np.array_equal(got, exp["u16"])
    v = renderer.volume
    # VolumeReader::CreateDeviceVolume (VolumeReader.cpp:174-185): bbox = +-0.5 * dim * spacing, invMaxMagnitude
    assert v.bbox.vmax.x == pytest.approx(0.5 * n[2] * spacing[0]) and v.bbox.vmin.z == pytest.approx(-0.5 * n[0] * spacing[2])
    assert v.invMaxMagnitude == pytest.approx(1.0 / max(exp["max_gradient_magnitude"], 1))
    assert (v.spacing.x, v.spacing.y, v.spacing.z) == tuple(np.float32(s) for s in spacing)
    return exp


CASES = [
    ("MET_SHORT", np.int16, -1000, 3000, False, False, ".mhd"),
    ("MET_SHORT", np.int16, -1000, 3000, True, False, ".mha"),     # big-endian, LOCAL data
    ("MET_USHORT", np.uint16, 0, 4000, False, True, ".mhd"),       # zlib-compressed .zraw
    ("MET_USHORT", np.uint16, 0, 60000, False, False, ".mha"),     # values above 32767 wrap in the cast to short
    ("MET_UCHAR", np.uint8, 0, 255, False, False, ".mhd"),
    ("MET_CHAR", np.int8, -100, 100, False, True, ".mha"),
    ("MET_INT", np.int32, -2000, 2000, True, False, ".mhd"),
    ("MET_UINT", np.uint32, 0, 5000, False, False, ".mhd"),
    ("MET_FLOAT", np.float32, -500.5, 1200.25, False, False, ".mhd"),
    ("MET_DOUBLE", np.float64, -3.75, 900.5, True, True, ".mhd"),
]


@pytest.mark.parametrize("et,dtype,lo,hi,msb,comp,ext", CASES)
def test_load_metaimage_matches_oracle(renderer, tmp_path, et, dtype, lo, hi, msb, comp, ext):
    rng = np.random.default_rng(11)
    n = 40
    data = _ct_like(n, dtype, rng, lo, hi)[:, : n - 7, : n - 3]    # ragged dims
    spacing = (0.7, 1.0, 1.3)
    path = M.write_metaimage(tmp_path / ("vol" + ext), data, spacing=spacing, element_type=et, msb=msb, compressed=comp)
    stats, hist = renderer.load_metaimage(path)
    _check(renderer, stats, hist, data, spacing)


def test_header_size_variants(renderer, tmp_path):
    rng = np.random.default_rng(3)
    data = _ct_like(24, np.int16, rng, -200, 900)
    for hs in (64, -1):
        path = M.write_metaimage(tmp_path / f"h{hs}.mhd", data, header_size=hs)
        stats, hist = renderer.load_metaimage(path)
        _check(renderer, stats, hist, data, (1.0, 1.0, 1.0))


def test_from_raw_equals_file_path_and_renders_like_a_plain_u16_volume(renderer, tmp_path):
    rng = np.random.default_rng(5)
    n = 48
    data = _ct_like(n, np.int16, rng, -1000, 2500)
    stats, hist = renderer.load_raw(data, L.MET_SHORT, (n, n, n))
    exp = _check(renderer, stats, hist, data, (1.0, 1.0, 1.0))
    renderer.set_transfer_function(S.tf_table("default"))
    renderer.set_camera(S.default_camera((n,) * 3, 96, 80))
    a = raycast_f32(renderer).clone()
    assert float(a[..., 3].max()) > 0.5
    # the same voxels through the generic builder give the same picture
    renderer.load_volume(exp["u16"], L.VOXEL_U16, (n, n, n), max_grad_mag=float(exp["max_gradient_magnitude"]))
    b = raycast_f32(renderer)
    assert torch.equal(a, b)


def test_constant_volume_and_errors(renderer, tmp_path):
    data = np.full((8, 8, 8), 7, np.int16)
    stats, hist = renderer.load_raw(data, L.MET_SHORT, (8, 8, 8))
    assert (stats.data_min, stats.data_max, stats.histogram_bins, stats.max_gradient_magnitude) == (7, 7, 0, 0)
    assert renderer.download_volume((8, 8, 8)).max() == 0          # extent 0: the reference divides by zero here; we store 0
    assert renderer.volume.invMaxMagnitude == 1.0
    # truncated data file
    path = M.write_metaimage(tmp_path / "t.mhd", np.zeros((4, 4, 4), np.int16))
    with open(tmp_path / "t.raw", "wb") as f:
        f.write(b"\0" * 10)
    with pytest.raises(L.SvrError, match="shorter"):
        renderer.load_metaimage(path)
    bad = tmp_path / "c.mhd"
    bad.write_text("NDims = 3\nDimSize = 2 2 2\nElementNumberOfChannels = 3\nElementType = MET_UCHAR\nElementDataFile = LOCAL\n" + "x" * 24)
    with pytest.raises(L.SvrError, match="single-channel"):
        renderer.load_metaimage(bad)
    # the renderer still works afterwards
    renderer.load_raw(_ct_like(16, np.uint8, np.random.default_rng(1), 0, 255), L.MET_UCHAR, (16, 16, 16))
