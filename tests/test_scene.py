"""Host logic of the Python side (camera / TF / lights / configs / work splits) -- no GPU."""
import math

import numpy as np
import pytest

from sunvolumerender_b200 import _lib as L
from sunvolumerender_b200 import scene as S


def test_default_camera_follows_canvas_rules():
    cam = S.default_camera((512.0, 512.0, 512.0), 1920, 1080)
    # Canvas::ZoomToExtent, gui/canvas.cpp:191-197
    assert cam.pos.z == pytest.approx(1.5 * 512 / (2 * math.tan(math.radians(22.5))), rel=1e-6)
    assert (cam.pos.x, cam.pos.y) == (0.0, 0.0)
    assert cam.aspectRatio == pytest.approx(1920 / 1080)
    assert cam.tanFovxOverTwo == pytest.approx(math.tan(math.radians(22.5)), rel=1e-6)
    assert (cam.imageW, cam.imageH) == (1920, 1080)
    assert cam.w.tuple() == (0.0, 0.0, 1.0)


def test_look_at_camera_is_orthonormal():
    cam = S.look_at_camera((3, 4, 5), (0, 0, 0), (0, 1, 0), image_w=64, image_h=64)
    u, v, w = (np.array(x.tuple()) for x in (cam.u, cam.v, cam.w))
    assert np.dot(u, w) == pytest.approx(0, abs=1e-6) and np.dot(v, w) == pytest.approx(0, abs=1e-6)
    assert np.linalg.norm(w) == pytest.approx(1, abs=1e-6)
    assert np.allclose(w, np.array([3, 4, 5]) / np.linalg.norm([3, 4, 5]), atol=1e-6)


def test_tf_tables():
    for kind, mx in (("default", 0.5), ("thin", 0.02), ("cloud", 0.5)):
        t = S.tf_table(kind)
        assert t.shape == (1024, 4) and t.dtype == np.float32
        assert t[:, 3].max() == pytest.approx(mx)
        assert t[0, 3] == 0.0
    d = S.tf_table("default")
    assert np.allclose(d[0, :3], [69 / 255, 199 / 255, 186 / 255])
    assert np.allclose(d[-1, :3], [183 / 255, 7 / 255, 140 / 255])
    i01 = int(round(0.1 * 1023))
    assert d[i01 + 1, 3] == pytest.approx(0.5)
    thin = S.tf_table("thin")
    assert (thin[: int(0.1 * 1023), 3] == 0).all()  # everything below intensity 0.1 is empty


def test_default_area_light():
    l = S.default_area_light((128.0, 128.0, 128.0))
    R = 0.5 * math.sqrt(3) * 128
    assert l.disk.radius == pytest.approx(10.0)
    assert l.disk.center.y == pytest.approx(1.5 * R + 1.0)
    assert l.disk.normal.tuple() == (0.0, -1.0, 0.0)
    assert l.intensity == 500.0
    assert S.default_area_light((512.0,) * 3).disk.radius == pytest.approx(40.0)


def test_step_size_and_volume_struct():
    assert S.raycast_step_size() == pytest.approx(0.5 * math.sqrt(3))
    v = S.host_volume_struct((128, 64, 32), spacing=(1.0, 2.0, 0.5), max_grad_mag=100.0)
    assert v.bbox.vmax.tuple() == (64.0, 64.0, 8.0)
    assert v.bbox.vmin.tuple() == (-64.0, -64.0, -8.0)
    assert v.bbox.invSize.x == pytest.approx(1 / 128)
    assert v.invMaxMagnitude == pytest.approx(0.01)
    assert v.x_clip.x == -1.0 and v.x_clip.y == 1.0
    assert v.densityScale == 1.0 and v.gradientFactor == 0.5


def test_sphere_volume():
    v = S.sphere_volume(16, L.VOXEL_U8)
    assert v.shape == (16, 16, 16) and v.dtype == np.uint8
    assert v[0, 0, 0] == 0 and v[8, 8, 8] > 200
    v16 = S.sphere_volume(16, L.VOXEL_U16)
    assert v16.dtype == np.uint16 and abs(int(v16[8, 8, 8]) - 257 * int(v[8, 8, 8])) <= 257


def test_configs_match_baseline_json():
    c = S.CONFIGS
    assert (c["C1"].n, c["C1"].width, c["C1"].height, c["C1"].spp) == (128, 512, 512, 16)
    assert (c["C2"].n, c["C2"].width, c["C2"].height) == (256, 1024, 1024)
    assert (c["C3"].n, c["C3"].fmt, c["C3"].width, c["C3"].height, c["C3"].spp, c["C3"].trace_depth) == (512, L.VOXEL_U16, 1920, 1080, 256, 1)
    assert (c["C4"].n, c["C4"].fmt, c["C4"].trace_depth, c["C4"].spp) == (1024, L.VOXEL_F16, 32, 512)
    assert (c["C5"].n, c["C5"].width, c["C5"].height, c["C5"].spp) == (2048, 3840, 2160, 1024)
    assert c["C5"].n ** 3 * c["C5"].voxel_bytes == 16 * 2 ** 30


@pytest.mark.parametrize("total,world", [(256, 1), (256, 8), (1024, 8), (10, 4), (3, 8)])
def test_split_samples_partitions_exactly(total, world):
    parts = S.split_samples(total, world)
    assert len(parts) == world
    assert sum(c for _, c in parts) == total
    pos = 0
    for first, cnt in parts:
        assert first == pos
        pos += cnt
    assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


@pytest.mark.parametrize("h,world", [(1080, 8), (1080, 3), (512, 2), (7, 4)])
def test_split_rows_partitions_exactly(h, world):
    parts = S.split_rows(h, world)
    assert parts[0][0] == 0 and parts[-1][1] == h
    for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
        assert a1 == b0 and a0 <= a1
