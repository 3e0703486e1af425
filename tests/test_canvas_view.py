"""CPU tests of the camera manipulation behind include/svr_canvas.h (svr_view_*: host arithmetic, no GPU) against
oracle/canvas_oracle.py, plus the closed forms the reference's handlers imply (gui/canvas.cpp:119-226)."""
import math

import numpy as np
import pytest

from oracle import canvas_oracle as O
from sunvolumerender_b200 import _lib as L
from sunvolumerender_b200 import scene as S
from sunvolumerender_b200.canvas import View

W, H = 640, 640   # the reference's WIDTH x HEIGHT (common.h:8-9)
SIZE = (180.0, 215.0, 126.5)


def _same(view, ov, tol=2e-6):
    m = view.matrix()
    assert np.allclose(m, ov.m, rtol=0, atol=tol), np.abs(m - ov.m).max()
    assert view.v.eyeDist == pytest.approx(float(ov.eye_dist), rel=1e-6)
    assert list(view.v.translate) == pytest.approx([float(x) for x in ov.translate], rel=1e-6, abs=1e-6)
    cam, oc = view.camera(), ov.camera()
    scale = max(1.0, float(np.abs(oc["pos"]).max()))
    assert np.allclose(cam.pos.tuple(), oc["pos"], rtol=0, atol=4e-6 * scale)
    for name in ("u", "v", "w"):
        assert np.allclose(getattr(cam, name).tuple(), oc[name], rtol=0, atol=tol)
    assert cam.tanFovxOverTwo == pytest.approx(float(oc["tan_half"]), rel=1e-7)
    assert cam.aspectRatio == pytest.approx(float(oc["aspect"]), rel=1e-7)


def test_fresh_view_is_the_camera_a_loaded_volume_gets():
    v, o = View(W, H), O.View(W, H)
    assert np.array_equal(v.matrix(), np.eye(4, dtype=np.float32)) and (v.v.fov, v.v.apeture, v.v.focalLength, v.v.exposure) == (45.0, 0.0, 1.0, 1.0)
    v.reset(SIZE)
    o.reset(SIZE)
    _same(v, o)
    # Canvas::LoadVolume: eye on +z at ZoomToExtent's distance, u v w = x y z (canvas.cpp:35-38, 179-188)
    cam = v.camera()
    d = 1.5 * max(SIZE) / (2.0 * math.tan(math.radians(22.5)))
    assert cam.pos.tuple() == pytest.approx((0.0, 0.0, d), rel=1e-6)
    assert (cam.u.tuple(), cam.v.tuple(), cam.w.tuple()) == ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0))
    ref = S.default_camera(SIZE, W, H)
    assert cam.tanFovxOverTwo == ref.tanFovxOverTwo and cam.aspectRatio == ref.aspectRatio
    assert (cam.imageW, cam.imageH, cam.exposure, cam.focalLength, cam.apeture) == (W, H, 1.0, 1.0, 0.0)
    # the view matrix keeps glm::lookAt's translation column
    assert v.matrix()[3][2] == pytest.approx(-d, rel=1e-6)


def test_pixel_to_view():
    lib = L.load()
    import ctypes as C

    out = (C.c_float * 2)()
    for px, py, exp in ((0, 0, (-1, 1)), (W, H, (1, -1)), (W / 2, H / 2, (0, 0)), (160, 480, (-0.5, -0.5))):
        lib.svr_view_pixel_to_view(W, H, px, py, C.byref(out))
        assert tuple(out) == pytest.approx(exp, abs=1e-7)


def test_scripted_interaction_follows_the_oracle():
    v, o = View(W, H), O.View(W, H)
    v.reset(SIZE)
    o.reset(SIZE)
    rng = np.random.default_rng(4)
    pos = np.array([320.0, 320.0])
    for step in range(200):
        kind = rng.integers(0, 10)
        if kind < 5:      # drag with the left and / or middle button
            buttons = int(rng.choice([L.BUTTON_LEFT, L.BUTTON_MID, L.BUTTON_LEFT | L.BUTTON_MID]))
            v.mouse_press(float(pos[0]), float(pos[1]), buttons)
            o.mouse_press(float(pos[0]), float(pos[1]), buttons)
            for _ in range(int(rng.integers(1, 5))):
                pos = np.clip(pos + rng.normal(0, 15, 2), 0, W)
                assert v.mouse_move(float(pos[0]), float(pos[1]), buttons, SIZE) == 1
                o.mouse_move(float(pos[0]), float(pos[1]), buttons, SIZE)
        elif kind < 7:
            delta = int(rng.choice([-240, -120, 120, 240]))
            v.wheel(delta, SIZE)
            o.wheel(delta, SIZE)
        elif kind < 9:
            key = int(rng.choice([L.KEY_LEFT, L.KEY_RIGHT, L.KEY_DOWN]))
            assert v.key(key) == 1
            o.key(key)
        else:             # a move with no button pressed changes nothing but the anchor
            pos = np.clip(pos + rng.normal(0, 15, 2), 0, W)
            assert v.mouse_move(float(pos[0]), float(pos[1]), 0, SIZE) == 0
            o.mouse_move(float(pos[0]), float(pos[1]), 0, SIZE)
        _same(v, o, tol=2e-5)   # 200 chained float32 rotations
    # the rotation part stays orthonormal
    R = v.matrix()[:3, :3].astype(np.float64)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-4)
    assert v.key(99) == 0


def test_rotation_conventions():
    # Key_Left turns the view by +90 degrees about y: glm::rotate post-multiplies, so the camera axes become
    # u = (0,0,-1), w = (1,0,0): the eye moves to +x
    v = View(W, H)
    v.reset(SIZE)
    v.key(L.KEY_LEFT)
    cam = v.camera()
    assert cam.u.tuple() == pytest.approx((0, 0, -1), abs=1e-6) and cam.w.tuple() == pytest.approx((1, 0, 0), abs=1e-6)
    assert cam.pos.x == pytest.approx(v.v.eyeDist, rel=1e-6) and abs(cam.pos.z) < 1e-3
    v.key(L.KEY_RIGHT)
    assert v.camera().w.tuple() == pytest.approx((0, 0, 1), abs=1e-6)
    # a left-button drag by a quarter of the widget height (delta.y = -0.5 view units... up is positive) rotates 50 degrees about x
    v.mouse_press(320, 320, L.BUTTON_LEFT)
    v.mouse_move(320, 160, L.BUTTON_LEFT, SIZE)
    w = v.camera().w.tuple()
    assert w[1] == pytest.approx(-math.sin(math.radians(50.0)), abs=1e-5) and w[2] == pytest.approx(math.cos(math.radians(50.0)), abs=1e-5)
    # middle-button drag: translation by delta * |size| / 2, the eye moves against it along u
    v2 = View(W, H)
    v2.reset(SIZE)
    v2.mouse_press(320, 320, L.BUTTON_MID)
    v2.mouse_move(352, 320, L.BUTTON_MID, SIZE)
    half_diag = 0.5 * math.sqrt(sum(s * s for s in SIZE))
    assert v2.v.translate[0] == pytest.approx(0.1 * half_diag, rel=1e-5) and v2.v.translate[1] == 0.0
    assert v2.camera().pos.x == pytest.approx(-0.1 * half_diag, rel=1e-5)
    # wheel: eyeDist += delta * |size| / 1000
    d0 = v2.v.eyeDist
    v2.wheel(120, SIZE)
    assert v2.v.eyeDist - d0 == pytest.approx(120 * 2 * half_diag * 0.001, rel=1e-5)
