"""Checker for include/svr_canvas.h (test infrastructure, never imported by the product): a numpy restatement of the
camera manipulation in the reference's Canvas.

Follows gui/canvas.cpp:119-226 (mouse / wheel / key handlers, UpdateCamera, ZoomToExtent), gui/canvas.h:160-164
(PixelPosToViewPos), core/cuda_camera.h:34-47 (cudaCamera::Setup), and -- for glm::lookAt / glm::rotate / glm::radians
-- the published formulas of GLM's glm/gtc/matrix_transform.inl.  GLM is an un-vendored, un-versioned dependency of the
reference (CMakeLists.txt:37) and is not installed here, and the reference has no tests: PARITY UNPINNED for this stage.

All arithmetic is done in float32 where the C++ expression is float, in float64 where Qt's QPointF (qreal = double) or a
double literal promotes it.  Matrices are numpy arrays m[c][r] (column-major, like glm::mat4).
"""
import math

import numpy as np

f32 = np.float32
BUTTON_LEFT, BUTTON_MID = 1, 4
KEY_LEFT, KEY_RIGHT, KEY_DOWN = 0, 1, 2


def _normalize(v):
    v = np.asarray(v, f32)
    return v * (f32(1) / np.sqrt(np.dot(v, v), dtype=f32))


def look_at(eye, center, up):
    eye, center, up = (np.asarray(a, f32) for a in (eye, center, up))
    f = _normalize(center - eye)
    s = _normalize(np.cross(f, up).astype(f32))
    u = np.cross(s, f).astype(f32)
    m = np.eye(4, dtype=f32)
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -f
    m[3][0], m[3][1], m[3][2] = -np.dot(s, eye), -np.dot(u, eye), np.dot(f, eye)
    return m


def rotate(m, angle, axis):
    angle = f32(angle)
    c, s = f32(math.cos(float(angle))), f32(math.sin(float(angle)))
    axis = _normalize(axis)
    temp = (f32(1) - c) * axis
    R = np.zeros((3, 3), f32)
    R[0][0] = c + temp[0] * axis[0]
    R[0][1] = temp[0] * axis[1] + s * axis[2]
    R[0][2] = temp[0] * axis[2] - s * axis[1]
    R[1][0] = temp[1] * axis[0] - s * axis[2]
    R[1][1] = c + temp[1] * axis[1]
    R[1][2] = temp[1] * axis[2] + s * axis[0]
    R[2][0] = temp[2] * axis[0] + s * axis[1]
    R[2][1] = temp[2] * axis[1] - s * axis[0]
    R[2][2] = c + temp[2] * axis[2]
    out = np.zeros((4, 4), f32)
    for col in range(3):
        out[col] = m[0] * R[col][0] + m[1] * R[col][1] + m[2] * R[col][2]
    out[3] = m[3]
    return out


class View:
    """The camera-related members of Canvas (gui/canvas.h:210-217) and its event handlers."""

    def __init__(self, width, height):
        self.w, self.h = width, height
        self.m = np.eye(4, dtype=f32)
        self.eye_dist = f32(0)
        self.translate = np.zeros(2, f32)
        self.fov, self.apeture, self.focal_length, self.exposure = f32(45), f32(0), f32(1), f32(1)
        self.mouse_start = np.zeros(2, np.float64)

    def pixel_to_view(self, px, py):  # canvas.h:160-164: float arithmetic, stored in a QPointF
        return np.array([f32(2) * f32(px) / f32(self.w) - f32(1), f32(1) - f32(2) * f32(py) / f32(self.h)], np.float64)

    def zoom_to_extent(self, size):  # canvas.cpp:191-197
        span = f32(max(f32(s) for s in size)) * f32(1.5)
        half = f32(self.fov * f32(0.5)) * f32(0.01745329251994329576923690768489)
        self.eye_dist = f32(span / (f32(2) * f32(math.tan(float(half)))))

    def reset(self, size):  # canvas.cpp:35-38
        self.zoom_to_extent(size)
        self.m = look_at((0, 0, self.eye_dist), (0, 0, 0), (0, 1, 0))

    def mouse_press(self, px, py, buttons):  # canvas.cpp:119-128
        if buttons & (BUTTON_LEFT | BUTTON_MID):
            self.mouse_start = self.pixel_to_view(px, py)

    def mouse_move(self, px, py, buttons, size):  # canvas.cpp:135-169
        now = self.pixel_to_view(px, py)
        dx, dy = now - self.mouse_start
        if buttons & BUTTON_LEFT:
            self.m = rotate(self.m, f32(math.radians(dy * 100.0)), (1, 0, 0))
            self.m = rotate(self.m, f32(math.radians(-dx * 100.0)), (0, 1, 0))
        if buttons & BUTTON_MID:
            base = np.sqrt(np.sum(np.asarray(size, f32) ** 2, dtype=f32), dtype=f32) * f32(0.5)
            self.translate[0] += f32(dx * float(base))
            self.translate[1] += f32(dy * float(base))
        self.mouse_start = now

    def wheel(self, delta, size):  # canvas.cpp:171-177
        self.eye_dist = f32(self.eye_dist + f32(delta) * np.sqrt(np.sum(np.asarray(size, f32) ** 2, dtype=f32), dtype=f32) * f32(0.001))

    def key(self, key):  # canvas.cpp:198-226
        deg = {KEY_DOWN: 180.0, KEY_LEFT: 90.0, KEY_RIGHT: -90.0}[key]
        self.m = rotate(self.m, f32(deg) * f32(0.01745329251994329576923690768489), (0, 1, 0))

    def camera(self):  # canvas.cpp:179-188 + cuda_camera.h:34-47
        u, v, w = self.m[0][:3], self.m[1][:3], self.m[2][:3]
        pos = w * self.eye_dist - u * self.translate[0] - v * self.translate[1]
        tan_half = f32(math.tan(float(f32(float(self.fov * f32(0.5)) * math.pi / 180.0))))
        return dict(pos=pos.astype(f32), u=u.copy(), v=v.copy(), w=w.copy(), aspect=f32(self.w) / f32(self.h), tan_half=tan_half)
