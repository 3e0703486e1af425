"""TEST INFRASTRUCTURE -- not product code.  numpy-side restatement of the transfer-function table the
reference builds with vtkPiecewiseFunction::GetTable / vtkColorTransferFunction::GetTable
(gui/transferfunction.cpp:17-29) and of the `.tf` file layout (gui/transferfunction.cpp:55-126).

PARITY UNPINNED: VTK is an un-vendored, un-versioned dependency (CMakeLists.txt:24) that is not installed
here, and the reference ships no `.tf` file; the node semantics (midpoint, sharpness, clamping) restate
VTK's documented behaviour.  The `.tf` layout is the reference's own code and is restated literally.
"""
import struct

import numpy as np


def _shape(s, y1, y2, midpoint, sharpness):
    midpoint = min(max(midpoint, 0.00001), 0.99999)
    s = 0.5 * s / midpoint if s < midpoint else 0.5 + 0.5 * (s - midpoint) / (1.0 - midpoint)
    if sharpness > 0.99:
        return y1 if s < 0.5 else y2
    if sharpness < 0.01:
        return (1.0 - s) * y1 + s * y2
    if s < 0.5:
        s = 0.5 * (s * 2.0) ** (1.0 + 10.0 * sharpness)
    elif s > 0.5:
        s = 1.0 - 0.5 * ((1.0 - s) * 2.0) ** (1.0 + 10.0 * sharpness)
    ss, sss = s * s, s * s * s
    h1, h2, h3, h4 = 2 * sss - 3 * ss + 1, -2 * sss + 3 * ss, sss - 2 * ss + s, sss - ss
    t = (1.0 - sharpness) * (y2 - y1)
    v = h1 * y1 + h2 * y2 + h3 * t + h4 * t
    return min(max(v, min(y1, y2)), max(y1, y2))


def _nodes(points, nvals):
    out = {}
    for p in points:
        p = tuple(p) + (0.5, 0.0)
        out[float(p[0])] = (float(p[0]), [float(v) for v in p[1:1 + nvals]], float(p[1 + nvals]), float(p[2 + nvals]))
    return [out[k] for k in sorted(out)]


def get_table(points, nvals, size):
    nodes = _nodes(points, nvals)
    table = np.zeros((size, nvals), np.float32)
    idx = 0
    for i in range(size):
        x = i / (size - 1)
        while idx < len(nodes) and x > nodes[idx][0]:
            idx += 1
        for c in range(nvals):
            if not nodes:
                v = 0.0
            elif idx >= len(nodes):
                v = nodes[-1][1][c]
            elif idx == 0:
                v = nodes[0][1][c]
            else:
                a, b = nodes[idx - 1], nodes[idx]
                v = _shape((x - a[0]) / (b[0] - a[0]), a[1][c], b[1][c], a[2], a[3])
            table[i, c] = v
    return table


def composite_table(opacity_points, color_points, size=1024):
    t = np.zeros((size, 4), np.float32)
    t[:, :3] = get_table(color_points, 3, size)
    t[:, 3] = get_table(opacity_points, 1, size)[:, 0]
    return t


def write_tf(path, opacity_points, color_points):
    with open(path, "wb") as f:
        f.write(struct.pack("=i", len(opacity_points)))
        for p in opacity_points:
            f.write(struct.pack("=4d", *((tuple(p) + (0.5, 0.0))[:4])))
        f.write(struct.pack("=i", len(color_points)))
        for p in color_points:
            f.write(struct.pack("=6d", *((tuple(p) + (0.5, 0.0))[:6])))


def read_tf(path):
    raw = open(path, "rb").read()
    n = struct.unpack_from("=i", raw, 0)[0]
    o = [struct.unpack_from("=4d", raw, 4 + 32 * i) for i in range(n)]
    off = 4 + 32 * n
    m = struct.unpack_from("=i", raw, off)[0]
    c = [struct.unpack_from("=6d", raw, off + 4 + 48 * i) for i in range(m)]
    return o, c
