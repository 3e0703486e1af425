"""The transfer-function input stage (include/svr_tf_io.h; gui/transferfunction.cpp:17-29, 55-126; host
code, no GPU) against oracle/tf_oracle.py and closed-form known answers."""
import ctypes as C

import numpy as np
import pytest

from oracle import tf_oracle as T
from sunvolumerender_b200 import _lib as L
from sunvolumerender_b200 import scene as S


def test_linear_nodes_are_plain_interpolation():
    op = [(0.0, 0.0), (0.1, 0.5), (1.0, 0.5)]
    col = [(0.0, 1.0, 0.0, 0.0), (1.0, 0.0, 0.0, 1.0)]
    table, mx = S.build_tf_table(op, col)
    x = np.linspace(0.0, 1.0, 1024)
    assert np.allclose(table[:, 3], np.interp(x, [0, 0.1, 1.0], [0, 0.5, 0.5]), atol=1e-7)
    assert np.allclose(table[:, 0], 1.0 - x, atol=1e-7) and np.allclose(table[:, 2], x, atol=1e-7) and (table[:, 1] == 0).all()
    assert mx == pytest.approx(0.5)
    # this is the table the synthetic configurations call "default" (scene.tf_table), colour nodes aside
    assert np.allclose(table[:, 3], S.tf_table("default")[:, 3], atol=1e-6)


def test_midpoint_and_sharpness():
    # step at the midpoint
    t, _ = S.build_tf_table([(0.0, 0.2, 0.25, 1.0), (1.0, 0.8)], [(0.0, 0, 0, 0), (1.0, 1, 1, 1)])
    x = np.linspace(0, 1, 1024)
    assert (t[x < 0.2499, 3] == np.float32(0.2)).all() and (t[x > 0.2501, 3] == np.float32(0.8)).all()
    # hermite blend: monotone, inside [y1, y2], one half of the way up exactly at the midpoint
    t, _ = S.build_tf_table([(0.0, 0.0, 0.5, 0.5), (1.0, 1.0)], [(0.0, 0, 0, 0), (1.0, 1, 1, 1)], n=1025)
    a = t[:, 3]
    assert (np.diff(a) >= 0).all() and a[0] == 0 and a[-1] == 1 and a[512] == pytest.approx(0.5, abs=1e-6)
    assert a[256] < 0.25 < 0.75 < a[768]                       # sharper than linear
    assert np.allclose(a + a[::-1], 1.0, atol=1e-6)              # symmetric about the centre for midpoint 0.5
    # a shifted midpoint moves the half-way point
    t, _ = S.build_tf_table([(0.0, 0.0, 0.25, 0.0), (1.0, 1.0)], [(0.0, 0, 0, 0), (1.0, 1, 1, 1)], n=1025)
    assert t[256, 3] == pytest.approx(0.5, abs=1e-6)


def test_clamping_outside_the_node_range_and_node_replacement():
    t, mx = S.build_tf_table([(0.3, 0.1), (0.6, 0.9)], [(0.5, 0.2, 0.4, 0.6)])
    x = np.linspace(0, 1, 1024)
    assert (t[x <= 0.3, 3] == np.float32(0.1)).all() and (t[x >= 0.6, 3] == np.float32(0.9)).all()
    assert np.allclose(t[:, :3], [0.2, 0.4, 0.6]) and mx == pytest.approx(0.9)
    # unsorted input, and a repeated x replaces the earlier node (vtkPiecewiseFunction::AddPoint)
    t2, _ = S.build_tf_table([(0.6, 0.5), (0.3, 0.1), (0.6, 0.9)], [(0.5, 0.2, 0.4, 0.6)])
    assert np.array_equal(t, t2)


def test_application_default_matches_oracle_and_reference_shape():
    op, col = S.default_tf_nodes()
    assert len(op) == 11 and len(col) == 6 and op[0] == (0.0, 0.0, 0.5, 0.5) and op[3] == pytest.approx((0.3, 0.5, 0.5, 0.5))
    assert col[1][:4] == pytest.approx((0.2, 172 / 255, 3 / 255, 57 / 255))
    table, mx = S.build_tf_table(op, col)
    exp = T.composite_table(op, col)
    assert np.abs(table - exp).max() <= 1e-7
    assert mx == 0.5 and table[0, 3] == 0 and (table[103:, 3] == 0.5).all()      # plateau from x = 0.1
    assert 0 < table[20, 3] < table[51, 3] < table[80, 3] < 0.5                   # sharpness-0.5 ramp below it


def test_random_node_sets_against_oracle():
    rng = np.random.default_rng(4)
    for _ in range(25):
        no, nc = rng.integers(1, 9), rng.integers(1, 7)
        op = [(float(x), float(rng.uniform(0, 1)), float(rng.uniform(0, 1)), float(rng.choice([0.0, 0.3, 0.7, 1.0, rng.uniform(0, 1)]))) for x in rng.uniform(-0.2, 1.2, no)]
        col = [(float(x),) + tuple(float(v) for v in rng.uniform(0, 1, 3)) + (float(rng.uniform(0, 1)), float(rng.choice([0.0, 0.5, 1.0]))) for x in rng.uniform(-0.2, 1.2, nc)]
        size = int(rng.choice([2, 17, 256, 1024]))
        table, mx = S.build_tf_table(op, col, n=size)
        exp = T.composite_table(op, col, size)
        assert np.abs(table - exp).max() <= 2e-7
        assert mx == exp[:, 3].max()


def test_tf_file_round_trip_and_layout(tmp_path):
    lib = L.load()
    op, col = S.default_tf_nodes()
    on = (L.TfOpacityNode * len(op))(*[L.TfOpacityNode(*p) for p in op])
    cn = (L.TfColorNode * len(col))(*[L.TfColorNode(*p) for p in col])
    path = tmp_path / "default.tf"
    assert lib.svr_tf_file_write(str(path).encode(), on, len(op), cn, len(col)) == 0
    # byte layout of gui/transferfunction.cpp:69-86: int, n x 4 doubles, int, m x 6 doubles
    assert path.stat().st_size == 4 + 11 * 32 + 4 + 6 * 48
    o2, c2 = T.read_tf(path)
    assert [tuple(p) for p in o2] == op and [tuple(p) for p in c2] == col
    # and a file written the reference's way reads back through the C ABI
    T.write_tf(tmp_path / "other.tf", [(0.0, 0.0, 0.5, 0.0), (0.5, 1.0, 0.3, 0.9)], [(0.0, 1, 0, 0, 0.5, 0), (1.0, 0, 0, 1, 0.5, 0)])
    ro, rc = (L.TfOpacityNode * 8)(), (L.TfColorNode * 8)()
    no, nc = C.c_uint32(8), C.c_uint32(8)
    assert lib.svr_tf_file_read(str(tmp_path / "other.tf").encode(), ro, C.byref(no), rc, C.byref(nc)) == 0
    assert (no.value, nc.value) == (2, 2) and (ro[1].x, ro[1].y, ro[1].midpoint, ro[1].sharpness) == (0.5, 1.0, 0.3, 0.9)
    assert (rc[1].x, rc[1].b) == (1.0, 1.0)
    # errors: capacity, truncation, missing file
    no, nc = C.c_uint32(1), C.c_uint32(8)
    assert lib.svr_tf_file_read(str(tmp_path / "other.tf").encode(), ro, C.byref(no), rc, C.byref(nc)) != 0
    assert b"capacity" in lib.svr_last_error()
    (tmp_path / "cut.tf").write_bytes(path.read_bytes()[:100])
    no, nc = C.c_uint32(8), C.c_uint32(8)
    assert lib.svr_tf_file_read(str(tmp_path / "cut.tf").encode(), ro, C.byref(no), rc, C.byref(nc)) != 0
    assert lib.svr_tf_file_read(str(tmp_path / "none.tf").encode(), ro, C.byref(no), rc, C.byref(nc)) != 0
