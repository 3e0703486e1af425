"""MetaImage header parsing through the C ABI (include/svr_volume_io.h) -- host code, no GPU -- against
the numpy-side writer/parser in oracle/metaimage_oracle.py."""
import ctypes as C

import numpy as np
import pytest

from oracle import metaimage_oracle as M
from sunvolumerender_b200 import _lib as L


def _header(path):
    lib = L.load()
    h = L.MetaImageHeader()
    rc = lib.svr_metaimage_read_header(str(path).encode(), C.byref(h))
    return rc, h, lib.svr_last_error().decode()


@pytest.mark.parametrize("et", sorted(M.MET))
def test_element_types_and_fields(tmp_path, et):
    data = np.arange(2 * 3 * 5).reshape(2, 3, 5)
    p = M.write_metaimage(tmp_path / "v.mhd", data, spacing=(0.5, 0.75, 2.0), element_type=et, msb=(et == "MET_SHORT"))
    rc, h, err = _header(p)
    assert rc == 0, err
    assert (h.ndims, list(h.dim)) == (3, [5, 3, 2])
    assert list(h.spacing) == [0.5, 0.75, 2.0]
    assert h.element_type == M.MET_INDEX[et] and h.channels == 1
    assert bool(h.msb) == (et == "MET_SHORT") and not h.compressed
    assert h.data_file.decode() == str(tmp_path / "v.raw") and h.header_size == 0


def test_local_data_offset_and_compression_fields(tmp_path):
    data = np.arange(4 * 4 * 4, dtype=np.int16).reshape(4, 4, 4)
    p = M.write_metaimage(tmp_path / "v.mha", data, compressed=True)
    rc, h, err = _header(p)
    assert rc == 0, err
    ref = M.parse_header(p)
    assert h.data_file == b"" and h.data_offset == ref["_data_offset"]
    assert h.compressed and h.compressed_size == int(ref["CompressedDataSize"])
    p2 = M.write_metaimage(tmp_path / "w.mhd", data, header_size=-1)
    rc, h, _ = _header(p2)
    assert rc == 0 and h.header_size == -1


def test_element_size_is_the_fallback_for_spacing(tmp_path):
    p = tmp_path / "s.mhd"
    p.write_text("ObjectType = Image\nNDims = 3\nDimSize = 2 2 2\nElementSize = 3 4 5\nElementType = MET_UCHAR\nElementDataFile = s.raw\n")
    rc, h, err = _header(p)
    assert rc == 0, err
    assert list(h.spacing) == [3.0, 4.0, 5.0]
    p.write_text("NDims = 2\nDimSize = 7 9\nElementType = MET_FLOAT\nElementDataFile = LOCAL\n")
    rc, h, err = _header(p)
    assert rc == 0 and list(h.dim) == [7, 9, 1] and list(h.spacing) == [1.0, 1.0, 1.0]


@pytest.mark.parametrize("text,needle", [
    ("NDims = 3\nDimSize = 2 2 2\nElementType = MET_LONG_LONG\nElementDataFile = a.raw\n", "ElementType"),
    ("NDims = 3\nDimSize = 2 2 2\nElementType = MET_SHORT\nBinaryData = False\nElementDataFile = a.raw\n", "ASCII"),
    ("NDims = 3\nDimSize = 2 2 2\nElementType = MET_SHORT\nElementDataFile = LIST\n", "multi-file"),
    ("NDims = 3\nDimSize = 2 2 2\nElementType = MET_SHORT\nElementDataFile = slice%03d.raw 1 2 1\n", "multi-file"),
    ("NDims = 3\nElementType = MET_SHORT\nElementDataFile = a.raw\n", "required"),
    ("NDims = 3\nDimSize = 2 0 2\nElementType = MET_SHORT\nElementDataFile = a.raw\n", "zero dimension"),
    ("NDims = 4\nDimSize = 2 2 2 2\nElementType = MET_SHORT\nElementDataFile = a.raw\n", "NDims"),
])
def test_rejected_headers(tmp_path, text, needle):
    p = tmp_path / "bad.mhd"
    p.write_text(text)
    rc, _, err = _header(p)
    assert rc != 0 and needle in err


def test_missing_file(tmp_path):
    rc, _, err = _header(tmp_path / "nope.mhd")
    assert rc != 0 and "cannot open" in err


def test_oracle_preprocess_known_answers():
    """Hand-checked values of the numpy restatement (VolumeReader.cpp:52-76)."""
    d = np.zeros((3, 3, 3), np.int16)
    d[1, 1, 1] = 100
    d[0, 0, 0] = -50
    r = M.preprocess(d)
    assert (r["data_min"], r["data_max"]) == (-50, 100)
    assert r["u16"][1, 1, 1] == 65535 and r["u16"][0, 0, 0] == 0
    assert r["u16"][2, 2, 2] == int(np.float32(50) / np.float32(150) * np.float32(65535))  # 21845
    # bins = 150; zero voxels ignored; the maximum (bin 150) falls outside the extent
    assert len(r["histogram"]) == 150 and r["histogram_total"] == 1 and r["histogram"][0] == 1
    # |grad| at a face neighbour of the centre: (100 - 0) * 0.5 = 50; at the corner: one-sided, (-50 - 0) * 0.5 on three axes
    assert r["max_gradient_magnitude"] == 50
    # unsigned data above 32767 wraps in the cast to short
    assert M.cast_to_short(np.array([40000], np.uint16))[0] == 40000 - 65536
    assert M.cast_to_short(np.array([-3.7, 2.9], np.float32)).tolist() == [-3, 2]
