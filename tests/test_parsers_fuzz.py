"""The host-side file readers (MetaImage header, .tf, Radiance .hdr) refuse malformed input with an error code:
no crash, no out-of-bounds write, whatever the bytes (CPU; seeded, so a failure reproduces)."""
import ctypes as C
import os
import random

from sunvolumerender_b200 import _lib as L


def test_metaimage_header_reader_survives_garbage(tmp_path):
    lib = L.load()
    rnd = random.Random(1)
    keys = ["ObjectType", "NDims", "DimSize", "ElementSpacing", "ElementSize", "ElementType", "ElementByteOrderMSB", "BinaryDataByteOrderMSB",
            "CompressedData", "CompressedDataSize", "HeaderSize", "ElementNumberOfChannels", "ElementDataFile", "BinaryData", "Offset", "junk"]
    vals = ["3", "2", "0", "-1", "4 4 4", "4 4", "4 4 4 4", "99999999999 2 2", "-3 4 4", "1.0 1.0 1.0", "0 0 0", "nan 1 1", "MET_SHORT", "MET_FOO", "True",
            "False", "LOCAL", "LIST", "x.raw", "/nonexistent/" + "a" * 2000, "", "=", "1e40", "0x10", "4294967296 1 1", "65536 65536 65536"]
    for it in range(600):
        lines = [rnd.choice(keys) + rnd.choice([" = ", "=", " ", ":", " =", "= "]) + rnd.choice(vals) for _ in range(rnd.randint(0, 12))]
        if rnd.random() < 0.5:
            lines.append("ElementDataFile = " + rnd.choice(["LOCAL", "x.raw", ""]))
        body = ("\n".join(lines) + "\n").encode()
        if rnd.random() < 0.3:
            body += bytes(rnd.getrandbits(8) for _ in range(rnd.randint(0, 64)))
        p = tmp_path / ("f.mha" if rnd.random() < 0.5 else "f.mhd")
        p.write_bytes(body)
        h = L.MetaImageHeader()
        rc = lib.svr_metaimage_read_header(str(p).encode(), C.byref(h))
        if rc == 0:  # whatever is accepted is sane
            assert 1 <= h.ndims <= 3 and all(0 < d < 1 << 20 for d in list(h.dim)[: h.ndims])
    # a well-formed header is still accepted afterwards
    good = tmp_path / "g.mhd"
    good.write_text("NDims = 3\nDimSize = 4 5 6\nElementType = MET_USHORT\nElementDataFile = g.raw\n")
    (tmp_path / "g.raw").write_bytes(b"\0" * 240)
    h = L.MetaImageHeader()
    assert lib.svr_metaimage_read_header(str(good).encode(), C.byref(h)) == 0 and list(h.dim) == [4, 5, 6]


def test_tf_file_reader_survives_garbage(tmp_path):
    lib = L.load()
    rnd = random.Random(2)
    on, cn = (L.TfOpacityNode * 256)(), (L.TfColorNode * 256)()
    p = tmp_path / "f.tf"
    for it in range(400):
        p.write_bytes(bytes(rnd.getrandbits(8) for _ in range(rnd.randint(0, 200))))
        no, nc = C.c_uint32(256), C.c_uint32(256)
        rc = lib.svr_tf_file_read(str(p).encode(), on, C.byref(no), cn, C.byref(nc))
        assert rc != 0 or (no.value <= 256 and nc.value <= 256)
    # capacity is respected: a file with more nodes than the caller has room for is an error, not an overrun
    import struct

    p.write_bytes(struct.pack("<i", 300) + struct.pack("<4d", 0, 0, 0.5, 0) * 300 + struct.pack("<i", 2) + struct.pack("<6d", 0, 0, 0, 0, 0.5, 0) * 2)
    small_o = (L.TfOpacityNode * 4)()
    no, nc = C.c_uint32(4), C.c_uint32(256)
    assert lib.svr_tf_file_read(str(p).encode(), small_o, C.byref(no), cn, C.byref(nc)) != 0


def test_hdr_reader_survives_garbage(tmp_path):
    lib = L.load()
    rnd = random.Random(3)
    heads = [b"#?RADIANCE\n", b"#?RGBE\n", b"", b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 4 +X 4\n", b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 99999 +X 99999\n",
             b"#?RADIANCE\n\n-Y -1 +X 8\n", b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 2 +X 16\n\x02\x02\x00\x10"]
    p = tmp_path / "f.hdr"
    for it in range(400):
        p.write_bytes(rnd.choice(heads) + bytes(rnd.getrandbits(8) for _ in range(rnd.randint(0, 300))))
        w, h = C.c_uint32(0), C.c_uint32(0)
        rc = lib.svr_hdr_read(str(p).encode(), None, C.byref(w), C.byref(h))
        if rc == 0 and 0 < w.value * h.value < 1 << 16:
            buf = (C.c_float * (3 * w.value * h.value))()
            lib.svr_hdr_read(str(p).encode(), buf, C.byref(w), C.byref(h))  # truncated pixel data: an error, not a crash
    assert os.path.exists(p)
