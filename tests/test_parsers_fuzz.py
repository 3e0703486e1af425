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


def test_headers_that_promise_absurd_sizes_are_errors_not_aborts(tmp_path):
    """DimSize 200000^3 with CompressedData, dims whose product wraps size_t, and a .hdr of 3e6 x 3e6 used to
    end in std::bad_alloc -> std::terminate across the C boundary.  Each is an error code now, before any allocation."""
    lib = L.load()
    cases = {
        "huge_compressed": b"ObjectType = Image\nNDims = 3\nDimSize = 200000 200000 200000\nElementType = MET_UCHAR\nCompressedData = True\nElementDataFile = LOCAL\n" + b"x" * 64,
        "wraps_size_t": b"NDims = 3\nDimSize = 4194304 2097152 2097152\nElementType = MET_USHORT\nElementDataFile = LOCAL\n" + b"x" * 64,
        "max_dims_tiny_zlib": b"NDims = 3\nDimSize = 16384 16384 16384\nElementType = MET_DOUBLE\nCompressedData = True\nElementDataFile = LOCAL\n" + b"x" * 64,
        "max_dims_raw": b"NDims = 3\nDimSize = 16384 16384 16384\nElementType = MET_DOUBLE\nElementDataFile = LOCAL\n" + b"x" * 64,
        "negative_dim": b"NDims = 3\nDimSize = -4 8 8\nElementType = MET_UCHAR\nElementDataFile = LOCAL\n" + b"x" * 600,
        "header_size_beyond_file": b"NDims = 3\nDimSize = 2 2 2\nElementType = MET_UCHAR\nHeaderSize = 99999999\nElementDataFile = d.raw\n",
    }
    (tmp_path / "d.raw").write_bytes(b"12345678")
    vol, stats = L.Volume(), L.VolumeStats()
    for name, blob in cases.items():
        p = tmp_path / (name + ".mha")
        p.write_bytes(blob)
        rc = lib.svr_volume_load_metaimage(str(p).encode(), C.byref(vol), C.byref(stats), None, 0)
        assert rc != 0, name
        assert lib.svr_last_error(), name
    for dims in (b"-Y 3000000 +X 3000000", b"-Y 32768 +X 32768", b"-Y 20000 +X 20000", b"-Y 4096 +X 4096"):
        p = tmp_path / "big.hdr"
        p.write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n" + dims + b"\n" + b"\x02\x02\x10\x00" * 8)
        w, h = C.c_uint32(0), C.c_uint32(0)
        assert lib.svr_hdr_read(str(p).encode(), None, C.byref(w), C.byref(h)) != 0
        env = L.EnvLight()
        assert lib.svr_env_load_hdr(str(p).encode(), C.byref(env)) != 0   # fails before it touches the GPU


def test_hdr_buffer_sized_for_another_picture_is_an_error(tmp_path):
    """svr_hdr_read's second call decodes into a caller buffer: *w, *h are in-out, so a file that changed between the
    sizing call and the decoding call cannot overrun the buffer."""
    lib = L.load()
    p = tmp_path / "a.hdr"

    def flat(w, h):
        return b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (h, w) + bytes([10, 20, 30, 128]) * (w * h)

    p.write_bytes(flat(4, 4))
    w, h = C.c_uint32(0), C.c_uint32(0)
    assert lib.svr_hdr_read(str(p).encode(), None, C.byref(w), C.byref(h)) == 0 and (w.value, h.value) == (4, 4)
    buf = (C.c_float * (3 * 16))()
    p.write_bytes(flat(6, 6))   # the file grows behind the caller's back
    assert lib.svr_hdr_read(str(p).encode(), buf, C.byref(w), C.byref(h)) != 0
    p.write_bytes(flat(4, 4))
    w, h = C.c_uint32(4), C.c_uint32(4)
    assert lib.svr_hdr_read(str(p).encode(), buf, C.byref(w), C.byref(h)) == 0
    assert buf[0] == 10 * 2.0 ** (128 - 136)
