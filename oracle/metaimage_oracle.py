"""TEST INFRASTRUCTURE -- not product code.  numpy restatement of the volume input stage of the reference
(core/VolumeReader.cpp:13-94, 124-136), the checker for include/svr_volume_io.h.

PARITY UNPINNED: the reference does this stage with VTK filters (vtkMetaImageReader, vtkImageCast,
vtkImageAccumulate, vtkImageGradientMagnitude); VTK is an un-vendored dependency
(find_package(VTK), CMakeLists.txt:24, no version pinned; the vtkSmartPointer/SetInput API dates it to
VTK 5.x) and is not installed here, and the reference ships no test data.  The filter semantics below
restate VTK 5's documented behaviour; VolumeReader::Rescale is the reference's own code and is
restated literally.
"""
import os
import zlib

import numpy as np

MET = {"MET_UCHAR": np.uint8, "MET_CHAR": np.int8, "MET_USHORT": np.uint16, "MET_SHORT": np.int16,
       "MET_UINT": np.uint32, "MET_INT": np.int32, "MET_FLOAT": np.float32, "MET_DOUBLE": np.float64}
MET_INDEX = {k: i for i, k in enumerate(MET)}


def write_metaimage(path, data, spacing=(1.0, 1.0, 1.0), element_type="MET_SHORT", msb=False, compressed=False, local=None,
                    header_size=None, extra_lines=()):
    """Writes `data` (z, y, x) as .mha (LOCAL data) or .mhd + .raw / .zraw.  Returns the header path."""
    data = np.ascontiguousarray(data, MET[element_type])
    if local is None:
        local = str(path).endswith(".mha")
    raw = data.astype(data.dtype.newbyteorder(">" if msb else "<")).tobytes()
    payload = zlib.compress(raw, 6) if compressed else raw
    lines = ["ObjectType = Image", "NDims = 3", "BinaryData = True", f"BinaryDataByteOrderMSB = {'True' if msb else 'False'}",
             f"CompressedData = {'True' if compressed else 'False'}"]
    if compressed:
        lines.append(f"CompressedDataSize = {len(payload)}")
    lines += ["TransformMatrix = 1 0 0 0 1 0 0 0 1", "Offset = 0 0 0", "CenterOfRotation = 0 0 0", "AnatomicalOrientation = RAI",
              f"ElementSpacing = {spacing[0]} {spacing[1]} {spacing[2]}", f"DimSize = {data.shape[2]} {data.shape[1]} {data.shape[0]}",
              f"ElementType = {element_type}"]
    lines += list(extra_lines)
    prefix = b""
    if header_size is not None:
        lines.append(f"HeaderSize = {header_size}")
        prefix = b"\x5a" * max(header_size, 0) if header_size >= 0 else b"\x5a" * 37
    if local:
        with open(path, "wb") as f:
            f.write(("\n".join(lines) + "\nElementDataFile = LOCAL\n").encode())
            f.write(payload)
    else:
        data_name = os.path.splitext(os.path.basename(str(path)))[0] + (".zraw" if compressed else ".raw")
        with open(path, "wb") as f:
            f.write(("\n".join(lines) + f"\nElementDataFile = {data_name}\n").encode())
        with open(os.path.join(os.path.dirname(str(path)), data_name), "wb") as f:
            f.write(prefix + payload)
    return path


def parse_header(path):
    """Key/value pairs up to and including ElementDataFile, and the offset of LOCAL data."""
    out = {}
    with open(path, "rb") as f:
        while True:
            line = f.readline()
            if not line:
                break
            text = line.decode("latin-1")
            if "=" not in text:
                continue
            k, v = text.split("=", 1)
            out[k.strip()] = v.strip()
            if k.strip() == "ElementDataFile":
                out["_data_offset"] = f.tell()
                break
    return out


def cast_to_short(a):
    """vtkImageCast, output short, ClampOverflow off: static_cast<short>(v)."""
    a = np.asarray(a)
    if a.dtype.kind == "f":
        with np.errstate(invalid="ignore"):
            i = np.trunc(a).clip(-2 ** 31, 2 ** 31 - 1).astype(np.int64)  # float -> int truncates toward zero
        return (i & 0xFFFF).astype(np.uint16).view(np.int16)
    return (a.astype(np.int64) & 0xFFFF).astype(np.uint16).view(np.int16)


def preprocess(data, spacing=(1.0, 1.0, 1.0)):
    """data: (z, y, x) array of any MetaImage element type.  Returns dict(u16, data_min, data_max,
    histogram, histogram_total, max_gradient_magnitude) as VolumeReader::Read leaves them."""
    s = cast_to_short(data)
    dmin, dmax = int(s.min()), int(s.max())
    # VolumeReader::Rescale<short, unsigned short> (VolumeReader.cpp:124-136): fp32 expression, truncation
    extent = np.float32(dmax) - np.float32(dmin)
    if extent > 0:
        v = (s.astype(np.float32) - np.float32(dmin)) / extent * np.float32(65535)
        u16 = v.astype(np.float32).astype(np.int64).astype(np.uint16)
    else:
        u16 = np.zeros(s.shape, np.uint16)
    # vtkImageAccumulate: extent [0, max-min-1], origin min, spacing 1, IgnoreZero on (VolumeReader.cpp:57-63)
    bins = dmax - dmin
    sv = s[s != 0].astype(np.int64) - dmin
    sv = sv[(sv >= 0) & (sv < bins)]
    hist = np.bincount(sv, minlength=max(bins, 0)).astype(np.uint32)[: max(bins, 0)]
    # vtkImageGradientMagnitude (3-D, HandleBoundaries on), output cast to short; VolumeReader.cpp:70-76 takes its max
    d = s.astype(np.float64)
    g2 = np.zeros(d.shape, np.float64)
    for axis, sp in ((2, spacing[0]), (1, spacing[1]), (0, spacing[2])):
        lo = np.take(d, np.maximum(np.arange(d.shape[axis]) - 1, 0), axis=axis)
        hi = np.take(d, np.minimum(np.arange(d.shape[axis]) + 1, d.shape[axis] - 1), axis=axis)
        g2 += ((lo - hi) * (0.5 / float(sp))) ** 2
    mag = (np.trunc(np.sqrt(g2)).astype(np.int64) & 0xFFFF).astype(np.uint16).view(np.int16)
    return dict(u16=u16, data_min=dmin, data_max=dmax, histogram=hist, histogram_total=int(hist.sum()), max_gradient_magnitude=int(mag.max()))
