"""Host-side scene construction -- the part of gui/canvas.cpp, gui/mainwindow.cpp and
core/VolumeReader.cpp that decides WHAT the render entry points are called with.  Pure numpy /
ctypes: no GPU needed, so the `-m "not gpu"` tests cover it.

Synthetic configurations C1..C5 are the ones SURVEY.md section 8(d) and BASELINE.json name.
"""
import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L

TF_TABLE_SIZE = 1024  # gui/transferfunction.h:29

# colour nodes of the default transfer function, gui/mainwindow.cpp:57-62
_COLOR_NODES = [
    (0.0, (69.0 / 255, 199.0 / 255, 186.0 / 255)),
    (0.2, (172.0 / 255, 3.0 / 255, 57.0 / 255)),
    (0.4, (169.0 / 255, 83.0 / 255, 58.0 / 255)),
    (0.6, (43.0 / 255, 32.0 / 255, 161.0 / 255)),
    (0.8, (247.0 / 255, 158.0 / 255, 97.0 / 255)),
    (1.0, (183.0 / 255, 7.0 / 255, 140.0 / 255)),
]


def make_camera(pos, u, v, w, fovx=45.0, apeture=0.0, focal_length=1.0, exposure=1.0, image_w=640, image_h=480):
    """cudaCamera::Setup(pos, u, v, w, ...) -- core/cuda_camera.h:34-47 (field `apeture` as spelled there)."""
    cam = L.Camera()
    cam.pos = L.Vec3(*pos)
    cam.u = L.Vec3(*u)
    cam.v = L.Vec3(*v)
    cam.w = L.Vec3(*w)
    cam.imageW = int(image_w)
    cam.imageH = int(image_h)
    cam.aspectRatio = np.float32(image_w) / np.float32(image_h)
    # tanf(fovx * 0.5f * M_PI / 180.f): the product is evaluated in double, then tanf of its float value
    cam.tanFovxOverTwo = float(np.tan(np.float32(np.float32(fovx) * np.float32(0.5) * math.pi / 180.0), dtype=np.float32))
    cam.exposure = exposure
    cam.focalLength = focal_length
    cam.apeture = apeture
    return cam


def look_at_camera(pos, target, up, **kw):
    """cudaCamera::Setup(pos, target, up, ...) -- core/cuda_camera.h:49-62."""
    pos = np.asarray(pos, np.float32)
    target = np.asarray(target, np.float32)
    up = np.asarray(up, np.float32)
    w = pos - target
    w = w / np.sqrt(np.dot(w, w))
    u = np.cross(up, w)
    v = np.cross(w, u)
    return make_camera(pos, u, v, w, **kw)


def eye_distance(extent, fov=45.0):
    """Canvas::ZoomToExtent, gui/canvas.cpp:191-197."""
    return 1.5 * max(extent) / (2.0 * math.tan(math.radians(fov * 0.5)))


def default_camera(extent, image_w, image_h, fov=45.0, exposure=1.0, apeture=0.0, focal_length=1.0):
    """Camera as Canvas frames a freshly loaded volume: on +z looking at the origin (gui/canvas.cpp:31-38, 179-188)."""
    d = eye_distance(extent, fov)
    return make_camera((0.0, 0.0, d), (1, 0, 0), (0, 1, 0), (0, 0, 1), fov, apeture, focal_length, exposure, image_w, image_h)


def default_area_light(extent, n_ref=128.0, intensity=500.0):
    """MainWindow::onAddLight, gui/mainwindow.cpp:229-240: disk of radius 10 facing -y, placed
    1.5 * boundingSphereRadius + 1 above the volume.  The radius scales with the volume so that the
    light subtends the same solid angle at every config size."""
    R = 0.5 * math.sqrt(sum(e * e for e in extent))
    light = L.AreaLight()
    light.disk.radius = 10.0 * (max(extent) / n_ref)
    light.disk.center = L.Vec3(0.0, 1.5 * R + 1.0, 0.0)
    light.disk.normal = L.Vec3(0.0, -1.0, 0.0)
    light.color = L.Vec3(1.0, 1.0, 1.0)
    light.intensity = intensity
    return light


def constant_env_light(radiance=(0.5, 0.5, 0.5), intensity=1.0):
    """Lights::SetEnvionmentLight(radiance), core/lights/lights.cpp:77-80; default radiance gui/canvas.cpp:11-12."""
    env = L.EnvLight()
    env.tex = 0
    env.defaultRadiance = L.Vec3(*radiance)
    env.intensity = intensity
    env.offset = L.Vec2(0.0, 0.0)
    return env


def _color_table(n=TF_TABLE_SIZE):
    x = np.linspace(0.0, 1.0, n)  # vtkColorTransferFunction::GetTable(0, 1, n, ...) samples both ends
    xs = np.array([c[0] for c in _COLOR_NODES])
    out = np.empty((n, 3), np.float32)
    for ch in range(3):
        out[:, ch] = np.interp(x, xs, np.array([c[1][ch] for c in _COLOR_NODES]))
    return out


def tf_table(kind="default", n=TF_TABLE_SIZE):
    """n x (r, g, b, opacity) float32 table in the layout TransferFunction uploads
    (gui/transferfunction.cpp:17-29).  Opacity ramps:
      default: 0 at 0 rising linearly to 0.5 at 0.1, then flat (stand-in for mainwindow.cpp:51-55)
      thin   : 0.02 * clamp((i - 0.1) / 0.9, 0, 1)
      cloud  : 0.5 * i, white
    """
    x = np.linspace(0.0, 1.0, n)
    t = np.empty((n, 4), np.float32)
    t[:, :3] = _color_table(n)
    if kind == "default":
        t[:, 3] = np.clip(x / 0.1, 0.0, 1.0) * 0.5
    elif kind == "thin":
        t[:, 3] = 0.02 * np.clip((x - 0.1) / 0.9, 0.0, 1.0)
    elif kind == "cloud":
        t[:, :3] = 1.0
        t[:, 3] = 0.5 * x
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(t)


def build_tf_table(opacity_nodes, color_nodes, n=TF_TABLE_SIZE):
    """TransferFunction's composite table from node lists (gui/transferfunction.cpp:17-29) through the C
    ABI (svr_tf_build_table): opacity nodes (x, y[, midpoint, sharpness]), colour nodes (x, r, g, b[,
    midpoint, sharpness]).  Returns (table n x 4 float32, maxOpacity)."""
    import ctypes as C

    lib = L.load()
    on = (L.TfOpacityNode * max(1, len(opacity_nodes)))(*[L.TfOpacityNode(*(tuple(p) + (0.5, 0.0))[:4]) for p in opacity_nodes])
    cn = (L.TfColorNode * max(1, len(color_nodes)))(*[L.TfColorNode(*(tuple(p) + (0.5, 0.0))[:6]) for p in color_nodes])
    table = np.zeros((n, 4), np.float32)
    mx = C.c_float()
    L.check(lib.svr_tf_build_table(on, len(opacity_nodes), cn, len(color_nodes), C.c_void_p(table.ctypes.data), n, C.byref(mx)), "svr_tf_build_table")
    return table, float(mx.value)


def default_tf_nodes():
    """The start-up transfer function of the application (gui/mainwindow.cpp:46-62) as node lists."""
    import ctypes as C

    lib = L.load()
    on, cn = (L.TfOpacityNode * 16)(), (L.TfColorNode * 16)()
    no, nc = C.c_uint32(16), C.c_uint32(16)
    L.check(lib.svr_tf_default_nodes(on, C.byref(no), cn, C.byref(nc)), "svr_tf_default_nodes")
    return ([(o.x, o.y, o.midpoint, o.sharpness) for o in on[: no.value]],
            [(c.x, c.r, c.g, c.b, c.midpoint, c.sharpness) for c in cn[: nc.value]])


def raycast_step_size(spacing=(1.0, 1.0, 1.0)):
    """VolumeReader::GetElementBoundingSphereRadius, core/VolumeReader.cpp:198-201 (passed at gui/canvas.cpp:92)."""
    return 0.5 * math.sqrt(sum(s * s for s in spacing))


def sphere_volume(n, fmt=L.VOXEL_U8):
    """C1 generator on the host (numpy) for CPU-only tests: rho = clamp(1 - r/(0.45 n), 0, 1)."""
    c = 0.5 * n
    ax = (np.arange(n, dtype=np.float32) + 0.5 - c).astype(np.float32)
    z, y, x = np.meshgrid(ax, ax, ax, indexing="ij")
    r = np.sqrt(x * x + y * y + z * z, dtype=np.float32)
    d = np.clip(1.0 - r / np.float32(0.45 * n), 0.0, 1.0).astype(np.float32)
    return encode_voxels(d, fmt)


def encode_voxels(d, fmt):
    if fmt == L.VOXEL_U8:
        return (d * 255.0 + 0.5).astype(np.uint8)
    if fmt == L.VOXEL_U16:
        return (d * 65535.0 + 0.5).astype(np.uint16)
    if fmt == L.VOXEL_F16:
        return d.astype(np.float16)
    return d.astype(np.float32)


VOXEL_DTYPES = {L.VOXEL_U8: np.uint8, L.VOXEL_U16: np.uint16, L.VOXEL_F16: np.float16, L.VOXEL_F32: np.float32}


def host_volume_struct(n_xyz, spacing=(1.0, 1.0, 1.0), max_grad_mag=1.0):
    """cudaVolume as VolumeReader::CreateDeviceVolume + Canvas::LoadVolume leave it, without a
    texture (core/VolumeReader.cpp:174-185; gui/canvas.cpp:31-32).  Used by the CPU oracle."""
    v = L.Volume()
    size = [np.float32(n) * np.float32(s) for n, s in zip(n_xyz, spacing)]
    vmax = [s - s * np.float32(0.5) for s in size]
    v.bbox.vmin = L.Vec3(*[-x for x in vmax])
    v.bbox.vmax = L.Vec3(*vmax)
    v.bbox.invSize = L.Vec3(*[np.float32(1.0) / (x + x) for x in vmax])
    v.densityScale = 1.0
    v.invMaxMagnitude = 1.0 / max_grad_mag
    v.gradientFactor = 0.5
    v.spacing = L.Vec3(*spacing)
    v.invSpacing = L.Vec3(*[1.0 / s for s in spacing])
    v.x_clip = L.Vec2(-1.0, 1.0)
    v.y_clip = L.Vec2(-1.0, 1.0)
    v.z_clip = L.Vec2(-1.0, 1.0)
    return v


@dataclass
class Config:
    """One BASELINE.json configuration."""
    name: str
    n: int                 # volume edge
    fmt: int               # voxel format
    gen: int               # svr_volume_kind
    width: int
    height: int
    tf: str
    trace_depth: int = 1
    spp: int = 16
    env: bool = False
    gen_seed: int = 1234
    notes: str = ""
    extent: tuple = field(init=False)

    def __post_init__(self):
        self.extent = (float(self.n),) * 3

    @property
    def voxel_bytes(self):
        return L.VOXEL_BYTES[self.fmt]


CONFIGS = {
    "C1": Config("C1", 128, L.VOXEL_U8, L.GEN_SPHERE, 512, 512, "default", 1, 16),
    "C2": Config("C2", 256, L.VOXEL_U8, L.GEN_CT, 1024, 1024, "thin", 1, 1),
    "C3": Config("C3", 512, L.VOXEL_U16, L.GEN_CT, 1920, 1080, "default", 1, 256, env=True),
    "C4": Config("C4", 1024, L.VOXEL_F16, L.GEN_CLOUD, 1920, 1080, "cloud", 32, 512, gen_seed=42),
    "C5": Config("C5", 2048, L.VOXEL_U16, L.GEN_CT, 3840, 2160, "default", 1, 1024, env=True),
}


def split_samples(total, world_size):
    """Sample-index ranges per rank: rank r renders [first, first+count)."""
    base, rem = divmod(total, world_size)
    out, first = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < rem else 0)
        out.append((first, cnt))
        first += cnt
    return out


def split_rows(height, world_size, align=4):
    """Row ranges per rank for the ray caster's image-tile split, aligned to the warp tile height."""
    rows = (height + align - 1) // align
    base, rem = divmod(rows, world_size)
    out, y = [], 0
    for r in range(world_size):
        cnt = (base + (1 if r < rem else 0)) * align
        y1 = min(height, y + cnt)
        out.append((y, y1))
        y = y1
    return out
