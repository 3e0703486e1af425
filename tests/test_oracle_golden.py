"""The CPU oracle against the golden fixtures (CPU only, no GPU needed).

The reference ships no golden vectors; tests/golden/*.npz are outputs of the
reference ITSELF -- its own unmodified kernels (oracle/_ref) and the texture hardware they sample
through -- captured on a B200 by tests/golden/make_golden.py.  This is what pins oracle/svr_oracle.cpp.

Tolerances: the reference is built -use_fast_math, the oracle uses IEEE libm, so floats agree to
fast-math rounding: ray casting within 1e-4 per channel on EVERY pixel (north_star's bound), path
tracing within 1e-4 on >= 99.9% of the pixels (a last-bit difference can flip one accept/reject of
the delta tracker and send that one pixel down another path) and within 0.01% on the image mean.
The texture filter is integer arithmetic for u8/u16 reads: bit-exact.
"""
import ctypes as C
import os

import numpy as np
import pytest

from sunvolumerender_b200 import _lib as L
from sunvolumerender_b200 import scene as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def filt():
    return np.load(os.path.join(GOLD, "texture_filter.npz"))


@pytest.fixture(scope="module")
def kern():
    return np.load(os.path.join(GOLD, "reference_kernels.npz"))


def _tex3d(B, vox, fmt, uvw):
    n = vox.shape[0]
    o = B.CpuOracle(vox, fmt, (n, n, n), S.host_volume_struct((n, n, n)), S.tf_table("default"), S.default_camera((n,) * 3, 16, 16))
    lib = B.cpu()
    return np.array([lib.svr_oracle_tex3d(C.byref(o.scene), float(a), float(b), float(c)) for a, b, c in uvw], np.float32)


@pytest.mark.parametrize("name,fmt", [("u16", L.VOXEL_U16), ("u8", L.VOXEL_U8)])
def test_integer_texture_filter_is_bit_exact(oracle_cpu, filt, name, fmt):
    got = filt[f"{name}_got"]
    mine = _tex3d(oracle_cpu, filt[f"{name}_vox"], fmt, filt[f"{name}_uvw"])
    assert np.array_equal(mine.view(np.uint32), got.view(np.uint32))
    assert (got > 0).mean() > 0.6 and (got == 0).any()  # interior and border samples both present


def test_trilinear_weights_including_rounding_ties(oracle_cpu, filt):
    """One-hot f32 volume sampled on the 1/256 lattice: the fetch IS the weight of texel (1,1,1);
    multiples of 1/256, and every rounding tie of the two-stage product lands on the hardware's side."""
    got = filt["onehot_got"]
    assert np.array_equal(got * 256, np.round(got * 256))
    mine = _tex3d(oracle_cpu, filt["onehot_vox"], L.VOXEL_F32, filt["onehot_uvw"])
    assert np.array_equal(mine, got)
    assert len(np.unique(got)) > 100


def test_f16_texture_filter(oracle_cpu, filt):
    got = filt["f16_got"]
    mine = _tex3d(oracle_cpu, filt["f16_vox"], L.VOXEL_F16, filt["f16_uvw"])
    assert (mine == got).mean() > 0.985           # results are fp16 values; the rest differ by one fp16 ulp
    assert np.abs(mine - got).max() <= 2.0 ** -11


def test_transfer_function_filter(oracle_cpu, filt):
    tab, x, got = filt["tf_tab"], filt["tf_x"], filt["tf_got"]
    o = oracle_cpu.CpuOracle(np.zeros((2, 2, 2), np.uint8), L.VOXEL_U8, (2, 2, 2), S.host_volume_struct((2, 2, 2)), tab, S.default_camera((2,) * 3, 16, 16))
    lib = oracle_cpu.cpu()
    out = np.zeros(4, np.float32)
    mine = np.zeros_like(got)
    for i, xv in enumerate(x):
        lib.svr_oracle_tf(C.byref(o.scene), float(xv), out.ctypes.data)
        mine[i] = out
    assert np.abs(mine - got).max() <= 6e-8
    assert (mine == got).mean() > 0.99


def _scene(B, g, name):
    n, fmt, W, H, depth, frames, nl = [int(v) for v in g[f"{name}_meta"]]
    vol = L.Volume.from_buffer_copy(bytes(g[f"{name}_volume"]))
    cam = L.Camera.from_buffer_copy(bytes(g[f"{name}_camera"]))
    lb = bytes(g[f"{name}_lights"])
    lights = [L.AreaLight.from_buffer_copy(lb[i * 44:(i + 1) * 44]) for i in range(nl)]
    o = B.CpuOracle(g[f"{name}_vox"], fmt, (n, n, n), vol, g[f"{name}_tf"], cam, lights)
    return o, depth, frames


SCENES = ["ct_u16", "sphere_u8", "ct_u8_thin"]


@pytest.mark.parametrize("name", SCENES)
def test_raycast_matches_reference_kernel(oracle_cpu, kern, name):
    o, _, _ = _scene(oracle_cpu, kern, name)
    rgba, u8, _ = o.raycast(float(kern[f"{name}_step"]))
    ref = kern[f"{name}_rc_f32x255"] / np.float32(255.0)   # float twin of kernel_raycasting (oracle/glm_shim)
    assert ref[..., 3].max() > 0.4 and (ref[..., 3] == 0).any()
    assert np.abs(rgba - ref).max() <= 1e-4
    assert np.abs(u8.astype(int) - kern[f"{name}_rc_u8"].astype(int)).max() <= 1


@pytest.mark.parametrize("name", SCENES)
def test_pathtrace_matches_reference_kernel(oracle_cpu, kern, name):
    o, depth, frames = _scene(oracle_cpu, kern, name)
    hdrs = kern[f"{name}_pt_hdr"]
    hdr = None
    for f in range(frames):
        hdr, _ = o.pathtrace(depth, f, 1, hdr=hdr)     # running mean continued frame by frame (pathtracer.cu:81-84)
        ref = hdrs[f]
        assert (ref.max(axis=2) > 0).mean() > 0.01
        d = np.abs(hdr - ref).max(axis=2)
        assert (d <= 1e-4).mean() >= 0.999, (f, d.max())
        assert abs(float(hdr.mean()) - float(ref.mean())) <= 1e-4 * float(ref.mean())
    got = o.tonemap(hdrs[-1])                              # hdr_to_ldr, pathtracer.cu:282-290
    assert np.abs(got.astype(int) - kern[f"{name}_pt_u8"].astype(int)).max() <= 1
