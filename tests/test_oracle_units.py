"""CPU tests of the oracle's building blocks (software texture units, RNG, host restatement).
Known answers come from the definitions the reference relies on: cuRAND's XORWOW
(curand_kernel.h: _curand_init_scratch, curand()), wangHash (pathtracer.cu:70-79), CUDA linear
filtering (normalised coordinates, x*N-0.5, border / clamp addressing)."""
import ctypes as C

import numpy as np
import pytest

from sunvolumerender_b200 import _lib as L
from sunvolumerender_b200 import scene as S


def _wang(a):
    a &= 0xFFFFFFFF
    a = (a ^ 61) ^ (a >> 16)
    a = (a + (a << 3)) & 0xFFFFFFFF
    a = a ^ (a >> 4)
    a = (a * 0x27D4EB2D) & 0xFFFFFFFF
    a = a ^ (a >> 15)
    return a


def _xorwow_py(seed, n):
    m = 0xFFFFFFFF
    s0 = (seed & m) ^ 0xAAD26B49
    s1 = ((seed >> 32) & m) ^ 0xF7DCEFDD
    t0 = (1099087573 * s0) & m
    t1 = (2591861531 * s1) & m
    d = (6615241 + t1 + t0) & m
    v = [(123456789 + t0) & m, 362436069 ^ t0, (521288629 + t1) & m, 88675123 ^ t1, (5783321 + t0) & m]
    out = []
    for _ in range(n):
        t = v[0] ^ (v[0] >> 2)
        v[0], v[1], v[2], v[3] = v[1], v[2], v[3], v[4]
        v[4] = ((v[4] ^ ((v[4] << 4) & m)) ^ (t ^ ((t << 1) & m))) & m
        d = (d + 362437) & m
        out.append((v[4] + d) & m)
    return out


def test_wang_hash(oracle_cpu):
    lib = oracle_cpu.cpu()
    for a in [0, 1, 2, 61, 255, 1 << 16, 0xDEADBEEF, 0xFFFFFFFF]:
        assert lib.svr_oracle_wang_hash(a) == _wang(a)
    # frameNo 0 and 1 seed different streams
    assert lib.svr_oracle_wang_hash(0) != lib.svr_oracle_wang_hash(1)


def test_xorwow_matches_curand_definition(oracle_cpu):
    lib = oracle_cpu.cpu()
    for seed in [0, 1, 12345, _wang(0) + 17, 0xFFFFFFFF]:
        out = np.zeros(64, np.float32)
        lib.svr_oracle_xorwow_uniforms(seed, 64, out.ctypes.data)
        ref = np.array(_xorwow_py(seed, 64), dtype=np.uint32)
        # curand_uniform: x * 2^-32 + 2^-33 in fp32 (curand_uniform.h:69-72)
        expect = ref.astype(np.float32) * np.float32(2.3283064e-10) + np.float32(2.3283064e-10 / 2.0)
        assert np.array_equal(out, expect)
        assert out.min() > 0.0 and out.max() <= 1.0


def _scene(oracle_cpu, vox, fmt, n, tf=None, filter_mode=0, W=16, H=16):
    vol = S.host_volume_struct((n, n, n))
    cam = S.default_camera((n, n, n), W, H)
    return oracle_cpu.CpuOracle(vox, fmt, (n, n, n), vol, tf if tf is not None else S.tf_table("default"), cam, [S.default_area_light((n, n, n))], filter_mode=filter_mode)


def test_tex3d_texel_centres_and_border(oracle_cpu):
    n = 8
    rng = np.random.default_rng(1)
    vox = rng.integers(0, 65536, size=(n, n, n), dtype=np.uint16)
    o = _scene(oracle_cpu, vox, L.VOXEL_U16, n)
    lib = oracle_cpu.cpu()
    for (i, j, k) in [(0, 0, 0), (3, 4, 5), (7, 7, 7), (1, 6, 2)]:
        v = lib.svr_oracle_tex3d(C.byref(o.scene), (i + 0.5) / n, (j + 0.5) / n, (k + 0.5) / n)
        assert v == pytest.approx(float(vox[k, j, i]) / 65535.0, abs=1e-7)
    # border addressing: far outside is exactly 0, half a texel outside blends with 0
    assert lib.svr_oracle_tex3d(C.byref(o.scene), -0.5, 0.5, 0.5) == 0.0
    assert lib.svr_oracle_tex3d(C.byref(o.scene), 1.5, 0.5, 0.5) == 0.0
    edge = lib.svr_oracle_tex3d(C.byref(o.scene), 0.0, 0.5 / n, 0.5 / n)
    # integer reads carry 16 bits: the blend is rounded half up to a 16-bit integer (svr_oracle.cpp)
    assert edge == np.float32(np.floor(0.5 * float(vox[0, 0, 0]) + 0.5)) / np.float32(65535.0)


def test_tex3d_formats(oracle_cpu):
    n = 4
    d = np.linspace(0, 1, n * n * n, dtype=np.float32).reshape(n, n, n)
    lib = oracle_cpu.cpu()
    for fmt in (L.VOXEL_U8, L.VOXEL_U16, L.VOXEL_F16, L.VOXEL_F32):
        vox = S.encode_voxels(d, fmt)
        o = _scene(oracle_cpu, vox, fmt, n)
        v = lib.svr_oracle_tex3d(C.byref(o.scene), 2.5 / n, 1.5 / n, 3.5 / n)
        tol = {L.VOXEL_U8: 1 / 255, L.VOXEL_U16: 1 / 65535, L.VOXEL_F16: 1e-3, L.VOXEL_F32: 1e-7}[fmt]
        assert abs(v - d[3, 1, 2]) <= tol


def test_tex3d_weights_are_quantised_to_8_bits(oracle_cpu):
    n = 4
    vox = np.zeros((n, n, n), np.float32)
    vox[:, :, :] = np.arange(n, dtype=np.float32)[None, None, :]
    lib = oracle_cpu.cpu()
    o0 = _scene(oracle_cpu, vox, L.VOXEL_F32, n, filter_mode=0)
    o2 = _scene(oracle_cpu, vox, L.VOXEL_F32, n, filter_mode=2)
    us = (np.arange(1000) / 1000.0 * 2.0 + 1.0) / n  # texel coordinate 0.5 .. 2.5
    q = np.array([lib.svr_oracle_tex3d(C.byref(o0.scene), u, 0.5 / n, 0.5 / n) for u in us])
    f = np.array([lib.svr_oracle_tex3d(C.byref(o2.scene), u, 0.5 / n, 0.5 / n) for u in us])
    assert np.allclose(q * 256, np.round(q * 256), atol=1e-4)  # steps of 1/256
    assert np.abs(q - f).max() <= 0.5 / 256 + 1e-6


def test_tf_lookup_clamps_and_interpolates(oracle_cpu):
    tf = S.tf_table("default")
    o = _scene(oracle_cpu, S.sphere_volume(8), L.VOXEL_U8, 8, tf=tf, filter_mode=2)
    lib = oracle_cpu.cpu()
    out = np.zeros(4, np.float32)
    lib.svr_oracle_tf(C.byref(o.scene), -1.0, out.ctypes.data)
    assert np.allclose(out, tf[0])
    lib.svr_oracle_tf(C.byref(o.scene), 2.0, out.ctypes.data)
    assert np.allclose(out, tf[-1])
    lib.svr_oracle_tf(C.byref(o.scene), (100 + 0.5) / 1024, out.ctypes.data)
    assert np.allclose(out, tf[100], atol=1e-7)
    lib.svr_oracle_tf(C.byref(o.scene), (100 + 1.0) / 1024, out.ctypes.data)
    assert np.allclose(out, 0.5 * (tf[100] + tf[101]), atol=1e-6)
    assert o.scene.tf.maxOpacity == pytest.approx(0.5)


def test_raycast_empty_and_opaque(oracle_cpu):
    n = 16
    empty = np.zeros((n, n, n), np.uint8)
    o = _scene(oracle_cpu, empty, L.VOXEL_U8, n)
    rgba, u8, cnt = o.raycast(S.raycast_step_size())
    assert rgba.max() == 0.0 and u8.max() == 0
    assert cnt[oracle_cpu.CNT["paths"]] == 16 * 16
    full = np.full((n, n, n), 255, np.uint8)
    o = _scene(oracle_cpu, full, L.VOXEL_U8, n)
    rgba, u8, cnt = o.raycast(S.raycast_step_size())
    centre = rgba[8, 8]
    assert centre[3] > 0.95  # early termination threshold, raycasting.cu:55
    assert centre[3] <= 1.0 and (u8[8, 8, 3] == int(255 * centre[3]))
    # rays that miss the box stay transparent black
    assert rgba[0, 0, 3] == 0.0


def test_raycast_rows_partition(oracle_cpu):
    n = 16
    o = _scene(oracle_cpu, S.sphere_volume(n), L.VOXEL_U8, n, W=24, H=20)
    full, _, _ = o.raycast(S.raycast_step_size())
    a, _, _ = o.raycast(S.raycast_step_size(), rows=(0, 8))
    b, _, _ = o.raycast(S.raycast_step_size(), rows=(8, 20))
    assert np.array_equal(full[:8], a[:8]) and np.array_equal(full[8:], b[8:])


def test_pathtrace_running_mean_and_determinism(oracle_cpu):
    n = 16
    o = _scene(oracle_cpu, S.sphere_volume(n), L.VOXEL_U8, n)
    h1, c1 = o.pathtrace(2, 0, 4)
    h2, c2 = o.pathtrace(2, 0, 4)
    assert np.array_equal(h1, h2) and np.array_equal(c1, c2)
    # frames 0..1 then 2..3 continue the same running mean (pathtracer.cu:81-84)
    ha, _ = o.pathtrace(2, 0, 2)
    hb, _ = o.pathtrace(2, 2, 2, hdr=ha.copy())
    assert np.allclose(hb, h1, atol=1e-6)
    # frameNo 0 clears whatever was in the buffer (pathtracer.cu:297-300)
    junk = np.full_like(h1, 7.0)
    hc, _ = o.pathtrace(2, 0, 4, hdr=junk)
    assert np.array_equal(hc, h1)
    assert np.isfinite(h1).all() and h1.min() >= 0.0


def test_pathtrace_no_lights_is_black_and_env_adds_light(oracle_cpu):
    n = 16
    vol = S.host_volume_struct((n, n, n))
    cam = S.default_camera((n, n, n), 16, 16)
    vox = S.sphere_volume(n)
    dark = oracle_cpu.CpuOracle(vox, L.VOXEL_U8, (n, n, n), vol, S.tf_table("default"), cam, [])
    h, _ = dark.pathtrace(2, 0, 2)
    assert h.max() == 0.0  # env contribution is commented out in the reference (pathtracer.cu:233)
    env = oracle_cpu.CpuOracle(vox, L.VOXEL_U8, (n, n, n), vol, S.tf_table("default"), cam, [], env=S.constant_env_light(), env_enabled=True)
    h, _ = env.pathtrace(2, 0, 2)
    assert h[0, 0, 0] == pytest.approx(0.5)  # corner ray misses the box and sees the constant sky


def test_tonemap_matches_formula(oracle_cpu):
    o = _scene(oracle_cpu, S.sphere_volume(8), L.VOXEL_U8, 8)
    hdr = np.array([[[0.0, 0.01, 0.05], [0.1, 0.5, 10.0]]], np.float32)
    out = o.tonemap(hdr)
    l = 1.0 - np.exp(-16.0 * hdr.astype(np.float64))
    expect = (np.power(l, 2.2) * 255).astype(np.uint8)  # tonemapping.h:13-27: exponent 1/gamma = 2.2
    assert np.abs(out[..., :3].astype(int) - expect.astype(int)).max() <= 1
    assert (out[..., 3] == 255).all()
