"""Generates the golden fixtures under tests/golden/ -- run on a GPU box:

    python tests/golden/make_golden.py --out gpurun_out/golden      (then copy *.npz into tests/golden/)

The reference ships no tests, golden vectors or fixtures, so the fixtures are
outputs of the reference ITSELF run on a B200:

  reference_kernels.npz  the reference's own unmodified kernel_raycasting / kernel_pathtracer /
                         hdr_to_ldr (oracle/_ref, compiled from /root/reference by oracle/Makefile) on
                         small scenes: voxels, scene structs, float and u8 images.
  texture_filter.npz     raw tex3D / tex1D fetches of the texture hardware those kernels sample
                         through (descriptors of VolumeReader.cpp:138-172, transferfunction.cpp:30-44),
                         on random and one-hot volumes -- the data that fixed the software sampler of
                         oracle/svr_oracle.cpp.

tests/test_oracle_golden.py (CPU, `-m "not gpu"`) holds the CPU oracle to both files; the GPU parity
tests then hold the product kernels to the oracle and to oracle/_ref directly.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as B  # noqa: E402
from sunvolumerender_b200 import _lib as L  # noqa: E402
from sunvolumerender_b200 import scene as S  # noqa: E402
from sunvolumerender_b200.render import Renderer  # noqa: E402


def struct_bytes(s):
    return np.frombuffer(bytes(s), dtype=np.uint8).copy()


def fetch3(r, uvw):
    m = uvw.shape[0]
    d_uvw = torch.from_numpy(np.ascontiguousarray(uvw, np.float32)).cuda()
    d_out = torch.zeros(m, dtype=torch.float32, device="cuda")
    L.check(r.lib.svr_debug_sample_volume(C.byref(r.volume), C.c_void_p(d_uvw.data_ptr()), m, C.c_void_p(d_out.data_ptr())))
    return d_out.cpu().numpy()


def texture_filter(r):
    out = {}
    rng = np.random.default_rng(2026)
    n, m = 16, 6000
    for name, dt, fmt in (("u16", np.uint16, L.VOXEL_U16), ("u8", np.uint8, L.VOXEL_U8), ("f16", np.float16, L.VOXEL_F16)):
        if dt == np.float16:
            vox = rng.uniform(0, 1, (n, n, n)).astype(np.float16)
        else:
            vox = rng.integers(0, np.iinfo(dt).max + 1, (n, n, n)).astype(dt)
        r.load_volume(vox, fmt, (n, n, n), max_grad_mag=1.0)
        uvw = rng.uniform(-0.1, 1.1, (m, 3)).astype(np.float32)
        out[f"{name}_vox"], out[f"{name}_uvw"], out[f"{name}_got"] = vox, uvw, fetch3(r, uvw)
    # one-hot f32 volume: the fetch returns the trilinear weight of texel (1,1,1) itself
    vox = np.zeros((4, 4, 4), np.float32)
    vox[1, 1, 1] = 1.0
    r.load_volume(vox, L.VOXEL_F32, (4, 4, 4), max_grad_mag=1.0)
    pts = rng.integers(0, 3 * 256, (m, 3)).astype(np.float32) / np.float32(256)  # exact 1/256 lattice: many rounding ties
    uvw = ((pts + np.float32(0.5)) / np.float32(4)).astype(np.float32)
    out["onehot_vox"], out["onehot_uvw"], out["onehot_got"] = vox, uvw, fetch3(r, uvw)
    tab = rng.uniform(0, 1, (1024, 4)).astype(np.float32)
    r.set_transfer_function(tab)
    x = rng.uniform(-0.05, 1.05, m).astype(np.float32)
    d_x = torch.from_numpy(x).cuda()
    d_o = torch.zeros(m * 4, dtype=torch.float32, device="cuda")
    L.check(r.lib.svr_debug_sample_tf(C.byref(r.tf), C.c_void_p(d_x.data_ptr()), m, C.c_void_p(d_o.data_ptr())))
    out["tf_tab"], out["tf_x"], out["tf_got"] = tab, x, d_o.view(m, 4).cpu().numpy()
    return out


SCENES = {
    # name: (n, fmt, generator, tf, W, H, depth, frames, nlights, raycast tf)
    "ct_u16": (32, L.VOXEL_U16, L.GEN_CT, "default", 64, 64, 3, 3, 1),
    "sphere_u8": (32, L.VOXEL_U8, L.GEN_SPHERE, "default", 64, 64, 1, 2, 2),
    "ct_u8_thin": (32, L.VOXEL_U8, L.GEN_CT, "thin", 64, 64, 2, 1, 1),
}


def reference_kernels(r):
    out = {}
    for name, (n, fmt, gen, tf, W, H, depth, frames, nl) in SCENES.items():
        cfg = S.Config(name, n, fmt, gen, W, H, tf, trace_depth=depth, spp=frames)
        vb = r.generate_volume(gen, fmt, n, 1234)
        r.load_volume(vb, fmt, (n,) * 3)
        r.set_transfer_function(S.tf_table(tf))
        r.set_camera(S.default_camera(cfg.extent, W, H))
        lights = [S.default_area_light(cfg.extent)]
        if nl == 2:
            l2 = S.default_area_light(cfg.extent)
            l2.disk.center = L.Vec3(40.0, 30.0, 25.0)
            nrm = -np.array([40.0, 30.0, 25.0]) / np.linalg.norm([40.0, 30.0, 25.0])
            l2.disk.normal = L.Vec3(*[float(x) for x in nrm])
            l2.color = L.Vec3(1.0, 0.6, 0.3)
            lights.append(l2)
        r.set_area_lights(lights)
        r.set_env_light(S.constant_env_light(), enabled=False)
        vox = vb.cpu().numpy().view(S.VOXEL_DTYPES[fmt]).reshape(n, n, n)
        out[f"{name}_vox"] = vox
        out[f"{name}_meta"] = np.array([n, fmt, W, H, depth, frames, len(lights)], np.int64)
        out[f"{name}_tf"] = S.tf_table(tf)
        out[f"{name}_volume"] = struct_bytes(r.volume)
        out[f"{name}_camera"] = struct_bytes(r.camera)
        out[f"{name}_lights"] = np.concatenate([struct_bytes(l) for l in lights])
        step = S.raycast_step_size()
        out[f"{name}_step"] = np.float32(step)
        ref = B.RefCuda(W, H)
        ref.setup(r.volume, r.tf, r.camera, lights, r.env)
        ref.render_raycasting(step)
        out[f"{name}_rc_u8"] = ref.ldr_image().cpu().numpy().copy()
        twin = B.RefCuda(W, H, f32=True)
        twin.setup(r.volume, r.tf, r.camera, lights, r.env)
        twin.render_raycasting(step)
        out[f"{name}_rc_f32x255"] = twin.ldr_image().cpu().numpy().copy()  # the four products handed to u8vec4
        ref.setup(r.volume, r.tf, r.camera, lights, r.env)
        hdrs = []
        for f in range(frames):
            ref.render_pathtracer(1, depth)
            hdrs.append(ref.hdr_image().cpu().numpy().copy())
        out[f"{name}_pt_hdr"] = np.stack(hdrs)  # running mean after frame 0, 1, ...
        out[f"{name}_pt_u8"] = ref.ldr_image().cpu().numpy().copy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    r = Renderer(0)
    tfi = texture_filter(r)
    np.savez_compressed(os.path.join(a.out, "texture_filter.npz"), **tfi)
    rk = reference_kernels(r)
    np.savez_compressed(os.path.join(a.out, "reference_kernels.npz"), **rk)
    for f in ("texture_filter.npz", "reference_kernels.npz"):
        print(f, os.path.getsize(os.path.join(a.out, f)), "bytes")
    print("device:", torch.cuda.get_device_name(0))


if __name__ == "__main__":
    main()
