"""How does the texture unit quantise linear-filter weights?  Fetches tex3D / tex1D at fine coordinate
steps on known data and dumps (coordinates, results) to gpurun_out/filter_probe.npz, so the model in
oracle/svr_oracle.cpp (filterMode) can be fitted offline.  Scratch tool; its result is committed as
tests/golden/texture_filter.npz by tests/golden/make_golden.py."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S  # noqa: E402
from sunvolumerender_b200.render import Renderer  # noqa: E402

r = Renderer(0)
out = {}


def fetch3(uvw):
    m = uvw.shape[0]
    d_uvw = torch.from_numpy(np.ascontiguousarray(uvw, np.float32)).cuda()
    d_out = torch.zeros(m, dtype=torch.float32, device="cuda")
    L.check(r.lib.svr_debug_sample_volume(C.byref(r.volume), C.c_void_p(d_uvw.data_ptr()), m, C.c_void_p(d_out.data_ptr())))
    return d_out.cpu().numpy()


rng = np.random.default_rng(0)
for n in (4, 64, 512):
    vox = np.zeros((4, 4, n), np.float32)
    vox[:] = np.arange(n, dtype=np.float32)[None, None, :]
    r.load_volume(vox, L.VOXEL_F32, (n, 4, 4), max_grad_mag=1.0)
    m = 1 << 16
    xs = rng.uniform(-1.0, n, m).astype(np.float32)
    u = ((xs + np.float32(0.5)) / np.float32(n)).astype(np.float32)
    uvw = np.stack([u, np.full(m, 0.5 / 4, np.float32), np.full(m, 0.5 / 4, np.float32)], 1)
    out[f"ramp{n}_u"] = u
    out[f"ramp{n}_got"] = fetch3(uvw)
for name, dt, fmt in (("u16", np.uint16, L.VOXEL_U16), ("u8", np.uint8, L.VOXEL_U8), ("f16", np.float16, L.VOXEL_F16)):
    n = 16
    if dt == np.float16:
        vox = rng.uniform(0, 1, (n, n, n)).astype(np.float16)
    else:
        vox = rng.integers(0, np.iinfo(dt).max + 1, (n, n, n)).astype(dt)
    r.load_volume(vox, fmt, (n, n, n), max_grad_mag=1.0)
    m = 1 << 15
    uvw = rng.uniform(-0.1, 1.1, (m, 3)).astype(np.float32)
    out[f"{name}_vox"] = vox
    out[f"{name}_uvw"] = uvw
    out[f"{name}_got"] = fetch3(uvw)
    # x only: y, z at texel centres
    uvw1 = uvw.copy()
    uvw1[:, 1] = (np.floor(uvw1[:, 1] * n).clip(0, n - 1) + 0.5) / n
    uvw1[:, 2] = (np.floor(uvw1[:, 2] * n).clip(0, n - 1) + 0.5) / n
    out[f"{name}_uvw1"] = uvw1
    out[f"{name}_got1"] = fetch3(uvw1)
tab = rng.uniform(0, 1, (1024, 4)).astype(np.float32)
tab[:, 3] = np.arange(1024)
r.set_transfer_function(tab)
m = 1 << 16
x = rng.uniform(-0.05, 1.05, m).astype(np.float32)
d_x = torch.from_numpy(x).cuda()
d_o = torch.zeros(m * 4, dtype=torch.float32, device="cuda")
L.check(r.lib.svr_debug_sample_tf(C.byref(r.tf), C.c_void_p(d_x.data_ptr()), m, C.c_void_p(d_o.data_ptr())))
out["tf_tab"] = tab
out["tf_x"] = x
out["tf_got"] = d_o.view(m, 4).cpu().numpy()
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/filter_probe.npz", **out)
print("wrote gpurun_out/filter_probe.npz", {k: v.shape for k, v in out.items()})
