"""How does the texture unit quantise linear-filter weights?  Dumps tex3D of a ramp volume at fine
coordinate steps and compares with candidate models; also the TF's tex1D.  Scratch tool whose result
pins oracle/svr_oracle.cpp's software sampler (filterMode)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S  # noqa: E402
from sunvolumerender_b200.render import Renderer  # noqa: E402

r = Renderer(0)
for n in (4, 64, 512):
    vox = np.zeros((4, 4, n), np.float32)
    vox[:] = np.arange(n, dtype=np.float32)[None, None, :]
    r.load_volume(vox, L.VOXEL_F32, (n, 4, 4), max_grad_mag=1.0)
    m = 1 << 16
    rng = np.random.default_rng(0)
    xs = rng.uniform(0.5, n - 1.5, m).astype(np.float32)  # texel-space position of the sample
    u = ((xs + np.float32(0.5)) / np.float32(n)).astype(np.float32)
    uvw = np.stack([u, np.full(m, 0.5 / 4, np.float32), np.full(m, 0.5 / 4, np.float32)], 1).astype(np.float32)
    d_uvw = torch.from_numpy(uvw).cuda()
    d_out = torch.zeros(m, dtype=torch.float32, device="cuda")
    L.check(r.lib.svr_debug_sample_volume(C.byref(r.volume), C.c_void_p(d_uvw.data_ptr()), m, C.c_void_p(d_out.data_ptr())))
    got = d_out.cpu().numpy().astype(np.float64)
    xb = (u.astype(np.float32) * np.float32(n) - np.float32(0.5)).astype(np.float32).astype(np.float64)  # as the oracle computes it
    xb_exact = u.astype(np.float64) * n - 0.5
    cands = {
        "round(frac*256)/256 (f32 xb)": np.floor(xb) + np.floor((xb - np.floor(xb)) * 256 + 0.5) / 256,
        "floor(frac*256)/256 (f32 xb)": np.floor(xb) + np.floor((xb - np.floor(xb)) * 256) / 256,
        "round(xb*256)/256 (exact xb)": np.floor(xb_exact * 256 + 0.5) / 256,
        "floor(xb*256)/256 (exact xb)": np.floor(xb_exact * 256) / 256,
        "round(u*n*256)/256 - 0.5": np.floor(u.astype(np.float64) * n * 256 + 0.5) / 256 - 0.5,
        "floor(u*n*256)/256 - 0.5": np.floor(u.astype(np.float64) * n * 256) / 256 - 0.5,
        "unquantised": xb_exact,
    }
    print(f"n={n}: result*256 integral? max dev {np.abs(got * 256 - np.round(got * 256)).max():.3e}")
    for k, v in cands.items():
        e = np.abs(got - v)
        print(f"   {k:36s} max err {e.max():.3e}  mismatches(>1e-6) {(e > 1e-6).mean():.5f}")
# TF: 1024 x float4 with .w = index
tab = np.zeros((1024, 4), np.float32)
tab[:, 3] = np.arange(1024)
r.set_transfer_function(tab)
m = 1 << 16
xs = np.random.default_rng(1).uniform(0.5, 1022.5, m).astype(np.float32)
x = ((xs + np.float32(0.5)) / np.float32(1024)).astype(np.float32)
d_x = torch.from_numpy(x).cuda()
d_o = torch.zeros(m * 4, dtype=torch.float32, device="cuda")
L.check(r.lib.svr_debug_sample_tf(C.byref(r.tf), C.c_void_p(d_x.data_ptr()), m, C.c_void_p(d_o.data_ptr())))
got = d_o.view(m, 4)[:, 3].cpu().numpy().astype(np.float64)
xe = x.astype(np.float64) * 1024 - 0.5
for k, v in {"round(xb*256)/256": np.floor(xe * 256 + 0.5) / 256, "floor(xb*256)/256": np.floor(xe * 256) / 256}.items():
    e = np.abs(got - v)
    print(f"TF {k:24s} max err {e.max():.3e} mismatches {(e > 1e-6).mean():.5f}")
