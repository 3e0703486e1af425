"""One-hot volume: tex3D returns the filter weight of that texel directly.  Scratch tool."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L  # noqa: E402
from sunvolumerender_b200.render import Renderer  # noqa: E402
r = Renderer(0)
out = {}
def fetch3(uvw):
    m = uvw.shape[0]
    d_uvw = torch.from_numpy(np.ascontiguousarray(uvw, np.float32)).cuda()
    d_out = torch.zeros(m, dtype=torch.float32, device="cuda")
    L.check(r.lib.svr_debug_sample_volume(C.byref(r.volume), C.c_void_p(d_uvw.data_ptr()), m, C.c_void_p(d_out.data_ptr())))
    return d_out.cpu().numpy()
rng = np.random.default_rng(3)
n = 4
m = 1 << 17
p = rng.uniform(0.0, 3.0, (m, 3)).astype(np.float32)      # texel-space xb
# a lattice of exact 1/256 positions too
g = (rng.integers(0, 3 * 256, (m, 3)).astype(np.float32) / np.float32(256))
for tag, pts in (("rand", p), ("grid", g)):
    uvw = ((pts + np.float32(0.5)) / np.float32(n)).astype(np.float32)
    out[f"{tag}_uvw"] = uvw
    for name, dt, fmt, one in (("f32", np.float32, L.VOXEL_F32, 1.0), ("u16", np.uint16, L.VOXEL_U16, 65535), ("u8", np.uint8, L.VOXEL_U8, 255), ("f16", np.float16, L.VOXEL_F16, 1.0)):
        vox = np.zeros((n, n, n), dt)
        vox[1, 1, 1] = one
        r.load_volume(vox, fmt, (n, n, n), max_grad_mag=1.0)
        out[f"{tag}_{name}"] = fetch3(uvw)
        vox[:] = 0
        vox[1, 1, 1] = one; vox[2, 2, 2] = one   # two opposite corners
        r.load_volume(vox, fmt, (n, n, n), max_grad_mag=1.0)
        out[f"{tag}_{name}_diag"] = fetch3(uvw)
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/filter_probe2.npz", **out)
print("ok")
