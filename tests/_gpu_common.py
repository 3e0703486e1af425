"""Shared helpers of the GPU parity tests: scene setup through the C ABI, the reference's own
kernels (oracle/_ref, test infrastructure) on the same texture objects, the CPU oracle on the same
voxels."""
import numpy as np
import torch

from oracle import binding as B
from sunvolumerender_b200 import _lib as L
from sunvolumerender_b200 import scene as S
from sunvolumerender_b200.render import setup_config


def small_config(name="S", n=64, w=128, h=128, fmt=L.VOXEL_U8, gen=L.GEN_SPHERE, tf="default", depth=1, env=False):
    return S.Config(name, n, fmt, gen, w, h, tf, trace_depth=depth, spp=1, env=env)


def setup(r, cfg):
    """Loads cfg into the renderer; returns the voxels as a host numpy array (z, y, x)."""
    r.set_option(L.OPT_PT_MODE, 2)
    r.set_option(L.OPT_PT_KERNEL, 2)
    r.set_option(L.OPT_PT_WARP_PIXELS, 0)
    r.set_option(L.OPT_PT_WARP_MIN_SPP, 32)
    r.set_option(L.OPT_PT_QUEUE_MIN_DEPTH, 8)
    r.set_option(L.OPT_SHADOW_ESTIMATOR, 0)
    r.set_option(L.OPT_RC_SKIP, 1)
    r.set_option(L.OPT_LEAP, 1)
    r.set_option(L.OPT_PT_ENTRY_CACHE, 1)
    r.set_option(L.OPT_PT_LIGHT_CULL, 1)
    r.set_option(L.OPT_PT_PROFILE, 0)
    r.set_option(L.OPT_PT_REFILL, 0)
    r.set_option(L.OPT_PT_POOL_PIXELS, 0)
    r.set_option(L.OPT_ENV_NEE, 0)
    r.set_option(L.OPT_PT_BLOCK_SPLIT, 0)
    r.set_option(L.OPT_PT_PIXEL_CACHE, 1)
    r.set_option(L.OPT_FUSED_UPLOAD, 1)
    r.set_option(L.OPT_PT_LOOKAHEAD, 32)
    r.set_option(L.OPT_MACROCELL_SIZE, 8)
    r.set_option(L.OPT_COUNTERS, 0)
    r.set_option(L.OPT_SEED, 0x5EED)
    vb = setup_config(r, cfg)
    vox = vb.cpu().numpy().view(S.VOXEL_DTYPES[cfg.fmt]).reshape(cfg.n, cfg.n, cfg.n)
    return vox


def reference(r, cfg, f32=False, r32=False, env=False):
    """env=True: the environment-light twin (pathtracer.cu:233 re-enabled at build time, oracle/Makefile)."""
    ref = B.RefCuda(cfg.width, cfg.height, r32=r32, f32=f32, env=env)
    ref.setup(r.volume, r.tf, r.camera, r.lights, r.env)
    return ref


def cpu_oracle(r, cfg, vox, env_enabled=False):
    return B.CpuOracle(vox, cfg.fmt, (cfg.n,) * 3, r.volume, S.tf_table(cfg.tf), r.camera, r.lights, env=r.env, env_enabled=env_enabled)


def raycast_f32(r, rows=None):
    H, W = r.camera.imageH, r.camera.imageW
    out = torch.zeros(H * W * 4, dtype=torch.float32, device="cuda")
    r.render_raycasting_f32(out, rows=rows, img=r.img)
    torch.cuda.synchronize()
    return out.view(H, W, 4)


def rmse(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).mean()))
