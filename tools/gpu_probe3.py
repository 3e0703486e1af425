"""Probe 3: parity re-check + kernel-shape / burst / cell-size sweep on C3 (normal and close view).  Scratch tool."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
from oracle import binding as B


def ev_time(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best


r = Renderer(0)
# ---- parity on a small scene
n, W, H = 64, 64, 64
cfg = S.Config('T', n, L.VOXEL_U8, L.GEN_SPHERE, W, H, 'default')
setup_config(r, cfg)
ref = B.RefCuda(W, H); ref.setup(r.volume, r.tf, r.camera, r.lights, r.env)
step = S.raycast_step_size()
ref.render_raycasting(step); ru8 = ref.ldr_image().cpu().numpy().astype(int)
r.render_raycasting(); mu8 = r.ldr_image().cpu().numpy().astype(int)
print('rc u8 maxdiff', np.abs(mu8 - ru8).max())
for shape in (0, 1):
    r.set_option(L.OPT_PT_KERNEL, shape)
    r.set_option(L.OPT_PT_MODE, 0)
    for depth in (1, 5):
        r.frame_no = 0; ref.frame_no = 0
        for _ in range(3):
            r.render_pathtracer(depth)
        ref.render_pathtracer(3, depth)
        a = r.hdr_image().cpu().numpy(); b = ref.hdr_image().cpu().numpy()
        d = np.abs(a - b).max(axis=2)
        print(f'shape {shape} compat depth={depth} 3 frames: frac within 1e-4: {(d < 1e-4).mean():.4f} max {d.max():.3e}')
    spp = 256
    ref.frame_no = 0; ref.render_pathtracer(spp, 4); b = ref.hdr_image().cpu().numpy()
    for mode in (0, 1, 2):
        for est in (0, 1):
            r.set_option(L.OPT_PT_MODE, mode); r.set_option(L.OPT_SHADOW_ESTIMATOR, est)
            r.frame_no = 0; r.render_pathtracer_spp(spp, 4)
            a = r.hdr_image().cpu().numpy()
            print(f'shape {shape} mode {mode} est {est} {spp}spp depth4: mean mine {a.mean():.5f} ref {b.mean():.5f} rmse {np.sqrt(((a-b)**2).mean()):.5f}')
    r.set_option(L.OPT_SHADOW_ESTIMATOR, 0)
del ref

# ---- C3: bit-identity of the acceleration toggles, then sweeps
cfg = S.CONFIGS['C3']
setup_config(r, cfg)
spp = 16
r.set_option(L.OPT_PT_MODE, 2)
imgs = {}
for shape in (0, 1):
    for leap in (1, 0):
        for cache in (1, 0):
            r.set_option(L.OPT_PT_KERNEL, shape); r.set_option(L.OPT_LEAP, leap); r.set_option(L.OPT_PT_ENTRY_CACHE, cache)
            r.frame_no = 0; r.render_pathtracer_spp(4, 3)
            imgs[(shape, leap, cache)] = r.hdr_image().clone()
base = imgs[(1, 0, 0)]
for k, v in imgs.items():
    print('identity shape/leap/cache', k, 'max abs diff vs (1,0,0):', (v - base).abs().max().item())
r.set_option(L.OPT_LEAP, 1); r.set_option(L.OPT_PT_ENTRY_CACHE, 1)

ref = B.RefCuda(cfg.width, cfg.height); ref.setup(r.volume, r.tf, r.camera, r.lights, r.env)
def refN():
    ref.frame_no = 0; ref.render_pathtracer(spp, 1)
tref = ev_time(refN, 2)
print(f'C3 ref {spp}spp: {tref*1e3:.2f} ms {cfg.width*cfg.height*spp/tref/1e6:.1f} Msamples/s')
def mineN():
    r.frame_no = 0; r.render_pathtracer_spp(spp, 1)
for shape in (0, 1):
    r.set_option(L.OPT_PT_KERNEL, shape)
    for mode in (0, 1, 2):
        r.set_option(L.OPT_PT_MODE, mode)
        t = ev_time(mineN)
        print(f'C3 shape={shape} mode={mode}: {t*1e3:.2f} ms {cfg.width*cfg.height*spp/t/1e6:.1f} Msamples/s  x{tref/t:.2f}')
r.set_option(L.OPT_PT_MODE, 2)
for shape in (0, 1):
    r.set_option(L.OPT_PT_KERNEL, shape)
    for leap, cache in ((0, 0), (1, 0), (1, 1)):
        r.set_option(L.OPT_LEAP, leap); r.set_option(L.OPT_PT_ENTRY_CACHE, cache)
        t = ev_time(mineN)
        print(f'C3 mode=2 shape={shape} leap={leap} cache={cache}: {t*1e3:.2f} ms  x{tref/t:.2f}')
        r.set_option(L.OPT_COUNTERS, 1); r.reset_counters(); mineN(); print('   counters', r.counters()); r.set_option(L.OPT_COUNTERS, 0)
r.set_option(L.OPT_PT_KERNEL, 0)
for burst in (1, 2, 4, 8, 16):
    r.set_option(L.OPT_PT_ROUNDS, burst)
    t = ev_time(mineN)
    print(f'C3 sched mode=2 burst={burst}: {t*1e3:.2f} ms  x{tref/t:.2f}')
r.set_option(L.OPT_PT_ROUNDS, 0)
for cell in (4, 8, 16):
    r.set_option(L.OPT_MACROCELL_SIZE, cell)
    mineN()
    for shape in (0, 1):
        r.set_option(L.OPT_PT_KERNEL, shape)
        for blk in (64, 128, 256):
            r.set_option(L.OPT_PT_BLOCK, blk)
            t = ev_time(mineN)
            print(f'C3 mode=2 cell={cell} shape={shape} blk={blk}: {t*1e3:.2f} ms  x{tref/t:.2f}')
    r.set_option(L.OPT_PT_BLOCK, 128)
r.set_option(L.OPT_MACROCELL_SIZE, 8)
# ---- close view
cam = r.camera
near = S.make_camera((0, 0, cam.pos.z * 0.45), (1, 0, 0), (0, 1, 0), (0, 0, 1), 45.0, 0.0, 1.0, 1.0, cfg.width, cfg.height)
r.set_camera(near)
ref.setup(r.volume, r.tf, r.camera, r.lights, r.env)
tref = ev_time(refN, 2)
print(f'close ref {spp}spp: {tref*1e3:.2f} ms')
for shape in (0, 1):
    r.set_option(L.OPT_PT_KERNEL, shape)
    t = ev_time(mineN)
    print(f'close mode=2 shape={shape}: {t*1e3:.2f} ms  x{tref/t:.2f}')
