"""Probe 4: decompose the C3 path-tracing time: all-miss (clipped away), no lights, env off, kernel shapes."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config


def ev_time(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best


r = Renderer(0)
cfg = S.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'C3']
setup_config(r, cfg)
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
depth = cfg.trace_depth
def mineN():
    r.frame_no = 0; r.render_pathtracer_spp(spp, depth)
def report(tag):
    for shape in (0, 1):
        r.set_option(L.OPT_PT_KERNEL, shape)
        t = ev_time(mineN)
        print(f'{tag} shape={shape}: {t*1e3:.3f} ms  {cfg.width*cfg.height*spp/t/1e6:.0f} Msamples/s')
    r.set_option(L.OPT_COUNTERS, 1); r.reset_counters(); mineN(); print('   counters', r.counters()); r.set_option(L.OPT_COUNTERS, 0)
r.set_option(L.OPT_PT_MODE, 2)
report('normal')
lights = r.lights
r.set_area_lights([]); report('no lights'); r.set_area_lights(lights)
r.set_option(L.OPT_ENV_ENABLED, 0); report('env off'); r.set_option(L.OPT_ENV_ENABLED, 1 if cfg.env else 0)
r.set_volume_params(x_clip=(0.0, 0.0)); report('all rays miss (x clip 0,0)'); r.set_volume_params(x_clip=(-1.0, 1.0))
cam = r.camera
far = S.make_camera((0, 0, cam.pos.z * 40), (1, 0, 0), (0, 1, 0), (0, 0, 1), 45.0, 0.0, 1.0, 1.0, cfg.width, cfg.height)
r.set_camera(far); report('camera 40x farther'); r.set_camera(cam)
near = S.make_camera((0, 0, cam.pos.z * 0.45), (1, 0, 0), (0, 1, 0), (0, 0, 1), 45.0, 0.0, 1.0, 1.0, cfg.width, cfg.height)
r.set_camera(near); report('camera close (volume fills view)'); r.set_camera(cam)
