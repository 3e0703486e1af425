"""A/B of library variants on the drop-in protocol: 256 x render_pathtracer (1 sample per call) on C3, default and close view.
   python tools/gpu_protocol_sweep.py [lib.so ...]"""
import os, subprocess, sys
CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
tag = sys.argv[1]
r = Renderer(0)
cfg = S.CONFIGS["C3"]; setup_config(r, cfg)
cam0 = r.camera
def run(name):
    best = 1e9
    for rep in range(4):
        r.frame_no = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(256):
            r.render_pathtracer(cfg.trace_depth)
        e1.record(); torch.cuda.synchronize()
        if rep: best = min(best, e0.elapsed_time(e1))
    print(f"{tag:10s} {name:16s} 256 x 1 spp: {best:8.3f} ms  {cfg.width*cfg.height*256/best/1e6:7.2f} Gs/s  mean {r.hdr_image().double().mean().item():.6f}", flush=True)
for ahead in (32, 0, -32):
    r.set_option(L.OPT_PT_LOOKAHEAD, ahead)
    r.set_camera(cam0)
    run(f"C3 la={ahead}")
    r.set_camera(S.make_camera((0, 0, cam0.pos.z * 0.45), (1, 0, 0), (0, 1, 0), (0, 0, 1), 45.0, 0.0, 1.0, 1.0, cfg.width, cfg.height))
    run(f"C3close la={ahead}")
'''
for lib in [a for a in sys.argv[1:] if a.endswith(".so") or a == "default"] or ["default"]:
    env = dict(os.environ)
    if lib != "default":
        env["SVR_B200_LIB"] = os.path.abspath(lib)
    subprocess.call([sys.executable, "-c", CHILD, os.path.basename(lib).replace("libsvr_", "").replace(".so", "") or "default"], env=env)
