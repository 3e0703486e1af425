"""C3 (default and close view): macrocell size x build variant for the sample-parallel kernel.  Scratch tool."""
import os, subprocess, sys
CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
tag = sys.argv[1]; spp = 256
r = Renderer(0); cfg = S.CONFIGS["C3"]; setup_config(r, cfg)
buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
def run(t):
    best = 1e9
    for i in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); e1.record(); torch.cuda.synchronize()
        if i: best = min(best, e0.elapsed_time(e1))
    print(f"{tag:16s} {t:28s} {best:8.3f} ms  {cfg.width*cfg.height*spp/best/1e6:8.2f} Gsamples/s", flush=True)
cam0 = r.camera
for view in ("default", "close"):
    if view == "close":
        r.set_camera(S.make_camera((0, 0, cam0.pos.z * 0.45), (1, 0, 0), (0, 1, 0), (0, 0, 1), 45.0, 0.0, 1.0, 1.0, cfg.width, cfg.height))
    for cell in (4, 8, 16) if tag == "default" else (8,):
        r.set_option(L.OPT_MACROCELL_SIZE, cell)
        run(f"{view} cell={cell}")
    r.set_option(L.OPT_MACROCELL_SIZE, 8)
'''
libs = sys.argv[1:] or [""]
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["SVR_B200_LIB"] = os.path.abspath(lib)
    subprocess.call([sys.executable, "-c", CHILD, os.path.basename(lib) or "default"], env=env)
