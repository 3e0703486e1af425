"""A/B of library variants (tools/build_variants.sh): C3 default view, C3 close view, C4 (32 spp).  Scratch tool.
   python tools/gpu_sweep4.py build/variants/libsvr_a.so build/variants/libsvr_b.so ..."""
import os, subprocess, sys
CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
tag = sys.argv[1]
r = Renderer(0)
def run(cfg, t, spp, reps=4):
    buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
    best = 1e9
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); e1.record(); torch.cuda.synchronize()
        if i: best = min(best, e0.elapsed_time(e1))
    chk = buf.view(-1, 4)[:, :3].double().sum().item() / (cfg.width * cfg.height * spp)
    print(f"{tag:14s} {t:14s} {best:8.3f} ms  {cfg.width*cfg.height*spp/best/1e6:8.2f} Gsamples/s  mean {chk:.7f}", flush=True)
cfg = S.CONFIGS["C3"]; setup_config(r, cfg)
cam0 = r.camera
run(cfg, "C3", 256)
r.set_camera(S.make_camera((0, 0, cam0.pos.z * 0.45), (1, 0, 0), (0, 1, 0), (0, 0, 1), 45.0, 0.0, 1.0, 1.0, cfg.width, cfg.height))
run(cfg, "C3close", 256, 3)
if "--c4" in sys.argv:
    cfg = S.CONFIGS["C4"]; setup_config(r, cfg)
    run(cfg, "C4", 32, 3)
'''
args = [a for a in sys.argv[1:] if not a.startswith("--")]
flags = [a for a in sys.argv[1:] if a.startswith("--")]
for lib in args or [""]:
    env = dict(os.environ)
    if lib:
        env["SVR_B200_LIB"] = os.path.abspath(lib)
    subprocess.call([sys.executable, "-c", CHILD, os.path.basename(lib).replace("libsvr_", "").replace(".so", "") or "default"] + flags, env=env)
