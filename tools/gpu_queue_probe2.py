"""Kernel shape 3 on C4 for library variants (register budgets).  Scratch tool."""
import os, subprocess, sys
CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
r = Renderer(0); cfg = S.CONFIGS["C4"]; setup_config(r, cfg); spp = 128
buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
for shape in (2, 3):
    r.set_option(L.OPT_PT_KERNEL, shape)
    best = 1e9
    for i in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); e1.record(); torch.cuda.synchronize()
        if i: best = min(best, e0.elapsed_time(e1))
    print(f"{sys.argv[1]:6s} C4 shape {shape}: {best:9.3f} ms  {cfg.width*cfg.height*spp/best/1e6:8.3f} Gsamples/s", flush=True)
'''
for lib in sys.argv[1:]:
    env = dict(os.environ); env["SVR_B200_LIB"] = os.path.abspath(lib)
    subprocess.call([sys.executable, "-c", CHILD, os.path.basename(lib).replace("libsvr_", "").replace(".so", "")], env=env)
