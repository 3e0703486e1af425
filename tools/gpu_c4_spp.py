"""C4 at full depth, 256 spp per launch: library variants.  Scratch tool."""
import os, subprocess, sys
CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
r = Renderer(0); cfg = S.CONFIGS["C4"]; setup_config(r, cfg)
buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
for spp in (64, 256):
    best = 1e9
    for i in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    chk = buf.view(-1, 4)[:, :3].double().sum().item() / (cfg.width * cfg.height * spp)
    print(f"{sys.argv[1]:10s} C4 spp {spp}: {best:9.2f} ms {cfg.width*cfg.height*spp/best/1e6:7.3f} Gsamples/s mean {chk:.7f}", flush=True)
'''
for lib in sys.argv[1:]:
    env = dict(os.environ); env["SVR_B200_LIB"] = os.path.abspath(lib)
    subprocess.call([sys.executable, "-c", CHILD, os.path.basename(lib).replace("libsvr_", "").replace(".so", "")], env=env)
