"""C3 path-tracing time over build variants (SVR_B200_LIB) and launch options.  Scratch tool.
usage: python tools/gpu_sweep.py [spp] [lib ...]"""
import os, subprocess, sys
CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
spp = int(sys.argv[1]); tag = sys.argv[2]
cfgname = sys.argv[3] if len(sys.argv) > 3 else "C3"
r = Renderer(0); cfg = S.CONFIGS[cfgname]; setup_config(r, cfg)
buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
for blk in (64, 128, 256):
    r.set_option(L.OPT_PT_BLOCK, blk)
    best = 1e9
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); e1.record(); torch.cuda.synchronize()
        if i: best = min(best, e0.elapsed_time(e1))
    print(f"{tag:28s} {cfgname} blk={blk:3d} spp={spp}: {best:8.3f} ms  {cfg.width*cfg.height*spp/best/1e6:8.2f} Gsamples/s  mean {float(buf.view(-1,4)[:,:3].mean())/spp:.6f}", flush=True)
'''
spp = sys.argv[1] if len(sys.argv) > 1 else "256"
cfgname = os.environ.get("SWEEP_CFG", "C3")
libs = sys.argv[2:] or [""]
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["SVR_B200_LIB"] = os.path.abspath(lib)
    subprocess.call([sys.executable, "-c", CHILD, spp, os.path.basename(lib) or "default", cfgname], env=env)
