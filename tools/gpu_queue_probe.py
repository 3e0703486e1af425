"""Kernel shape 3 (scatter queue) against shape 2 on C3 and C4.  Scratch tool."""
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
r = Renderer(0)
for name, spp in (("C3", 256), ("C4", 128)):
    cfg = S.CONFIGS[name]; setup_config(r, cfg)
    buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
    for shape in (2, 3):
        r.set_option(L.OPT_PT_KERNEL, shape)
        best = 1e9
        for i in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); e1.record(); torch.cuda.synchronize()
            if i: best = min(best, e0.elapsed_time(e1))
        mean = buf.view(-1, 4)[:, :3].double().sum().item() / (cfg.width * cfg.height * spp)
        print(f"{name} shape {shape}: {best:9.3f} ms  {cfg.width*cfg.height*spp/best/1e6:8.3f} Gsamples/s  mean {mean:.7f}", flush=True)
    r.set_option(L.OPT_PT_KERNEL, 2)
