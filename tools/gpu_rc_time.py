"""Ray caster frame times on C2 (thin / default) and C1.  Scratch tool."""
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
r = Renderer(0)
for name, cfg in (("C2 thin", S.Config("C2", 256, 0, 1, 1024, 1024, "thin")), ("C2 default", S.Config("C2", 256, 0, 1, 1024, 1024, "default")), ("C1", S.CONFIGS["C1"])):
    setup_config(r, cfg); step = S.raycast_step_size()
    r.render_raycasting(step); torch.cuda.synchronize()
    best = 1e9
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.render_raycasting(step); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name}: {best:.4f} ms  {cfg.width*cfg.height/best/1e3:.1f} Mrays/s", flush=True)
