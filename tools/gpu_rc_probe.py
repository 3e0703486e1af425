"""Ray caster: time vs macrocell size and block size on C1/C2.  Scratch tool."""
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
r = Renderer(0)
step = S.raycast_step_size()
for name, cfg in (("C2thin", S.Config("C2", 256, 0, 1, 1024, 1024, "thin")), ("C2default", S.Config("C2", 256, 0, 1, 1024, 1024, "default")), ("C1", S.CONFIGS["C1"])):
    setup_config(r, cfg)
    for cell in (2, 4, 8, 16, 0):
        r.set_option(L.OPT_MACROCELL_SIZE, cell)
        for blk in (64, 128, 256):
            r.set_option(L.OPT_RC_BLOCK, blk)
            r.render_raycasting(step); torch.cuda.synchronize()
            best = 1e9
            for i in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r.render_raycasting(step); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"{name:10s} cell={cell:2d} blk={blk:3d}: {best:.3f} ms  {cfg.width*cfg.height/best/1e3:.1f} Mrays/s", flush=True)
    r.set_option(L.OPT_RC_BLOCK, 128); r.set_option(L.OPT_MACROCELL_SIZE, 0)
