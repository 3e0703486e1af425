"""Pixels per warp run (SVR_OPT_PT_WARP_PIXELS) against samples per launch, C3 default and close view: ms per launch."""
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
r = Renderer(0)
cfg = S.CONFIGS["C3"]; setup_config(r, cfg)
buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
cam0 = r.camera
for view in ("default", "close"):
    if view == "close":
        r.set_camera(S.make_camera((0, 0, cam0.pos.z * 0.45), (1, 0, 0), (0, 1, 0), (0, 0, 1), 45.0, 0.0, 1.0, 1.0, cfg.width, cfg.height))
    for spp in (32, 64, 128, 256):
        row = []
        for wp in (1, 2, 4):
            r.set_option(L.OPT_PT_WARP_PIXELS, wp)
            best = 1e9
            for i in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r.accumulate(buf, cfg.trace_depth, i * spp, spp, clear=True); e1.record(); torch.cuda.synchronize()
                if i: best = min(best, e0.elapsed_time(e1))
            row.append(f"wp={wp}: {best:7.3f}")
        print(f"{view:8s} {spp:4d} spp per launch   " + "   ".join(row), flush=True)
