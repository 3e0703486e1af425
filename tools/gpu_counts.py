"""Counted work of one C3 launch (256 spp) for the default and close views.  Scratch tool."""
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
r = Renderer(0)
cfg = S.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3"]; setup_config(r, cfg)
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 256
buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
cam0 = r.camera
for view in ("default", "close"):
    if view == "close":
        r.set_camera(S.make_camera((0, 0, cam0.pos.z * 0.45), (1, 0, 0), (0, 1, 0), (0, 0, 1), 45.0, 0.0, 1.0, 1.0, cfg.width, cfg.height))
    for lights in (1, 0):
        if not lights:
            r.set_area_lights([])
        r.set_option(L.OPT_COUNTERS, 1); r.reset_counters()
        r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); torch.cuda.synchronize()
        c = r.counters(); r.set_option(L.OPT_COUNTERS, 0)
        sc = max(c["scatters"], 1)
        print(view, "lights" if lights else "nolights", c, "| per scatter: cells %.2f track %.2f shadow %.2f" % (c["cells"] / sc, c["track_taps"] / sc, c["shadow_taps"] / sc), flush=True)
    r.set_area_lights([S.default_area_light(cfg.extent)])
