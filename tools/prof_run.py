"""Short, fixed workload for ncu: `python tools/prof_run.py <pt|rc> <config> [mode] [spp] [reps]`.
Sets up one BASELINE.json configuration and launches the chosen kernel a few times."""
import sys

import torch

sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S  # noqa: E402
from sunvolumerender_b200.render import Renderer, setup_config  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "pt"
cfg = S.CONFIGS[sys.argv[2] if len(sys.argv) > 2 else "C3"]
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 2
spp = int(sys.argv[4]) if len(sys.argv) > 4 else 8
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3

r = Renderer(0)
setup_config(r, cfg)
r.set_option(L.OPT_PT_MODE, mode)
for _ in range(reps):
    if what == "pt":
        r.frame_no = 0
        r.render_pathtracer_spp(spp, cfg.trace_depth)
    else:
        r.render_raycasting()
torch.cuda.synchronize()
print("done", what, cfg.name, mode, spp, r.launch_count())
