"""Short, fixed workload for ncu:
    python tools/prof_run.py <pt|proto|rc|ref> <config> [mode] [spp] [reps] [shape] [view]
Sets up one BASELINE.json configuration and launches the chosen kernel a few times.  `ref` runs the
reference's own kernels (oracle/_ref) on the same scene, for side-by-side profiles.  view = close
moves the camera in so that the volume fills the frame."""
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
shape = int(sys.argv[6]) if len(sys.argv) > 6 else 2
view = sys.argv[7] if len(sys.argv) > 7 else "default"

r = Renderer(0)
setup_config(r, cfg)
if view == "close":
    cam = r.camera
    r.set_camera(S.make_camera((0, 0, cam.pos.z * 0.45), (1, 0, 0), (0, 1, 0), (0, 0, 1), 45.0, 0.0, 1.0, 1.0, cfg.width, cfg.height))
r.set_option(L.OPT_PT_MODE, mode)
r.set_option(L.OPT_PT_KERNEL, shape)
import os  # noqa: E402
if os.environ.get("SVR_LOOKAHEAD"):  # e.g. -32: batches forced (under ncu the per-launch overhead makes the library's own timing switch them off)
    r.set_option(L.OPT_PT_LOOKAHEAD, int(os.environ["SVR_LOOKAHEAD"]))
if what == "ref":
    from oracle import binding as B  # noqa: E402  (profiling aid, not a product path)

    ref = B.RefCuda(cfg.width, cfg.height)
    ref.setup(r.volume, r.tf, r.camera, r.lights, r.env)
    for _ in range(reps):
        ref.frame_no = 0
        ref.render_pathtracer(spp, cfg.trace_depth)
else:
    for _ in range(reps):
        if what == "proto":   # the drop-in protocol: spp calls of one sample each (sample look-ahead from frame 16)
            r.frame_no = 0
            for _ in range(spp):
                r.render_pathtracer(cfg.trace_depth)
        elif what == "pt":
            r.frame_no = 0
            r.render_pathtracer_spp(spp, cfg.trace_depth)
        else:
            r.render_raycasting()
torch.cuda.synchronize()
print("done", what, cfg.name, mode, spp, r.launch_count())
