"""A/B of library variants x option sets (round 2): C3 default view, C3 close view, optionally C4 / C1.  Scratch tool.
   python tools/gpu_sweep5.py [--c4] [--opts "profile=0;profile=1;profile=1,refill=4"] [lib.so ...]"""
import os, subprocess, sys
CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
tag, optsets = sys.argv[1], sys.argv[2].split(";")
KEYS = {"profile": L.OPT_PT_PROFILE, "refill": L.OPT_PT_REFILL, "kernel": L.OPT_PT_KERNEL, "cell": L.OPT_MACROCELL_SIZE, "wp": L.OPT_PT_WARP_PIXELS,
        "qdepth": L.OPT_PT_QUEUE_MIN_DEPTH, "block": L.OPT_PT_BLOCK, "shadow": L.OPT_SHADOW_ESTIMATOR, "split": L.OPT_PT_BLOCK_SPLIT}
r = Renderer(0)
def run(cfg, t, spp, reps=4):
    buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
    for o in optsets:
        defaults = {"profile": 0, "refill": 0, "kernel": 2, "cell": 0, "wp": 0, "qdepth": 8, "block": 128, "shadow": 0, "split": 0}
        for kv in o.split(","):
            if kv:
                k, v = kv.split("="); defaults[k] = int(v)
        for k, v in defaults.items():
            r.set_option(KEYS[k], v)
        best = 1e9
        for i in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r.accumulate(buf, cfg.trace_depth, i * spp, spp, clear=True); e1.record(); torch.cuda.synchronize()
            if i: best = min(best, e0.elapsed_time(e1))
        chk = buf.view(-1, 4)[:, :3].double().sum().item() / (cfg.width * cfg.height * spp)
        r.set_option(L.OPT_COUNTERS, 1); r.reset_counters(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); torch.cuda.synchronize()
        c = r.counters(); r.set_option(L.OPT_COUNTERS, 0)
        print(f"{tag:10s} {t:8s} {o:28s} {best:8.3f} ms {cfg.width*cfg.height*spp/best/1e6:7.2f} Gs/s mean {chk:.6f} taps/path trk {c['track_taps']/c['paths']:.3f} shd {c['shadow_taps']/c['paths']:.3f} cells {c['cells']/c['paths']:.2f} scat {c['scatters']/c['paths']:.4f}", flush=True)
if "--noc3" not in sys.argv:
    cfg = S.CONFIGS["C3"]; setup_config(r, cfg)
    cam0 = r.camera
    run(cfg, "C3", 256)
    r.set_camera(S.make_camera((0, 0, cam0.pos.z * 0.45), (1, 0, 0), (0, 1, 0), (0, 0, 1), 45.0, 0.0, 1.0, 1.0, cfg.width, cfg.height))
    run(cfg, "C3close", 256, 3)
if "--c1" in sys.argv:
    cfg = S.CONFIGS["C1"]; setup_config(r, cfg)
    run(cfg, "C1x256", 256, 3)
if "--c4" in sys.argv:
    cfg = S.CONFIGS["C4"]; setup_config(r, cfg)
    for c4spp in ([512, 128] if "--c4big" in sys.argv else [128 if "--c4spp128" in sys.argv else 32]):
        run(cfg, f"C4x{c4spp}", c4spp, 2 if c4spp > 128 else 3)
'''
args = [a for a in sys.argv[1:] if not a.startswith("--") and a.endswith(".so")]
flags = [a for a in sys.argv[1:] if a.startswith("--") and a != "--opts"]
opts = "profile=1"
if "--opts" in sys.argv:
    opts = sys.argv[sys.argv.index("--opts") + 1]
for lib in args or [""]:
    env = dict(os.environ)
    if lib:
        env["SVR_B200_LIB"] = os.path.abspath(lib)
    subprocess.call([sys.executable, "-c", CHILD, os.path.basename(lib).replace("libsvr_", "").replace(".so", "") or "default", opts] + flags, env=env)
