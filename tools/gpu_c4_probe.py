"""C4 (1024^3 f16 cloud, depth 32) and C5-like checks: ours vs the reference's kernels.  Scratch tool."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
from oracle import binding as B
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
r = Renderer(0)
cfg = S.CONFIGS[name]
t = time.time(); setup_config(r, cfg); torch.cuda.synchronize(); print(f"{name} setup {time.time()-t:.2f}s  invMaxMag {r.volume.invMaxMagnitude:.3e}  mem {torch.cuda.memory_allocated()/2**30:.2f} GiB", flush=True)
W, H = cfg.width, cfg.height
buf = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
def run(tag, n=2):
    best = 1e9
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    img = buf.view(H, W, 4)[..., :3] / spp
    print(f"{tag:34s} {best:9.3f} ms  {W*H*spp/best/1e6:8.3f} Gsamples/s  mean {float(img.mean()):.6f} finite {bool(torch.isfinite(img).all())}", flush=True)
    return img.clone()
r.set_option(L.OPT_ENV_ENABLED, 0)
imgs = {}
for shape in (2, 1):
    r.set_option(L.OPT_PT_KERNEL, shape); r.set_option(L.OPT_PT_MODE, 2)
    imgs[shape] = run(f"ours mode2 shape{shape}")
r.set_option(L.OPT_PT_KERNEL, 2)
r.set_option(L.OPT_COUNTERS, 1); r.reset_counters(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); torch.cuda.synchronize(); print("counters", r.counters()); r.set_option(L.OPT_COUNTERS, 0)
for cell in (4, 16):
    r.set_option(L.OPT_MACROCELL_SIZE, cell); run(f"ours mode2 shape2 cell={cell}")
r.set_option(L.OPT_MACROCELL_SIZE, 8)
r.set_option(L.OPT_SHADOW_ESTIMATOR, 1); run("ours mode2 ratio shadow"); r.set_option(L.OPT_SHADOW_ESTIMATOR, 0)
r.set_option(L.OPT_PT_MODE, 1); run("ours mode1 (global majorant)", 1)
try:
    ref = B.RefCuda(W, H)
    ref.setup(r.volume, r.tf, r.camera, r.lights, r.env)
    ref.render_pathtracer(1, cfg.trace_depth); torch.cuda.synchronize()
    ref.frame_no = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ref.render_pathtracer(spp, cfg.trace_depth); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); a = ref.hdr_image()
    print(f"reference {spp} frames: {ms:.3f} ms  {W*H*spp/ms/1e6:.3f} Gsamples/s  mean {float(a.mean()):.6f}")
    d = (imgs[2] - a)
    print("mean diff rel", float(d.mean() / a.mean()), "rmse", float((d ** 2).mean().sqrt()))
except Exception as e:
    print("reference unavailable:", e)
