"""C3 path-tracing time over kernel shape / warp pixels / block size.  Scratch tool."""
import sys, torch, itertools
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfgname = sys.argv[2] if len(sys.argv) > 2 else "C3"
r = Renderer(0); cfg = S.CONFIGS[cfgname]; setup_config(r, cfg)
if len(sys.argv) > 3 and sys.argv[3] == "close":
    cam = r.camera
    r.set_camera(S.make_camera((0, 0, cam.pos.z * 0.45), (1, 0, 0), (0, 1, 0), (0, 0, 1), 45.0, 0.0, 1.0, 1.0, cfg.width, cfg.height))
buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
def run(tag):
    best = 1e9
    for i in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); e1.record(); torch.cuda.synchronize()
        if i: best = min(best, e0.elapsed_time(e1))
    print(f"{tag:40s} {best:8.3f} ms  {cfg.width*cfg.height*spp/best/1e6:8.2f} Gsamples/s  mean {float(buf.view(-1,4)[:,:3].mean())/spp:.6f}", flush=True)
r.set_option(L.OPT_PT_KERNEL, 1)
for blk in (128,):
    r.set_option(L.OPT_PT_BLOCK, blk); run(f"mega blk={blk}")
r.set_option(L.OPT_PT_KERNEL, 2)
for blk, wp in itertools.product((64, 128, 256), (1, 2, 4, 8, 16, 32)):
    r.set_option(L.OPT_PT_BLOCK, blk); r.set_option(L.OPT_PT_WARP_PIXELS, wp); run(f"warp blk={blk} warpPixels={wp}")
