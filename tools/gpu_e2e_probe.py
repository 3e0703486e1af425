"""Where does the e2e step of bench.py spend its time?  Scratch tool."""
import sys, time
import torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config

r = Renderer(0)
cfg = S.CONFIGS["C3"]
vb = setup_config(r, cfg)
W, H = cfg.width, cfg.height
host_vox = torch.empty(vb.numel(), dtype=torch.uint8, pin_memory=True); host_vox.copy_(vb)
host_img = torch.empty(W * H * 4, dtype=torch.uint8, pin_memory=True)
host_hdr = torch.empty(W * H * 3, dtype=torch.float32, pin_memory=True)
sum_buf = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
tf = S.tf_table(cfg.tf)

def timed(name, fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    print(f"{name:50s} min {min(ts):8.3f} ms  median {sorted(ts)[len(ts)//2]:8.3f} ms")

timed("upload_volume (H2D pinned -> cudaArray, drop grid)", lambda: r.upload_volume(host_vox))
timed("upload_volume from device (D2D -> cudaArray)", lambda: r.upload_volume(vb))
stage = torch.empty_like(vb)
timed("H2D pinned -> linear device buffer", lambda: stage.copy_(host_vox, non_blocking=True))
timed("set_transfer_function", lambda: r.set_transfer_function(tf))
timed("set_camera+lights+env", lambda: (r.set_camera(S.default_camera(cfg.extent, W, H)), r.set_area_lights([S.default_area_light(cfg.extent)]), r.set_env_light(S.constant_env_light(), enabled=True)))
def first_render():
    r.upload_volume(vb)
    r.accumulate(sum_buf, 1, 0, 1, clear=True)
timed("upload D2D + 1 spp (grid rebuild incl.)", first_render)
timed("1 spp, grid cached", lambda: r.accumulate(sum_buf, 1, 0, 1, clear=True))
timed("256 spp, grid cached", lambda: r.accumulate(sum_buf, 1, 0, 256, clear=True), 3)
timed("resolve", lambda: r.resolve(sum_buf))
timed("D2H img+hdr", lambda: (host_img.copy_(r.img, non_blocking=True), host_hdr.copy_(r.hdr, non_blocking=True)))
