"""Time of the macrocell range stage, texture path vs linear-source path (C3 volume).  Scratch tool."""
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
r = Renderer(0); cfg = S.CONFIGS["C3"]; vb = setup_config(r, cfg)
buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
host = torch.empty(vb.numel(), dtype=torch.uint8, pin_memory=True); host.copy_(vb)
r.accumulate(buf, 1, 0, 32, clear=True); torch.cuda.synchronize()
def t(fn, n=5):
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
def up_dev():
    r.upload_volume(vb)
def up_host():
    r.upload_volume(host)
def render1():
    r.accumulate(buf, 1, 0, 1, clear=True)   # 1 spp: grid refresh + a cheap launch
print(f"upload from device (D2D -> array + linear range): {t(up_dev):.3f} ms; then 1-spp render incl. majorants: {t(lambda: (up_dev(), render1())) - t(up_dev):.3f} ms")
print(f"upload from pinned host (H2D -> array): {t(up_host):.3f} ms; then 1-spp render incl. texture range + majorants: {t(lambda: (up_host(), render1())) - t(up_host):.3f} ms")
print(f"1-spp render alone: {t(render1):.3f} ms")
