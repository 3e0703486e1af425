"""Which host-side call between prefetch and accumulate delays the render?  Scratch tool."""
import sys, time, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config, VolumeStream
r = Renderer(0); cfg = S.CONFIGS["C3"]; vb = setup_config(r, cfg)
W, H, spp = cfg.width, cfg.height, cfg.spp; npix = W * H
sum_buf = torch.zeros(npix * 4, dtype=torch.float32, device="cuda")
host_vox = torch.empty(vb.numel(), dtype=torch.uint8, pin_memory=True); host_vox.copy_(vb)
host_img = torch.empty(npix * 4, dtype=torch.uint8, pin_memory=True)
host_hdr = torch.empty(npix * 3, dtype=torch.float32, pin_memory=True)
tf_table = S.tf_table(cfg.tf); cam = S.default_camera(cfg.extent, W, H); lights = [S.default_area_light(cfg.extent)]; env = S.constant_env_light()
vs = VolumeStream(r, vb.numel())
def loop(mode, K=10):
    torch.cuda.synchronize(); vs.prefetch(host_vox); out = []
    for i in range(K):
        t0 = time.perf_counter()
        vs.bind()
        r.set_transfer_function(tf_table); r.set_camera(cam); r.set_area_lights(lights); r.set_env_light(env, enabled=cfg.env)
        if mode == "event_before":
            e = torch.cuda.Event(enable_timing=True); e.record()
        if i + 1 < K: vs.prefetch(host_vox)
        if mode == "event_after":
            e = torch.cuda.Event(enable_timing=True); e.record()
        if mode == "event_after_notiming":
            e = torch.cuda.Event(); e.record()
        r.accumulate(sum_buf, cfg.trace_depth, 0, spp, clear=True)
        r.resolve(sum_buf); host_img.copy_(r.img, non_blocking=True); host_hdr.copy_(r.hdr, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out.append((time.perf_counter() - t0) * 1e3)
    print(f"{mode:22s}", " ".join(f"{x:.2f}" for x in out), flush=True)
for m in ("none", "event_before", "event_after", "event_after_notiming", "none"):
    loop(m)
