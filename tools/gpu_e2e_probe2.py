"""Where the e2e step's time goes (N = 1): device events and host clocks around each phase of bench.py's streamed loop.  Scratch tool."""
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
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
rows = []
vs.prefetch(host_vox)
K = 12
for i in range(K):
    h = [time.perf_counter()]; e = [ev()]
    vs.bind(); h.append(time.perf_counter()); e.append(ev())
    r.set_transfer_function(tf_table); r.set_camera(cam); r.set_area_lights(lights); r.set_env_light(env, enabled=cfg.env); h.append(time.perf_counter()); e.append(ev())
    if i + 1 < K: vs.prefetch(host_vox)
    h.append(time.perf_counter()); e.append(ev())
    r.accumulate(sum_buf, cfg.trace_depth, 0, spp, clear=True); h.append(time.perf_counter()); e.append(ev())
    r.resolve(sum_buf); host_img.copy_(r.img, non_blocking=True); host_hdr.copy_(r.hdr, non_blocking=True); e.append(ev())
    torch.cuda.current_stream().synchronize(); h.append(time.perf_counter())
    rows.append((h, e))
names = ["bind", "setup_*", "prefetch", "accumulate", "resolve+D2H"]
for i, (h, e) in enumerate(rows[2:], 2):
    dev = [e[j].elapsed_time(e[j + 1]) for j in range(5)]
    host = [(h[j + 1] - h[j]) * 1e3 for j in range(5)]
    print(i, "dev ms:", " ".join(f"{n}={x:.3f}" for n, x in zip(names, dev)), "| host ms:", " ".join(f"{x:.3f}" for x in host), "| step host", f"{(h[-1]-h[0])*1e3:.3f}", flush=True)
