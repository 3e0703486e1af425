"""Times svr_volume_upload from a device buffer (the streamed-volume path: one pass that fills the cudaArray and reduces
the macrocell ranges) against the copy + range-kernel path, on the C3 volume.  Run under ncu to capture the kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config

cfg = S.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
r = Renderer(0)
vb = setup_config(r, cfg)
r.render_pathtracer_spp(32, cfg.trace_depth)
torch.cuda.synchronize()
src = vb.clone()
for fused in (1, 0, 1):
    r.set_option(L.OPT_FUSED_UPLOAD, fused)
    ts = []
    for i in range(4):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        r.upload_volume(src)
        e[1].record()
        r.frame_no = 0
        r.render_pathtracer_spp(32, cfg.trace_depth)
        e[2].record()
        torch.cuda.synchronize()
        ts.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    print(f"fused={fused}: upload {min(t[0] for t in ts):.3f} ms, grid build + 32-spp render {min(t[1] for t in ts):.3f} ms")

# the same upload right behind a long render kernel, with and without a device-to-host copy running beside it
sum_buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
host = torch.empty(cfg.width * cfg.height * 16, dtype=torch.uint8).pin_memory()
side = torch.cuda.Stream()
r.set_option(L.OPT_FUSED_UPLOAD, 1)
for d2h in (0, 1, 0, 1):
    ts = []
    for i in range(4):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        r.accumulate(sum_buf, cfg.trace_depth, i * 256, 256, clear=True)
        r.resolve(sum_buf)
        e[0].record()
        if d2h:
            with torch.cuda.stream(side):
                side.wait_event(e[0])
                host.copy_(sum_buf.view(torch.uint8), non_blocking=True)
        r.upload_volume(src)
        e[1].record()
        torch.cuda.synchronize()
        ts.append(e[0].elapsed_time(e[1]))
    print(f"behind a 256-spp render, D2H beside it={d2h}: upload {min(ts):.3f} .. {max(ts):.3f} ms")

# source freshly written over PCIe on another stream (what VolumeStream.bind consumes), with and without a full
# synchronize between the transfer and the upload
hsrc = torch.empty(src.numel(), dtype=torch.uint8).pin_memory()
hsrc.copy_(src.cpu())
stage = torch.empty_like(src)
for sync in (0, 1, 0, 1):
    ts = []
    for i in range(4):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ready = torch.cuda.Event()
        r.accumulate(sum_buf, cfg.trace_depth, i * 256, 256, clear=True)
        with torch.cuda.stream(side):
            stage.copy_(hsrc, non_blocking=True)
            ready.record(side)
        r.resolve(sum_buf)
        if sync:
            torch.cuda.synchronize()
        e[0].record()
        torch.cuda.current_stream().wait_event(ready)
        r.upload_volume(stage)
        e[1].record()
        torch.cuda.synchronize()
        ts.append(e[0].elapsed_time(e[1]))
    print(f"source written by an H2D copy beside the previous render, synchronize before the upload={sync}: upload {min(ts):.3f} .. {max(ts):.3f} ms")
