"""Best macrocell size per scene vs statistics of the cell-8 majorant grid.  Scratch tool."""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
r = Renderer(0)
def stats():
    dims = (C.c_int32 * 3)(); cell = C.c_int32()
    L.check(r.lib.svr_grid_info(dims, C.byref(cell)))
    gx, gy, gz = dims
    maj = np.zeros((gz, gy, gx), np.float32)
    L.check(r.lib.svr_grid_copy(C.c_void_p(maj.ctypes.data), None))
    occ = maj[maj > 0]
    return f"cells {maj.size} occupied {occ.size/maj.size:.3f} mean {occ.mean():.4f} median {np.median(occ):.4f} p10 {np.percentile(occ,10):.4f} p90 {np.percentile(occ,90):.4f}"
scenes = [("C3", S.CONFIGS["C3"], None), ("C3close", S.CONFIGS["C3"], 0.45), ("C4", S.CONFIGS["C4"], None), ("C1", S.CONFIGS["C1"], None),
          ("C3thin", S.Config("C3thin", 512, L.VOXEL_U16, L.GEN_CT, 1920, 1080, "thin", 1, 256), None),
          ("C3d8", S.Config("C3d8", 512, L.VOXEL_U16, L.GEN_CT, 1920, 1080, "default", 8, 256), None)]
for name, cfg, close in scenes:
    setup_config(r, cfg)
    r.set_option(L.OPT_ENV_ENABLED, 0)
    if close:
        cam = r.camera
        r.set_camera(S.make_camera((0, 0, cam.pos.z * close), (1, 0, 0), (0, 1, 0), (0, 0, 1), 45.0, 0.0, 1.0, 1.0, cfg.width, cfg.height))
    buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
    spp = 64 if cfg.trace_depth < 8 else 32
    res = []
    for cell in (4, 8, 16, 32):
        r.set_option(L.OPT_MACROCELL_SIZE, cell)
        best = 1e9
        for i in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); e1.record(); torch.cuda.synchronize()
            if i: best = min(best, e0.elapsed_time(e1))
        res.append(f"cell{cell}: {best:.2f}ms")
        if cell == 8: st = stats()
    print(name, " ".join(res), "|", st, flush=True)
    r.set_option(L.OPT_MACROCELL_SIZE, 8)
