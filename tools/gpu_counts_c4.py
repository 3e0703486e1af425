"""Counted work of one C4 launch (32 spp) at several macrocell sizes and shadow estimators.  Scratch tool."""
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
r = Renderer(0)
cfg = S.CONFIGS["C4"]; setup_config(r, cfg)
spp = 32
buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
for cell in (0, 8, 16, 32):
    for est in (0, 1):
        r.set_option(L.OPT_MACROCELL_SIZE, cell); r.set_option(L.OPT_SHADOW_ESTIMATOR, est)
        best = 1e9
        for i in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); e1.record(); torch.cuda.synchronize()
            if i: best = min(best, e0.elapsed_time(e1))
        r.set_option(L.OPT_COUNTERS, 1); r.reset_counters()
        r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); torch.cuda.synchronize()
        c = r.counters(); r.set_option(L.OPT_COUNTERS, 0)
        sc = max(c["scatters"], 1)
        mean = buf.view(-1, 4)[:, :3].double().sum().item() / (cfg.width * cfg.height * spp)
        print(f"cell {cell} shadow_est {est}: {best:.2f} ms mean {mean:.6f} | paths {c['paths']} scatters {sc} ({sc/c['paths']:.2f}/path) | per scatter: cells {c['cells']/sc:.2f} track {c['track_taps']/sc:.2f} shadow {c['shadow_taps']/sc:.2f}", flush=True)
