"""C3: macrocell edge 2 vs 4 (counts and time).  Scratch tool."""
import sys, torch
sys.path.insert(0, ".")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
r = Renderer(0); cfg = S.CONFIGS["C3"]; setup_config(r, cfg); spp = 256
buf = torch.zeros(cfg.width * cfg.height * 4, dtype=torch.float32, device="cuda")
for cell in (4, 2):
    r.set_option(L.OPT_MACROCELL_SIZE, cell)
    best = 1e9
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); e1.record(); torch.cuda.synchronize()
        if i: best = min(best, e0.elapsed_time(e1))
    r.set_option(L.OPT_COUNTERS, 1); r.reset_counters(); r.accumulate(buf, cfg.trace_depth, 0, spp, clear=True); torch.cuda.synchronize()
    c = r.counters(); r.set_option(L.OPT_COUNTERS, 0); sc = c["scatters"]
    print(f"cell {cell}: {best:.3f} ms | per scatter: cells {c['cells']/sc:.2f} track {c['track_taps']/sc:.2f} shadow {c['shadow_taps']/sc:.2f}", flush=True)
