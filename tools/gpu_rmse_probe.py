"""RMSE(new, ref)/RMSE(ref, ref) over seeds and estimator modes.  Scratch tool."""
import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer
from _gpu_common import reference, rmse, setup, small_config
from test_gpu_pathtrace import _reference_batches, _product_batches
r = Renderer(0)
depth, K, per = 4, 8, 64
cfg = small_config(gen=L.GEN_CT, fmt=L.VOXEL_U16, depth=depth)
setup(r, cfg)
ref = reference(r, cfg)
rb, _ = _reference_batches(ref, 4 * K, per, depth)
halves = [rb[i * K:(i + 1) * K].mean(axis=0) for i in range(4)]
floors = [rmse(halves[i], halves[j]) for i in range(4) for j in range(i + 1, 4)]
ref_all = rb.mean(axis=0)
cap = float(np.percentile(ref_all[ref_all > 0], 99.5))
_r = rmse
rmse = lambda a, b: _r(np.minimum(a, cap), np.minimum(b, cap))
floors = [rmse(halves[i], halves[j]) for i in range(4) for j in range(i + 1, 4)]
print("cap", cap, "ref-vs-ref floors:", np.round(floors, 5), "median", np.median(floors))
for mode, est, shape in [(2, 0, 2), (2, 1, 2), (1, 0, 2), (1, 1, 0), (1, 1, 1), (2, 1, 1)]:
    out = []
    for seed in (0x5EED, 1, 2, 3):
        r.set_option(L.OPT_SEED, seed); r.set_option(L.OPT_PT_MODE, mode); r.set_option(L.OPT_SHADOW_ESTIMATOR, est); r.set_option(L.OPT_PT_KERNEL, shape)
        mine = _product_batches(r, K, per, depth).mean(axis=0)
        out.append([rmse(mine, h) for h in halves])
    print(mode, est, shape, "median rmse per seed / floor:", np.round(np.median(np.array(out), axis=1) / np.median(floors), 3).tolist())
