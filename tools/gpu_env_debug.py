import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import hdr_oracle as H
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer
from _gpu_common import reference, setup, small_config
from test_gpu_pathtrace import _reference_batches, _product_batches, _tile_means
import pathlib, tempfile
r = Renderer(0)
cfg = small_config(gen=L.GEN_CT, fmt=L.VOXEL_U16, depth=3, env=True)
setup(r, cfg)
tmp = pathlib.Path(tempfile.mkdtemp())
w, h = 128, 64
v, u = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
img = np.stack([0.2 + 0.6 * u, 0.3 + 0.3 * np.sin(6.28 * u) ** 2, 0.9 - 0.7 * v], axis=2)
img[h // 4: h // 4 + 4, w // 2: w // 2 + 6] = [60.0, 50.0, 30.0]
r.load_env_map(H.write_hdr(tmp / "sky.hdr", H.float_to_rgbe(img)), intensity=1.5, offset=(0.13, -0.04))
r.set_area_lights([])
K, per, depth = 8, 32, 3
ref = reference(r, cfg, env=True)
rb, ref_all = _reference_batches(ref, 4 * K, per, depth)
for mode in (0, 1, 2):
    r.set_option(L.OPT_PT_MODE, mode)
    mb = _product_batches(r, K, per, depth)
    tm = np.stack([_tile_means(b) for b in mb]); tr = np.stack([_tile_means(b) for b in rb])
    se = np.sqrt(tm.var(axis=0, ddof=1) / K + tr.var(axis=0, ddof=1) / (4 * K))
    z = np.abs(tm.mean(axis=0) - tr.mean(axis=0)) / np.maximum(se, 1e-30)
    print("mode", mode, "frac z<4.5", (z < 4.5).mean(), "mean", mb.mean(), ref_all.mean())
    print(np.round(z.max(axis=2), 1))
    rel = (tm.mean(axis=0) - tr.mean(axis=0)) / np.maximum(tr.mean(axis=0), 1e-9)
    print(np.round(rel[..., 0] * 100, 2))
print("---- reference vs reference (first 16 batches vs last 16)")
ta, tb = tr[:16], tr[16:]
se = np.sqrt(ta.var(axis=0, ddof=1) / 16 + tb.var(axis=0, ddof=1) / 16)
z = np.abs(ta.mean(axis=0) - tb.mean(axis=0)) / np.maximum(se, 1e-30)
print("frac z<4.5", (z < 4.5).mean()); print(np.round(z.max(axis=2), 1))
print("---- product mode 2, batches 0-7 vs another seed")
r.set_option(L.OPT_PT_MODE, 2)
m1 = np.stack([_tile_means(b) for b in _product_batches(r, 16, per, depth)])
r.set_option(L.OPT_SEED, 777)
m2 = np.stack([_tile_means(b) for b in _product_batches(r, 16, per, depth)])
se = np.sqrt(m1.var(axis=0, ddof=1) / 16 + m2.var(axis=0, ddof=1) / 16)
z = np.abs(m1.mean(axis=0) - m2.mean(axis=0)) / np.maximum(se, 1e-30)
print("frac z<4.5", (z < 4.5).mean()); print(np.round(z.max(axis=2), 1))
print("top-left tile, channel 0: ref mean %.8f  mine(seed A) %.8f mine(seed B) %.8f ; per-batch std ref %.3e mine %.3e" % (tr.mean(axis=0)[0,0,0], m1.mean(axis=0)[0,0,0], m2.mean(axis=0)[0,0,0], tr.std(axis=0)[0,0,0], m1.std(axis=0)[0,0,0]))
