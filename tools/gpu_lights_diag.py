"""Diagnostic (round 2): product vs reference on the light-in-view scenes -- noise floors of each side and per-pixel z-scores."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from sunvolumerender_b200 import _lib as L  # noqa: E402
from sunvolumerender_b200.render import Renderer  # noqa: E402
import test_gpu_lights_in_view as T  # noqa: E402
from test_gpu_pathtrace import _product_batches, _reference_batches  # noqa: E402
from _gpu_common import reference, rmse  # noqa: E402

r = Renderer(0)
for name in sys.argv[1:] or ["ring_of_small_disks", "front_facing"]:
    cfg = T._setup(r, name, 3, False)
    ref = reference(r, cfg)
    rb, ref_all = _reference_batches(ref, 32, 32, 3)
    del ref
    for mode, cull in ((2, 1), (2, 0), (1, 1)):
        r.set_option(L.OPT_PT_MODE, mode)
        r.set_option(L.OPT_PT_LIGHT_CULL, cull)
        mb = _product_batches(r, 32, 32, 3)
        cap = float(np.percentile(ref_all[ref_all > 0], 99.5))
        cr = lambda a, b: rmse(np.minimum(a, cap), np.minimum(b, cap))
        hr = [rb[i * 8:(i + 1) * 8].mean(axis=0) for i in range(4)]
        hm = [mb[i * 8:(i + 1) * 8].mean(axis=0) for i in range(4)]
        f_rr = np.median([cr(hr[i], hr[j]) for i in range(4) for j in range(i + 1, 4)])
        f_mm = np.median([cr(hm[i], hm[j]) for i in range(4) for j in range(i + 1, 4)])
        f_rm = np.median([cr(hm[i], hr[j]) for i in range(4) for j in range(4)])
        # per-pixel z (channel 0)
        m_r, m_m = rb.mean(axis=0)[..., 0], mb.mean(axis=0)[..., 0]
        se = np.sqrt(rb.var(axis=0, ddof=1)[..., 0] / 32 + mb.var(axis=0, ddof=1)[..., 0] / 32)
        z = np.where(se > 0, (m_m - m_r) / np.maximum(se, 1e-30), 0)
        bad = np.argwhere(np.abs(z) > 5)
        print(f"{name} mode {mode} cull {cull}: cap {cap:.3f} floor ref/ref {f_rr:.4f} prod/prod {f_mm:.4f} prod/ref {f_rm:.4f}; mean ref {ref_all.mean():.6f} prod {mb.mean():.6f}; "
              f"pixels |z|>5: {len(bad)} of {z.size}; var ratio prod/ref (sum over pixels) {mb.var(axis=0).sum() / rb.var(axis=0).sum():.3f}")
        for y, x in bad[:12]:
            print(f"   px ({x},{y}) ref {m_r[y, x]:.4f} prod {m_m[y, x]:.4f} z {z[y, x]:.1f}")
    r.set_option(L.OPT_PT_LIGHT_CULL, 1)
