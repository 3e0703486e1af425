"""Environment lighting (core/lights/cuda_environment_light.h:58-72; core/lights/lights.cpp:31-90) through
the C ABI.  The reference left its only call site commented out (pathtracer.cu:233); the checker is the
environment-light twin of the reference -- the same sources with that one line re-enabled at build time
(oracle/Makefile) -- so constant skies and HDR maps are compared path for path in the twin mode and
statistically in the product mode."""
import numpy as np
import pytest
import torch

from oracle import hdr_oracle as H
from sunvolumerender_b200 import _lib as L
from sunvolumerender_b200 import scene as S

from _gpu_common import reference, setup, small_config
from test_gpu_pathtrace import _frames, _statistical_parity

pytestmark = pytest.mark.gpu


def _sky_file(tmp_path, w=128, h=64):
    v, u = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    img = np.stack([0.2 + 0.6 * u, 0.3 + 0.3 * np.sin(6.28 * u) ** 2, 0.9 - 0.7 * v], axis=2)
    img[h // 4: h // 4 + 4, w // 2: w // 2 + 6] = [60.0, 50.0, 30.0]
    return H.write_hdr(tmp_path / "sky.hdr", H.float_to_rgbe(img))


def _twin_check(renderer, cfg, depth, frames=3):
    renderer.set_option(L.OPT_PT_MODE, 0)
    ref = reference(renderer, cfg, env=True)
    mine = _frames(renderer, frames, depth)
    ref.render_pathtracer(frames, depth)
    theirs = ref.hdr_image().cpu().numpy()
    d = np.abs(mine - theirs).max(axis=2)
    tol = 1e-4 * np.maximum(1.0, theirs.max(axis=2))
    assert (d <= tol).mean() >= 0.998, (d.max(), (d <= tol).mean())
    assert abs(mine.mean() - theirs.mean()) <= 2e-3 * theirs.mean()
    return mine


def test_constant_sky_matches_the_reference_twin(renderer):
    cfg = small_config(gen=L.GEN_CT, fmt=L.VOXEL_U16, depth=3, env=True)
    setup(renderer, cfg)
    mine = _twin_check(renderer, cfg, 3)
    assert mine[0, 0] == pytest.approx([0.5, 0.5, 0.5])          # a corner ray sees the sky (gui/canvas.cpp:11-12)
    env = S.constant_env_light((0.1, 0.4, 0.9), intensity=2.0)
    renderer.set_env_light(env, enabled=True)
    mine = _twin_check(renderer, cfg, 3)
    assert mine[0, 0] == pytest.approx([0.2, 0.8, 1.8])
    _statistical_parity(renderer, cfg, 3, 8, 32, lambda: renderer.set_option(L.OPT_PT_MODE, 2), mean_tol=0.01, env=True)


def test_hdr_map_matches_the_reference_twin(renderer, tmp_path):
    cfg = small_config(gen=L.GEN_CT, fmt=L.VOXEL_U16, depth=3, env=True)
    setup(renderer, cfg)
    renderer.load_env_map(_sky_file(tmp_path), intensity=1.5, offset=(0.13, -0.04))
    mine = _twin_check(renderer, cfg, 3)
    assert np.ptp(mine[0, :, 2]) > 0.005                         # the map, not a constant, is what the corner rays see
    renderer.set_area_lights([])                                 # lit by the sky alone
    _twin_check(renderer, cfg, 2)
    _statistical_parity(renderer, cfg, 3, 8, 32, lambda: renderer.set_option(L.OPT_PT_MODE, 2), mean_tol=0.01, env=True)
    # offset wraps (cudaAddressModeWrap, lights.cpp:62-63): a whole turn changes nothing
    renderer.set_option(L.OPT_PT_MODE, 2)
    a = renderer.env
    renderer.frame_no = 0
    renderer.render_pathtracer_spp(4, 2)
    torch.cuda.synchronize()
    img0 = renderer.hdr_image().clone()
    a.offset = L.Vec2(a.offset.x + 1.0, a.offset.y)
    renderer.set_env_light(a, enabled=True)
    renderer.frame_no = 0
    renderer.render_pathtracer_spp(4, 2)
    torch.cuda.synchronize()
    assert torch.allclose(renderer.hdr_image(), img0, rtol=1e-3, atol=1e-4)
