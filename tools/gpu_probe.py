"""First-contact GPU probe: parity of the new kernels against the CPU oracle and the reference's
own CUDA kernels on a tiny scene, plus rough timings.  Scratch tool, not a test."""
import sys, time, json
import numpy as np
import torch
sys.path.insert(0, '.')
from sunvolumerender_b200 import _lib as L, scene as S
from sunvolumerender_b200.render import Renderer, setup_config
from oracle import binding as B

def ev_time(fn, reps=3):
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best

r = Renderer(0)
print('lib version', r.lib.svr_version(), torch.cuda.get_device_name(0))
n, W, H = 64, 64, 64
cfg = S.Config('T', n, L.VOXEL_U8, L.GEN_SPHERE, W, H, 'default')
vb = setup_config(r, cfg)
vox = vb.cpu().numpy().reshape(n, n, n)
print('maxgrad inv', r.volume.invMaxMagnitude)
step = S.raycast_step_size()
# --- ray cast
out = torch.zeros(H * W * 4, dtype=torch.float32, device='cuda')
for skip in (0, 1):
    r.set_option(L.OPT_RC_SKIP, skip)
    r.render_raycasting_f32(out, img=r.img)
    torch.cuda.synchronize()
    mine = out.view(H, W, 4).cpu().numpy(); mine_u8 = r.ldr_image().cpu().numpy().astype(int)
    for fm in (0, 1, 2):
        o = B.CpuOracle(vox, cfg.fmt, (n, n, n), r.volume, S.tf_table('default'), r.camera, r.lights, filter_mode=fm)
        rgba, u8, cnt = o.raycast(step)
        print(f'rc skip={skip} filter={fm}: max|mine-cpu|={np.abs(mine - rgba).max():.3e}  u8 maxdiff={np.abs(mine_u8 - u8.astype(int)).max()}')
ref = B.RefCuda(W, H)
ref.setup(r.volume, r.tf, r.camera, r.lights, r.env)
ref.render_raycasting(step)
ru8 = ref.ldr_image().cpu().numpy().astype(int)
print('rc mine vs refcuda u8 maxdiff', np.abs(mine_u8 - ru8).max(), 'n diff', (mine_u8 != ru8).sum())
# --- path trace compat
r.set_option(L.OPT_PT_MODE, 0)
for depth in (1, 4):
    r.frame_no = 0; ref.frame_no = 0
    r.render_pathtracer(depth); ref.render_pathtracer(1, depth)
    a = r.hdr_image().cpu().numpy(); b = ref.hdr_image().cpu().numpy()
    d = np.abs(a - b).max(axis=2)
    print(f'pt compat depth={depth} 1spp: frac pixels within 1e-4: {(d < 1e-4).mean():.4f}  max {d.max():.3e} mean a {a.mean():.5f} b {b.mean():.5f}')
    o = B.CpuOracle(vox, cfg.fmt, (n, n, n), r.volume, S.tf_table('default'), r.camera, r.lights)
    c, _ = o.pathtrace(depth, 0, 1)
    d = np.abs(b - c).max(axis=2)
    print(f'   cpu oracle vs refcuda: frac within 1e-4: {(d < 1e-4).mean():.4f} mean cpu {c.mean():.5f}')
# --- statistical
spp = 256
ref.frame_no = 0; ref.render_pathtracer(spp, 4); b = ref.hdr_image().cpu().numpy()
for mode in (0, 1, 2):
    r.set_option(L.OPT_PT_MODE, mode)
    r.frame_no = 0; r.render_pathtracer_spp(spp, 4)
    a = r.hdr_image().cpu().numpy()
    print(f'mode {mode} {spp}spp depth4: mean mine {a.mean():.5f} ref {b.mean():.5f}  rmse {np.sqrt(((a-b)**2).mean()):.5f}')
# --- timings on C1
cfg = S.CONFIGS['C1']
vb = setup_config(r, cfg)
ref = B.RefCuda(cfg.width, cfg.height)
ref.setup(r.volume, r.tf, r.camera, r.lights, r.env)
out = torch.zeros(cfg.height * cfg.width * 4, dtype=torch.float32, device='cuda')
print('C1 rc ref ms', 1e3 * ev_time(lambda: ref.render_raycasting(step)))
for skip in (0, 1):
    r.set_option(L.OPT_RC_SKIP, skip)
    print(f'C1 rc mine skip={skip} ms', 1e3 * ev_time(lambda: r.render_raycasting()))
def ref16():
    ref.frame_no = 0; ref.render_pathtracer(16, 1)
print('C1 pt ref 16spp ms', 1e3 * ev_time(ref16))
for mode in (0, 1, 2):
    r.set_option(L.OPT_PT_MODE, mode)
    def mine16():
        r.frame_no = 0; r.render_pathtracer_spp(16, 1)
    print(f'C1 pt mine mode={mode} 16spp ms', 1e3 * ev_time(mine16))
    r.set_option(L.OPT_COUNTERS, 1); r.reset_counters(); mine16(); print('   counters', r.counters()); r.set_option(L.OPT_COUNTERS, 0)
