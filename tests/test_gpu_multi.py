"""Two-GPU tests (skipped on a one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
the sample split + NCCL sum-reduce of bench.py against a single-GPU render of the same samples, the ray caster's
row bands written by both ranks into rank 0's image through peer mappings (distributed.PeerFrame), and
render.VolumeStream's three fan-outs (copy-engine pushes into peer-mapped staging buffers, NCCL broadcast, one
upload per rank) against direct uploads."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sunvolumerender_b200 import _lib as L
from sunvolumerender_b200 import scene as S

pytestmark = pytest.mark.gpu

N, W, H, DEPTH, SPP = 48, 96, 64, 2, 64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _frames():
    base = S.sphere_volume(N, L.VOXEL_U16)
    return [base.copy(), (base[::-1, :, ::-1] // 2).copy(), (base[:, ::-1, :] // 3 * 2).copy()]


def _setup(r, vox):
    cfg = S.Config("mp", N, L.VOXEL_U16, L.GEN_SPHERE, W, H, "default", trace_depth=DEPTH, spp=SPP)
    r.set_option(L.OPT_PT_MODE, 2)
    r.set_option(L.OPT_MACROCELL_SIZE, 8)
    r.set_option(L.OPT_SEED, 0x5EED)
    r.load_volume(vox, cfg.fmt, (N,) * 3, max_grad_mag=3000.0)
    r.set_transfer_function(S.tf_table(cfg.tf))
    r.set_camera(S.default_camera(cfg.extent, W, H))
    r.set_area_lights([S.default_area_light(cfg.extent)])
    r.set_env_light(S.constant_env_light(), enabled=False)


def _worker(rank, world_size, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=dev)
    from sunvolumerender_b200 import distributed as D
    from sunvolumerender_b200.render import Renderer, VolumeStream

    try:
        r = Renderer(rank)
        frames = _frames()
        _setup(r, frames[0])
        buf = torch.zeros(H * W * 4, dtype=torch.float32, device=dev)

        # ---- sample split + NCCL reduce + resolve
        first, count = D.pathtrace_distributed(lambda b, f, c: r.accumulate(b, DEPTH, f, c, clear=True), lambda b: r.resolve(b), buf, SPP)
        torch.cuda.synchronize()
        if rank == 0:
            np.save(os.path.join(out_dir, "split_hdr.npy"), r.hdr_image().cpu().numpy())
            r.accumulate(buf, DEPTH, 0, SPP, clear=True)
            r.resolve(buf)
            torch.cuda.synchronize()
            np.save(os.path.join(out_dir, "single_hdr.npy"), r.hdr_image().cpu().numpy())
        np.save(os.path.join(out_dir, f"range_{rank}.npy"), np.array([first, count]))

        # ---- VolumeStream: every rank ends up with every frame's voxels, whatever the fan-out
        def render():
            r.accumulate(buf, DEPTH, 0, 32, clear=True)
            torch.cuda.synchronize()
            return buf.clone()

        direct = []
        for f in frames:
            r.upload_volume(f)
            direct.append(render())
        pinned = [torch.from_numpy(f).view(torch.uint8).reshape(-1).pin_memory() for f in frames]
        used = {}
        for fanout in ("p2p", "nvlink", "pcie"):
            vs = VolumeStream(r, fanout=fanout)
            used[fanout] = vs.fanout
            # only rank 0 holds the data for the one-upload fan-outs: the others pass garbage
            mine = pinned if (rank == 0 or fanout == "pcie") else [torch.zeros_like(p).pin_memory() for p in pinned]
            vs.prefetch(mine[0])
            for i in range(len(frames)):
                vs.bind()
                r.set_transfer_function(S.tf_table("default"))
                if i + 1 < len(frames):
                    vs.prefetch(mine[i + 1])
                img = render()
                assert torch.equal(img, direct[i]), (fanout, rank, i)
            vs.close()
        if rank == 0:
            np.save(os.path.join(out_dir, "fanouts.npy"), np.array([used[k] for k in ("p2p", "nvlink", "pcie")]))
        dist.barrier()

        # ---- ray casting split into row bands, every rank writing STRAIGHT into rank 0's image (distributed.PeerFrame)
        r.upload_volume(frames[0])
        r.set_transfer_function(S.tf_table("thin"))
        r.render_raycasting()
        torch.cuda.synchronize()
        whole = r.ldr_image().clone()
        pf = D.PeerFrame(r, W * H * 4)
        for it in range(3):   # three frames into the same buffer, separated by barriers
            r.render_raycasting_bands(rank, world_size, img_ptr=pf.img_ptr)
            pf.frame_done()
            torch.cuda.synchronize()
            if rank == 0:
                assert torch.equal(pf.image().view(H, W, 4), whole), it
            dist.barrier()
        pf.close()

        # ---- the path tracer's IMAGE split: both ranks render all samples of their row bands straight into rank 0's
        # accumulator; the frame is the single-GPU frame bit for bit
        r.set_transfer_function(S.tf_table("default"))
        pf = D.PeerFrame(r, W * H * 16)
        r.accumulate_bands(pf.img_ptr, DEPTH, 0, SPP, rank, world_size, clear=True)
        pf.frame_done()
        torch.cuda.synchronize()
        if rank == 0:
            got = pf.image().view(torch.float32)
            r.accumulate(buf, DEPTH, 0, SPP, clear=True)
            torch.cuda.synchronize()
            assert torch.equal(got, buf)
        dist.barrier()
        pf.close()
        if rank == 0:
            np.save(os.path.join(out_dir, "peer_frame_ok.npy"), np.array([1]))
    finally:
        dist.destroy_process_group()


def test_two_gpu_split_and_volume_stream(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "range_0.npy"), np.load(tmp_path / "range_1.npy")
    assert (r0[0], r0[0] + r0[1]) == (0, r1[0]) and r1[0] + r1[1] == SPP
    split, single = np.load(tmp_path / "split_hdr.npy"), np.load(tmp_path / "single_hdr.npy")
    # the same samples summed in a different order (two partial sums instead of one)
    assert np.allclose(split, single, rtol=2e-5, atol=1e-6)
    assert single.mean() > 0
    used = list(np.load(tmp_path / "fanouts.npy"))
    assert used[1:] == ["nvlink", "pcie"] and used[0] in ("p2p", "nvlink")  # p2p falls back where CUDA IPC is unavailable
    assert os.path.exists(tmp_path / "peer_frame_ok.npy")
