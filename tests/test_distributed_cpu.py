"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in sunvolumerender_b200/distributed.py.
The per-rank "kernels" are the CPU oracle here (test infrastructure); on GPUs bench.py plugs in
Renderer.accumulate / Renderer.resolve and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sunvolumerender_b200 import _lib as L
from sunvolumerender_b200 import distributed as D
from sunvolumerender_b200 import scene as S

N, W, H, DEPTH, SPP = 16, 24, 20, 2, 6


def _oracle():
    from oracle import binding as B

    vox = S.sphere_volume(N, L.VOXEL_U8)
    vol = S.host_volume_struct((N, N, N), max_grad_mag=3000.0)
    cam = S.default_camera((N, N, N), W, H)
    return B.CpuOracle(vox, L.VOXEL_U8, (N, N, N), vol, S.tf_table("default"), cam, [S.default_area_light((N, N, N))])


def _frame_radiance(o, f):
    # one reference frame into a zeroed buffer leaves L / (f + 1) (running_estimate, pathtracer.cu:81-84)
    hdr, _ = o.pathtrace(DEPTH, f, 1, hdr=np.zeros((H, W, 3), np.float32))
    return hdr.astype(np.float64) * (f + 1)


def _accumulate_factory(o):
    def accumulate(buf, first, count):
        acc = np.zeros((H, W, 4), np.float64)
        for f in range(first, first + count):
            acc[..., :3] += _frame_radiance(o, f)
            acc[..., 3] += 1
        buf.copy_(torch.from_numpy(acc.astype(np.float32)))

    return accumulate


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, weak, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        o = _oracle()
        result = {}

        def resolve(buf):
            b = buf.numpy()
            result["hdr"] = b[..., :3] / b[..., 3:4]
            result["count"] = b[..., 3].copy()

        buf = torch.zeros(H, W, 4, dtype=torch.float32)
        first, count = D.pathtrace_distributed(_accumulate_factory(o), resolve, buf, SPP, weak=weak)
        frame = torch.zeros(H, W, 4, dtype=torch.float32)

        def render_rows(fr, y0, y1):
            rgba, _, _ = o.raycast(S.raycast_step_size(), rows=(y0, y1))
            fr[y0:y1] = torch.from_numpy(rgba[y0:y1])

        rows = D.raycast_distributed(render_rows, frame, H)
        t = D.max_over_ranks(10.0 + rank)
        np.save(os.path.join(out_dir, f"range_{rank}.npy"), np.array([first, count, rows[0], rows[1], t]))
        if rank == 0:
            np.save(os.path.join(out_dir, "hdr.npy"), result["hdr"])
            np.save(os.path.join(out_dir, "count.npy"), result["count"])
            np.save(os.path.join(out_dir, "rc.npy"), frame.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("weak", [False, True])
def test_two_rank_sample_split_and_reduce(tmp_path, oracle_cpu, weak):
    world_size = 2
    mp.spawn(_worker, args=(world_size, _free_port(), weak, str(tmp_path)), nprocs=world_size, join=True)
    r0 = np.load(tmp_path / "range_0.npy")
    r1 = np.load(tmp_path / "range_1.npy")
    total = SPP * world_size if weak else SPP
    # the ranks' sample ranges tile [0, total) without overlap
    assert (r0[0], r0[0] + r0[1]) == (0, r1[0]) and r1[0] + r1[1] == total
    assert r0[4] == r1[4] == 11.0  # max over ranks
    # single-process result over the same samples
    o = _oracle()
    ref = sum(_frame_radiance(o, f) for f in range(total)) / total
    hdr = np.load(tmp_path / "hdr.npy")
    assert np.allclose(hdr, ref, rtol=1e-5, atol=1e-6)
    assert (np.load(tmp_path / "count.npy") == total).all()
    # ray-cast row split: disjoint rows, gather == full frame
    full, _, _ = o.raycast(S.raycast_step_size())
    assert np.array_equal(np.load(tmp_path / "rc.npy"), full)
    assert (r0[2], r1[3]) == (0, H) and r0[3] == r1[2]


def test_single_process_is_the_identity():
    assert D.world() == (0, 1)
    assert D.sample_range(10, first_sample=5) == (5, 10)
    assert D.sample_range(10, rank=1, world_size=4) == (3, 3)
    assert D.sample_range(10, rank=3, world_size=4, weak=True) == (30, 10)
    assert D.max_over_ranks(3.5) == 3.5
