"""Multi-GPU split of the render path: one process per GPU, the scene replicated,
the SAMPLES of every pixel partitioned across ranks, and one exchange step -- a sum-reduce of the
per-rank float4 (rgb sum, sample count) buffers to rank 0, which resolves (divide, tone map) in one
pass.  Ray casting (one deterministic pass) splits image ROWS instead and needs only a gather.

The functions take the per-rank work as callables so that the same logic runs over NCCL on GPUs
(bench.py, with Renderer.accumulate / Renderer.resolve) and over gloo on CPU in the tests.
"""
import torch
import torch.distributed as dist

from . import scene as S


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def sample_range(total_samples, first_sample=0, rank=None, world_size=None, weak=False):
    """Sample indices [first, first + count) this rank renders.  weak=True: every rank renders
    `total_samples` samples of its own (the job grows with the number of GPUs); otherwise the
    `total_samples` are divided."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    if weak:
        return first_sample + rank * total_samples, total_samples
    first, count = S.split_samples(total_samples, world_size)[rank]
    return first_sample + first, count


def pathtrace_distributed(accumulate, resolve, sum_buf, total_samples, first_sample=0, weak=False, group=None):
    """accumulate(sum_buf, first, count) fills this rank's partial sums (clearing first);
    the partials are sum-reduced onto rank 0, where resolve(sum_buf) produces the image.
    Returns (first, count) rendered by this rank."""
    rank, world_size = world()
    first, count = sample_range(total_samples, first_sample, rank, world_size, weak)
    accumulate(sum_buf, first, count)
    if world_size > 1:
        dist.reduce(sum_buf, dst=0, op=dist.ReduceOp.SUM, group=group)
    if rank == 0:
        resolve(sum_buf)
    return first, count


def raycast_distributed(render_rows, frame, height, group=None):
    """render_rows(frame, y0, y1) writes rows [y0, y1) of the full-frame tensor `frame` (zero
    elsewhere); the row blocks are disjoint, so a sum-reduce onto rank 0 is a gather."""
    rank, world_size = world()
    y0, y1 = S.split_rows(height, world_size)[rank]
    frame.zero_()
    if y1 > y0:
        render_rows(frame, y0, y1)
    if world_size > 1:
        dist.reduce(frame, dst=0, op=dist.ReduceOp.SUM, group=group)
    return y0, y1


def max_over_ranks(value, device=None):
    """Device-timed milliseconds -> the slowest rank's, as every multi-GPU number is reported."""
    rank, world_size = world()
    if world_size == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
