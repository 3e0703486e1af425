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


class PeerFrame:
    """One image owned by rank `root`, written by every rank of the box through peer mappings (CUDA IPC over NVLink): the
    multi-GPU ray caster's gather without a collective.  Every rank renders its row bands with `img_ptr` as the image
    (root: its own buffer; the others: the peer mapping) and calls frame_done(); when root's stream has passed
    frame_done() the whole frame is in root's buffer.  Completion travels as a flag in root's memory that the other
    ranks' streams raise and root's stream waits on (svr_peer_signal / svr_peer_wait): no NCCL call, no host sync.
    Successive frames reuse the buffer: the caller separates them (bench.py: a barrier between frames).

    Collective: every rank of `group` constructs it and closes it."""

    def __init__(self, renderer, nbytes, root=0, group=None, timeout_ms=2000):
        import ctypes as C

        from . import _lib as L

        self.C, self.L = C, L
        self.r, self.lib = renderer, renderer.lib
        self.rank, self.world = world()
        self.root, self.group, self.timeout_ms = root, group, timeout_ms
        self.nbytes, self.frames = int(nbytes), 0
        self.own = self.rank == root
        handles = None
        if self.own:
            self.img_ptr, self.flag_ptr = C.c_void_p(0), C.c_void_p(0)
            L.check(self.lib.svr_stage_alloc(C.byref(self.img_ptr), self.nbytes), "svr_stage_alloc")
            L.check(self.lib.svr_stage_alloc(C.byref(self.flag_ptr), 8), "svr_stage_alloc")
            zero = torch.zeros(2, dtype=torch.int32, device=renderer.device)
            L.check(self.lib.svr_stage_copy(self.flag_ptr, C.c_void_p(zero.data_ptr()), 8, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "svr_stage_copy")
            torch.cuda.synchronize()
            hs = []
            for p in (self.img_ptr, self.flag_ptr):
                h = (C.c_ubyte * 64)()
                L.check(self.lib.svr_stage_export(p, C.byref(h)), "svr_stage_export")
                hs.append(bytes(h))
            handles = hs
        box = [handles]
        if self.world > 1:
            dist.broadcast_object_list(box, src=root, group=group)
        if not self.own:
            ptrs = []
            for hb in box[0]:
                q = C.c_void_p(0)
                h = (C.c_ubyte * 64).from_buffer_copy(hb)
                L.check(self.lib.svr_stage_import(C.byref(h), C.byref(q)), "svr_stage_import")
                ptrs.append(q)
            self.img_ptr, self.flag_ptr = ptrs

    def frame_done(self):
        """Stream-ordered: the other ranks raise root's flag after their writes, root waits for all of them."""
        self.frames += 1
        if self.world == 1:
            return
        if self.own:
            self.L.check(self.lib.svr_peer_wait(self.flag_ptr, self.frames * (self.world - 1), self.timeout_ms), "svr_peer_wait")
        else:
            self.L.check(self.lib.svr_peer_signal(self.flag_ptr), "svr_peer_signal")

    def image(self):
        """root: a copy of the frame as a uint8 tensor; also checks that no wait ever timed out."""
        assert self.own
        out = torch.empty(self.nbytes, dtype=torch.uint8, device=self.r.device)
        flag = torch.zeros(2, dtype=torch.int32, device=self.r.device)
        st = self.C.c_void_p(torch.cuda.current_stream().cuda_stream)
        self.L.check(self.lib.svr_stage_copy(self.C.c_void_p(out.data_ptr()), self.img_ptr, self.nbytes, st), "svr_stage_copy")
        self.L.check(self.lib.svr_stage_copy(self.C.c_void_p(flag.data_ptr()), self.flag_ptr, 8, st), "svr_stage_copy")
        torch.cuda.synchronize()
        if int(flag[1].item()) != 0:
            raise RuntimeError("PeerFrame: a wait for the other ranks' bands timed out")
        return out

    def close(self):
        torch.cuda.synchronize()
        if self.world > 1:
            if not self.own:
                self.lib.svr_stage_release(self.img_ptr)
                self.lib.svr_stage_release(self.flag_ptr)
            dist.barrier(group=self.group)
        if self.own:
            self.lib.svr_stage_free(self.img_ptr)
            self.lib.svr_stage_free(self.flag_ptr)


def max_over_ranks(value, device=None):
    """Device-timed milliseconds -> the slowest rank's, as every multi-GPU number is reported."""
    rank, world_size = world()
    if world_size == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
