"""Can rank 0 push a staging buffer into the other ranks' device memory with copy engines (CUDA IPC + P2P over NVLink),
and signal them with an interprocess event?  Launch with torchrun, >= 2 GPUs.  Scratch tool."""
import os, sys, time, torch, torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
gl = dist.new_group(backend="gloo")
n = 256 << 20
stage = torch.zeros(n, dtype=torch.uint8, device=dev)
handles = [None] * world
dist.all_gather_object(handles, reduce_tensor(stage) if rank != 0 else None, group=gl)
ev = torch.cuda.Event(interprocess=True) if rank == 0 else None
evh = [ev.ipc_handle() if rank == 0 else None]
dist.broadcast_object_list(evh, src=0, group=gl)
copy_stream = torch.cuda.Stream()
if rank == 0:
    t0 = time.time()
    peers = {r: handles[r][0](*handles[r][1]) for r in range(1, world)}
    print("opened", len(peers), "peer buffers in", round(time.time() - t0, 2), "s; devices:", [str(p.device) for p in peers.values()], flush=True)
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True); host.random_(0, 255)
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        with torch.cuda.stream(copy_stream):
            e0.record(); stage.copy_(host, non_blocking=True); e1.record()
            for r, p in peers.items():
                p.copy_(stage, non_blocking=True)
            e2.record(); ev.record()
        dist.barrier(group=gl)      # the record has been CALLED: peers may now wait on the event
        torch.cuda.synchronize()
        print(f"it {it}: H2D {e0.elapsed_time(e1):.3f} ms, push to {world - 1} peers {e1.elapsed_time(e2):.3f} ms", flush=True)
        dist.barrier(group=gl)
    chk = int(host.to(torch.int64).sum())
    obj = [chk]
else:
    pev = torch.cuda.Event.from_ipc_handle(dev, evh[0])
    for it in range(3):
        dist.barrier(group=gl)
        torch.cuda.current_stream().wait_event(pev)
        s = int(stage.to(torch.int64).sum())     # reads after the event
        torch.cuda.synchronize()
        dist.barrier(group=gl)
    obj = [None]
dist.broadcast_object_list(obj, src=0, group=gl)
if rank != 0:
    print(f"rank {rank}: checksum {'OK' if s == obj[0] else 'MISMATCH'} ({s} vs {obj[0]})", flush=True)
dist.barrier(group=gl)
dist.destroy_process_group()
