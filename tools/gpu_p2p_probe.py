"""Peer-to-peer copy bandwidth between two GPUs of the box, one process.  Scratch tool."""
import subprocess, torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
print("can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1))
n = 256 << 20
a = torch.zeros(n, dtype=torch.uint8, device="cuda:0"); b = torch.zeros(n, dtype=torch.uint8, device="cuda:1")
for it in range(4):
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    with torch.cuda.device(0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.copy_(a, non_blocking=True); e1.record()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    ms = e0.elapsed_time(e1)
    print(f"push 0->1: {ms:.3f} ms = {n / ms / 1e6:.1f} GB/s")
