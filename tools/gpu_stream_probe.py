"""Does an in-flight H2D on a side stream delay event records / tiny kernels on the main stream?  Scratch tool."""
import ctypes, torch
x = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()
rt = ctypes.CDLL("libcudart.so.12")
fl = ctypes.c_uint(99); rt.cudaStreamGetFlags(ctypes.c_void_p(side.cuda_stream), ctypes.byref(fl)); print("side stream flags (1 = non-blocking):", fl.value)
def trial(main, what):
    torch.cuda.synchronize()
    with torch.cuda.stream(main):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.cuda.stream(side):
            d.copy_(x, non_blocking=True)
        if what == "kernel":
            t = torch.zeros(16, device="cuda"); t += 1
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for name, main in (("legacy default", torch.cuda.default_stream()), ("explicit", torch.cuda.Stream())):
    for what in ("event only", "kernel"):
        print(name, what, [round(trial(main, what), 3) for _ in range(3)])
