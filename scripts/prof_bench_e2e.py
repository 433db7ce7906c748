import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mentpy_b200 as mb
B, T = 65536, 10
dev = torch.device("cuda")
gs = mb.templates.grid_cluster(2, 6)
ps = mb.PatternSimulator(gs, backend="cuda-sv")
def measure(label, bufs, n=50):
    for i in range(3): ps.run_batch(bufs[i % len(bufs)], copy=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): ps.run_batch(bufs[i % len(bufs)], copy=False)
    torch.cuda.synchronize(); print(f"{label}: {(time.perf_counter()-t0)/n*1e6:8.1f} us/step", flush=True)
h1 = [torch.empty((B, T), dtype=torch.float64).pin_memory().uniform_(0, 6.28)]
measure("1 pinned buffer", h1)
h4 = [torch.empty((B, T), dtype=torch.float64).pin_memory().uniform_(0, 6.28) for _ in range(4)]
measure("4 pinned buffers", h4)
big = torch.rand((21, B, T), device=dev, dtype=torch.float64)
outs = torch.empty((21, B, 4), dtype=torch.complex128, device=dev)
measure("after 190 MiB device pool", h4)
h4b = [torch.empty((B, T), dtype=torch.float64).pin_memory() for _ in range(4)]
for j, h in enumerate(h4b): h.copy_(big[j].cpu())
measure("buffers filled via copy_ from .cpu()", h4b)
g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
x = torch.zeros(10, device=dev)
with torch.cuda.graph(g, stream=s):
    x += 1
g.replay(); torch.cuda.synchronize()
measure("after a CUDA graph capture", h4)
import threading
try:
    import pynvml; pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    stop = threading.Event()
    def loop():
        while not stop.is_set():
            pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM); time.sleep(0.002)
    th = threading.Thread(target=loop, daemon=True); th.start()
    measure("while an NVML sampler thread runs", h4)
    stop.set(); th.join()
    measure("after the sampler thread stopped", h4)
except Exception as e:
    print("nvml", e)
