"""dev helper: pinned host -> device copy rate of one bench batch (the ceiling of the e2e figure), and the e2e
pipeline at several chunk sizes."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import graphics_project_b200 as gp
H, W, B = 1080, 1920, 256
host = torch.empty((B, H, W, 3), dtype=torch.uint8, pin_memory=True)
u = gp.synth.make_batch(8, H, W)
for i in range(B): host[i] = torch.from_numpy(u[i % 8])
dev = torch.empty_like(host, device="cuda")
for _ in range(2): dev.copy_(host, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): dev.copy_(host, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("plain H2D", round(host.numel() / dt / 1e9, 2), "GB/s =", round(B / dt), "frames/s ceiling")
del dev
ol = torch.empty((B, 1024, 2), dtype=torch.float32, pin_memory=True); oc = torch.empty((B,), dtype=torch.int32, pin_memory=True)
for chunk in (8, 16, 32, 64):
    pipe = gp.pipeline.CannyHoughHost(H, W, 0, chunk=chunk)
    for _ in range(2): pipe.run(host, ol, oc)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): pipe.run(host, ol, oc)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 4
    print("chunk", chunk, round(B / dt), "frames/s")
    del pipe; torch.cuda.empty_cache()
