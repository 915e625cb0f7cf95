"""Experiment: raw PCIe ceilings of the box (pinned host memory), the bound of the e2e leg.
  python profiles/exp_pcie.py     # run under gpurun
"""
import time
import torch

dev = torch.device("cuda", 0)
for mb in (8.3, 33.2):
    n = int(mb * 1e6) // 4
    src = torch.empty(n, dtype=torch.float32, device=dev).normal_()
    for nbuf, nstream in ((1, 1), (8, 1), (8, 2), (8, 4)):
        hosts = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(nbuf)]
        streams = [torch.cuda.Stream(dev) for _ in range(nstream)]
        reps = 200
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for r in range(reps):
            with torch.cuda.stream(streams[r % nstream]):
                hosts[r % nbuf].copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        d2h = reps * n * 4 / dt / 1e9
        t0 = time.perf_counter()
        for r in range(reps):
            with torch.cuda.stream(streams[r % nstream]):
                src.copy_(hosts[r % nbuf], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h2d = reps * n * 4 / dt / 1e9
        print("%.1f MB x %d buffers, %d streams: D2H %.1f GB/s  H2D %.1f GB/s" % (mb, nbuf, nstream, d2h, h2d), flush=True)
