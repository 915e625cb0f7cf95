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

# ---- D2H under load: the resident PD solver running on another stream; many pinned buffers ----
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flame_b200 import _lib, synth
W, H, V, iters = synth.CONFIGS["D"]
g = synth.config_graph("D")
ctx = _lib.Context(0, 15)
for i in range(15):
    ctx.graph_init(i, g["pos"], g["data_term"], g["edges"])
    ctx.mesh_bind(i, g["tris"], H, W)
ctx.sync()
p = _lib.default_pd_params()
n = W * H
src = torch.empty(n, dtype=torch.float32, device=dev).normal_()
for nbuf, load in ((8, "none"), (96, "none"), (8, "solver"), (8, "solver+raster"), (96, "solver+raster")):
    hosts = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    cs = torch.cuda.Stream(dev)
    reps = 300
    torch.cuda.synchronize()
    if load != "none":
        for _ in range(60):  # ~70-90 ms of GPU work queued
            ctx.pd_run(p, iters, 0, 15, 0)
            if load == "solver+raster":
                ctx.mesh_project(0, 15, 1.0)
    t0 = time.perf_counter()
    with torch.cuda.stream(cs):
        for r in range(reps):
            hosts[r % nbuf].copy_(src, non_blocking=True)
    cs.synchronize()
    dt = time.perf_counter() - t0
    ctx.sync()
    print("D2H 8.3 MB x %d buffers, load=%s: %.1f GB/s" % (nbuf, load, reps * n * 4 / dt / 1e9), flush=True)
    del hosts
ctx.close()
