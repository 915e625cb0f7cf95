"""Experiment: one context, dense projection on the context's second stream (flame_b200_ctx_set_async_projection)
vs in-order, for several batch sizes. Run under gpurun. (An earlier variant also split the projection
launches into sub-batches of 4 / 7 frames to keep the owner words in L2: slower, removed.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from flame_b200 import _lib, synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    W, H, V, iters = synth.CONFIGS["D"]
    NF = 56
    frames = [synth.config_graph("D", frame=i % 8) for i in range(NF)]
    p = _lib.default_pd_params()
    ctx = _lib.Context(0, NF)
    for i in range(NF):
        g = frames[i]
        ctx.graph_init(i, g["pos"], g["data_term"], g["edges"])
        ctx.mesh_bind(i, g["tris"], H, W)
    ctx.sync()
    st = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for B in (7, 14, 28, 35, 42, 56):
        for on in (0, 1):
            ctx.set_async_projection(on)

            def step():
                ctx.graph_reset(0, B)
                ctx.pd_run(p, iters, 0, B, 0)
                ctx.mesh_project(0, B, 1.0)

            for _ in range(3):
                step()
            ctx.sync()
            K = 10
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            for _ in range(K):
                step()
            ctx.join()
            b.record(st)
            ctx.sync()
            ms = a.elapsed_time(b)
            print("  B=%2d async=%d  %.3f ms/step  %.0f frames/s" % (B, on, ms / K, B * K / (ms * 1e-3)))
    ctx.close()


if __name__ == "__main__":
    main()
