"""Experiment (not part of the bench contract): how many frame clusters run concurrently, and how much of
the dense projection hides behind the PD solver when two contexts (streams) alternate.

  python profiles/exp_waves.py            # prints a table; run under gpurun
"""
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
    NF = 48
    frames = [synth.config_graph("D", frame=i % 8) for i in range(NF)]
    p = _lib.default_pd_params()

    def make_ctx(n, first=0):
        ctx = _lib.Context(0, n)
        for i in range(n):
            g = frames[(first + i) % NF]
            ctx.graph_init(i, g["pos"], g["data_term"], g["edges"])
            ctx.mesh_bind(i, g["tris"], H, W)
        ctx.sync()
        return ctx, torch.cuda.ExternalStream(ctx.stream, device=dev)

    # ---- E1: PD time vs number of frames in one launch ----
    ctx, st = make_ctx(NF)
    print("E1: pd_run(200 iters) vs frames per launch")
    for n in list(range(1, 25)) + [28, 32, 35, 36, 40, 42, 48]:
        ts = []
        for rep in range(4):
            ctx.graph_reset(0, n)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            ctx.pd_run(p, iters, 0, n, 0)
            b.record(st)
            ctx.sync()
            ts.append(a.elapsed_time(b))
        t = float(np.median(ts[1:]))
        print("  n=%2d  %.3f ms  %.4f ms/frame" % (n, t, t / n))
    # raster alone
    for n in (8, 32, 48):
        ts = []
        for rep in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            ctx.mesh_project(0, n, 1.0)
            b.record(st)
            ctx.sync()
            ts.append(a.elapsed_time(b))
        print("  raster n=%2d  %.3f ms" % (n, float(np.median(ts[1:]))))
    ctx.close()

    # ---- E2: S contexts x B frames, steps interleaved across streams ----
    print("E2: S streams x B frames per stream, 8 steps each; frames/s")
    for S, B in ((1, 28), (1, 32), (1, 35), (1, 42), (2, 14), (2, 16), (2, 21), (2, 28), (2, 32), (2, 35),
                 (3, 14), (4, 7), (4, 8), (4, 14)):
        cs = [make_ctx(B, first=k * B) for k in range(S)]

        def step():
            for c, s in cs:
                c.graph_reset(0, B)
                c.pd_run(p, iters, 0, B, 0)
                c.mesh_project(0, B, 1.0)

        for _ in range(3):
            step()
        torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in cs]
        K = 8
        e0.record(cs[0][1])
        for c, s in cs[1:]:
            s.wait_event(e0)
        for _ in range(K):
            step()
        for (c, s), e in zip(cs, ends):
            e.record(s)
        torch.cuda.synchronize(dev)
        ms = max(e0.elapsed_time(e) for e in ends)
        print("  S=%d B=%2d  %.3f ms/step  %.0f frames/s" % (S, B, ms / K, S * B * K / (ms * 1e-3)))
        for c, s in cs:
            c.close()


if __name__ == "__main__":
    main()
