"""Experiment: the dense projection alone (config D meshes): cached bins vs rebuilt bins, per frame.
  python profiles/exp_raster.py     # run under gpurun
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
    cfg = os.environ.get("EXP_CFG", "D")
    W, H, V, iters = synth.CONFIGS[cfg]
    NF = 32
    frames = [synth.config_graph(cfg, frame=i % 4) for i in range(4)]
    ctx = _lib.Context(0, NF)
    for i in range(NF):
        g = frames[i % 4]
        ctx.graph_init(i, g["pos"], g["data_term"], g["edges"])
        ctx.mesh_bind(i, g["tris"], H, W)
    ctx.sync()
    st = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def timed(n, rebind, reps=6):
        ts = []
        for _ in range(reps):
            if rebind:
                for i in range(n):
                    ctx.mesh_bind(i, frames[i % 4]["tris"], H, W)
                ctx.sync()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            ctx.mesh_project(0, n, 1.0)
            b.record(st)
            ctx.sync()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts[1:]))

    for n in (1, 8, 32):
        tc = timed(n, False)
        tr = timed(n, True)
        print("config %s  n=%2d  cached bins %.3f ms (%.1f us/frame)   rebuilt bins %.3f ms (%.1f us/frame)"
              % (cfg, n, tc, 1e3 * tc / n, tr, 1e3 * tr / n), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
