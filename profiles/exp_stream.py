"""Streaming PD kernel (state in HBM, one kernel per iteration): effective algorithmic bandwidth vs batch size.
Run under gpurun. FLAME_B200_STREAM_U selects the kernel variant (default 2)."""
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
    NF = 128
    frames = [synth.config_graph("D", frame=i % 8) for i in range(NF)]
    E = frames[0]["edges"].shape[0]
    p = _lib.default_pd_params()
    ctx = _lib.Context(0, NF)
    for i in range(NF):
        g = frames[i]
        ctx.graph_init(i, g["pos"], g["data_term"], g["edges"])
    ctx.sync()
    st = torch.cuda.ExternalStream(ctx.stream, device=dev)
    peak = 6570.0
    print("variant", os.environ.get("FLAME_B200_STREAM_U", "2 (default)"))
    for B in (32, 64, 96, 128):
        ts = []
        for rep in range(4):
            ctx.graph_reset(0, B)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            ctx.pd_run(p, iters, 0, B, _lib.PD_STREAM)
            b.record(st)
            ctx.sync()
            ts.append(a.elapsed_time(b))
        t = float(np.median(ts[1:]))
        gbs = 76.0 * (V + E) * iters * B / (t * 1e-3) / 1e9
        print("  B=%3d  %.3f ms  %.2f us/iter  effective %.0f GB/s = %.3f of %.0f" % (B, t, 1e3 * t / iters, gbs, gbs / peak, peak))
    ctx.close()


if __name__ == "__main__":
    main()
