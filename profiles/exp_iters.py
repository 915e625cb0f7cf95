"""Experiment: per-iteration cost of the resident solver separated from its load / write-back phases.

  python profiles/exp_iters.py [cluster] [variant]        # run under gpurun

Times one wave at 200 and at 1000 iterations; slope = time per iteration, intercept = load + write-back."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from flame_b200 import _lib, synth
    if len(sys.argv) > 1:
        os.environ["FLAME_B200_CLUSTER"] = sys.argv[1]
    if len(sys.argv) > 2:
        os.environ["FLAME_B200_VARIANT"] = sys.argv[2]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    NF = 30
    frames = [synth.config_graph("D", frame=i % 6) for i in range(6)]
    p = _lib.default_pd_params()
    ctx = _lib.Context(0, NF)
    for i in range(NF):
        g = frames[i % 6]
        ctx.graph_init(i, g["pos"], g["data_term"], g["edges"])
    ctx.sync()
    st = torch.cuda.ExternalStream(ctx.stream, device=dev)
    wave = ctx.pd_wave(0)

    def timed(n, iters, reps=7):
        ts = []
        for _ in range(reps):
            ctx.graph_reset(0, n)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            ctx.pd_run(p, iters, 0, n, _lib.PD_RESIDENT)
            b.record(st)
            ctx.sync()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts[1:]))

    for n in (1, wave):
        t200, t1000 = timed(n, 200), timed(n, 1000)
        per_it = (t1000 - t200) / 800.0
        print("frames=%2d  200 it %.3f ms  1000 it %.3f ms  -> %.3f us/iteration, load+store %.1f us" % (
            n, t200, t1000, per_it * 1e3, (t200 - 200 * per_it) * 1e3))


if __name__ == "__main__":
    main()
