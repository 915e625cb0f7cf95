"""Experiment (not part of the bench contract): the resident PD solver per cluster size / kernel build.

  python profiles/exp_variants.py [C:variant:0 ...]     # run under gpurun; C = cluster size, variant = build index
                                                        # (0 <640,4,1>, 1 <640,6,2>, 2 <1024,4,2>); the third field is
                                                        # unused (it selected a relaxed second barrier in an earlier build);
                                                        # FLAME_B200_THREADS=<n> additionally forces the block size

For every configuration: frames resident at once (pd_wave), time of a 200-iteration solve of 1 frame and
of one full wave, frames/s of back-to-back waves, and a bit-comparison of the result with the first
configuration (all builds must agree bit for bit).
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
    cfgname = os.environ.get("EXP_CFG", "D")
    W, H, V, iters = synth.CONFIGS[cfgname]
    NF = 60
    frames = [synth.config_graph(cfgname, frame=i % 6) for i in range(6)]
    p = _lib.default_pd_params()
    combos = sys.argv[1:] or ["9:1:0", "16:0:0", "9:2:0", "10:1:0"]
    ref = None
    for combo in combos:
        C_, var, rel = combo.split(":")
        os.environ["FLAME_B200_CLUSTER"] = C_
        os.environ["FLAME_B200_VARIANT"] = var
        ctx = _lib.Context(0, NF)
        try:
            for i in range(NF):
                g = frames[i % 6]
                ctx.graph_init(i, g["pos"], g["data_term"], g["edges"])
            ctx.sync()
            st = torch.cuda.ExternalStream(ctx.stream, device=dev)
            wave = ctx.pd_wave(0)
            if wave <= 0:
                print("%-10s does not fit" % combo)
                continue

            def timed(n, reps=5):
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

            t1 = timed(1)
            tw = timed(min(wave, NF))
            nb = (NF // wave) * wave
            tb = timed(nb, reps=4)
            got = ctx.graph_download(0, V, frames[0]["edges"].shape[0])
            same = ""
            if ref is None:
                ref = got
            else:
                same = "bit-equal" if all(np.array_equal(got[k].view(np.uint32), ref[k].view(np.uint32))
                                          for k in ("x", "w1", "w2", "q1", "q2", "q3", "x_bar", "x_prev")) else "MISMATCH"
            print("%-10s wave=%2d  1 frame %.3f ms  wave %.3f ms (%.4f ms/frame)  %d frames %.3f ms -> %.0f frames/s  %s"
                  % (combo, wave, t1, tw, tw / min(wave, NF), nb, tb, nb / tb * 1e3, same), flush=True)
        finally:
            ctx.close()


if __name__ == "__main__":
    main()
