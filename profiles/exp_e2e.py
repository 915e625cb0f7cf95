"""Experiment: what bounds the end-to-end (host buffers in/out) path? Runs the bench's e2e loop with and
without the dense-map download, for several pipeline counts. Run under gpurun."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from flame_b200 import _lib, synth
    W, H, V, iters = synth.CONFIGS["D"]
    frames = [synth.config_graph("D", frame=i % 8) for i in range(64)]
    p = _lib.default_pd_params()
    for P in (8, 16):
        for with_map in (True, False):
            ctxs = [_lib.Context(0, 1) for _ in range(P)]
            outs = [_lib.PinnedArray((H, W)) for _ in range(P)]
            vout = [tuple(np.empty(V, np.float32) for _ in range(3)) for _ in range(P)]

            def worker(k, n):
                for s in range(n):
                    for i in range(k, len(frames), P):
                        g = frames[i]
                        x, w1, w2 = vout[k]
                        ctxs[k].process_frame(0, g["pos"], g["data_term"], g["edges"], g["tris"], H, W, p, iters,
                                              outs[k].array if with_map else None, x, w1, w2)
                return 0

            pool = ThreadPoolExecutor(P)
            list(pool.map(lambda k: worker(k, 1), range(P)))
            t0 = time.perf_counter()
            n = 6
            list(pool.map(lambda k: worker(k, n), range(P)))
            dt = time.perf_counter() - t0
            print("P=%2d map_download=%d  %.0f frames/s" % (P, with_map, n * len(frames) / dt))
            for c in ctxs:
                c.close()


if __name__ == "__main__":
    main()
