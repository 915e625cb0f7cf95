"""Experiment: does pinning the process to the GPU's local CPUs (NUMA node) before allocating the pinned
buffers raise the end-to-end (PCIe D2H-bound) rate? Run under gpurun."""
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gpu_local_cpus(index=0):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    n = os.cpu_count()
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
    cpus = [i for i in range(n) if (words[i // 64] >> (i % 64)) & 1]
    return cpus


def run(P, label):
    from flame_b200 import _lib, synth
    W, H, V, iters = synth.CONFIGS["D"]
    frames = [synth.config_graph("D", frame=i % 8) for i in range(64)]
    p = _lib.default_pd_params()
    ctxs = [_lib.Context(0, 1) for _ in range(P)]
    outs = [_lib.PinnedArray((H, W)) for _ in range(P)]
    vout = [tuple(np.empty(V, np.float32) for _ in range(3)) for _ in range(P)]

    def worker(k, n):
        for s in range(n):
            for i in range(k, len(frames), P):
                g = frames[i]
                x, w1, w2 = vout[k]
                ctxs[k].process_frame(0, g["pos"], g["data_term"], g["edges"], g["tris"], H, W, p, iters,
                                      outs[k].array, x, w1, w2)

    pool = ThreadPoolExecutor(P)
    list(pool.map(lambda k: worker(k, 1), range(P)))
    t0 = time.perf_counter()
    n = 6
    list(pool.map(lambda k: worker(k, n), range(P)))
    dt = time.perf_counter() - t0
    print("%s P=%d  %.0f frames/s  (%.1f GB/s D2H)" % (label, P, n * len(frames) / dt, n * len(frames) / dt * 4 * W * H / 1e9))
    for c in ctxs:
        c.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pin":
        cpus = gpu_local_cpus(0)
        print("GPU0 local cpus: %d..%d (%d)" % (cpus[0], cpus[-1], len(cpus)))
        os.sched_setaffinity(0, cpus)
        run(16, "pinned-to-gpu-node")
    else:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[-1200:])
        run(16, "default-affinity")
