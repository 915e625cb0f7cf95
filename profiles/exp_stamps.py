"""Experiment: where the resident solver spends its time outside the iterations. Needs a library built with
FLAME_B200_NVCC_EXTRA=-DFB_PROFILE_LOAD (globaltimer stamps of block 0, flame_b200_debug_stamps).

  python profiles/exp_stamps.py [cluster]        # run under gpurun
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FLAME_B200_CLUSTER"] = sys.argv[1] if len(sys.argv) > 1 else "9"
import torch  # noqa: F401 (loads the CUDA runtime the library links against)
from flame_b200 import _lib, synth
ctx = _lib.Context(0, 15)
for i in range(15):
    g = synth.config_graph("D", frame=i % 6)
    ctx.graph_init(i, g["pos"], g["data_term"], g["edges"])
ctx.sync()
p = _lib.default_pd_params()
names = ["koff+zero", "vertex regs", "count+scan", "vertex-major fill", "rank-major", "balanced dealing", "edge setup", "p1 fill+cluster sync", "200 iterations", "write back"]
for rep in range(3):
    ctx.graph_reset(0, 15)
    ctx.pd_run(p, 200, 0, 15, _lib.PD_RESIDENT)
    ctx.sync()
    out = (ctypes.c_ulonglong * 32)()
    _lib.lib().flame_b200_debug_stamps(out)
    t = [out[i] for i in range(10)]
    print(" | ".join("%s %.1f" % (names[i + 1], (t[i + 1] - t[i]) / 1e3) for i in range(9)), "| total %.1f us" % ((t[9] - t[0]) / 1e3))
