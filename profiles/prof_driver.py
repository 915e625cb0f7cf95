"""Small driver for ncu captures (run under gpurun): a few launches of each hot kernel at config D.
usage: python profiles/prof_driver.py [batch] [mode] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flame_b200 import _lib, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
W, H, V, _ = synth.CONFIGS["D"]
ctx = _lib.Context(0, B)
for i in range(B):
    g = synth.config_graph("D", frame=i)
    ctx.graph_init(i, g["pos"], g["data_term"], g["edges"])
    ctx.mesh_bind(i, g["tris"], H, W)
p = _lib.default_pd_params()
for rep in range(3):
    ctx.graph_reset(0, B)
    ctx.pd_run(p, iters, 0, B, mode)
    ctx.mesh_project(0, B, 1.0)
ctx.sync()
print("done", B, mode, iters, _lib.kernel_launches())
