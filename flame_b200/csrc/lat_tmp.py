import sys, os, time
sys.path.insert(0,'/root/repo')
import torch
from flame_b200 import _lib, synth
W,H,V,_=synth.CONFIGS["D"]
ctx=_lib.Context(0,1)
g=synth.config_graph("D")
ctx.graph_init(0,g["pos"],g["data_term"],g["edges"])
p=_lib.default_pd_params()
st=torch.cuda.ExternalStream(ctx.stream)
for it in (200,):
    for rep in range(3):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(st); ctx.pd_run(p,it); b.record(st); ctx.sync()
    print("flags",os.environ.get("FLAME_B200_PDFLAGS","0"),"iters",it,"ms",a.elapsed_time(b), "cycles/iter", a.elapsed_time(b)*1e-3*1.965e9/it)
