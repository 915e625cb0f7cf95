"""Small driver for ncu captures of the epipolar search (run under gpurun): config C, 5 000 resident features.
usage: python profiles/prof_search.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flame_b200 import _lib, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
W, H, V, _ = synth.CONFIGS["C"]
K, Kinv, frames = synth.make_stereo_pair(W, H)
feats = synth.make_features(synth.config_graph("C")["pos"], W, H)
ctx = _lib.Context(0, 1)
ctx.set_camera(W, H, K, Kinv)
for i, fr in enumerate(frames):
    ctx.frame_upload(i, fr["img_pad"], fr["gradx_pad"], fr["grady_pad"], synth.PAD, fr["pose"])
tp = _lib.default_track_params()
ctx.features_upload(feats)
for _ in range(reps):
    ctx.features_track(tp, 1, 0)
n, f, d = ctx.features_download(commit=False)
print("done", len(feats), "features,", n, "updates,", _lib.kernel_launches(), "launches")
