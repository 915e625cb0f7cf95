"""Regenerates the committed fixtures under tests/golden/. Run HERE (needs /root/reference):

    python tests/golden/make_golden.py

1. graph_<W>x<H>_v<V>_s<seed>.npz — vertex positions (flame_b200.synth.make_vertices) and the index
   arrays the REFERENCE's own Shewchuk Triangle returns for them, called exactly as
   src/flame/utils/delaunay.cc:31-77 does (oracle/_ref, compiled verbatim from /root/reference).
   These pin the "index arrays bit-exact" contract: the product passes them through unchanged.
2. RasterizationTest_DrawShadedTriangleBarycentric{1,2}.png — the reference's own golden test
   vectors for the rasteriser (test/data/, used by test/utils/rasterization_test.cc:514-598),
   copied byte-for-byte as test vectors (they are data, not source).
"""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from flame_b200 import synth  # noqa: E402
from oracle import pyoracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    pyoracle.build()
    for name, (W, H, V, _) in synth.CONFIGS.items():
        pos = synth.make_vertices(W, H, V, seed=1)
        tris, edges, _ = pyoracle.ref_delaunay(pos)
        out = os.path.join(HERE, "graph_%dx%d_v%d_s1.npz" % (W, H, V))
        np.savez_compressed(out, pos=pos, tris=tris, edges=edges)
        print(name, out, "V", V, "T", tris.shape[0], "E", edges.shape[0])
    for k in (1, 2):
        fn = "RasterizationTest_DrawShadedTriangleBarycentric%d.png" % k
        shutil.copyfile(os.path.join(REF, "test", "data", fn), os.path.join(HERE, fn))
        print("copied", fn)


if __name__ == "__main__":
    main()
