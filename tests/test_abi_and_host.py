"""CPU-only checks: the C ABI library loads and exports every symbol include/flame_b200.h declares;
host-side logic (synthetic generator, oracle properties, struct layouts); and the product path
fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from flame_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "flame_b200.h")).read()
    declared = set(re.findall(r"\b(flame_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    L = _lib.lib()
    for s in declared:
        assert getattr(L, s) is not None


def test_version_and_default_params():
    assert _lib.lib().flame_b200_version() == 100
    p = _lib.default_pd_params()
    # nltgv2_l1_graph_regularizer.h:121-129
    assert (p.data_factor, p.step_x, p.step_q, p.theta, p.x_min, p.x_max) == (
        np.float32(0.1), np.float32(0.001), 125.0, 0.25, 0.0, 10.0)
    t = _lib.default_track_params()
    assert t.fparams.win_size == 5 and t.fparams.sparams.max_cost == 1300.0
    assert t.idepth_var_max == 0.25 and t.max_dropouts == 5 and t.outlier_sigma_thresh == 3.0


def test_struct_layouts_match_oracle_and_numpy():
    from oracle import pyoracle
    assert ctypes.sizeof(_lib.TrackParams) == ctypes.sizeof(pyoracle.TrackParams)
    assert ctypes.sizeof(_lib.PDParams) == 24
    assert _lib.FEATURE_DTYPE.itemsize == 40 and _lib.DIAG_DTYPE.itemsize == 20
    assert _lib.FEATURE_DTYPE == pyoracle.FEATURE_DTYPE == synth.FEATURE_DTYPE
    assert ctypes.sizeof(_lib.TriFilterParams) == ctypes.sizeof(pyoracle.TriFilterParams) == 32
    assert ctypes.sizeof(_lib.GraphSync) == 72  # 2 x int32 + 7 pointers + int32 (+ padding)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.FlameB200Error) as e:
        _lib.Context(0, 1)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "flame_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".sh")):
                src = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in src and "flame_oracle" not in src and "libflame_ref" not in src, f


def test_synthetic_graph_is_a_valid_triangulation():
    for cfg in ("A", "B"):
        g = synth.config_graph(cfg)
        V, E, T = g["pos"].shape[0], g["edges"].shape[0], g["tris"].shape[0]
        assert g["source"] == "reference-triangle-fixture"
        assert E == V + T - 1  # Euler for a triangulated convex hull: V - E + (T + 1) = 2
        assert g["edges"].min() == 0 and g["edges"].max() == V - 1
        # every edge of every triangle is in the edge list (as an unordered pair)
        es = set(map(tuple, np.sort(g["edges"], 1)))
        for a, b in ((0, 1), (1, 2), (2, 0)):
            te = np.sort(g["tris"][:, [a, b]], 1)
            assert all(tuple(x) in es for x in te[:200])
    g2 = synth.make_graph(320, 240, 200, seed=5, triangulator="scipy")
    assert g2["edges"].shape[0] == 200 + g2["tris"].shape[0] - 1


def test_oracle_pd_properties(oracle):
    """Reference has no PD test: pin the restatement with properties — constant data is a
    fixed point; the objective decreases; planar data drives w to the plane gradient."""
    g = synth.config_graph("A")
    alpha, beta = oracle.edge_weights(g["edges"], g["pos"])
    p = oracle.pd_default_params()
    s = synth.init_state(dict(g, data_term=np.full(g["pos"].shape[0], 0.5, np.float32)), alpha, beta)
    oracle.pd_run(s, p, 30)
    assert np.all(s["x"] == np.float32(0.5)) and not s["q1"].any()
    s = synth.init_state(g, alpha, beta)
    s0, d0 = oracle.pd_costs(s, p)
    oracle.pd_run(s, p, 50)
    s1, d1 = oracle.pd_costs(s, p)
    lam = p.data_factor
    assert s1 / lam + lam * d1 < s0 / lam + lam * d0
    plane = (0.3 + 0.0004 * g["pos"][:, 0] - 0.0002 * g["pos"][:, 1]).astype(np.float32)
    s = synth.init_state(dict(g, data_term=plane), alpha, beta)
    oracle.pd_run(s, p, 3000)
    assert abs(np.median(s["w1"]) - 0.0004) < 1e-4 and abs(np.median(s["w2"]) + 0.0002) < 1e-4


def test_oracle_edge_weights(oracle):
    pos = np.array([[0, 0], [3, 4], [6, 8]], np.float32)
    edges = np.array([[0, 1], [2, 0]], np.int32)
    a, b = oracle.edge_weights(edges, pos)
    assert np.allclose(a, [0.2, 0.1]) and np.all(b == 1)


def test_sync_topology_bookkeeping():
    """synth.sync_topology mirrors Flame::syncGraph's index bookkeeping (flame.cc:1984-2123): kept vertices
    first in their old order, surviving edges keep order and orientation, new edges are appended."""
    import numpy as np
    from flame_b200 import synth
    g = synth.make_graph(640, 480, 400, seed=3, noise_seed=4, triangulator="scipy")
    rng = np.random.default_rng(0)
    keep = rng.uniform(size=400) > 0.2
    pos_kept = (g["pos"][keep] + rng.normal(0, 0.3, (int(keep.sum()), 2))).astype(np.float32)
    pos_new = synth.make_vertices(640, 480, 50, seed=9)
    pos, tris, edges, vtx_from, edge_from = synth.sync_topology(g["edges"], keep, pos_kept, pos_new)
    nk = int(keep.sum())
    assert pos.shape[0] == nk + 50 and np.array_equal(vtx_from[:nk], np.flatnonzero(keep)) and (vtx_from[nk:] == -1).all()
    carried = edge_from >= 0
    # carried edges: same endpoints (through vtx_from) in the same orientation, old relative order preserved
    assert np.array_equal(vtx_from[edges[carried]], g["edges"][edge_from[carried]])
    assert (np.diff(edge_from[carried]) > 0).all()
    # carried edges come first, new ones after
    assert not carried[np.argmax(~carried):].any()
    # the edge set is exactly the triangulation's
    und = {tuple(sorted(e)) for e in edges.tolist()}
    tri_e = {tuple(sorted((int(t[i]), int(t[(i + 1) % 3])))) for t in tris for i in range(3)}
    assert und == tri_e and len(und) == edges.shape[0]
    # host-side state edits: carried values move, new ones get the defaults
    old = {k: rng.uniform(size=400).astype(np.float32) for k in
           ("x", "w1", "w2", "x_bar", "w1_bar", "w2_bar", "x_prev", "w1_prev", "w2_prev")}
    old.update({k: rng.uniform(size=g["edges"].shape[0]).astype(np.float32) for k in ("q1", "q2", "q3")})
    data = rng.uniform(0.2, 1.0, pos.shape[0]).astype(np.float32)
    s = synth.sync_state_host(old, pos, edges, vtx_from, edge_from, data)
    assert np.array_equal(s["w1"][:nk], old["w1"][keep]) and (s["w1"][nk:] == 0).all()
    assert np.array_equal(s["x"][nk:], data[nk:]) and np.array_equal(s["x_prev"][nk:], data[nk:])
    assert np.array_equal(s["q2"][carried], old["q2"][edge_from[carried]]) and (s["q2"][~carried] == 0).all()


def test_graph_plan_vertex_order_and_launch_plan():
    """flame_b200_graph_plan (host only): the vertex order is a permutation, cut into Z-curve chunks that are sorted
    by (degree, owned edges) descending; koff / chunk_base address exactly the incidences that exist; and the
    launch plan of a config-D graph on a B200's cluster-occupancy table is the one DESIGN.md quotes."""
    from flame_b200 import _lib, synth
    rng = np.random.default_rng(8)
    graphs = [(synth.config_graph(c)["pos"], synth.config_graph(c)["edges"]) for c in ("A", "D")]
    V = 3000
    e = rng.integers(0, V, (3 * V, 2)).astype(np.int32)
    graphs.append((rng.uniform(0, 800, (V, 2)).astype(np.float32), np.ascontiguousarray(e[e[:, 0] != e[:, 1]])))
    for pos, edges in graphs:
        for pref in (_lib.PREF_THROUGHPUT, _lib.PREF_LATENCY):
            lay, vperm, koff, chunk_base = _lib.graph_plan(pos, edges, pref)
            V, E = pos.shape[0], edges.shape[0]
            deg = np.bincount(edges.ravel(), minlength=V)
            own = np.bincount(edges[:, 0], minlength=V)
            assert sorted(vperm.tolist()) == list(range(V))
            assert lay["max_degree"] == deg.max() and lay["n_chunks"] == -(-V // lay["chunk_width"])
            assert chunk_base[0] == 0 and chunk_base[-1] == 2 * E
            for c in range(lay["n_chunks"]):
                ids = vperm[c * lay["chunk_width"]:(c + 1) * lay["chunk_width"]]
                d, o = deg[ids], own[ids]
                assert np.all(np.diff(d) <= 0)                      # degree descending
                assert np.all((np.diff(o) <= 0) | (np.diff(d) < 0))  # owned count descending inside a degree class
                # incidence k exists for the first #(deg > k) vertices of the chunk; the blocks are packed
                starts = np.array([(d > k).sum() for k in range(lay["max_degree"] + 1)])
                assert np.array_equal(koff[c], np.concatenate([[0], np.cumsum(starts)[:-1]]))
                assert chunk_base[c + 1] - chunk_base[c] == d.sum()
            if lay["cluster"]:
                assert lay["max_incidences"] == np.diff(chunk_base).max()
                assert lay["shared_bytes"] <= 227 * 1024
    lay, _, _, _ = _lib.graph_plan(*graphs[1][:2], _lib.PREF_THROUGHPUT)
    assert (lay["cluster"], lay["variant"], lay["threads"]) == (9, 1, 576)
    lay, _, _, _ = _lib.graph_plan(*graphs[1][:2], _lib.PREF_LATENCY)
    assert (lay["cluster"], lay["variant"], lay["threads"]) == (16, 0, 640)
    # edges are validated before anything else
    with pytest.raises(_lib.FlameB200Error):
        _lib.graph_plan(graphs[0][0], np.array([[0, 10 ** 6]], np.int32))
