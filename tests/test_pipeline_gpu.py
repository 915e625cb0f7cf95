"""GPU: the asynchronous frame pipeline (flame_b200_process_frame_submit / _wait: frames submitted to
consecutive slots are uploaded on their own streams and solved wave by wave) returns, for every frame,
exactly what the CPU oracle computes for it - bit for bit - whatever the ring size / wave split."""
import numpy as np
import pytest

from flame_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _want(oracle, g, iters):
    alpha, beta = oracle.edge_weights(g["edges"], g["pos"])
    s = synth.init_state(g, alpha, beta)
    oracle.pd_run(s, oracle.pd_default_params(), iters)
    img = oracle.interpolate_mesh(g["tris"], g["pos"], s["x"], g["H"], g["W"])
    return s, img


def _same_img(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(
        a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32))


@pytest.mark.parametrize("ring", [1, 5, 16])
def test_pipeline_matches_oracle(oracle, ring):
    iters = 15
    graphs = [synth.make_graph(640, 480, 300 + 90 * i, seed=300 + i, noise_seed=400 + i, triangulator="scipy")
              for i in range(4)]
    wants = [_want(oracle, g, iters) for g in graphs]
    H, W = 480, 640
    p = _lib.default_pd_params()
    n_frames = 23
    with _lib.Context(0, ring) as ctx:
        outs = [_lib.PinnedArray((H, W)) for _ in range(ring)]
        busy = [None] * ring
        checked = 0

        def collect(slot):
            nonlocal checked
            gi = busy[slot]
            V = graphs[gi]["pos"].shape[0]
            x, w1, w2 = (np.empty(V, np.float32) for _ in range(3))
            ctx.process_frame_wait(slot, x, w1, w2)
            s, img = wants[gi]
            for got, k in ((x, "x"), (w1, "w1"), (w2, "w2")):
                assert np.array_equal(got.view(np.uint32), s[k].view(np.uint32)), (slot, gi, k)
            assert _same_img(outs[slot].array, img), (slot, gi)
            busy[slot] = None
            checked += 1

        for f in range(n_frames):
            slot = f % ring
            if busy[slot] is not None:
                collect(slot)
            g = graphs[f % 4]
            ctx.process_frame_submit(slot, g["pos"], g["data_term"], g["edges"], g["tris"], H, W, p, iters,
                                     outs[slot].array)
            busy[slot] = f % 4
        for slot in range(ring):
            if busy[slot] is not None:
                collect(slot)
        assert checked == n_frames
        st = ctx.stats()
        assert st["frames"] == n_frames and st["d2h_bytes"] >= n_frames * H * W * 4
        # the slot keeps the frame: the ordinary calls work on it afterwards
        V, E = graphs[(n_frames - 1) % 4]["pos"].shape[0], graphs[(n_frames - 1) % 4]["edges"].shape[0]
        got = ctx.graph_download((n_frames - 1) % ring, V, E)
        assert np.array_equal(got["x"].view(np.uint32), wants[(n_frames - 1) % 4][0]["x"].view(np.uint32))
        for o in outs:
            o.free()


def test_pipeline_errors_and_mixed_use(oracle):
    g = synth.make_graph(640, 480, 400, seed=7, noise_seed=8, triangulator="scipy")
    p = _lib.default_pd_params()
    with _lib.Context(0, 4) as ctx:
        out = _lib.PinnedArray((480, 640))
        with pytest.raises(_lib.FlameB200Error):
            ctx.process_frame_wait(0)  # nothing in flight
        ctx.process_frame_submit(0, g["pos"], g["data_term"], g["edges"], g["tris"], 480, 640, p, 5, out.array)
        with pytest.raises(_lib.FlameB200Error):  # one frame in flight per slot
            ctx.process_frame_submit(0, g["pos"], g["data_term"], g["edges"], g["tris"], 480, 640, p, 5, out.array)
        # different parameters in the next slot: a new wave, both results correct
        p2 = _lib.PDParams(0.2, 0.002, 100.0, 0.3, 0.0, 5.0)
        out2 = _lib.PinnedArray((480, 640))
        ctx.process_frame_submit(1, g["pos"], g["data_term"], g["edges"], g["tris"], 480, 640, p2, 7, out2.array)
        x0, x1 = np.empty(400, np.float32), np.empty(400, np.float32)
        ctx.process_frame_wait(1, x1)
        ctx.process_frame_wait(0, x0)
        alpha, beta = oracle.edge_weights(g["edges"], g["pos"])
        for x, params, n in ((x0, oracle.pd_default_params(), 5), (x1, None, 7)):
            s = synth.init_state(g, alpha, beta)
            if params is None:
                params = oracle.pd_default_params()
                params.data_factor, params.step_x, params.step_q, params.theta, params.x_min, params.x_max = (
                    0.2, 0.002, 100.0, 0.3, 0.0, 5.0)
            oracle.pd_run(s, params, n)
            assert np.array_equal(x.view(np.uint32), s["x"].view(np.uint32))
        ctx.sync()
        out.free()
        out2.free()


def test_pipeline_triangulates_when_no_topology_is_passed(oracle):
    """edges = tris = None: process_frame_submit triangulates the vertices itself (flame_b200_delaunay); the
    result equals the oracle run on the arrays process_frame_topology hands back, and those arrays are what
    flame_b200_delaunay returns for the same points."""
    iters = 12
    H, W = 480, 640
    p = _lib.default_pd_params()
    graphs = [synth.make_graph(W, H, 250 + 70 * i, seed=900 + i, noise_seed=950 + i, triangulator="scipy")
              for i in range(3)]
    with _lib.Context(0, 4) as ctx:
        outs = [_lib.PinnedArray((H, W)) for _ in range(3)]
        for i, g in enumerate(graphs):
            ctx.process_frame_submit(i, g["pos"], g["data_term"], None, None, H, W, p, iters, outs[i].array)
        for i, g in enumerate(graphs):
            V = g["pos"].shape[0]
            x, w1, w2 = (np.empty(V, np.float32) for _ in range(3))
            ctx.process_frame_wait(i, x, w1, w2)
            tris, edges = ctx.process_frame_topology(i)
            t2, e2 = _lib.delaunay(g["pos"])
            assert np.array_equal(tris, t2) and np.array_equal(edges, e2)
            g2 = dict(g, tris=tris, edges=edges)
            s, img = _want(oracle, g2, iters)
            assert np.array_equal(x.view(np.uint32), s["x"].view(np.uint32))
            assert np.array_equal(w1.view(np.uint32), s["w1"].view(np.uint32))
            assert _same_img(outs[i].array, img)
        # a slot whose frame came with the caller's arrays has no topology to hand back
        g = graphs[0]
        ctx.process_frame_submit(3, g["pos"], g["data_term"], g["edges"], g["tris"], H, W, p, iters, None)
        ctx.process_frame_wait(3)
        with pytest.raises(_lib.FlameB200Error):
            ctx.process_frame_topology(3)
        for o in outs:
            o.free()
