"""GPU parity: libflame_b200's PD solver vs the CPU oracle (bit-exact), through the C ABI."""
import numpy as np
import pytest

from flame_b200 import _lib, synth

pytestmark = pytest.mark.gpu

STATE = ("x", "w1", "w2", "x_bar", "w1_bar", "w2_bar", "x_prev", "w1_prev", "w2_prev", "q1", "q2", "q3")


def _oracle_state(oracle, g, n_iters, params=None):
    alpha, beta = oracle.edge_weights(g["edges"], g["pos"])
    s = synth.init_state(g, alpha, beta)
    p = oracle.pd_default_params() if params is None else params
    oracle.pd_run(s, p, n_iters)
    return s


def _assert_bit_equal(got, want, names):
    for k in names:
        a, b = got[k].view(np.uint32), want[k].view(np.uint32)
        bad = np.flatnonzero(a != b)
        assert bad.size == 0, "%s differs at %d/%d entries (first %d: %r vs %r)" % (
            k, bad.size, a.size, bad[0], got[k][bad[0]], want[k][bad[0]])


@pytest.mark.parametrize("cfg", ["A", "B", "C", "D"])
@pytest.mark.parametrize("mode", [_lib.PD_STREAM, _lib.PD_RESIDENT, _lib.PD_AUTO])
def test_pd_bit_exact(gpu_ctx, oracle, cfg, mode):
    g = synth.config_graph(cfg)
    n = g["pd_iters"]
    want = _oracle_state(oracle, g, n)
    V, E = g["pos"].shape[0], g["edges"].shape[0]
    gpu_ctx.graph_init(0, g["pos"], g["data_term"], g["edges"])
    gpu_ctx.pd_run(_lib.default_pd_params(), n, mode=mode)
    got = gpu_ctx.graph_download(0, V, E)
    _assert_bit_equal(got, want, STATE + ("alpha", "beta"))


def test_pd_full_state_upload_and_resume(gpu_ctx, oracle):
    """Upload an arbitrary mid-run state (all 19 arrays), run more iterations, compare."""
    g = synth.config_graph("B")
    mid = _oracle_state(oracle, g, 7)
    want = {k: v.copy() if isinstance(v, np.ndarray) else v for k, v in mid.items()}
    oracle.pd_run(want, oracle.pd_default_params(), 13)
    gpu_ctx.graph_upload(1, mid)
    gpu_ctx.pd_run(_lib.default_pd_params(), 13, slot0=1)
    got = gpu_ctx.graph_download(1, g["pos"].shape[0], g["edges"].shape[0])
    _assert_bit_equal(got, want, STATE)


def test_pd_zero_and_one_iteration(gpu_ctx, oracle):
    g = synth.config_graph("A")
    V, E = g["pos"].shape[0], g["edges"].shape[0]
    for n in (0, 1):
        want = _oracle_state(oracle, g, n)
        gpu_ctx.graph_init(0, g["pos"], g["data_term"], g["edges"])
        gpu_ctx.pd_run(_lib.default_pd_params(), n)
        got = gpu_ctx.graph_download(0, V, E)
        _assert_bit_equal(got, want, STATE)


def test_pd_nondefault_params_weights_and_clamp(gpu_ctx, oracle):
    g = synth.make_graph(640, 480, 800, seed=11, noise_seed=12, triangulator="scipy")
    rng = np.random.default_rng(5)
    g["data_weight"] = rng.uniform(0.2, 5.0, g["pos"].shape[0]).astype(np.float32)
    g["data_term"][::7] = 9.99  # pushes x against x_max
    g["data_term"][3::11] = 0.0
    alpha, beta = oracle.edge_weights(g["edges"], g["pos"])
    beta = rng.uniform(0.5, 2.0, beta.shape[0]).astype(np.float32)
    s = synth.init_state(g, alpha, beta)
    p = oracle.pd_default_params()
    p.data_factor, p.step_x, p.step_q, p.theta, p.x_min, p.x_max = 0.35, 0.004, 60.0, 0.5, 0.05, 4.0
    want = {k: v.copy() if isinstance(v, np.ndarray) else v for k, v in s.items()}
    oracle.pd_run(want, p, 40)
    gp = _lib.PDParams(0.35, 0.004, 60.0, 0.5, 0.05, 4.0)
    for mode in (_lib.PD_STREAM, _lib.PD_RESIDENT):
        gpu_ctx.graph_upload(2, s)
        gpu_ctx.pd_run(gp, 40, slot0=2, mode=mode)
        got = gpu_ctx.graph_download(2, g["pos"].shape[0], g["edges"].shape[0])
        _assert_bit_equal(got, want, STATE)


def test_pd_batch_of_different_graphs(gpu_ctx, oracle):
    """Slots 0..5 hold different graphs (different V/E); one batched launch per mode."""
    graphs = [synth.make_graph(640, 480, 300 + 170 * i, seed=20 + i, noise_seed=40 + i,
                               triangulator="scipy") for i in range(6)]
    wants = [_oracle_state(oracle, g, 25) for g in graphs]
    for mode in (_lib.PD_STREAM, _lib.PD_RESIDENT):
        for i, g in enumerate(graphs):
            gpu_ctx.graph_init(i, g["pos"], g["data_term"], g["edges"])
        gpu_ctx.pd_run(_lib.default_pd_params(), 25, slot0=0, n_slots=6, mode=mode)
        for i, g in enumerate(graphs):
            got = gpu_ctx.graph_download(i, g["pos"].shape[0], g["edges"].shape[0])
            _assert_bit_equal(got, wants[i], STATE)


def test_pd_cost(gpu_ctx, oracle):
    g = synth.config_graph("C")
    want = _oracle_state(oracle, g, 30)
    ws, wd = oracle.pd_costs(want, oracle.pd_default_params())
    gpu_ctx.graph_init(0, g["pos"], g["data_term"], g["edges"])
    gpu_ctx.pd_run(_lib.default_pd_params(), 30)
    s, d = gpu_ctx.pd_cost(0, _lib.default_pd_params(), exact_order=True)
    assert np.float32(s) == np.float32(ws) and np.float32(d) == np.float32(wd)
    s, d = gpu_ctx.pd_cost(0, _lib.default_pd_params(), exact_order=False)
    assert abs(s - ws) <= 1e-4 * abs(ws) and abs(d - wd) <= 1e-4 * abs(wd)


def test_pd_energy_decreases_and_fixed_point(gpu_ctx):
    """Properties at full size (config D): the minimised objective J = NLTGV2(x) + lambda*|x - data|_1
    (= smoothnessCost/lambda + lambda*dataCost with the reference's cost definitions, cc:51-85)
    drops well below its initial value after 200 iterations; a constant data term is a fixed point
    (x stays, q stays 0)."""
    g = synth.config_graph("D")
    V, E = g["pos"].shape[0], g["edges"].shape[0]
    p = _lib.default_pd_params()
    gpu_ctx.graph_init(0, g["pos"], g["data_term"], g["edges"])
    s0, d0 = gpu_ctx.pd_cost(0, p)
    gpu_ctx.pd_run(p, 200)
    s1, d1 = gpu_ctx.pd_cost(0, p)
    lam = p.data_factor
    assert d0 == 0.0 and s1 / lam + lam * d1 < 0.5 * (s0 / lam) and s1 < 0.5 * s0
    const = np.full(V, 0.7, np.float32)
    gpu_ctx.graph_init(0, g["pos"], const, g["edges"])
    gpu_ctx.pd_run(p, 50)
    got = gpu_ctx.graph_download(0, V, E)
    assert np.all(got["x"] == np.float32(0.7)) and not got["q1"].any() and not got["w1"].any()


def test_pd_errors(gpu_ctx):
    with pytest.raises(_lib.FlameB200Error):
        gpu_ctx.pd_run(_lib.default_pd_params(), 1, slot0=7, n_slots=3)  # out of range
    g = synth.config_graph("A")
    bad = g["edges"].copy()
    bad[0, 0] = 10 ** 6
    with pytest.raises(_lib.FlameB200Error):
        gpu_ctx.graph_init(0, g["pos"], g["data_term"], bad)


def test_cpp_adapter_against_reference_code():
    """include/flame_b200_adapter.hpp (the reference's call signatures over the C ABI) vs the
    reference's own dgraph::step / DrawShadedTriangleBarycentric compiled verbatim, in one C++ binary."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "adapter_test")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/adapter_test not built (needs /root/reference at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ADAPTER_OK" in out.stdout, out.stdout + out.stderr


def test_pd_tiny_graphs(gpu_ctx, oracle):
    """Graphs far smaller than a CTA (3 vertices / 2 edges, an isolated vertex, a self-contained
    triangle) in both modes."""
    cases = [
        (np.array([[0, 0], [10, 3], [20, 12]], np.float32), np.array([[0, 1], [2, 1]], np.int32)),
        (np.array([[5, 5], [9, 1], [2, 8], [30, 30]], np.float32), np.array([[0, 1], [1, 2], [2, 0]], np.int32)),
        (np.array([[1, 1], [4, 4]], np.float32), np.array([[1, 0]], np.int32)),
    ]
    for pos, edges in cases:
        V = pos.shape[0]
        g = {"pos": pos, "edges": edges, "data_term": (0.4 + 0.1 * np.arange(V)).astype(np.float32),
             "data_weight": np.ones(V, np.float32)}
        alpha, beta = oracle.edge_weights(edges, pos)
        want = synth.init_state(g, alpha, beta)
        oracle.pd_run(want, oracle.pd_default_params(), 9)
        for mode in (_lib.PD_STREAM, _lib.PD_RESIDENT):
            gpu_ctx.graph_init(3, pos, g["data_term"], edges)
            gpu_ctx.pd_run(_lib.default_pd_params(), 9, slot0=3, mode=mode)
            got = gpu_ctx.graph_download(3, V, edges.shape[0])
            _assert_bit_equal(got, want, STATE)


def test_pd_empty_and_edgeless_graphs(gpu_ctx, oracle):
    """V = 0 and E = 0 (isolated vertices only: prox + extragradient with no coupling)."""
    gpu_ctx.graph_init(5, np.zeros((0, 2), np.float32), np.zeros(0, np.float32), np.zeros((0, 2), np.int32))
    gpu_ctx.pd_run(_lib.default_pd_params(), 5, slot0=5)
    assert gpu_ctx.graph_download(5, 0, 0)["x"].shape == (0,)
    pos = np.array([[3, 4], [50, 60], [7, 90]], np.float32)
    data = np.array([0.2, 11.0, -1.0], np.float32)  # outside [x_min, x_max]: clamped by the prox
    g = {"pos": pos, "edges": np.zeros((0, 2), np.int32), "data_term": data, "data_weight": np.ones(3, np.float32)}
    want = synth.init_state(g, np.zeros(0, np.float32), np.zeros(0, np.float32))
    oracle.pd_run(want, oracle.pd_default_params(), 4)
    for mode in (_lib.PD_STREAM, _lib.PD_RESIDENT):
        gpu_ctx.graph_init(5, pos, data, g["edges"])
        gpu_ctx.pd_run(_lib.default_pd_params(), 4, slot0=5, mode=mode)
        got = gpu_ctx.graph_download(5, 3, 0)
        _assert_bit_equal(got, want, ("x", "w1", "w2", "x_bar", "w1_bar", "w2_bar", "x_prev"))


def test_pd_graph_larger_than_a_cluster_streams(gpu_ctx, oracle):
    """A graph beyond the resident kernel's capacity (V = 40k): AUTO must fall back to the streaming
    kernel and stay bit-exact; RESIDENT must refuse loudly."""
    g = synth.make_graph(2560, 1440, 40000, seed=3, noise_seed=4, triangulator="scipy")
    V, E = g["pos"].shape[0], g["edges"].shape[0]
    alpha, beta = oracle.edge_weights(g["edges"], g["pos"])
    want = synth.init_state(g, alpha, beta)
    oracle.pd_run(want, oracle.pd_default_params(), 12)
    gpu_ctx.graph_init(6, g["pos"], g["data_term"], g["edges"])
    with pytest.raises(_lib.FlameB200Error):
        gpu_ctx.pd_run(_lib.default_pd_params(), 1, slot0=6, mode=_lib.PD_RESIDENT)
    gpu_ctx.graph_init(6, g["pos"], g["data_term"], g["edges"])
    gpu_ctx.pd_run(_lib.default_pd_params(), 12, slot0=6, mode=_lib.PD_AUTO)
    got = gpu_ctx.graph_download(6, V, E)
    _assert_bit_equal(got, want, STATE)
    # dense projection at 2560x1440 (larger than any BASELINE config; still < 2896 px per side)
    img_w = oracle.interpolate_mesh(g["tris"], g["pos"], want["x"], 1440, 2560)
    gpu_ctx.mesh_bind(6, g["tris"], 1440, 2560)
    gpu_ctx.mesh_project(6, 1, 1.0)
    out = np.empty((1440, 2560), np.float32)
    gpu_ctx.mesh_download(6, out)
    gpu_ctx.sync()
    assert np.array_equal(np.isnan(out), np.isnan(img_w))
    m = ~np.isnan(out)
    assert np.array_equal(out[m].view(np.uint32), img_w[m].view(np.uint32))


@pytest.mark.parametrize("cfg", ["A", "B", "D"])
@pytest.mark.parametrize("pref", [_lib.PREF_THROUGHPUT, _lib.PREF_LATENCY])
def test_pd_cluster_preference(gpu_ctx, oracle, cfg, pref):
    """Throughput layout (smallest cluster: 9 CTAs x 1024 threads for config D) and latency layout
    (16 CTAs x 640 threads) give the same bits."""
    g = synth.config_graph(cfg)
    n = 40
    want = _oracle_state(oracle, g, n)
    V, E = g["pos"].shape[0], g["edges"].shape[0]
    gpu_ctx.set_pd_preference(pref)
    try:
        gpu_ctx.graph_init(0, g["pos"], g["data_term"], g["edges"])
        gpu_ctx.pd_run(_lib.default_pd_params(), n, mode=_lib.PD_RESIDENT)
        got = gpu_ctx.graph_download(0, V, E)
    finally:
        gpu_ctx.set_pd_preference(_lib.PREF_THROUGHPUT)
    _assert_bit_equal(got, want, STATE)


@pytest.mark.parametrize("V", [1400, 3000, 5600, 7000])
def test_pd_cluster_sizes_not_powers_of_two(gpu_ctx, oracle, V):
    """The planner takes the smallest cluster that holds the graph (any size, not only powers of two)."""
    g = synth.make_graph(1280, 720, V, seed=70 + V, noise_seed=71 + V, triangulator="scipy")
    want = _oracle_state(oracle, g, 30)
    E = g["edges"].shape[0]
    for pref in (_lib.PREF_THROUGHPUT, _lib.PREF_LATENCY):
        gpu_ctx.set_pd_preference(pref)
        try:
            gpu_ctx.graph_init(6, g["pos"], g["data_term"], g["edges"])
            gpu_ctx.pd_run(_lib.default_pd_params(), 30, slot0=6, mode=_lib.PD_RESIDENT)
            got = gpu_ctx.graph_download(6, V, E)
        finally:
            gpu_ctx.set_pd_preference(_lib.PREF_THROUGHPUT)
        _assert_bit_equal(got, want, STATE)


def test_pd_layout_with_multi_edges_and_self_loops(gpu_ctx, oracle):
    """The device-built incidence layout orders a vertex's incidences by edge id: a random multigraph with
    parallel edges (same and reversed orientation) keeps the reference's summation order."""
    rng = np.random.default_rng(77)
    V = 700
    pos = rng.uniform(4, 600, (V, 2)).astype(np.float32)
    edges = rng.integers(0, V, (2600, 2)).astype(np.int32)
    edges[5] = edges[4]          # parallel edge, same orientation
    edges[9] = edges[8][::-1]    # parallel edge, reversed
    keep = edges[:, 0] != edges[:, 1]
    edges = np.ascontiguousarray(edges[keep])
    g = {"pos": pos, "edges": edges, "data_term": rng.uniform(0.1, 2.0, V).astype(np.float32),
         "data_weight": np.ones(V, np.float32)}
    alpha, beta = oracle.edge_weights(edges, pos)
    want = synth.init_state(g, alpha, beta)
    oracle.pd_run(want, oracle.pd_default_params(), 15)
    for mode in (_lib.PD_STREAM, _lib.PD_RESIDENT):
        gpu_ctx.graph_init(3, pos, g["data_term"], edges)
        gpu_ctx.pd_run(_lib.default_pd_params(), 15, slot0=3, mode=mode)
        got = gpu_ctx.graph_download(3, V, edges.shape[0])
        _assert_bit_equal(got, want, STATE)


def test_pd_vertex_degree_beyond_the_resident_limit(gpu_ctx, oracle):
    """A hub vertex with more than 64 incident edges (FB_RESIDENT_MAX_DEG): AUTO takes the streaming kernel,
    RESIDENT refuses, results stay bit-exact."""
    rng = np.random.default_rng(12)
    V = 300
    pos = rng.uniform(4, 600, (V, 2)).astype(np.float32)
    hub = np.stack([np.zeros(90, np.int32), np.arange(1, 91, dtype=np.int32)], 1)
    hub[::2] = hub[::2, ::-1]  # the hub is source of some, target of others
    ring = np.stack([np.arange(1, V - 1, dtype=np.int32), np.arange(2, V, dtype=np.int32)], 1)
    edges = np.ascontiguousarray(np.concatenate([hub, ring]).astype(np.int32))
    g = {"pos": pos, "edges": edges, "data_term": rng.uniform(0.1, 2.0, V).astype(np.float32),
         "data_weight": np.ones(V, np.float32)}
    alpha, beta = oracle.edge_weights(edges, pos)
    want = synth.init_state(g, alpha, beta)
    oracle.pd_run(want, oracle.pd_default_params(), 20)
    gpu_ctx.graph_init(4, pos, g["data_term"], edges)
    with pytest.raises(_lib.FlameB200Error):
        gpu_ctx.pd_run(_lib.default_pd_params(), 1, slot0=4, mode=_lib.PD_RESIDENT)
    gpu_ctx.graph_init(4, pos, g["data_term"], edges)
    gpu_ctx.pd_run(_lib.default_pd_params(), 20, slot0=4, mode=_lib.PD_AUTO)
    _assert_bit_equal(gpu_ctx.graph_download(4, V, edges.shape[0]), want, STATE)


@pytest.mark.parametrize("V,cluster", [(260, None), (2600, None), (6000, "13")])
def test_pd_hub_vertices_inside_the_resident_limit(gpu_ctx, oracle, V, cluster, monkeypatch):
    """Hubs of degree 40-60 (all-source, all-target and mixed) on top of a sparse random graph, solved by the
    resident kernel: many owned-edge ranks with one vertex each, record blocks of very different sizes, source
    groups whose supply of target groups is lopsided - the rank-major list and the balanced dealing must not
    change a bit. (cluster = a forced cluster size: chunks that cut through the hubs' neighbourhoods.)"""
    if cluster:
        monkeypatch.setenv("FLAME_B200_CLUSTER", cluster)
    rng = np.random.default_rng(V)
    pos = rng.uniform(4, 900, (V, 2)).astype(np.float32)
    hubs = []
    for h, (deg, flip) in enumerate([(60, "none"), (55, "all"), (48, "half"), (40, "none")]):
        nb = rng.choice(np.arange(8, V), deg, replace=False).astype(np.int32)
        e = np.stack([np.full(deg, h, np.int32), nb], 1)
        if flip == "all":
            e = e[:, ::-1]
        elif flip == "half":
            e[::2] = e[::2, ::-1]
        hubs.append(e)
    sparse = rng.integers(8, V, (2 * V, 2)).astype(np.int32)
    sparse = sparse[sparse[:, 0] != sparse[:, 1]]
    edges = np.concatenate(hubs + [sparse])
    edges = np.ascontiguousarray(edges[rng.permutation(edges.shape[0])])
    deg = np.bincount(edges.ravel(), minlength=V)
    assert 40 <= deg.max() <= 64
    g = {"pos": pos, "edges": edges, "data_term": rng.uniform(0.1, 2.0, V).astype(np.float32),
         "data_weight": np.ones(V, np.float32)}
    alpha, beta = oracle.edge_weights(edges, pos)
    want = synth.init_state(g, alpha, beta)
    oracle.pd_run(want, oracle.pd_default_params(), 12)
    gpu_ctx.graph_init(5, pos, g["data_term"], edges)
    gpu_ctx.pd_run(_lib.default_pd_params(), 12, slot0=5, mode=_lib.PD_RESIDENT)
    _assert_bit_equal(gpu_ctx.graph_download(5, V, edges.shape[0]), want, STATE)
