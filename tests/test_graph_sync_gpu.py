"""GPU parity for the device-side graph carry-over: flame_b200_graph_project
(Flame::projectGraph's vertex loop, flame.cc:1892-1899) against the oracle's EpipolarGeometry, and
flame_b200_graph_resync (Flame::syncGraph's graph edits, flame.cc:1984-2162) against the same edits done
on the host followed by the oracle solver. Bit-exact, through the C ABI."""
import numpy as np
import pytest

from flame_b200 import _lib, synth

pytestmark = pytest.mark.gpu

STATE = ("x", "w1", "w2", "x_bar", "w1_bar", "w2_bar", "x_prev", "w1_prev", "w2_prev", "q1", "q2", "q3")


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def _assert_state(got, want, names=STATE):
    for k in names:
        assert _bits_equal(got[k], want[k]), "%s differs at %d entries" % (
            k, int((got[k].view(np.uint32) != want[k].view(np.uint32)).sum()))


def _warm_state(oracle, g, n_iters):
    alpha, beta = oracle.edge_weights(g["edges"], g["pos"])
    s = synth.init_state(g, alpha, beta)
    oracle.pd_run(s, oracle.pd_default_params(), n_iters)
    return s


def _churn(g, rng, drop_frac, n_new):
    """Drop some vertices, jitter the rest, add new ones; returns the sync_topology() tuple + data."""
    V = g["pos"].shape[0]
    keep = rng.uniform(size=V) >= drop_frac
    pos_kept = (g["pos"][keep] + rng.normal(0, 0.4, (int(keep.sum()), 2))).astype(np.float32)
    pos_new = synth.make_vertices(g["W"], g["H"], n_new, seed=int(rng.integers(1 << 30)))
    pos, tris, edges, vtx_from, edge_from = synth.sync_topology(g["edges"], keep, pos_kept, pos_new)
    data = synth.scene_idepth(pos[:, 0], pos[:, 1], g["W"], g["H"]).astype(np.float32)
    data = (data + rng.normal(0, 0.02, data.shape[0])).astype(np.float32)
    return pos, tris, edges, vtx_from, edge_from, data


def test_graph_project_matches_epipolar_geometry(gpu_ctx, oracle):
    g = synth.config_graph("B")
    W, H = g["W"], g["H"]
    V, E = g["pos"].shape[0], g["edges"].shape[0]
    K, Kinv = synth.camera(W, H)
    s = _warm_state(oracle, g, 20)
    s["x"][::53] = 0.0  # idepth == 0 -> maxDepthProjection branch (epipolar_geometry.h:156-162)
    q = synth.quat_from_yaw(0.8).astype(np.float32)
    t = np.array([0.04, -0.01, 0.03], np.float32)
    epi = oracle.Epi(K, Kinv)
    epi.load(q, t)
    for scale in (1.0, 0.7):
        gpu_ctx.set_camera(W, H, K, Kinv)
        gpu_ctx.graph_upload(0, s)
        pos_got, id_got = gpu_ctx.graph_project(0, V, q, t, scale)
        pos_want = np.empty_like(pos_got)
        id_want = np.empty_like(id_got)
        for v in range(V):
            xs = float(np.float32(s["x"][v]) * np.float32(scale))
            (pos_want[v, 0], pos_want[v, 1]), id_want[v] = epi.project_idepth(s["pos"][v], xs)
        assert _bits_equal(pos_got, pos_want)
        assert _bits_equal(id_got, id_want)
        st = gpu_ctx.graph_download(0, V, E)
        assert _bits_equal(st["x"], id_want / np.float32(scale))
        _assert_state(st, s, [k for k in STATE if k != "x"])  # x_bar / x_prev / w / q untouched
        # the slot is stale until it is re-synchronised
        with pytest.raises(_lib.FlameB200Error):
            gpu_ctx.pd_run(_lib.default_pd_params(), 1)
        ident_v = np.arange(V, dtype=np.int32)
        ident_e = np.arange(E, dtype=np.int32)
        gpu_ctx.graph_resync(0, pos_got, s["data_term"], s["edges"], ident_v, ident_e)
        gpu_ctx.pd_run(_lib.default_pd_params(), 5)
        got = gpu_ctx.graph_download(0, V, E)
        want = dict(s)
        for k in STATE:
            want[k] = s[k].copy()
        want["pos"] = pos_want
        want["x"] = id_want / np.float32(scale)
        want["alpha"], want["beta"] = oracle.edge_weights(s["edges"], pos_want)
        oracle.pd_run(want, oracle.pd_default_params(), 5)
        _assert_state(got, want, STATE + ("alpha", "beta"))


@pytest.mark.parametrize("mode", [_lib.PD_STREAM, _lib.PD_RESIDENT])
@pytest.mark.parametrize("variant", ["plain", "x_init+weights+sticky"])
def test_graph_resync_equals_host_edits(gpu_ctx, oracle, mode, variant):
    g = synth.config_graph("B")
    V0, E0 = g["pos"].shape[0], g["edges"].shape[0]
    rng = np.random.default_rng(77)
    gpu_ctx.graph_init(3, g["pos"], g["data_term"], g["edges"])
    gpu_ctx.pd_run(_lib.default_pd_params(), 30, slot0=3, mode=mode)
    old = gpu_ctx.graph_download(3, V0, E0)
    _assert_state(old, _warm_state(oracle, g, 30))

    pos, tris, edges, vtx_from, edge_from, data = _churn(g, rng, 0.15, 300)
    V, E = pos.shape[0], edges.shape[0]
    assert (vtx_from < 0).sum() == 300 and (edge_from >= 0).sum() > E // 2 and (edge_from < 0).sum() > 100
    kw = {}
    if variant != "plain":
        kw["data_weight"] = rng.uniform(0.3, 3.0, V).astype(np.float32)
        kw["x_init"] = (data + np.float32(0.05)).astype(np.float32)
        kw["check_sticky_obstacles"] = True
        data[::9] -= np.float32(0.5)  # x - data_term > 0.25 on carried vertices -> reset
    gpu_ctx.graph_resync(3, pos, data, edges, vtx_from, edge_from, **kw)
    want = synth.sync_state_host(old, pos, edges, vtx_from, edge_from, data, **kw)
    want["alpha"], want["beta"] = oracle.edge_weights(edges, pos)
    got0 = gpu_ctx.graph_download(3, V, E)
    _assert_state(got0, want, STATE + ("alpha", "beta"))
    if variant != "plain":
        carried = vtx_from >= 0
        assert ((old["x"][vtx_from[carried]] - data[carried]) > 0.25).any()

    gpu_ctx.pd_run(_lib.default_pd_params(), 25, slot0=3, mode=mode)
    oracle.pd_run(want, oracle.pd_default_params(), 25)
    _assert_state(gpu_ctx.graph_download(3, V, E), want)

    # second sync on top of the first (state carried twice), shrinking the graph
    g2 = {"pos": pos, "edges": edges, "W": g["W"], "H": g["H"]}
    old2 = {k: want[k] for k in STATE}
    pos2, tris2, edges2, vf2, ef2, data2 = _churn(g2, rng, 0.5, 40)
    gpu_ctx.graph_resync(3, pos2, data2, edges2, vf2, ef2)
    want2 = synth.sync_state_host(old2, pos2, edges2, vf2, ef2, data2)
    want2["alpha"], want2["beta"] = oracle.edge_weights(edges2, pos2)
    gpu_ctx.pd_run(_lib.default_pd_params(), 10, slot0=3, mode=mode)
    oracle.pd_run(want2, oracle.pd_default_params(), 10)
    _assert_state(gpu_ctx.graph_download(3, pos2.shape[0], edges2.shape[0]), want2, STATE + ("alpha", "beta"))


def test_graph_resync_grow_shrink_and_errors(gpu_ctx, oracle):
    small = synth.make_graph(640, 480, 300, seed=20, noise_seed=40, triangulator="scipy")
    big = synth.config_graph("D")
    p = _lib.default_pd_params()
    # unloaded slot -> ESTATE
    with _lib.Context(0, 1) as fresh:
        with pytest.raises(_lib.FlameB200Error):
            fresh.graph_resync(0, small["pos"], small["data_term"], small["edges"],
                               np.full(300, -1, np.int32), np.full(small["edges"].shape[0], -1, np.int32))
    # small -> big (buffers grow while the old state is carried): first 300 vertices carried
    gpu_ctx.graph_init(4, small["pos"], small["data_term"], small["edges"])
    gpu_ctx.pd_run(p, 12, slot0=4)
    old = gpu_ctx.graph_download(4, 300, small["edges"].shape[0])
    V, E = big["pos"].shape[0], big["edges"].shape[0]
    vtx_from = np.full(V, -1, np.int32)
    vtx_from[:300] = np.arange(300)[::-1]  # any mapping is legal
    edge_from = np.full(E, -1, np.int32)
    edge_from[5:5 + small["edges"].shape[0]:2] = np.arange(small["edges"].shape[0], dtype=np.int32)[::2]
    gpu_ctx.graph_resync(4, big["pos"], big["data_term"], big["edges"], vtx_from, edge_from)
    want = synth.sync_state_host(old, big["pos"], big["edges"], vtx_from, edge_from, big["data_term"])
    want["alpha"], want["beta"] = oracle.edge_weights(big["edges"], big["pos"])
    gpu_ctx.pd_run(p, 6, slot0=4)
    oracle.pd_run(want, oracle.pd_default_params(), 6)
    _assert_state(gpu_ctx.graph_download(4, V, E), want, STATE + ("alpha", "beta"))
    # a mesh bound before the sync refers to the old vertex set: projecting it must be refused
    gpu_ctx.mesh_bind(4, big["tris"], big["H"], big["W"])
    gpu_ctx.mesh_project(4, 1, 1.0)
    gpu_ctx.graph_resync(4, small["pos"], small["data_term"], small["edges"], np.full(300, -1, np.int32),
                         np.full(small["edges"].shape[0], -1, np.int32))
    with pytest.raises(_lib.FlameB200Error):
        gpu_ctx.mesh_project(4, 1, 1.0)
    gpu_ctx.graph_resync(4, big["pos"], big["data_term"], big["edges"], vtx_from, edge_from)  # not carried: new state
    gpu_ctx.graph_upload(4, dict(want, data_weight=np.ones(V, np.float32)))
    # out-of-range maps are rejected and leave the slot usable
    bad = vtx_from.copy()
    bad[7] = V
    with pytest.raises(_lib.FlameB200Error):
        gpu_ctx.graph_resync(4, big["pos"], big["data_term"], big["edges"], bad, edge_from)
    bad_e = edge_from.copy()
    bad_e[0] = E
    with pytest.raises(_lib.FlameB200Error):
        gpu_ctx.graph_resync(4, big["pos"], big["data_term"], big["edges"], vtx_from, bad_e)
    _assert_state(gpu_ctx.graph_download(4, V, E), want)
    # big -> empty -> small (all new)
    e0 = np.zeros((0, 2), np.int32)
    gpu_ctx.graph_resync(4, np.zeros((0, 2), np.float32), np.zeros(0, np.float32), e0,
                         np.zeros(0, np.int32), np.zeros(0, np.int32))
    gpu_ctx.pd_run(p, 3, slot0=4)
    gpu_ctx.graph_resync(4, small["pos"], small["data_term"], small["edges"], np.full(300, -1, np.int32),
                         np.full(small["edges"].shape[0], -1, np.int32))
    gpu_ctx.pd_run(p, 12, slot0=4)
    _assert_state(gpu_ctx.graph_download(4, 300, small["edges"].shape[0]), old)


@pytest.mark.parametrize("grad_check,letterbox", [(False, False), (True, False), (True, True)])
def test_graph_project_marked_matches_project_graph(gpu_ctx, oracle, grad_check, letterbox):
    """flame_b200_graph_project_marked vs the oracle's Flame::projectGraph (pinned on the verbatim slice):
    projected positions / idepths bit-exact and the same removal marks - outside valid_region (float rectangle),
    behind the camera, and the gradient check on the new frame's (device-resident, padded) gradients."""
    W, H, V, _ = synth.CONFIGS["A"]
    K, Kinv, frames = synth.make_stereo_pair(W, H, t_world=(0.10, 0.02, 0.05), yaw_deg=1.5)
    g = synth.config_graph("A")
    pad = synth.PAD
    gx = np.ascontiguousarray(frames[1]["gradx_pad"][pad:-pad, pad:-pad])
    gy = np.ascontiguousarray(frames[1]["grady_pad"][pad:-pad, pad:-pad])
    s = _warm_state(oracle, g, 10)
    s["x"][::31] = 0.0
    s["x"][5::47] = 1.9
    q = synth.quat_from_yaw(2.0).astype(np.float32)
    gpu_ctx.set_camera(W, H, K, Kinv)
    for i, fr in enumerate(frames):
        gpu_ctx.frame_upload(i, fr["img_pad"], fr["gradx_pad"], fr["grady_pad"], pad, fr["pose"])
    n_removed = 0
    for scale, t in ((1.0, np.array([0.3, -0.05, 0.1], np.float32)), (0.7, np.array([0.05, 0.0, -0.9], np.float32))):
        rules_o = oracle.project_rules(grad_check, letterbox, 6.0)
        rules_g = _lib.ProjectRules(1.4, 5, int(letterbox), int(grad_check), 6.0)
        gpu_ctx.graph_upload(0, s)
        pos_g, id_g, rem_g = gpu_ctx.graph_project_marked(0, V, q, t, scale, rules_g, fnew=1)
        pos_w, x_w, id_w, rem_w = oracle.project_graph(K, Kinv, q, t, W, H, rules_o, scale, s["pos"], s["x"], gx, gy)
        assert _bits_equal(pos_g, pos_w) and _bits_equal(id_g, id_w)
        assert np.array_equal(rem_g, rem_w)
        n_removed += int(rem_w.sum())
        assert _bits_equal(gpu_ctx.graph_download(0, V, g["edges"].shape[0])["x"], x_w)
    assert 0 < n_removed < 2 * V
