"""GPU parity: dense projection (interpolateMesh) vs the CPU oracle — identical NaN mask, values
bit-exact — through the C ABI."""
import numpy as np
import pytest

from flame_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _assert_same_image(got, want):
    gn, wn = np.isnan(got), np.isnan(want)
    assert np.array_equal(gn, wn), "NaN coverage mask differs at %d pixels" % int((gn != wn).sum())
    a, b = got[~gn].view(np.uint32), want[~wn].view(np.uint32)
    assert np.array_equal(a, b), "values differ at %d covered pixels" % int((a != b).sum())


@pytest.mark.parametrize("cfg", ["A", "B", "C", "D"])
def test_interpolate_mesh_bit_exact(gpu_ctx, oracle, cfg):
    g = synth.config_graph(cfg)
    vals = g["data_term"]
    want = oracle.interpolate_mesh(g["tris"], g["pos"], vals, g["H"], g["W"])
    got = gpu_ctx.interpolate_mesh(0, g["tris"], g["pos"], vals, g["H"], g["W"])
    _assert_same_image(got, want)


def test_reference_known_answers(gpu_ctx):
    """test/utils/image_utils_test.cc:754-783 (interpolateMeshTest): one triangle, values 1,2,3."""
    pos = np.array([[10, 10], [20, 10], [10, 20]], np.float32)   # must be CW as Triangle emits
    tris = np.array([[0, 1, 2]], np.int32)
    vals = np.array([1, 2, 3], np.float32)
    img = gpu_ctx.interpolate_mesh(0, tris, pos, vals, 64, 64)
    if np.isnan(img[10, 10]):  # wrong winding for this synthetic triangle: flip
        tris = np.array([[0, 2, 1]], np.int32)
        img = gpu_ctx.interpolate_mesh(0, tris, pos, vals, 64, 64)
    assert img[10, 10] == 1 and img[10, 20] == 2 and img[20, 10] == 3
    assert abs(img[15, 15] - 2.5) < 1e-6


def test_validity_masks_and_preinitialised_image(gpu_ctx, oracle):
    g = synth.config_graph("B")
    rng = np.random.default_rng(3)
    vv = (rng.uniform(size=g["pos"].shape[0]) > 0.1).astype(np.uint8)
    tv = (rng.uniform(size=g["tris"].shape[0]) > 0.2).astype(np.uint8)
    base = rng.uniform(size=(g["H"], g["W"])).astype(np.float32)
    want = oracle.interpolate_mesh(g["tris"], g["pos"], g["data_term"], g["H"], g["W"], vv, tv, base.copy())
    got = gpu_ctx.interpolate_mesh(0, g["tris"], g["pos"], g["data_term"], g["H"], g["W"], vv, tv, base.copy())
    _assert_same_image(got, want)


def test_degenerate_and_overlapping_triangles(gpu_ctx, oracle):
    """Collinear (after rounding) triangles write NaN along their line incl. the 4-px sweep overshoot;
    overlapping triangles: the LAST one in the list wins."""
    pos = np.array([[5, 5], [30, 5], [5, 30], [30, 30], [12.4, 17.5], [20.5, 17.49], [27, 17.6],
                    [40, 40], [40.2, 40.3], [39.8, 39.9]], np.float32)
    tris = np.array([[0, 1, 2], [0, 2, 1], [1, 3, 2], [2, 3, 1], [4, 5, 6], [6, 5, 4], [7, 8, 9],
                     [0, 3, 2], [2, 3, 0]], np.int32)
    vals = np.arange(1, 11, dtype=np.float32)
    want = oracle.interpolate_mesh(tris, pos, vals, 48, 48)
    got = gpu_ctx.interpolate_mesh(0, tris, pos, vals, 48, 48)
    _assert_same_image(got, want)
    assert np.isnan(want).sum() < 48 * 48  # something was drawn


def test_empty_mesh(gpu_ctx):
    img = gpu_ctx.interpolate_mesh(0, np.zeros((0, 3), np.int32), np.zeros((3, 2), np.float32),
                                   np.zeros(3, np.float32), 16, 16)
    assert np.isnan(img).all()


def test_project_from_graph_state(gpu_ctx, oracle):
    """mesh_bind + mesh_project use the slot's smoothed x (vtx_idepths_ = x * scale, flame.cc:376)."""
    g = synth.config_graph("C")
    alpha, beta = oracle.edge_weights(g["edges"], g["pos"])
    s = synth.init_state(g, alpha, beta)
    oracle.pd_run(s, oracle.pd_default_params(), 20)
    scale = np.float32(1.25)
    want = oracle.interpolate_mesh(g["tris"], g["pos"], (s["x"] * scale).astype(np.float32), g["H"], g["W"])
    gpu_ctx.graph_init(0, g["pos"], g["data_term"], g["edges"])
    gpu_ctx.mesh_bind(0, g["tris"], g["H"], g["W"])
    gpu_ctx.pd_run(_lib.default_pd_params(), 20)
    gpu_ctx.mesh_project(0, 1, float(scale))
    out = _lib.PinnedArray((g["H"], g["W"]))
    gpu_ctx.mesh_download(0, out.array)
    gpu_ctx.sync()
    _assert_same_image(out.array.copy(), want)


def test_process_frame_one_call(gpu_ctx, oracle):
    """flame_b200_process_frame == graph_init + pd_run + interpolateMesh of the oracle, bit-exact."""
    g = synth.config_graph("B")
    alpha, beta = oracle.edge_weights(g["edges"], g["pos"])
    s = synth.init_state(g, alpha, beta)
    oracle.pd_run(s, oracle.pd_default_params(), 30)
    want = oracle.interpolate_mesh(g["tris"], g["pos"], s["x"], g["H"], g["W"])
    out = _lib.PinnedArray((g["H"], g["W"]))
    V = g["pos"].shape[0]
    x, w1, w2 = (np.empty(V, np.float32) for _ in range(3))
    gpu_ctx.process_frame(4, g["pos"], g["data_term"], g["edges"], g["tris"], g["H"], g["W"],
                          _lib.default_pd_params(), 30, out.array, x, w1, w2)
    _assert_same_image(out.array.copy(), want)
    for a, b in ((x, s["x"]), (w1, s["w1"]), (w2, s["w2"])):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
