"""GPU parity for the display filters (flame.cc:389-407, 2207-2356), the vertex normals (flame.cc:2554-2640)
and the coverage statistic (flame.cc:434-443): flame_b200_mesh_filter / mesh_normals / mesh_coverage vs the
oracle, which tests/test_oracle_kat.py pins on the reference's own functions sliced verbatim out of flame.cc.
Everything is bit-exact except the oblique filter's angle, which goes through acosf (libdevice vs libm, both
within a few ulp of the true value): the tolerance is stated ON THE ANGLE - every triangle's device angle is
within ANGLE_TOL rad of the oracle's - and a filter decision may differ only for a triangle whose angle lies
within that tolerance of the threshold."""
import numpy as np
import pytest

from flame_b200 import _lib, synth

pytestmark = pytest.mark.gpu

ANGLE_TOL = 1e-6  # rad; acosf results are O(1): 1e-6 is ~8 ulp at 1.4 rad


def _same_image(got, want):
    gn, wn = np.isnan(got), np.isnan(want)
    return np.array_equal(gn, wn) and np.array_equal(got[~gn].view(np.uint32), want[~wn].view(np.uint32))


def _oracle_params(oracle, gp):
    import ctypes
    op = oracle.tri_filter_default_params()
    ctypes.memmove(ctypes.byref(op), ctypes.byref(gp), ctypes.sizeof(op))
    return op


def _default():
    p = _lib.TriFilterParams()
    _lib.lib().flame_b200_tri_filter_default_params(p)
    return p


@pytest.mark.parametrize("variant", ["default", "oblique_only", "edge_only", "idepth_only", "scaled"])
def test_filters_filtered_map_and_coverage(gpu_ctx, oracle, variant):
    g = synth.config_graph("C")
    W, H = g["W"], g["H"]
    V, E, T = g["pos"].shape[0], g["edges"].shape[0], g["tris"].shape[0]
    K, Kinv = synth.camera(W, H)
    data = g["data_term"].copy()
    data[::97] = 0.004   # some triangles below min_triangle_idepth
    gp = _default()
    scale = 1.0
    if variant == "oblique_only":
        gp.do_edge_length_filter = gp.do_idepth_triangle_filter = 0
        gp.oblique_normal_thresh = 1.2
    elif variant == "edge_only":
        gp.do_oblique_triangle_filter = gp.do_idepth_triangle_filter = 0
        gp.edge_length_thresh = 0.02
    elif variant == "idepth_only":
        gp.do_oblique_triangle_filter = gp.do_edge_length_filter = 0
        gp.min_triangle_idepth = 0.4
    elif variant == "scaled":
        scale = 0.6
    gpu_ctx.graph_init(2, g["pos"], data, g["edges"])
    gpu_ctx.mesh_bind(2, g["tris"], H, W)
    gpu_ctx.pd_run(_lib.default_pd_params(), 25, slot0=2)
    x = gpu_ctx.graph_download(2, V, E, fields=("x",))["x"]
    idepths = (x * np.float32(scale)).astype(np.float32)

    got = gpu_ctx.mesh_filter(2, T, Kinv, gp, scale=scale, apply=True)
    want, angle = oracle.triangle_filters(_oracle_params(oracle, gp), Kinv, W, g["tris"], g["pos"], idepths)
    diff = np.flatnonzero(got != want)
    assert 0 < want.sum() < T, "the filter should remove some, not all, triangles in this scene"
    if gp.do_oblique_triangle_filter:
        dev_angle = gpu_ctx.mesh_filter_angles(2, T, Kinv, scale)
        ok = ~np.isnan(angle)
        assert np.array_equal(np.isnan(dev_angle), ~ok)
        assert np.max(np.abs(dev_angle[ok] - angle[ok])) <= ANGLE_TOL
        assert np.all(np.abs(angle[diff] - gp.oblique_normal_thresh) <= ANGLE_TOL), \
            "%d decisions differ away from the acos threshold" % diff.size
    else:
        assert diff.size == 0

    # getFilteredInverseDepthMap (flame.h:217-228): the projection now honours the validity on the device
    gpu_ctx.mesh_project(2, 1, scale)
    out = _lib.PinnedArray((H, W))
    gpu_ctx.mesh_download(2, out.array)
    gpu_ctx.sync()
    img = out.array.copy()
    assert _same_image(img, oracle.interpolate_mesh(g["tris"], g["pos"], idepths, H, W, None, got))
    assert gpu_ctx.mesh_coverage(2) == oracle.coverage(img) == int((~np.isnan(img)).sum())

    # compute-only call leaves the installed validity alone; a new bind removes it
    gp2 = _default()
    gp2.do_oblique_triangle_filter = gp2.do_idepth_triangle_filter = 0
    gp2.edge_length_thresh = 1e-4
    none_valid = gpu_ctx.mesh_filter(2, T, Kinv, gp2, scale=scale, apply=False)
    assert none_valid.sum() == 0
    gpu_ctx.mesh_filter(2, T, Kinv, gp, scale=scale, apply=True, want=False)
    gpu_ctx.mesh_project(2, 1, scale)
    gpu_ctx.mesh_download(2, out.array)
    gpu_ctx.sync()
    assert _same_image(out.array, img)
    gpu_ctx.mesh_bind(2, g["tris"], H, W)
    gpu_ctx.mesh_project(2, 1, scale)
    gpu_ctx.mesh_download(2, out.array)
    gpu_ctx.sync()
    assert _same_image(out.array.copy(), oracle.interpolate_mesh(g["tris"], g["pos"], idepths, H, W))
    assert gpu_ctx.mesh_coverage(2) > oracle.coverage(img)


def test_filter_errors(gpu_ctx):
    K, Kinv = synth.camera(640, 480)
    with _lib.Context(0, 1) as ctx:
        with pytest.raises(_lib.FlameB200Error):
            ctx.mesh_filter(0, 1, Kinv)
        with pytest.raises(_lib.FlameB200Error):
            ctx.mesh_coverage(0)


@pytest.mark.parametrize("iters", [0, 25])
def test_vertex_normals_match_oracle(gpu_ctx, oracle, iters):
    """flame_b200_mesh_normals vs Flame::getVertexNormals (flame.cc:2554-2640; oracle pinned on the verbatim
    slice): only IEEE mul/add/div/sqrt are involved, so the comparison is bit-for-bit (NaN == NaN)."""
    g = synth.config_graph("C")
    W, H = g["W"], g["H"]
    V, E = g["pos"].shape[0], g["edges"].shape[0]
    K, Kinv = synth.camera(W, H)
    data = g["data_term"].copy()
    data[::41] = 0.0  # triangles with a non-positive corner idepth are skipped (flame.cc:2586-2588)
    for scale in (1.0, 0.8):
        gpu_ctx.graph_init(2, g["pos"], data, g["edges"])
        gpu_ctx.mesh_bind(2, g["tris"], H, W)
        gpu_ctx.pd_run(_lib.default_pd_params(), iters, slot0=2)
        x = gpu_ctx.graph_download(2, V, E, fields=("x",))["x"]
        got = gpu_ctx.mesh_normals(2, V, Kinv, scale)
        want = oracle.vertex_normals(Kinv, g["pos"], (x * np.float32(scale)).astype(np.float32), g["tris"])
        gn, wn = np.isnan(got), np.isnan(want)
        assert np.array_equal(gn, wn)
        assert np.array_equal(got[~gn].view(np.uint32), want[~wn].view(np.uint32))
        if iters == 0:
            assert (np.abs(want).sum(1) == 0).any()  # vertices all of whose triangles were skipped
        norms = np.linalg.norm(want[~wn.any(1) & (np.abs(want).sum(1) > 0)], axis=1)
        assert np.allclose(norms, 1.0, atol=1e-5)
