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


def test_large_footprints_slivers_and_wrapping_lines(gpu_ctx, oracle):
    """Triangles with a large sweep footprint take the CTA-per-triangle path (row spans clipped by the
    edge functions): long hull slivers, big interior triangles, long degenerate lines (diagonal,
    horizontal, vertical), and degenerate lines whose 4-px sweep overshoot runs past the right border and
    lands on the next row (linear addressing, rasterization.cc:206-236), incl. on the last row."""
    W, H = 200, 120
    pos = np.array([[2, 3], [197, 5], [100, 7],          # 0-2 thin sliver across the image
                    [10, 20], [190, 25], [90, 110],      # 3-5 big interior triangle
                    [5, 100], [150, 30], [77.5, 65],     # 6-8 diagonal line (set to exact collinear points below)
                    [3, 50], [199, 50], [120, 50],       # 9-11 horizontal line touching the right border
                    [60, 2], [60, 118], [60, 40],        # 12-14 vertical line
                    [150, 119], [199, 119], [170, 119],  # 15-17 horizontal line on the LAST row, wraps off-image
                    [193, 80], [198, 80], [196, 80],     # 18-20 short line at the border (small path, wraps)
                    [0, 0], [199, 0], [0, 119], [199, 119]], np.float32)
    pos[6], pos[7], pos[8] = (4, 100), (152, 26), (78, 63)   # integer collinear points on one line
    tris = np.array([[21, 22, 23], [23, 22, 24], [21, 23, 22], [24, 22, 23],   # both windings: full cover
                     [0, 1, 2], [0, 2, 1], [3, 4, 5], [3, 5, 4], [6, 7, 8], [8, 7, 6], [9, 10, 11],
                     [11, 10, 9], [12, 13, 14], [15, 16, 17], [18, 19, 20], [20, 19, 18]], np.int32)
    rng = np.random.default_rng(9)
    vals = rng.uniform(0.1, 2.0, pos.shape[0]).astype(np.float32)
    want = oracle.interpolate_mesh(tris, pos, vals, H, W)
    got = gpu_ctx.interpolate_mesh(0, tris, pos, vals, H, W)
    _assert_same_image(got, want)
    assert np.isnan(want).sum() < W * H
    # the same with a background image instead of the NaN fill, and width not a multiple of 4
    for (w2, h2) in ((199, 120), (201, 77)):
        p2 = np.minimum(pos, np.array([w2 - 1, h2 - 1], np.float32))
        base = rng.uniform(size=(h2, w2)).astype(np.float32)
        want = oracle.interpolate_mesh(tris, p2, vals, h2, w2, None, None, base.copy())
        got = gpu_ctx.interpolate_mesh(0, tris, p2, vals, h2, w2, None, None, base.copy())
        _assert_same_image(got, want)


def test_owner_epoch_wrap_and_rebinding(gpu_ctx, oracle):
    """The per-pixel owner words are tagged with a 12-bit projection epoch instead of being cleared:
    run past the wrap, re-bind the slot to other image sizes in between, and compare with the oracle."""
    rng = np.random.default_rng(4)
    g = synth.make_graph(64, 48, 60, seed=31, noise_seed=32, triangulator="scipy")
    big = synth.config_graph("A")
    tris_half = np.ascontiguousarray(g["tris"][::2])
    for it in range(4200):
        if it in (1000, 4090):  # other size in the same slot, then back
            v = big["data_term"]
            got = gpu_ctx.interpolate_mesh(5, big["tris"], big["pos"], v, big["H"], big["W"])
            _assert_same_image(got, oracle.interpolate_mesh(big["tris"], big["pos"], v, big["H"], big["W"]))
        vals = rng.uniform(0.1, 2.0, g["pos"].shape[0]).astype(np.float32)
        tris = g["tris"] if it % 3 else tris_half  # coverage shrinks and grows again
        got = gpu_ctx.interpolate_mesh(5, tris, g["pos"], vals, 48, 64)
        if it % 500 == 0 or it > 4080:
            _assert_same_image(got, oracle.interpolate_mesh(tris, g["pos"], vals, 48, 64))


def test_async_projection_pipeline(oracle):
    """flame_b200_ctx_set_async_projection: the projection runs on the context's second stream while the
    next solver launches proceed; downloads / re-binds / uploads order themselves after it."""
    gs = [synth.make_graph(640, 480, 500 + 100 * i, seed=50 + i, noise_seed=60 + i, triangulator="scipy")
          for i in range(3)]
    p = _lib.default_pd_params()
    with _lib.Context(0, 3) as ctx:
        ctx.set_async_projection(True)
        for i, g in enumerate(gs):
            ctx.graph_init(i, g["pos"], g["data_term"], g["edges"])
            ctx.mesh_bind(i, g["tris"], g["H"], g["W"])
        outs = [_lib.PinnedArray((g["H"], g["W"])) for g in gs]
        for rep, iters in enumerate((5, 9, 14)):
            ctx.graph_reset(0, 3)
            ctx.pd_run(p, iters, 0, 3)
            ctx.mesh_project(0, 3, 1.0)
            # no sync: the next round's reset + solver overlap this projection
        for i in range(3):
            ctx.mesh_download(i, outs[i].array)
        ctx.sync()
        for i, g in enumerate(gs):
            alpha, beta = oracle.edge_weights(g["edges"], g["pos"])
            s = synth.init_state(g, alpha, beta)
            oracle.pd_run(s, oracle.pd_default_params(), 14)
            _assert_same_image(outs[i].array.copy(), oracle.interpolate_mesh(g["tris"], g["pos"], s["x"], g["H"], g["W"]))
        # a re-bind + upload while a projection is pending must not corrupt it
        ctx.graph_reset(0, 3)
        ctx.pd_run(p, 4, 0, 3)
        ctx.mesh_project(0, 3, 2.0)
        ctx.mesh_download(1, outs[1].array)
        ctx.graph_init(1, gs[0]["pos"], gs[0]["data_term"], gs[0]["edges"])
        ctx.mesh_bind(1, gs[0]["tris"], gs[0]["H"], gs[0]["W"])
        ctx.sync()
        g = gs[1]
        alpha, beta = oracle.edge_weights(g["edges"], g["pos"])
        s = synth.init_state(g, alpha, beta)
        oracle.pd_run(s, oracle.pd_default_params(), 4)
        want = oracle.interpolate_mesh(g["tris"], g["pos"], (s["x"] * np.float32(2.0)).astype(np.float32), g["H"], g["W"])
        _assert_same_image(outs[1].array.copy(), want)


def test_random_triangle_soups(oracle):
    """The soups on which the oracle is pinned against rasterization.cc compiled verbatim
    (tests/test_oracle_kat.py::test_oracle_raster_equals_reference_on_random_soups), on the device."""
    with _lib.Context(0, 1) as ctx:
        for c in synth.triangle_soups():
            W, H, base = c["W"], c["H"], c["base"]
            want = oracle.interpolate_mesh(c["tris"], c["pos"], c["vals"], H, W, c["vv"], c["tv"],
                                           None if base is None else base.copy())
            got = ctx.interpolate_mesh(0, c["tris"], c["pos"], c["vals"], H, W, c["vv"], c["tv"],
                                       None if base is None else base.copy())
            _assert_same_image(got, want)


def test_limits_are_guarded(gpu_ctx):
    """The limits the kernels are built around fail loudly at the boundary instead of silently misbehaving:
    image sides above 2896 px (the reference's fp32 edge functions stop being exact), frame slots beyond 64,
    search parameters that allow more SSD positions than a warp holds, non-finite feature records."""
    g = synth.config_graph("A")
    gpu_ctx.graph_init(0, g["pos"], g["data_term"], g["edges"])
    gpu_ctx.mesh_bind(0, g["tris"], 2896, 2896)          # the largest image still exact
    with pytest.raises(_lib.FlameB200Error):
        gpu_ctx.mesh_bind(0, g["tris"], 480, 2897)
    with pytest.raises(_lib.FlameB200Error):
        gpu_ctx.mesh_bind(0, g["tris"], 2897, 640)
    gpu_ctx.mesh_bind(0, g["tris"], g["H"], g["W"])
    W, H, _, _ = synth.CONFIGS["A"]
    K, Kinv, frames = synth.make_stereo_pair(W, H)
    gpu_ctx.set_camera(W, H, K, Kinv)
    fr = frames[0]
    gpu_ctx.frame_upload(63, fr["img_pad"], fr["gradx_pad"], fr["grady_pad"], synth.PAD, fr["pose"])  # last slot
    with pytest.raises(_lib.FlameB200Error):
        gpu_ctx.frame_upload(64, fr["img_pad"], fr["gradx_pad"], fr["grady_pad"], synth.PAD, fr["pose"])
    gpu_ctx.frame_upload(0, fr["img_pad"], fr["gradx_pad"], fr["grady_pad"], synth.PAD, fr["pose"])
    gpu_ctx.frame_upload(1, frames[1]["img_pad"], frames[1]["gradx_pad"], frames[1]["grady_pad"], synth.PAD,
                         frames[1]["pose"])
    feats = synth.make_features(g["pos"], W, H)
    tp = _lib.default_track_params()
    tp.fparams.epilength_max = 124.0      # 124 / 1 + 4 = 128 positions: still fits
    gpu_ctx.update_feature_idepths(tp, 1, 0, feats.copy())
    tp.fparams.epilength_max = 125.0
    with pytest.raises(_lib.FlameB200Error):
        gpu_ctx.update_feature_idepths(tp, 1, 0, feats.copy())
    tp = _lib.default_track_params()
    bad = feats.copy()
    bad["x"][3] = np.nan
    with pytest.raises(_lib.FlameB200Error):
        gpu_ctx.update_feature_idepths(tp, 1, 0, bad)


def test_projection_redirected_into_another_buffer(gpu_ctx, oracle):
    """flame_b200_mesh_redirect: the projection stores the map (and copies the vertex idepths) straight into a
    caller-supplied device buffer - here on the same GPU; tests/test_multi_gpu.py does it across GPUs."""
    import torch
    g = synth.config_graph("B")
    H, W, V = g["H"], g["W"], g["pos"].shape[0]
    gpu_ctx.graph_init(1, g["pos"], g["data_term"], g["edges"])
    gpu_ctx.mesh_bind(1, g["tris"], H, W)
    gpu_ctx.pd_run(_lib.default_pd_params(), 20, slot0=1)
    x = gpu_ctx.graph_download(1, V, g["edges"].shape[0], fields=("x",))["x"]
    want = oracle.interpolate_mesh(g["tris"], g["pos"], (x * np.float32(0.5)).astype(np.float32), H, W)
    buf = _lib.DeviceBuffer(0, H * W * 4 + V * 4)
    try:
        buf.tensor((H * W + V,)).fill_(-7.0)
        torch.cuda.synchronize()
        gpu_ctx.mesh_redirect(1, buf.ptr, buf.ptr + H * W * 4)
        gpu_ctx.mesh_project(1, 1, 0.5)
        gpu_ctx.sync()
        got = buf.tensor((H, W)).cpu().numpy()
        vals = buf.tensor((V,), H * W * 4).cpu().numpy()
        gn, wn = np.isnan(got), np.isnan(want)
        assert np.array_equal(gn, wn) and np.array_equal(got[~gn].view(np.uint32), want[~wn].view(np.uint32))
        assert np.array_equal(vals.view(np.uint32), (x * np.float32(0.5)).astype(np.float32).view(np.uint32))
        gpu_ctx.mesh_redirect(1, None, None)   # back to the slot's own buffer
        gpu_ctx.mesh_project(1, 1, 0.5)
        out = np.empty((H, W), np.float32)
        gpu_ctx.mesh_download(1, out)
        gpu_ctx.sync()
        assert np.array_equal(np.isnan(out), wn) and np.array_equal(out[~wn].view(np.uint32), want[~wn].view(np.uint32))
    finally:
        gpu_ctx.mesh_redirect(1, None, None)
        gpu_ctx.sync()
        buf.close()
