"""Pins the CPU oracle against every known-answer test the reference holds for this path
(SURVEY.md §8c): test/stereo/epipolar_geometry_test.cc (all 21 cases), test/utils/image_utils_test.cc
(bilinearInterp x4, Liang-Barsky x5, interpolateMesh), test/utils/rasterization_test.cc
(DrawShadedTriangleBarycentric1/2 vs the reference's golden PNGs, byte-exact) — and, when the
reference checkout was available at build time, against the reference sources compiled verbatim
(oracle/_ref). CPU only."""
import math
import os

import numpy as np
import pytest

from flame_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
K525 = np.array([[525, 0, 320], [0, 525, 240], [0, 0, 1]], np.float32)
I4 = (1.0, 0.0, 0.0, 0.0)


def aa_quat(angle, axis):
    """Eigen::Quaternionf(Eigen::AngleAxisf(angle, axis)) as (w, x, y, z)."""
    h = np.float32(angle) * np.float32(0.5)
    s = np.float32(math.sin(float(h)))
    c = np.float32(math.cos(float(h)))
    ax = np.asarray(axis, np.float32)
    return np.array([c, s * ax[0], s * ax[1], s * ax[2]], np.float32)


def qrot(q, v):
    w, x, y, z = [float(a) for a in q]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    return R @ np.asarray(v, np.float64)


def qconj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]], np.float32)


def epi(oracle, K=K525):
    return oracle.Epi(K)


# ---------------------------------------------------------------- epipolar_geometry_test.cc:47-131
@pytest.mark.parametrize("u_ref,axis", [((320, 240), 0), ((320, 0), 0), ((320, 240), 1), ((0, 240), 1)])
def test_min_depth_projection_translate(oracle, u_ref, axis):
    e = epi(oracle)
    t = [0.0, 0.0, 0.0]
    t[axis] = 2.0
    e.load(I4, t)
    m = e.min_depth_projection(u_ref)
    assert m[axis] > K525[axis, 2] * 2
    assert abs(m[1 - axis] - u_ref[1 - axis]) < 1e-3
    t[axis] = -2.0
    e.load(I4, t)
    m = e.min_depth_projection(u_ref)
    assert m[axis] < 0
    assert abs(m[1 - axis] - u_ref[1 - axis]) < 1e-3


def test_min_depth_projection_60_yaw(oracle):  # :161-196
    q21 = aa_quat(-math.pi / 3, (0, 1, 0))
    t21 = np.array([2, 0, 0], np.float32)
    q12 = qconj(q21)
    t12 = (-qrot(q12, t21)).astype(np.float32)
    e = epi(oracle)
    e.load(q12, t12)
    m = e.min_depth_projection((320, 240))
    assert abs(m[0] - 16.8910904) < 1e-4 and abs(m[1] - 240) < 1e-4
    e.load(q21, t21)
    m = e.min_depth_projection((320, 240))
    assert abs(m[0] - 1049999424) < 1e-4 and abs(m[1] - 240) < 1e-4


KREAL = np.array([[535.43310546875, 0, 320.106652814575], [0, 539.212524414062, 247.632132204719],
                  [0, 0, 1]], np.float32)


def test_min_depth_projection_real_data(oracle):  # :201-243
    e = oracle.Epi(KREAL)
    e.load((0.999138, -0.000878, 0.041493, 0.000386), (-0.221092, -0.036134, 0.084099))
    m = e.min_depth_projection((320, 240))
    assert abs(m[0] - -1087.525391) < 1e-2 and abs(m[1] - 15.954912) < 1e-2
    e.load((-0.999853, 0.014856, -0.005249, -0.006822), (-0.258187, 0.040849, -0.054990))
    m = e.min_depth_projection((320, 240))
    assert abs(m[0] - 187.65597534179688) < 1e-1 and abs(m[1] - 278.55392456054688) < 1e-1


def test_max_depth_projection(oracle):  # :248-311
    e = epi(oracle)
    e.load(I4, (0, 0, 0))
    m = e.max_depth_projection((320, 240))
    assert abs(m[0] - 320) < 1e-3 and abs(m[1] - 240) < 1e-3
    e.load(aa_quat(-math.pi / 6, (0, 1, 0)), (0, 0, 0))
    m = e.max_depth_projection((320, 240))
    assert abs(m[0] - 16.891090393066406) < 1e-4 and abs(m[1] - 240) < 1e-4
    e.load(aa_quat(-math.pi / 6, (1, 0, 0)), (0, 0, 0))
    m = e.max_depth_projection((320, 240))
    assert abs(m[0] - 320) < 1e-4 and abs(m[1] - 543.10888671875) < 1e-4


def test_epiline(oracle):  # :313-361
    q = aa_quat(-math.pi / 3, (0, 1, 0))
    e = epi(oracle)
    e.load(qconj(q), (-qrot(q, (2, 0, 0))).astype(np.float32))
    _, ep = e.epiline((320, 240))
    assert abs(ep[0] - 1) < 1e-4 and abs(ep[1]) < 1e-4
    q = aa_quat(math.pi / 3, (1, 0, 0))
    e.load(qconj(q), (-qrot(q, (0, 2, 0))).astype(np.float32))
    _, ep = e.epiline((320, 240))
    assert abs(ep[0]) < 1e-4 and abs(ep[1] - 1) < 1e-4


def _two_cameras(oracle, angle, p_world):
    """T1 = yaw `angle` at the origin, T2 = identity rotation at (1,0,0); poses are camera-to-world."""
    T1 = np.concatenate([aa_quat(angle, (0, 1, 0)), [0, 0, 0]]).astype(np.float32)
    T2 = np.array([1, 0, 0, 0, 1, 0, 0], np.float32)
    T12 = oracle.se3_inv_mul(T2, T1)  # T2^-1 * T1
    T21 = oracle.se3_inv_mul(T1, T2)
    u1 = oracle.project_world(K525, T1[:4], T1[4:], p_world)
    u2 = oracle.project_world(K525, T2[:4], T2[4:], p_world)
    z1 = qrot(qconj(T1[:4]), np.asarray(p_world, np.float64) - T1[4:])[2]
    z2 = qrot(qconj(T2[:4]), np.asarray(p_world, np.float64) - T2[4:])[2]
    return T12, T21, u1, u2, z1, z2


CASES = [(-math.pi / 12, (1, 0, 10)), (-math.pi / 12, (-1, 0, 10)), (math.pi / 12, (0, 1, 10)),
         (math.pi / 12, (0, -1, 10))]


@pytest.mark.parametrize("angle,p_world", CASES)
def test_disparity_to_depth(oracle, angle, p_world):  # :370-560
    T12, T21, u1, u2, z1, z2 = _two_cameras(oracle, angle, p_world)
    e1 = epi(oracle)
    e1.load(T12[:4], T12[4:])
    d, u_inf, ep = e1.disparity(u1, u2)
    assert abs(e1.disparity_to_depth(u1, u_inf, ep, d) - z1) < 1e-4 * max(1.0, abs(z1)) + 1e-4
    e2 = epi(oracle)
    e2.load(T21[:4], T21[4:])
    d, u_inf, ep = e2.disparity(u2, u1)
    assert abs(e2.disparity_to_depth(u2, u_inf, ep, d) - z2) < 1e-4 * max(1.0, abs(z2)) + 1e-4


@pytest.mark.parametrize("angle,p_world,tol1", [(c[0], c[1], t) for c, t in zip(CASES, (1e-4, 1e-4, 1e-2, 1e-2))])
def test_disparity_to_inverse_depth(oracle, angle, p_world, tol1):  # :570-760
    T12, T21, u1, u2, z1, z2 = _two_cameras(oracle, angle, p_world)
    e1 = epi(oracle)
    e1.load(T12[:4], T12[4:])
    d, u_inf, ep = e1.disparity(u1, u2)
    assert abs(e1.disparity_to_idepth(u1, u_inf, ep, d) - 1.0 / z1) < tol1
    e2 = epi(oracle)
    e2.load(T21[:4], T21[4:])
    d, u_inf, ep = e2.disparity(u2, u1)
    assert abs(e2.disparity_to_idepth(u2, u_inf, ep, d) - 1.0 / z2) < 1e-4


def test_project_round_trip(oracle):  # :773-805
    T12, T21, u1, u2, _, _ = _two_cameras(oracle, -math.pi / 12, (1, 0, 10))
    e = epi(oracle)
    e.load(T21[:4], T21[4:])
    uc = e.project(u2, 1.0 / 10.0)
    assert abs(uc[0] - u1[0]) < 1e-4 and abs(uc[1] - u1[1]) < 1e-4


# ---------------------------------------------------------------- image_utils_test.cc
def test_bilinear_interp_kats(oracle):  # :150-166
    img = np.array([[91, 210], [162, 95]], np.uint8)
    for (x, y), want in (((0.5, 0.2), 146.1), ((0.2, 0.5), 131.70001), ((0.5, 0.5), 139.5),
                         ((0.2, 0.2), 121.5600052)):
        assert abs(oracle.bilinear_u8(img, x, y) - want) < 1e-5 * 10  # EXPECT_NEAR(…, 1e-5) on fp32
    f = img.astype(np.float32)
    assert abs(oracle.bilinear_f32(f, 0.5, 0.5) - 139.5) < 1e-5


def test_liang_barsky_cases(oracle):  # :644-749
    ok, c = oracle.clip_line(0, 3, 0, 3, 1, 1, 2, 2)
    assert ok and np.allclose(c, (1, 1, 2, 2), atol=1e-6)
    ok, c = oracle.clip_line(1.1, 3, 1.2, 3, 1, 1, 2, 2)
    assert ok and np.allclose(c, (1.2, 1.2, 2, 2), atol=1e-6)
    ok, c = oracle.clip_line(1, 1.5, 1, 1.25, 1, 1, 2, 2)
    assert ok and np.allclose(c, (1, 1, 1.25, 1.25), atol=1e-6)
    ok, c = oracle.clip_line(1.1, 1.5, 1.15, 1.25, 1, 1, 2, 2)
    assert ok and np.allclose(c, (1.15, 1.15, 1.25, 1.25), atol=1e-6)
    lo = np.float32(2 + 1e-3 + 1)
    ok, _ = oracle.clip_line(lo, np.float32(640 - 2 - 1e-3 - 2), lo, np.float32(480 - 2 - 1e-3 - 2),
                             -3.40414, 1.95745, -3.26889, 2.17682)
    assert not ok


def test_interpolate_mesh_kat(oracle):  # :754-783
    tris = np.array([[0, 1, 2]], np.int32)
    pos = np.array([[10, 10], [20, 10], [10, 20]], np.float32)
    vals = np.array([1, 2, 3], np.float32)
    for use_ref in ([False, True] if oracle.have_ref() else [False]):
        img = oracle.interpolate_mesh(tris, pos, vals, 40, 40, use_ref=use_ref)
        assert abs(img[10, 10] - 1) < 1e-3 and abs(img[10, 20] - 2) < 1e-3 and abs(img[20, 10] - 3) < 1e-3
        assert abs(img[15, 15] - 2.5) < 1e-3


# ---------------------------------------------------------------- rasterization_test.cc:514-598
def _normalize_u8(img):
    mn, mx = float(img.min()), float(img.max())
    return np.rint((img - mn) * (255.0 / (mx - mn))).astype(np.uint8)  # cv::normalize MINMAX + convertTo


@pytest.mark.parametrize("k,pts,vals", [(1, ((150, 100), (100, 200), (200, 300)), (0.0, 1.0, 0.0)),
                                         (2, ((150, 100), (100, 300), (200, 200)), (0.0, 0.0, 1.0))])
def test_rasteriser_golden_png(oracle, k, pts, vals):
    cv2 = pytest.importorskip("cv2")
    truth = cv2.imread(os.path.join(GOLDEN, "RasterizationTest_DrawShadedTriangleBarycentric%d.png" % k),
                       cv2.IMREAD_GRAYSCALE)
    assert truth is not None and truth.shape == (480, 640)
    for use_ref in ([False, True] if oracle.have_ref() else [False]):
        img = np.zeros((480, 640), np.float32)
        oracle.draw_triangle(img, pts[0], pts[1], pts[2], *vals, use_ref=use_ref)
        assert np.array_equal(_normalize_u8(img), truth)  # byte-exact, as cvCompare


# ---------------------------------------------------------------- oracle vs reference-compiled-verbatim
def _need_ref(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")


@pytest.mark.parametrize("cfg", ["A", "B"])
def test_oracle_raster_equals_reference_verbatim(oracle, cfg):
    _need_ref(oracle)
    g = synth.config_graph(cfg)
    a = oracle.interpolate_mesh(g["tris"], g["pos"], g["data_term"], g["H"], g["W"])
    b = oracle.interpolate_mesh(g["tris"], g["pos"], g["data_term"], g["H"], g["W"], use_ref=True)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(a)
    assert np.array_equal(a[m].view(np.uint32), b[m].view(np.uint32))


def test_golden_graph_fixtures_match_reference_triangle(oracle):
    """The committed index arrays are what the reference's own Triangle returns today (bit-exact)."""
    _need_ref(oracle)
    for name, (W, H, V, _) in synth.CONFIGS.items():
        z = np.load(os.path.join(GOLDEN, "graph_%dx%d_v%d_s1.npz" % (W, H, V)))
        assert np.array_equal(z["pos"], synth.make_vertices(W, H, V, seed=1))
        tris, edges, _ = oracle.ref_delaunay(z["pos"])
        assert np.array_equal(tris, z["tris"]) and np.array_equal(edges, z["edges"])
        if name == "B":
            break  # A and B are enough for the CPU suite's time budget


@pytest.mark.parametrize("cfg", ["A", "B"])
def test_oracle_pd_equals_reference_verbatim(oracle, cfg):
    """nltgv2_l1_graph_regularizer.cc compiled verbatim (Boost.Graph container shimmed) vs the flat
    restatement: every state array bit-identical after the config's iteration count; costs equal."""
    _need_ref(oracle)
    g = synth.config_graph(cfg)
    alpha, beta = oracle.edge_weights(g["edges"], g["pos"])
    rng = np.random.default_rng(1)
    g["data_weight"] = rng.uniform(0.5, 2.0, g["pos"].shape[0]).astype(np.float32)
    a = synth.init_state(g, alpha, beta)
    b = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    p = oracle.pd_default_params()
    oracle.pd_run(a, p, g["pd_iters"])
    oracle.ref_pd_run(b, p, g["pd_iters"])
    for k in oracle.GRAPH_STATE:
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    assert oracle.pd_costs(a, p) == oracle.ref_pd_costs(b, p)


@pytest.mark.parametrize("t_world,yaw", [((0.10, 0.02, 0.01), 1.0), ((0.05, 0.01, 0.45), 0.5),
                                         ((-0.08, 0.03, 0.0), -0.7), ((0.1, 0.02, -0.05), 1.0)])
def test_oracle_stereo_equals_reference_verbatim(oracle, t_world, yaw):
    """stereo/inverse_depth_filter.cc, inverse_depth_meas_model.cc, line_stereo.h, epipolar_geometry.h
    compiled verbatim (types shimmed) vs the C restatement, per feature, bit-identical — covers
    t_z > 0, t_z == 0, t_z < 0 and the re-parenting path."""
    _need_ref(oracle)
    W, H, V, _ = synth.CONFIGS["A"]
    K, Kinv, frames = synth.make_stereo_pair(W, H, t_world=t_world, yaw_deg=yaw)
    g = synth.config_graph("A")
    feats = synth.make_features(g["pos"], W, H, rel_sigma=0.1)
    p = oracle.track_default_params()
    f1, f2 = feats.copy(), feats.copy()
    n1, d1 = oracle.update_feature_idepths(p, K, Kinv, W, H, synth.PAD, frames, 1, 0, f1)
    n2, d2 = oracle.ref_update_feature_idepths(p, K, Kinv, W, H, synth.PAD, frames, 1, 0, f2)
    assert n1 == n2 and n1 > 0 and np.array_equal(d1["stage"], d2["stage"])
    assert f1.tobytes() == f2.tobytes()
    for k in ("flow_x", "flow_y", "mu_meas", "var_meas"):
        m = ~np.isnan(d1[k])
        assert np.array_equal(np.isnan(d1[k]), np.isnan(d2[k]))
        assert np.array_equal(d1[k][m].view(np.uint32), d2[k][m].view(np.uint32)), k


def _border_features(W, H, border=4):
    """Features within +-0.4 px of the valid-region border (cv::Rect::contains rounds the Point2f first)."""
    from flame_b200 import synth as _s
    g = _s.config_graph("A")
    feats = _s.make_features(g["pos"], W, H, rel_sigma=0.1)
    n = len(feats)
    offs = np.array([-0.6, -0.5, -0.4, -0.1, 0.0, 0.4, 0.5, 0.6], np.float32)
    q = n // 4
    feats["x"][:q] = border + offs[np.arange(q) % 8]
    feats["x"][q:2 * q] = (W - border) + offs[np.arange(q) % 8]
    feats["y"][2 * q:3 * q] = border + offs[np.arange(q) % 8]
    feats["y"][3 * q:4 * q] = (H - border) + offs[np.arange(q) % 8]
    return feats


def test_valid_region_rounds_like_opencv(oracle):
    """cv::Rect_<int>::contains(Point_<int>) — a Point2f is saturate_cast to int (round half to even)
    BEFORE the comparison (flame.cc:1624,1680): x = 3.6 with border 4 is inside, x = 3.4 is not; the
    right edge W - 4 excludes x = W - 4.4 .. (rounds to W - 4) but keeps W - 4.6."""
    W, H, V, _ = synth.CONFIGS["A"]
    K, Kinv, frames = synth.make_stereo_pair(W, H)
    feats = _border_features(W, H)
    p = oracle.track_default_params()
    f1 = feats.copy()
    n1, d1 = oracle.update_feature_idepths(p, K, Kinv, W, H, synth.PAD, frames, 1, 0, f1)
    # border = int(rescale_factor_max * win_size / 2 + 1) = 4 (flame.cc:1576); stage 5 = "outside the valid region"
    ix, iy = np.rint(feats["x"]).astype(int), np.rint(feats["y"]).astype(int)  # np.rint: half to even
    inside = (ix >= 4) & (ix < W - 4) & (iy >= 4) & (iy < H - 4)
    reached = d1["stage"] >= 5  # earlier exits (predict / search-region failures) never get to the test
    assert np.array_equal(d1["stage"][reached] != 5, inside[reached])
    near = (np.abs(feats["x"] - (W - 4)) < 0.7) | (np.abs(feats["y"] - 4) < 0.7) | (np.abs(feats["y"] - (H - 4)) < 0.7)
    assert (inside & reached & near).sum() > 20 and (~inside & reached & near).sum() > 20
    # the unrounded comparison would decide differently for some of them
    naive = (feats["x"] >= 4) & (feats["x"] < W - 4) & (feats["y"] >= 4) & (feats["y"] < H - 4)
    assert (naive != inside)[reached].any()
    if oracle.have_ref():
        f2 = feats.copy()
        n2, d2 = oracle.ref_update_feature_idepths(p, K, Kinv, W, H, synth.PAD, frames, 1, 0, f2)
        assert n1 == n2 and np.array_equal(d1["stage"], d2["stage"]) and f1.tobytes() == f2.tobytes()


def test_epigeo_oracle_equals_reference_verbatim(oracle):
    _need_ref(oracle)
    K, Kinv = synth.camera(640, 480)
    rng = np.random.default_rng(0)
    for q, t in ((aa_quat(0.02, (0, 1, 0)), (0.1, 0.02, 0.01)), (aa_quat(-0.3, (1, 0, 0)), (0.2, -0.1, -0.3)),
                 (np.array(I4, np.float32), (-0.5, 0.0, 0.0))):
        e = oracle.Epi(K, Kinv)
        e.load(q, t)
        for _ in range(200):
            u = (float(np.float32(rng.uniform(5, 635))), float(np.float32(rng.uniform(5, 475))))
            idp = float(np.float32(rng.uniform(0.05, 1.5)))
            r = oracle.ref_epigeo_query(K, Kinv, q, t, u, idp)
            (ui, ep) = e.epiline(u)
            o = np.array([*e.min_depth_projection(u), *e.max_depth_projection(u), *ui, *ep,
                          *e.project(u, idp), *e.reference_epiline(u)], np.float32)
            assert np.array_equal(o.view(np.uint32), r.view(np.uint32))


# ---------------------------------------------------------------- frame preparation
@pytest.mark.parametrize("axis", [0, 1])
def test_central_gradient_ramp_kat(oracle, axis):
    """test/utils/image_utils_test.cc:171-330 (getCentralGradientHorizontal/Vertical): a 640x480 ramp,
    0.5*(right-left) inside, forward/backward differences on the first/last column/row."""
    W, H = 640, 480
    slope = np.float32(255.0) / np.float32(W if axis == 0 else H)
    idx = np.arange(W if axis == 0 else H, dtype=np.float32)
    ramp = (slope * idx + np.float32(0.5)).astype(np.uint8)  # fast_roundf for positive values
    img = np.tile(ramp, (H, 1)) if axis == 0 else np.tile(ramp[:, None], (1, W))
    for use_ref in ([False, True] if oracle.have_ref() else [False]):
        gx, gy = oracle.central_gradient_u8(img, use_ref=use_ref)
        f = img.astype(np.float32)
        assert np.allclose(gx[:, 1:-1], 0.5 * (f[:, 2:] - f[:, :-2]), atol=1e-6)
        assert np.allclose(gx[:, 0], f[:, 1] - f[:, 0]) and np.allclose(gx[:, -1], f[:, -1] - f[:, -2])
        assert np.allclose(gy[1:-1, :], 0.5 * (f[2:, :] - f[:-2, :]), atol=1e-6)
        assert np.allclose(gy[0, :], f[1, :] - f[0, :]) and np.allclose(gy[-1, :], f[-1, :] - f[-2, :])


def test_oracle_frame_prep_equals_reference_and_opencv(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(4)
    for shape in ((97, 131), (480, 640), (6, 7)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        ip, gx, gy = oracle.frame_prep(img, 5)
        assert np.array_equal(ip, cv2.copyMakeBorder(img, 5, 5, 5, 5, cv2.BORDER_REFLECT_101))
        g0 = oracle.central_gradient_u8(img)
        if oracle.have_ref():  # utils/image_utils.cc compiled verbatim
            r0 = oracle.central_gradient_u8(img, use_ref=True)
            assert all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(g0, r0))
        assert np.array_equal(gx, cv2.copyMakeBorder(g0[0], 5, 5, 5, 5, cv2.BORDER_CONSTANT, value=0))
        assert np.array_equal(gy, cv2.copyMakeBorder(g0[1], 5, 5, 5, 5, cv2.BORDER_CONSTANT, value=0))


def test_oracle_clip_line_equals_reference_verbatim(oracle):
    _need_ref(oracle)
    rng = np.random.default_rng(2)
    for _ in range(3000):
        v = [float(np.float32(x)) for x in rng.uniform(-80, 720, 4)]
        o, r = oracle.clip_line(1, 639, 1, 479, *v), oracle.ref_clip_line(1, 639, 1, 479, *v)
        assert o[0] == r[0]
        if o[0]:
            assert np.array_equal(np.float32(o[1]).view(np.uint32), np.float32(r[1]).view(np.uint32))


@pytest.mark.parametrize("cfg", ["A", "C"])
def test_oracle_filters_and_normals_equal_reference_verbatim(oracle, cfg):
    """obliqueTriangleFilter / edgeLengthFilter / idepthTriangleFilter / getVertexNormals sliced verbatim out
    of the reference's flame.cc (oracle/ref_capi/slice_flame.py) and compiled against oracle/shim vs the C
    restatement: identical validity vectors and bit-identical normals, on smooth, noisy and partly
    non-positive idepths, default and non-default thresholds."""
    _need_ref(oracle)
    g = synth.config_graph(cfg)
    W, H = g["W"], g["H"]
    K, Kinv = synth.camera(W, H)
    rng = np.random.default_rng(31)
    V = g["pos"].shape[0]
    base = synth.scene_idepth(g["pos"][:, 0], g["pos"][:, 1], W, H).astype(np.float32)
    cases = [base, (base + rng.normal(0, 0.05, V)).astype(np.float32), g["data_term"].copy()]
    cases[2][::37] = 0.004
    cases.append(cases[1].copy())
    cases[3][::29] = 0.0
    cases[3][5::31] = -0.1
    n_removed = 0
    for k, idepths in enumerate(cases):
        for variant in range(3):
            p = oracle.tri_filter_default_params()
            if variant == 1:
                p.oblique_normal_thresh, p.oblique_idepth_diff_factor, p.edge_length_thresh = 1.1, 0.2, 0.05
                p.min_triangle_idepth = 0.3
            if variant == 2:
                p.do_oblique_triangle_filter = 0
            if k == 3 and p.do_oblique_triangle_filter:
                continue  # max_id == 0 -> FLAME_ASSERT-free but 0/0: the reference would divide by zero here too
            want = oracle.ref_triangle_filters(p, Kinv, W, H, g["tris"], g["pos"], idepths)
            got, _ = oracle.triangle_filters(p, Kinv, W, g["tris"], g["pos"], idepths)
            assert np.array_equal(got, want), (cfg, k, variant, int((got != want).sum()))
            n_removed += int((want == 0).sum())
        a = oracle.vertex_normals(Kinv, g["pos"], idepths, g["tris"])
        b = oracle.ref_vertex_normals(Kinv, g["pos"], idepths, g["tris"])
        an, bn = np.isnan(a), np.isnan(b)
        assert np.array_equal(an, bn) and np.array_equal(a[~an].view(np.uint32), b[~bn].view(np.uint32)), (cfg, k)
    assert n_removed > 0


def test_display_filter_restatement_sanity(oracle):
    """flame.cc:2207-2356: hand-checkable cases of the restatement."""
    from flame_b200 import synth
    K, Kinv = synth.camera(640, 480)
    pos = np.array([[100, 100], [140, 100], [100, 140], [400, 100], [100, 300]], np.float32)
    tris = np.array([[0, 1, 2], [0, 3, 4], [0, 1, 2]], np.int32)
    p = oracle.tri_filter_default_params()
    # fronto-parallel plane: the inward normal is the viewing direction up to the pixel's bearing
    v, a = oracle.triangle_filters(p, Kinv, 640, tris, pos, np.full(5, 0.5, np.float32))
    assert v.tolist() == [1, 0, 1]            # the second triangle has a 300-px edge > 0.333 * 640
    for t in range(2):  # angle between the optical axis and the ray through the triangle's centroid
        cx, cy = pos[tris[t]].mean(0)
        want = np.arctan(np.hypot((cx - 320.0) / 525.0, (cy - 240.0) / 525.0))
        assert abs(a[t] - want) < 1e-4
    # one corner much closer: relative and absolute idepth spread exceed 0.35 / 0.1
    v, _ = oracle.triangle_filters(p, Kinv, 640, tris, pos, np.array([0.5, 0.9, 0.5, 0.5, 0.5], np.float32))
    assert v.tolist() == [0, 0, 0]
    # mean idepth below min_triangle_idepth
    v, _ = oracle.triangle_filters(p, Kinv, 640, tris, pos, np.full(5, 0.005, np.float32))
    assert v.tolist() == [0, 0, 0]
    p.do_idepth_triangle_filter = 0
    v, _ = oracle.triangle_filters(p, Kinv, 640, tris, pos, np.full(5, 0.005, np.float32))
    assert v.tolist() == [1, 0, 1]
    img = np.full((4, 5), np.nan, np.float32)
    img[1, 2] = 0.0
    img[3, 4] = 7.0
    assert oracle.coverage(img) == 2


def test_vertex_normals_restatement_sanity(oracle):
    """flame.cc:2554-2640, restatement: a fronto-parallel mesh has outward normals
    (0, 0, -1); a plane tilted about the vertical axis tilts them in x only."""
    from flame_b200 import synth
    g = synth.config_graph("A")
    K, Kinv = synth.camera(g["W"], g["H"])
    V = g["pos"].shape[0]
    n = oracle.vertex_normals(Kinv, g["pos"], np.full(V, 0.5, np.float32), g["tris"])
    assert np.allclose(n, [0, 0, -1], atol=1e-4)
    # plane z = z0 + s * X  <=>  idepth(u) = (1 - s * (u - cx) / f) / z0 : normal proportional to (s, 0, -1)
    s_, z0 = 0.3, 2.0
    idepth = ((1.0 - s_ * (g["pos"][:, 0] - 320.0) / 525.0) / z0).astype(np.float32)
    n = oracle.vertex_normals(Kinv, g["pos"], idepth, g["tris"])
    want = np.array([s_, 0.0, -1.0]) / np.hypot(s_, 1.0)
    assert np.allclose(n, want, atol=2e-3)
    idepth[::3] = 0.0  # every triangle touching such a vertex is skipped
    n = oracle.vertex_normals(Kinv, g["pos"], idepth, g["tris"])
    assert np.all(n[::3] == 0)


def test_oracle_raster_equals_reference_on_random_soups(oracle):
    """Arbitrary triangle lists - overlapping, both windings, collinear, repeated vertices, lines touching the
    right border (sweep overshoot), with validity masks and a pre-initialised image: the restatement and
    rasterization.cc / interpolateMesh compiled verbatim agree pixel for pixel."""
    _need_ref(oracle)
    for c in synth.triangle_soups():
        W, H, base = c["W"], c["H"], c["base"]
        a = oracle.interpolate_mesh(c["tris"], c["pos"], c["vals"], H, W, c["vv"], c["tv"],
                                    None if base is None else base.copy())
        b = oracle.interpolate_mesh(c["tris"], c["pos"], c["vals"], H, W, c["vv"], c["tv"],
                                    None if base is None else base.copy(), use_ref=True)
        assert np.array_equal(np.isnan(a), np.isnan(b)), c["case"]
        m = ~np.isnan(a)
        assert np.array_equal(a[m].view(np.uint32), b[m].view(np.uint32)), c["case"]


def test_oracle_pd_equals_reference_on_random_graphs(oracle):
    """Random multigraphs (parallel edges, isolated vertices, mixed orientations), random positions,
    weights, parameters and a random warm state: restatement == nltgv2_l1_graph_regularizer.cc verbatim."""
    _need_ref(oracle)
    rng = np.random.default_rng(321)
    for case in range(12):
        V, E = int(rng.integers(2, 60)), int(rng.integers(0, 200))
        pos = rng.uniform(0, 300, (V, 2)).astype(np.float32)
        src = rng.integers(0, V, E)
        dst = (src + rng.integers(1, V, E)) % V            # no self loops (Delaunay never produces them)
        edges = np.ascontiguousarray(np.stack([src, dst], 1).astype(np.int32)).reshape(-1, 2)
        g = {"pos": pos, "edges": edges,
             "data_term": rng.uniform(0.0, 2.0, V).astype(np.float32),
             "data_weight": rng.uniform(0.0, 3.0, V).astype(np.float32)}
        a = synth.init_state(g, rng.uniform(0.01, 2.0, E).astype(np.float32),
                             rng.uniform(0.1, 2.0, E).astype(np.float32))
        for k in ("x", "w1", "w2", "x_bar", "w1_bar", "w2_bar", "x_prev", "w1_prev", "w2_prev"):
            a[k] = rng.normal(0.5, 0.5, V).astype(np.float32)
        for k in ("q1", "q2", "q3"):
            a[k] = rng.uniform(-1, 1, E).astype(np.float32)
        b = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
        p = oracle.pd_default_params()
        p.data_factor, p.step_x, p.step_q, p.theta = (float(rng.uniform(0.05, 1.0)), float(rng.uniform(1e-3, 0.05)),
                                                      float(rng.uniform(1.0, 200.0)), float(rng.uniform(0.0, 1.0)))
        p.x_min, p.x_max = 0.05, 1.8
        n = int(rng.integers(1, 40))
        oracle.pd_run(a, p, n)
        oracle.ref_pd_run(b, p, n)
        for k in oracle.GRAPH_STATE:
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), (case, k)
        assert oracle.pd_costs(a, p) == oracle.ref_pd_costs(b, p), case


def _sync_scene(cfg="A", t_world=(0.10, 0.02, 0.05), yaw=1.5):
    W, H, V, _ = synth.CONFIGS[cfg]
    K, Kinv, frames = synth.make_stereo_pair(W, H, t_world=t_world, yaw_deg=yaw)
    g = synth.config_graph(cfg)
    pad = synth.PAD
    gx = np.ascontiguousarray(frames[1]["gradx_pad"][pad:-pad, pad:-pad])
    gy = np.ascontiguousarray(frames[1]["grady_pad"][pad:-pad, pad:-pad])
    return W, H, K, Kinv, frames, g, gx, gy


@pytest.mark.parametrize("grad_check,letterbox", [(False, False), (True, False), (True, True)])
def test_oracle_project_graph_equals_reference_verbatim(oracle, grad_check, letterbox):
    """Flame::projectGraph sliced verbatim out of flame.cc (Boost graph of V vertices, utils::Frame gradients)
    vs the restatement: same removal set, bit-identical projected state of the surviving vertices."""
    _need_ref(oracle)
    W, H, K, Kinv, frames, g, gx, gy = _sync_scene()
    rng = np.random.default_rng(3)
    x = synth.scene_idepth(g["pos"][:, 0], g["pos"][:, 1], W, H).astype(np.float32)
    x[::31] = 0.0          # idepth 0 -> max-depth projection (a negative idepth trips the reference's FLAME_ASSERT)
    x[5::47] = 1.9         # close points: behind the camera after the backward motion below
    q = synth.quat_from_yaw(2.0).astype(np.float32)
    for scale, t in ((1.0, np.array([0.3, -0.05, 0.1], np.float32)),     # pushes vertices out of the valid region
                     (0.7, np.array([0.05, 0.0, -0.9], np.float32))):    # and some behind the camera (idepth_new < 0)
        rules = oracle.project_rules(grad_check, letterbox, 6.0)
        pos1, x1, id1, rem1 = oracle.project_graph(K, Kinv, q, t, W, H, rules, scale, g["pos"], x, gx, gy)
        pos2, x2, rem2 = oracle.ref_project_graph(K, Kinv, q, t, W, H, rules, scale, g["pos"], x, gx, gy)
        assert np.array_equal(rem1, rem2)
        keep = rem1 == 0
        assert 0 < keep.sum() < len(keep)
        assert np.array_equal(pos1[keep].view(np.uint32), pos2[keep].view(np.uint32))
        assert np.array_equal(x1[keep].view(np.uint32), x2[keep].view(np.uint32))
        if scale != 1.0:
            assert (id1 < 0).any()
        if grad_check:  # the gradient rule removes vertices the geometric rules keep
            _, _, _, rem0 = oracle.project_graph(K, Kinv, q, t, W, H, oracle.project_rules(False, letterbox), scale,
                                                 g["pos"], x, gx, gy)
            assert rem1.sum() > rem0.sum()


@pytest.mark.parametrize("win,letterbox", [(16, False), (8, False), (16, True), (20, False)])
def test_oracle_detect_features_equals_reference_verbatim(oracle, win, letterbox):
    """Flame::detectFeatures sliced verbatim vs the restatement: the same detections in the same order."""
    _need_ref(oracle)
    W, H, K, Kinv, frames, g, _, _ = _sync_scene()
    pad = synth.PAD
    gx = np.ascontiguousarray(frames[0]["gradx_pad"][pad:-pad, pad:-pad])
    gy = np.ascontiguousarray(frames[0]["grady_pad"][pad:-pad, pad:-pad])
    curr = g["pos"][:150]
    p = oracle.detect_params(win, letterbox, 5.0)
    a, na = oracle.detect_features(K, Kinv, frames[0]["pose"], frames[1]["pose"], W, H, p, gx, gy, curr)
    b, nb = oracle.ref_detect_features(K, Kinv, frames[0]["pose"], frames[1]["pose"], W, H, p, gx, gy, curr)
    assert na == nb and na > 50
    assert np.array_equal(a, b)
    # masked cells hold no detection
    cells = set((int(x / win), int(y / win)) for x, y in curr)
    assert not any((int(x / win), int(y / win)) in cells for x, y in a)
