"""GPU parity: epipolar search + measurement + fusion (Flame::updateFeatureIDepths loop body,
flame.cc:1280-1752) vs the CPU oracle, per feature, through the C ABI. Status codes, stages and
counters must match exactly; floats bit-for-bit (NaN == NaN)."""
import numpy as np
import pytest

from flame_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _same_float(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(
        a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32))


def _oracle_params(oracle, gp):
    """Copy the ctypes param struct field-by-field into the oracle's (same layout, separate types)."""
    import ctypes
    op = oracle.track_default_params()
    ctypes.memmove(ctypes.byref(op), ctypes.byref(gp), ctypes.sizeof(op))
    return op


def _run_both(ctx, oracle, W, H, K, Kinv, frames, feats, fnew, curr_pf, params=None):
    gp = _lib.default_track_params() if params is None else params
    ctx.set_camera(W, H, K, Kinv)
    for i, fr in enumerate(frames):
        ctx.frame_upload(i, fr["img_pad"], fr["gradx_pad"], fr["grady_pad"], synth.PAD, fr["pose"])
    got = feats.copy()
    n_got, d_got = ctx.update_feature_idepths(gp, fnew, curr_pf, got)
    want = feats.copy()
    n_want, d_want = oracle.update_feature_idepths(_oracle_params(oracle, gp), K, Kinv, W, H, synth.PAD,
                                                   frames, fnew, curr_pf, want)
    return (n_got, got, d_got), (n_want, want, d_want)


def _assert_equal(g, w):
    (n_got, got, d_got), (n_want, want, d_want) = g, w
    assert n_got == n_want
    assert np.array_equal(d_got["stage"], d_want["stage"])
    for k in ("id", "frame_id", "valid", "num_updates", "num_dropouts", "search_status"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("x", "y", "idepth_mu", "idepth_var"):
        assert _same_float(got[k], want[k]), k
    for k in ("flow_x", "flow_y", "mu_meas", "var_meas"):
        assert _same_float(d_got[k], d_want[k]), k


@pytest.mark.parametrize("cfg", ["A", "C"])
def test_update_feature_idepths_matches_oracle(gpu_ctx, oracle, cfg):
    W, H, V, _ = synth.CONFIGS[cfg]
    K, Kinv, frames = synth.make_stereo_pair(W, H)
    g = synth.config_graph(cfg)
    feats = synth.make_features(g["pos"], W, H)
    got, want = _run_both(gpu_ctx, oracle, W, H, K, Kinv, frames, feats, 1, 0)
    _assert_equal(got, want)
    stages = np.bincount(want[2]["stage"], minlength=10)
    assert stages[9] > 0.6 * V            # most features are measured and fused
    assert stages[6] > 0 and stages[4] >= 0  # and the failure paths are exercised too
    # the measurement is accurate where it succeeds (known synthetic scene)
    ok = want[2]["stage"] == 9
    rho = synth.scene_idepth(feats["x"], feats["y"], W, H)
    assert np.median(np.abs(got[2]["mu_meas"][ok] - rho[ok])) < 5e-3


def test_failure_and_reparent_paths(gpu_ctx, oracle):
    """Forward motion (t_z != 0) with wrong priors drives the rescale-factor gate (re-parenting to the
    current pose-frame, flame.cc:1591-1659), predict failures, tiny baselines, the chi-square gate and
    the dropout / variance bookkeeping."""
    W, H, V, _ = synth.CONFIGS["A"]
    K, Kinv, frames = synth.make_stereo_pair(W, H, t_world=(0.05, 0.01, 0.45), yaw_deg=0.5)
    # third frame: the "current pose-frame", between the two
    K2, _, fr2 = synth.make_stereo_pair(W, H, t_world=(0.02, 0.0, 0.2), yaw_deg=0.2)
    frames = [frames[0], frames[1], fr2[1]]
    g = synth.config_graph("A")
    feats = synth.make_features(g["pos"], W, H, rel_sigma=0.3, var=0.05)
    rng = np.random.default_rng(9)
    feats["idepth_mu"][::9] = 1.9            # near points: rescale factor out of range
    feats["idepth_mu"][4::17] = 0.0          # mu == 0 path of predict / project
    feats["idepth_var"][2::13] = 0.24        # one failure away from idepth_var_max
    feats["num_dropouts"][5::19] = 5         # one failure away from max_dropouts
    feats["frame_id"][7::23] = 1             # parented to the new frame itself: zero baseline -> skipped
    feats["x"][11::29] = 2.0                 # outside the valid region
    got, want = _run_both(gpu_ctx, oracle, W, H, K, Kinv, frames, feats, 1, 2)
    _assert_equal(got, want)
    stages = np.bincount(want[2]["stage"], minlength=10)
    assert stages[0] > 0 and stages[2] + stages[3] > 0 and stages[9] > 0
    assert (want[1]["valid"] == 0).any()


def test_valid_region_border_rounding(gpu_ctx, oracle):
    """Features within +-0.6 px of the valid-region border: cv::Rect::contains tests the ROUNDED point."""
    from test_oracle_kat import _border_features
    W, H, V, _ = synth.CONFIGS["A"]
    K, Kinv, frames = synth.make_stereo_pair(W, H)
    feats = _border_features(W, H)
    got, want = _run_both(gpu_ctx, oracle, W, H, K, Kinv, frames, feats, 1, 0)
    _assert_equal(got, want)
    assert (want[2]["stage"] == 5).any()


def test_nondefault_params(gpu_ctx, oracle):
    W, H, V, _ = synth.CONFIGS["A"]
    K, Kinv, frames = synth.make_stereo_pair(W, H, t_world=(-0.08, 0.03, 0.0), yaw_deg=-0.7)  # t_z == 0 branch
    g = synth.config_graph("A")
    feats = synth.make_features(g["pos"], W, H, rel_sigma=0.1, var=0.02)
    p = _lib.default_track_params()
    p.fparams.search_sigma = 3.0
    p.fparams.epilength_max = 20.0
    p.fparams.min_grad_mag = 8.0
    p.fparams.sparams.do_subpixel = 0
    p.fparams.sparams.max_cost = 900.0
    p.zparams.pixel_var = 9.0
    p.outlier_sigma_thresh = 2.0
    p.do_meas_fusion = 0
    p.do_letterbox = 1
    got, want = _run_both(gpu_ctx, oracle, W, H, K, Kinv, frames, feats, 1, 0, p)
    _assert_equal(got, want)


def test_empty_and_errors(gpu_ctx):
    W, H, _, _ = synth.CONFIGS["A"]
    K, Kinv, frames = synth.make_stereo_pair(W, H)
    gpu_ctx.set_camera(W, H, K, Kinv)
    for i, fr in enumerate(frames):
        gpu_ctx.frame_upload(i, fr["img_pad"], fr["gradx_pad"], fr["grady_pad"], synth.PAD, fr["pose"])
    n, _ = gpu_ctx.update_feature_idepths(_lib.default_track_params(), 1, 0, np.zeros(0, _lib.FEATURE_DTYPE))
    assert n == 0
    bad = np.zeros(3, _lib.FEATURE_DTYPE)
    bad["frame_id"] = 40  # never uploaded
    with pytest.raises(_lib.FlameB200Error):
        gpu_ctx.update_feature_idepths(_lib.default_track_params(), 1, 0, bad)


def test_frame_prep_on_device(gpu_ctx, oracle):
    """flame_b200_frame_upload_raw: padded image (REFLECT_101) + zero-padded central gradients computed
    by k_frame_prep == utils::Frame::create's level-0 buffers (oracle, pinned on the verbatim
    getCentralGradient and on cv2.copyMakeBorder), bit-exact; the search gives identical results
    whether the frames were prepared on the host or on the device."""
    W, H, V, _ = synth.CONFIGS["A"]
    K, Kinv, frames = synth.make_stereo_pair(W, H)
    gpu_ctx.set_camera(W, H, K, Kinv)
    for i, fr in enumerate(frames):
        gpu_ctx.frame_upload_raw(10 + i, fr["img"], synth.PAD, fr["pose"])
        ip, gx, gy = gpu_ctx.frame_download(10 + i, H, W, synth.PAD)
        wp, wx, wy = oracle.frame_prep(fr["img"], synth.PAD)
        assert np.array_equal(ip, wp)
        assert np.array_equal(gx.view(np.uint32), wx.view(np.uint32))
        assert np.array_equal(gy.view(np.uint32), wy.view(np.uint32))
        gpu_ctx.frame_upload(i, fr["img_pad"], fr["gradx_pad"], fr["grady_pad"], synth.PAD, fr["pose"])
    g = synth.config_graph("A")
    feats = synth.make_features(g["pos"], W, H)
    a = feats.copy()
    b = feats.copy()
    b["frame_id"] = 10
    na, da = gpu_ctx.update_feature_idepths(_lib.default_track_params(), 1, 0, a)
    nb, db = gpu_ctx.update_feature_idepths(_lib.default_track_params(), 11, 10, b)
    assert na == nb and np.array_equal(da["stage"], db["stage"])
    for k in ("x", "y", "idepth_mu", "idepth_var"):
        assert _same_float(a[k], b[k])
    # odd sizes / small images
    rng = np.random.default_rng(1)
    with _lib.Context(0, 1) as small:
        small.set_camera(37, 23, K, Kinv)
        img = rng.integers(0, 256, (23, 37), dtype=np.uint8)
        small.frame_upload_raw(20, img, 3, frames[0]["pose"])
        ip, gx, gy = small.frame_download(20, 23, 37, 3)
        wp, wx, wy = oracle.frame_prep(img, 3)
        assert np.array_equal(ip, wp) and np.array_equal(gx, wx) and np.array_equal(gy, wy)


@pytest.mark.parametrize("cfg,win,letterbox", [("A", 16, False), ("A", 8, True), ("C", 16, False), ("A", 20, False)])
def test_detect_features_matches_oracle(gpu_ctx, oracle, cfg, win, letterbox):
    """flame_b200_detect_features vs the oracle's grid detector (pinned on Flame::detectFeatures sliced verbatim):
    identical detections in identical (row-major cell) order, masked cells honoured, truncation reported."""
    W, H, V, _ = synth.CONFIGS[cfg]
    K, Kinv, frames = synth.make_stereo_pair(W, H)
    pad = synth.PAD
    gx = np.ascontiguousarray(frames[0]["gradx_pad"][pad:-pad, pad:-pad])
    gy = np.ascontiguousarray(frames[0]["grady_pad"][pad:-pad, pad:-pad])
    curr = synth.config_graph(cfg)["pos"][:200]
    gpu_ctx.set_camera(W, H, K, Kinv)
    for i, fr in enumerate(frames):
        gpu_ctx.frame_upload(i, fr["img_pad"], fr["gradx_pad"], fr["grady_pad"], pad, fr["pose"])
    gp = _lib.DetectParams(1.4, 5, int(letterbox), 5.0, win)
    want, n_want = oracle.detect_features(K, Kinv, frames[0]["pose"], frames[1]["pose"], W, H,
                                          oracle.detect_params(win, letterbox, 5.0), gx, gy, curr)
    got, n_got = gpu_ctx.detect_features(gp, 0, 1, curr)
    assert n_got == n_want and n_want > 100
    assert np.array_equal(got, want)
    few, n_few = gpu_ctx.detect_features(gp, 0, 1, curr, max_features=7)   # truncated output, full count
    assert n_few == n_want and np.array_equal(few, want[:7])
    none, n_none = gpu_ctx.detect_features(gp, 0, 1, np.zeros((0, 2), np.float32))  # nothing masked
    w2, n2 = oracle.detect_features(K, Kinv, frames[0]["pose"], frames[1]["pose"], W, H,
                                    oracle.detect_params(win, letterbox, 5.0), gx, gy, np.zeros((0, 2), np.float32))
    assert n_none == n2 >= n_want and np.array_equal(none, w2)
