"""ctypes bindings for the CPU oracle (oracle/libflame_oracle.so) and, when built, for the
reference sources compiled verbatim (oracle/_ref/libflame_ref.so).

TEST INFRASTRUCTURE ONLY. May be imported from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py — never from flame_b200/ (the product).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "libflame_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libflame_ref.so")

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)


def build(force=False):
    """Compile the oracle's C restatement (and oracle/_ref when /root/reference is present)."""
    src = os.path.join(_HERE, "flame_oracle.c")
    if force or not os.path.exists(_ORACLE_SO) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fPIC", "-shared", "-o", _ORACLE_SO, src, "-lm"])
    if os.path.isdir("/root/reference/src/flame"):
        newest = max(os.path.getmtime(os.path.join(dp, f))
                     for dp, _, fs in os.walk(_HERE) if "_ref" not in dp
                     for f in fs if f.endswith((".cc", ".hpp", ".h", ".sh")) or "shim" in dp)
        if force or not os.path.exists(_REF_SO) or os.path.getmtime(_REF_SO) < newest:
            subprocess.check_call([os.path.join(_HERE, "build_ref.sh"), "/root/reference"])


def _f(a):
    return a.ctypes.data_as(f32p)


def _i(a):
    return a.ctypes.data_as(i32p)


def _u8(a):
    return None if a is None else a.ctypes.data_as(u8p)


class PDParams(C.Structure):
    _fields_ = [("data_factor", C.c_float), ("step_x", C.c_float), ("step_q", C.c_float),
                ("theta", C.c_float), ("x_min", C.c_float), ("x_max", C.c_float)]


class _Graph(C.Structure):
    _fields_ = [("V", C.c_int32), ("E", C.c_int32), ("pos", f32p),
                ("x", f32p), ("w1", f32p), ("w2", f32p),
                ("x_bar", f32p), ("w1_bar", f32p), ("w2_bar", f32p),
                ("x_prev", f32p), ("w1_prev", f32p), ("w2_prev", f32p),
                ("data_term", f32p), ("data_weight", f32p), ("edges", i32p),
                ("alpha", f32p), ("beta", f32p), ("q1", f32p), ("q2", f32p), ("q3", f32p)]


class EpiGeo(C.Structure):
    _fields_ = [("K", C.c_float * 9), ("Kinv", C.c_float * 9), ("q", C.c_float * 4),
                ("t", C.c_float * 3), ("t_cmp_to_ref", C.c_float * 3), ("KRKinv", C.c_float * 9),
                ("KRKinv3", C.c_float * 3), ("Kt", C.c_float * 3), ("epipole", C.c_float * 2)]


class LSParams(C.Structure):
    _fields_ = [("max_cost", C.c_float), ("do_subpixel", C.c_int), ("sample_dist", C.c_float),
                ("second_best_factor", C.c_float)]


class FilterParams(C.Structure):
    _fields_ = [("win_size", C.c_int), ("search_sigma", C.c_float), ("min_grad_mag", C.c_float),
                ("idepth_min", C.c_float), ("idepth_max", C.c_float),
                ("epilength_min", C.c_float), ("epilength_max", C.c_float),
                ("process_var_factor", C.c_float), ("process_fail_var_factor", C.c_float),
                ("sparams", LSParams)]


class MeasParams(C.Structure):
    _fields_ = [("win_size", C.c_int), ("pixel_var", C.c_float), ("epipolar_line_var", C.c_float)]


class TrackParams(C.Structure):
    _fields_ = [("fparams", FilterParams), ("zparams", MeasParams), ("min_baseline", C.c_float),
                ("rescale_factor_min", C.c_float), ("rescale_factor_max", C.c_float),
                ("idepth_var_max", C.c_float), ("max_dropouts", C.c_int),
                ("outlier_sigma_thresh", C.c_float), ("do_meas_fusion", C.c_int),
                ("do_letterbox", C.c_int)]


class Frame(C.Structure):
    _fields_ = [("pose", C.c_float * 7), ("img_pad", u8p), ("gradx_pad", f32p), ("grady_pad", f32p)]


# numpy mirror of oracle_feature / oracle_track_diag (and of the product's flame_b200_feature)
FEATURE_DTYPE = np.dtype([("id", "<u4"), ("frame_id", "<u4"), ("x", "<f4"), ("y", "<f4"),
                          ("idepth_mu", "<f4"), ("idepth_var", "<f4"), ("valid", "<i4"),
                          ("num_updates", "<u4"), ("num_dropouts", "<u4"), ("search_status", "<i4")])
DIAG_DTYPE = np.dtype([("flow_x", "<f4"), ("flow_y", "<f4"), ("mu_meas", "<f4"),
                       ("var_meas", "<f4"), ("stage", "<i4")])

_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_ORACLE_SO):
            build()
        _lib = C.CDLL(_ORACLE_SO)
        _lib.oracle_pd_smoothness_cost.restype = C.c_float
        _lib.oracle_pd_data_cost.restype = C.c_float
        _lib.oracle_bilinear_u8.restype = C.c_float
        _lib.oracle_bilinear_u8.argtypes = [C.c_int, C.c_int, C.c_int, u8p, C.c_float, C.c_float]
        _lib.oracle_bilinear_f32.restype = C.c_float
        _lib.oracle_bilinear_f32.argtypes = [C.c_int, C.c_int, C.c_int, f32p, C.c_float, C.c_float]
        _lib.oracle_clip_line_liang_barsky.argtypes = [C.c_float] * 8 + [f32p] * 4
        _lib.oracle_draw_shaded_triangle_barycentric.argtypes = (
            [C.c_int] * 6 + [C.c_float] * 3 + [f32p, C.c_int, C.c_int])
        gp = C.POINTER(EpiGeo)
        _lib.oracle_epigeo_project.argtypes = [gp, C.c_float, C.c_float, C.c_float, f32p, f32p]
        _lib.oracle_epigeo_project_idepth.argtypes = [gp, C.c_float, C.c_float, C.c_float, f32p, f32p, f32p]
        _lib.oracle_epigeo_max_depth_projection.argtypes = [gp, C.c_float, C.c_float, f32p, f32p]
        _lib.oracle_epigeo_min_depth_projection.argtypes = [gp, C.c_float, C.c_float, f32p, f32p]
        _lib.oracle_epigeo_epiline.argtypes = [gp, C.c_float, C.c_float, f32p, f32p, f32p, f32p]
        _lib.oracle_epigeo_reference_epiline.argtypes = [gp, C.c_float, C.c_float, f32p, f32p]
        _lib.oracle_epigeo_disparity.restype = C.c_float
        _lib.oracle_epigeo_disparity.argtypes = [gp] + [C.c_float] * 4 + [f32p] * 4
        _lib.oracle_epigeo_disparity_to_depth.restype = C.c_float
        _lib.oracle_epigeo_disparity_to_depth.argtypes = [gp] + [C.c_float] * 7
        _lib.oracle_epigeo_disparity_to_idepth.restype = C.c_float
        _lib.oracle_epigeo_disparity_to_idepth.argtypes = [gp] + [C.c_float] * 7
        _lib.oracle_filter_update.argtypes = [C.c_float] * 4 + [f32p, f32p, C.c_float]
    return _lib


def ref():
    """The reference sources compiled verbatim (oracle/_ref). None if not built."""
    global _ref
    if _ref is None and os.path.exists(_REF_SO):
        _ref = C.CDLL(_REF_SO)
    return _ref


def have_ref():
    return ref() is not None


# ---------------------------------------------------------------------------- PD solver
def pd_default_params():
    p = PDParams()
    lib().oracle_pd_default_params(C.byref(p))
    return p


GRAPH_STATE = ("x", "w1", "w2", "x_bar", "w1_bar", "w2_bar", "x_prev", "w1_prev", "w2_prev",
               "q1", "q2", "q3")


def _graph_struct(g):
    s = _Graph()
    s.V = g["pos"].shape[0]
    s.E = g["edges"].shape[0]
    s.pos = _f(g["pos"])
    s.edges = _i(g["edges"])
    for k in GRAPH_STATE + ("data_term", "data_weight", "alpha", "beta"):
        setattr(s, k, _f(g[k]))
    return s


def pd_run(g, params, n_iters):
    """In-place: step() x n_iters on graph dict g (contiguous float32/int32 numpy arrays)."""
    s = _graph_struct(g)
    lib().oracle_pd_run(C.byref(params), C.byref(s), C.c_int(n_iters))
    return g


def pd_costs(g, params):
    s = _graph_struct(g)
    return (float(lib().oracle_pd_smoothness_cost(C.byref(params), C.byref(s))),
            float(lib().oracle_pd_data_cost(C.byref(params), C.byref(s))))


def edge_weights(edges, pos):
    E = edges.shape[0]
    alpha = np.empty(E, np.float32)
    beta = np.empty(E, np.float32)
    lib().oracle_edge_weights(C.c_int32(E), _i(edges), _f(pos), _f(alpha), _f(beta))
    return alpha, beta


# ---------------------------------------------------------------------------- dense projection
def draw_triangle(img, p1, p2, p3, v1, v2, v3, use_ref=False):
    rows, cols = img.shape
    fn = (ref().ref_draw_shaded_triangle_barycentric if use_ref
          else lib().oracle_draw_shaded_triangle_barycentric)
    fn.argtypes = [C.c_int] * 6 + [C.c_float] * 3 + [f32p, C.c_int, C.c_int]
    fn(p1[0], p1[1], p2[0], p2[1], p3[0], p3[1], v1, v2, v3, _f(img), rows, cols)
    return img


def interpolate_mesh(tris, pos, values, rows, cols, vtx_valid=None, tri_valid=None, img=None,
                     use_ref=False):
    if img is None:
        img = np.full((rows, cols), np.nan, np.float32)
    if use_ref:
        # the reference's 4-wide sweep may touch up to 3 floats past the buffer: give it slack
        buf = np.empty(rows * cols + 8, np.float32)
        buf[:rows * cols] = img.ravel()
        ref().ref_interpolate_mesh(C.c_int32(tris.shape[0]), _i(tris), _f(pos), _f(values),
                                   _u8(vtx_valid), _u8(tri_valid), _f(buf), rows, cols)
        img[...] = buf[:rows * cols].reshape(rows, cols)
    else:
        lib().oracle_interpolate_mesh(C.c_int32(tris.shape[0]), _i(tris), C.c_int32(pos.shape[0]),
                                      _f(pos), _f(values), _u8(vtx_valid), _u8(tri_valid), _f(img),
                                      rows, cols)
    return img


class TriFilterParams(C.Structure):
    _fields_ = [("do_oblique_triangle_filter", C.c_int32), ("oblique_normal_thresh", C.c_float),
                ("oblique_idepth_diff_factor", C.c_float), ("oblique_idepth_diff_abs", C.c_float),
                ("do_edge_length_filter", C.c_int32), ("edge_length_thresh", C.c_float),
                ("do_idepth_triangle_filter", C.c_int32), ("min_triangle_idepth", C.c_float)]


def tri_filter_default_params():
    p = TriFilterParams()
    lib().oracle_tri_filter_default_params(C.byref(p))
    return p


def triangle_filters(params, Kinv, width, tris, pos, idepths):
    """flame.cc:389-407, C restatement (pinned on ref_triangle_filters): returns (validity u8[T], angle f32[T])."""
    T = tris.shape[0]
    valid = np.empty(T, np.uint8)
    angle = np.empty(T, np.float32)
    Kinv = np.ascontiguousarray(Kinv, np.float32).ravel()
    lib().oracle_triangle_filters(C.byref(params), _f(Kinv), C.c_int(width), C.c_int32(T), _i(tris), _f(pos),
                                  _f(idepths), _u8(valid), _f(angle))
    return valid, angle


def ref_triangle_filters(params, Kinv, width, height, tris, pos, idepths):
    """The reference's obliqueTriangleFilter / edgeLengthFilter / idepthTriangleFilter, sliced verbatim out of
    flame.cc and combined as at flame.cc:389-407 (oracle/_ref): validity u8[T]."""
    T = tris.shape[0]
    valid = np.empty(T, np.uint8)
    Kinv = np.ascontiguousarray(Kinv, np.float32).ravel()
    ref().ref_triangle_filters(C.byref(params), _f(Kinv), C.c_int(width), C.c_int(height), C.c_int32(pos.shape[0]),
                               C.c_int32(T), _i(tris), _f(pos), _f(idepths), _u8(valid))
    return valid


def ref_vertex_normals(Kinv, pos, idepths, tris):
    """The reference's Flame::getVertexNormals (flame.cc:2554-2640), sliced verbatim (oracle/_ref)."""
    V = pos.shape[0]
    out = np.empty((V, 3), np.float32)
    Kinv = np.ascontiguousarray(Kinv, np.float32).ravel()
    ref().ref_vertex_normals(_f(Kinv), C.c_int32(V), _f(pos), _f(idepths), C.c_int32(tris.shape[0]), _i(tris),
                             _f(out))
    return out


def vertex_normals(Kinv, pos, idepths, tris):
    """Flame::getVertexNormals (flame.cc:2554-2640), C restatement (pinned on ref_vertex_normals): [V, 3] float32."""
    V = pos.shape[0]
    out = np.empty((V, 3), np.float32)
    Kinv = np.ascontiguousarray(Kinv, np.float32).ravel()
    lib().oracle_vertex_normals(_f(Kinv), C.c_int32(V), _f(pos), _f(idepths), C.c_int32(tris.shape[0]), _i(tris),
                                _f(out))
    return out


class ProjectRules(C.Structure):
    _fields_ = [("rescale_factor_max", C.c_float), ("win_size", C.c_int32), ("do_letterbox", C.c_int32),
                ("do_grad_check_after_projection", C.c_int32), ("min_grad_mag", C.c_float)]


class DetectParams(C.Structure):
    _fields_ = [("rescale_factor_max", C.c_float), ("win_size", C.c_int32), ("do_letterbox", C.c_int32),
                ("min_grad_mag", C.c_float), ("detection_win_size", C.c_int32)]


def project_rules(grad_check=False, letterbox=False, min_grad_mag=5.0):
    return ProjectRules(1.4, 5, int(letterbox), int(grad_check), min_grad_mag)   # params.h defaults


def detect_params(win=16, letterbox=False, min_grad_mag=5.0):
    return DetectParams(1.4, 5, int(letterbox), min_grad_mag, win)


def _project_graph(fn, K, Kinv, q, t, W, H, rules, scale, pos, x, gradx, grady, with_idepth):
    V = pos.shape[0]
    pos2, x2 = np.ascontiguousarray(pos, np.float32).copy(), np.ascontiguousarray(x, np.float32).copy()
    remove = np.zeros(V, np.uint8)
    K = np.ascontiguousarray(K, np.float32).ravel()
    Kinv = np.ascontiguousarray(Kinv, np.float32).ravel()
    q, t = np.ascontiguousarray(q, np.float32), np.ascontiguousarray(t, np.float32)
    gx = np.ascontiguousarray(gradx, np.float32) if gradx is not None else np.zeros((H, W), np.float32)
    gy = np.ascontiguousarray(grady, np.float32) if grady is not None else np.zeros((H, W), np.float32)
    args = [_f(K), _f(Kinv), _f(q), _f(t), C.c_int(W), C.c_int(H), C.byref(rules), C.c_float(scale), C.c_int32(V),
            _f(pos2), _f(x2), _f(gx), _f(gy)]
    idepth = np.zeros(V, np.float32)
    if with_idepth:
        args.append(_f(idepth))
    args.append(_u8(remove))
    fn(*args)
    return pos2, x2, idepth, remove


def project_graph(K, Kinv, q, t, W, H, rules, scale, pos, x, gradx=None, grady=None):
    """Flame::projectGraph's vertex loop + removal marks, C restatement: (pos', x', idepth_new, remove)."""
    return _project_graph(lib().oracle_project_graph, K, Kinv, q, t, W, H, rules, scale, pos, x, gradx, grady, True)


def ref_project_graph(K, Kinv, q, t, W, H, rules, scale, pos, x, gradx=None, grady=None):
    """The reference's Flame::projectGraph sliced verbatim (oracle/_ref); pos'/x' only for surviving vertices."""
    p, xx, _, rem = _project_graph(ref().ref_project_graph, K, Kinv, q, t, W, H, rules, scale, pos, x, gradx, grady,
                                   False)
    return p, xx, rem


def _detect(fn, K, Kinv, pose_ref, pose_prev, W, H, params, gradx, grady, curr_xy, max_out):
    K = np.ascontiguousarray(K, np.float32).ravel()
    Kinv = np.ascontiguousarray(Kinv, np.float32).ravel()
    curr = np.ascontiguousarray(curr_xy, np.float32).reshape(-1, 2)
    out = np.zeros((max_out, 2), np.float32)
    fn.restype = C.c_int32
    n = fn(_f(K), _f(Kinv), _f(np.ascontiguousarray(pose_ref, np.float32)),
           _f(np.ascontiguousarray(pose_prev, np.float32)), C.c_int(W), C.c_int(H), C.byref(params),
           _f(np.ascontiguousarray(gradx, np.float32)), _f(np.ascontiguousarray(grady, np.float32)),
           _f(curr), C.c_int32(curr.shape[0]), _f(out), C.c_int32(max_out))
    return out[:min(n, max_out)].copy(), n


def detect_features(K, Kinv, pose_ref, pose_prev, W, H, params, gradx, grady, curr_xy, max_out=20000):
    """Flame::detectFeatures' grid detector (flame.cc:1192-1277), C restatement: ((n, 2) features, n)."""
    return _detect(lib().oracle_detect_features, K, Kinv, pose_ref, pose_prev, W, H, params, gradx, grady, curr_xy,
                   max_out)


def ref_detect_features(K, Kinv, pose_ref, pose_prev, W, H, params, gradx, grady, curr_xy, max_out=20000):
    """The reference's Flame::detectFeatures sliced verbatim (oracle/_ref)."""
    return _detect(ref().ref_detect_features, K, Kinv, pose_ref, pose_prev, W, H, params, gradx, grady, curr_xy,
                   max_out)


def coverage(img):
    lib().oracle_coverage.restype = C.c_int64
    img = np.ascontiguousarray(img, np.float32)
    return int(lib().oracle_coverage(C.c_int(img.shape[0]), C.c_int(img.shape[1]), _f(img)))


# ---------------------------------------------------------------------------- Delaunay (reference Triangle)
def ref_delaunay(xy):
    """Shewchuk Triangle exactly as utils/delaunay.cc calls it. Returns (tris, edges, neighbors)."""
    r = ref()
    if r is None:
        raise RuntimeError("oracle/_ref/libflame_ref.so not built (needs /root/reference)")
    xy = np.ascontiguousarray(xy, np.float32)
    nt, ne = C.c_int32(), C.c_int32()
    pt, pe, pn = i32p(), i32p(), i32p()
    r.ref_delaunay(C.c_int32(xy.shape[0]), _f(xy), C.byref(nt), C.byref(pt), C.byref(ne),
                   C.byref(pe), C.byref(pn))
    tris = np.ctypeslib.as_array(pt, (nt.value, 3)).copy()
    edges = np.ctypeslib.as_array(pe, (ne.value, 2)).copy()
    neigh = np.ctypeslib.as_array(pn, (nt.value, 3)).copy()
    r.ref_free.argtypes = [C.c_void_p]
    for p in (pt, pe, pn):
        r.ref_free(C.cast(p, C.c_void_p))
    return tris, edges, neigh


# ---------------------------------------------------------------------------- image helpers
def bilinear_u8(img, x, y):
    return float(lib().oracle_bilinear_u8(img.shape[0], img.shape[1], img.strides[0], _u8(img), x, y))


def bilinear_f32(img, x, y):
    return float(lib().oracle_bilinear_f32(img.shape[0], img.shape[1], img.strides[0], _f(img), x, y))


def clip_line(xmin, xmax, ymin, ymax, x0, y0, x1, y1):
    o = (C.c_float * 4)()
    ok = lib().oracle_clip_line_liang_barsky(xmin, xmax, ymin, ymax, x0, y0, x1, y1,
                                             C.cast(C.byref(o, 0), f32p), C.cast(C.byref(o, 4), f32p),
                                             C.cast(C.byref(o, 8), f32p), C.cast(C.byref(o, 12), f32p))
    return bool(ok), tuple(o)


# ---------------------------------------------------------------------------- epipolar geometry
class Epi:
    """Thin OO wrapper over oracle_epigeo mirroring stereo::EpipolarGeometry<float>."""

    def __init__(self, K, Kinv=None):
        K = np.asarray(K, np.float32).reshape(3, 3)
        if Kinv is None:
            Kinv = np.linalg.inv(K.astype(np.float64)).astype(np.float32)
        self.g = EpiGeo()
        lib().oracle_epigeo_init(C.byref(self.g), _f(np.ascontiguousarray(K)),
                                 _f(np.ascontiguousarray(Kinv, np.float32)))

    def load(self, q_wxyz, t):
        q = np.ascontiguousarray(q_wxyz, np.float32)
        t = np.ascontiguousarray(t, np.float32)
        lib().oracle_epigeo_load(C.byref(self.g), _f(q), _f(t))

    def _two(self, fn, *a):
        ox, oy = C.c_float(), C.c_float()
        fn(C.byref(self.g), *a, C.byref(ox), C.byref(oy))
        return ox.value, oy.value

    def project(self, u, idepth):
        return self._two(lib().oracle_epigeo_project, u[0], u[1], idepth)

    def project_idepth(self, u, idepth):
        ox, oy, oi = C.c_float(), C.c_float(), C.c_float()
        lib().oracle_epigeo_project_idepth(C.byref(self.g), u[0], u[1], idepth, C.byref(ox),
                                           C.byref(oy), C.byref(oi))
        return (ox.value, oy.value), oi.value

    def max_depth_projection(self, u):
        return self._two(lib().oracle_epigeo_max_depth_projection, u[0], u[1])

    def min_depth_projection(self, u):
        return self._two(lib().oracle_epigeo_min_depth_projection, u[0], u[1])

    def epiline(self, u):
        ix, iy, ex, ey = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        lib().oracle_epigeo_epiline(C.byref(self.g), u[0], u[1], C.byref(ix), C.byref(iy),
                                    C.byref(ex), C.byref(ey))
        return (ix.value, iy.value), (ex.value, ey.value)

    def reference_epiline(self, u):
        return self._two(lib().oracle_epigeo_reference_epiline, u[0], u[1])

    def disparity(self, u_ref, u_cmp):
        ix, iy, ex, ey = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        d = lib().oracle_epigeo_disparity(C.byref(self.g), u_ref[0], u_ref[1], u_cmp[0], u_cmp[1],
                                          C.byref(ix), C.byref(iy), C.byref(ex), C.byref(ey))
        return float(d), (ix.value, iy.value), (ex.value, ey.value)

    def disparity_to_depth(self, u_ref, u_inf, epi, disp):
        return float(lib().oracle_epigeo_disparity_to_depth(
            C.byref(self.g), u_ref[0], u_ref[1], u_inf[0], u_inf[1], epi[0], epi[1], disp))

    def disparity_to_idepth(self, u_ref, u_inf, epi, disp):
        return float(lib().oracle_epigeo_disparity_to_idepth(
            C.byref(self.g), u_ref[0], u_ref[1], u_inf[0], u_inf[1], epi[0], epi[1], disp))


def project_world(K, q_wxyz, t, p):
    K = np.ascontiguousarray(K, np.float32)
    q = np.ascontiguousarray(q_wxyz, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    p = np.ascontiguousarray(p, np.float32)
    ox, oy = C.c_float(), C.c_float()
    lib().oracle_epigeo_project_world(_f(K), _f(q), _f(t), _f(p), C.byref(ox), C.byref(oy))
    return ox.value, oy.value


def se3_inv_mul(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    out = np.empty(7, np.float32)
    lib().oracle_se3_inv_mul(_f(a), _f(b), _f(out))
    return out


# ---------------------------------------------------------------------------- per-feature driver
def track_default_params():
    p = TrackParams()
    lib().oracle_track_default_params(C.byref(p))
    return p


def update_feature_idepths(params, K, Kinv, width, height, pad, frames, fnew, curr_pf, feats):
    """frames: list of dicts {pose(7), img_pad u8, gradx_pad f32, grady_pad f32}. feats: FEATURE_DTYPE
    array (updated in place). Returns (n_updates, diag array)."""
    K = np.ascontiguousarray(K, np.float32).ravel()
    Kinv = np.ascontiguousarray(Kinv, np.float32).ravel()
    arr = (Frame * len(frames))()
    keep = []
    for i, fr in enumerate(frames):
        pose = np.ascontiguousarray(fr["pose"], np.float32)
        for k in range(7):
            arr[i].pose[k] = pose[k]
        img = np.ascontiguousarray(fr["img_pad"], np.uint8)
        gx = np.ascontiguousarray(fr["gradx_pad"], np.float32)
        gy = np.ascontiguousarray(fr["grady_pad"], np.float32)
        keep += [img, gx, gy]
        arr[i].img_pad = _u8(img)
        arr[i].gradx_pad = _f(gx)
        arr[i].grady_pad = _f(gy)
    assert feats.dtype == FEATURE_DTYPE and feats.flags["C_CONTIGUOUS"]
    diag = np.zeros(feats.shape[0], DIAG_DTYPE)
    n = lib().oracle_update_feature_idepths(
        C.byref(params), _f(K), _f(Kinv), width, height, pad, arr, len(frames), fnew, curr_pf,
        feats.ctypes.data_as(C.c_void_p), feats.shape[0], diag.ctypes.data_as(C.c_void_p))
    return int(n), diag


# ---------------------------------------------------------------------------- reference (verbatim) arms
class _RefGraph(C.Structure):
    _fields_ = [("V", C.c_int32), ("E", C.c_int32), ("pos", f32p),
                ("x", f32p), ("w1", f32p), ("w2", f32p),
                ("x_bar", f32p), ("w1_bar", f32p), ("w2_bar", f32p),
                ("x_prev", f32p), ("w1_prev", f32p), ("w2_prev", f32p),
                ("data_term", f32p), ("data_weight", f32p), ("edges", i32p),
                ("alpha", f32p), ("beta", f32p), ("q1", f32p), ("q2", f32p), ("q3", f32p)]


def _ref_graph_struct(g):
    s = _RefGraph()
    s.V = g["pos"].shape[0]
    s.E = g["edges"].shape[0]
    s.pos = _f(g["pos"])
    s.edges = _i(g["edges"])
    for k in GRAPH_STATE + ("data_term", "data_weight", "alpha", "beta"):
        setattr(s, k, _f(g[k]))
    return s


def _params6(p):
    return np.array([p.data_factor, p.step_x, p.step_q, p.theta, p.x_min, p.x_max], np.float32)


def ref_pd_run(g, params, n_iters):
    """nltgv2_l1_graph_regularizer.cc compiled verbatim (Boost.Graph container shimmed): step() x n."""
    s = _ref_graph_struct(g)
    p6 = _params6(params)
    ref().ref_pd_run(_f(p6), C.byref(s), C.c_int(n_iters))
    return g


def ref_pd_costs(g, params):
    s = _ref_graph_struct(g)
    p6 = _params6(params)
    a, b = C.c_float(), C.c_float()
    ref().ref_pd_costs(_f(p6), C.byref(s), C.byref(a), C.byref(b))
    return a.value, b.value


def ref_update_feature_idepths(params, K, Kinv, width, height, pad, frames, fnew, curr_pf, feats):
    """The reference's stereo sources compiled verbatim (Eigen/Sophus/OpenCV types shimmed)."""
    K = np.ascontiguousarray(K, np.float32).ravel()
    Kinv = np.ascontiguousarray(Kinv, np.float32).ravel()
    arr = (Frame * len(frames))()
    keep = []
    for i, fr in enumerate(frames):
        pose = np.ascontiguousarray(fr["pose"], np.float32)
        for k in range(7):
            arr[i].pose[k] = pose[k]
        img = np.ascontiguousarray(fr["img_pad"], np.uint8)
        gx = np.ascontiguousarray(fr["gradx_pad"], np.float32)
        gy = np.ascontiguousarray(fr["grady_pad"], np.float32)
        keep += [img, gx, gy]
        arr[i].img_pad = _u8(img)
        arr[i].gradx_pad = _f(gx)
        arr[i].grady_pad = _f(gy)
    diag = np.zeros(feats.shape[0], DIAG_DTYPE)
    n = ref().ref_update_feature_idepths(
        C.byref(params), _f(K), _f(Kinv), width, height, pad, arr, len(frames), fnew, curr_pf,
        feats.ctypes.data_as(C.c_void_p), feats.shape[0], diag.ctypes.data_as(C.c_void_p))
    return int(n), diag


def ref_epigeo_query(K, Kinv, q_wxyz, t, u, idepth):
    """min/max depth projection, epiline (u_inf, epi), project(u, idepth), referenceEpiline."""
    K = np.ascontiguousarray(K, np.float32).ravel()
    Kinv = np.ascontiguousarray(Kinv, np.float32).ravel()
    q = np.ascontiguousarray(q_wxyz, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    out = np.zeros(12, np.float32)
    fn = ref().ref_epigeo_query
    fn.argtypes = [f32p, f32p, f32p, f32p, C.c_float, C.c_float, C.c_float, f32p]
    fn(_f(K), _f(Kinv), _f(q), _f(t), u[0], u[1], idepth, _f(out))
    return out


# ---------------------------------------------------------------------------- frame preparation
def central_gradient_u8(img, use_ref=False):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    gx = np.empty((h, w), np.float32)
    gy = np.empty((h, w), np.float32)
    fn = ref().ref_central_gradient_u8 if use_ref else lib().oracle_central_gradient_u8
    fn(C.c_int(w), C.c_int(h), _u8(img), _f(gx), _f(gy))
    return gx, gy


def frame_prep(img, border):
    """Padded level-0 buffers of utils::Frame::create (img REFLECT_101, gradients zero-padded)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    ip = np.empty((h + 2 * border, w + 2 * border), np.uint8)
    gx = np.empty(ip.shape, np.float32)
    gy = np.empty(ip.shape, np.float32)
    lib().oracle_frame_prep(C.c_int(w), C.c_int(h), C.c_int(border), _u8(img), _u8(ip), _f(gx), _f(gy))
    return ip, gx, gy


def ref_clip_line(xmin, xmax, ymin, ymax, x0, y0, x1, y1):
    out = np.zeros(4, np.float32)
    fn = ref().ref_clip_line
    fn.argtypes = [C.c_float] * 8 + [f32p]
    ok = fn(xmin, xmax, ymin, ymax, x0, y0, x1, y1, _f(out))
    return bool(ok), tuple(float(v) for v in out)
