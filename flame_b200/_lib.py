"""ctypes loader for libflame_b200.so (the C ABI declared in include/flame_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is visible when a
context is created, this module raises — it never routes to the oracle or to numpy.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libflame_b200.so")

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)


class FlameB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("flame_b200 error %d: %s" % (code, msg))
        self.code = code


class PDParams(C.Structure):
    """nltgv2_l1_graph_regularizer::Params (nltgv2_l1_graph_regularizer.h:121-129)."""
    _fields_ = [("data_factor", C.c_float), ("step_x", C.c_float), ("step_q", C.c_float),
                ("theta", C.c_float), ("x_min", C.c_float), ("x_max", C.c_float)]


class Stats(C.Structure):
    """flame_b200_stats: host-side accounting of the frame pipeline."""
    _fields_ = [("frames", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("prep_seconds", C.c_double), ("stage_seconds", C.c_double),
                ("enqueue_seconds", C.c_double), ("wait_seconds", C.c_double),
                ("triangulation_seconds", C.c_double)]


class GraphLayout(C.Structure):
    """flame_b200_graph_layout: what flame_b200_graph_plan reports."""
    _fields_ = [("chunk_width", C.c_int32), ("n_chunks", C.c_int32), ("max_degree", C.c_int32),
                ("cluster", C.c_int32), ("variant", C.c_int32), ("threads", C.c_int32),
                ("max_incidences", C.c_int32), ("max_owned_edges", C.c_int32), ("shared_bytes", C.c_int64)]


class ProjectRules(C.Structure):
    """flame_b200_project_rules: the flame::Params fields Flame::projectGraph reads."""
    _fields_ = [("rescale_factor_max", C.c_float), ("win_size", C.c_int32), ("do_letterbox", C.c_int32),
                ("do_grad_check_after_projection", C.c_int32), ("min_grad_mag", C.c_float)]


class DetectParams(C.Structure):
    """flame_b200_detect_params: the flame::Params fields the grid detector reads."""
    _fields_ = [("rescale_factor_max", C.c_float), ("win_size", C.c_int32), ("do_letterbox", C.c_int32),
                ("min_grad_mag", C.c_float), ("detection_win_size", C.c_int32)]


class GraphIn(C.Structure):
    _fields_ = [("V", C.c_int32), ("E", C.c_int32), ("pos", f32p), ("edges", i32p),
                ("data_term", f32p), ("data_weight", f32p),
                ("x", f32p), ("w1", f32p), ("w2", f32p),
                ("x_bar", f32p), ("w1_bar", f32p), ("w2_bar", f32p),
                ("x_prev", f32p), ("w1_prev", f32p), ("w2_prev", f32p),
                ("alpha", f32p), ("beta", f32p), ("q1", f32p), ("q2", f32p), ("q3", f32p)]


class GraphSync(C.Structure):
    """flame_b200_graph_sync."""
    _fields_ = [("V", C.c_int32), ("E", C.c_int32), ("pos", f32p), ("data_term", f32p),
                ("data_weight", f32p), ("vtx_from", i32p), ("x_init", f32p), ("edges", i32p),
                ("edge_from", i32p), ("check_sticky_obstacles", C.c_int32)]


class TriFilterParams(C.Structure):
    """flame_b200_tri_filter_params (params.h:70-85)."""
    _fields_ = [("do_oblique_triangle_filter", C.c_int32), ("oblique_normal_thresh", C.c_float),
                ("oblique_idepth_diff_factor", C.c_float), ("oblique_idepth_diff_abs", C.c_float),
                ("do_edge_length_filter", C.c_int32), ("edge_length_thresh", C.c_float),
                ("do_idepth_triangle_filter", C.c_int32), ("min_triangle_idepth", C.c_float)]


class GraphOut(C.Structure):
    _fields_ = [(k, f32p) for k in ("x", "w1", "w2", "x_bar", "w1_bar", "w2_bar", "x_prev",
                                    "w1_prev", "w2_prev", "alpha", "beta", "q1", "q2", "q3")]


class LSParams(C.Structure):
    _fields_ = [("max_cost", C.c_float), ("do_subpixel", C.c_int32), ("sample_dist", C.c_float),
                ("second_best_factor", C.c_float)]


class FilterParams(C.Structure):
    _fields_ = [("win_size", C.c_int32), ("search_sigma", C.c_float), ("min_grad_mag", C.c_float),
                ("idepth_min", C.c_float), ("idepth_max", C.c_float),
                ("epilength_min", C.c_float), ("epilength_max", C.c_float),
                ("process_var_factor", C.c_float), ("process_fail_var_factor", C.c_float),
                ("sparams", LSParams)]


class MeasParams(C.Structure):
    _fields_ = [("win_size", C.c_int32), ("pixel_var", C.c_float), ("epipolar_line_var", C.c_float)]


class TrackParams(C.Structure):
    _fields_ = [("fparams", FilterParams), ("zparams", MeasParams), ("min_baseline", C.c_float),
                ("rescale_factor_min", C.c_float), ("rescale_factor_max", C.c_float),
                ("idepth_var_max", C.c_float), ("max_dropouts", C.c_int32),
                ("outlier_sigma_thresh", C.c_float), ("do_meas_fusion", C.c_int32),
                ("do_letterbox", C.c_int32)]


FEATURE_DTYPE = np.dtype([("id", "<u4"), ("frame_id", "<u4"), ("x", "<f4"), ("y", "<f4"),
                          ("idepth_mu", "<f4"), ("idepth_var", "<f4"), ("valid", "<i4"),
                          ("num_updates", "<u4"), ("num_dropouts", "<u4"), ("search_status", "<i4")])
DIAG_DTYPE = np.dtype([("flow_x", "<f4"), ("flow_y", "<f4"), ("mu_meas", "<f4"),
                       ("var_meas", "<f4"), ("stage", "<i4")])

PD_AUTO, PD_STREAM, PD_RESIDENT = 0, 1, 2
PREF_THROUGHPUT, PREF_LATENCY = 0, 1

_lib = None

# every symbol include/flame_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "flame_b200_last_error", "flame_b200_version", "flame_b200_kernel_launches",
    "flame_b200_ctx_create", "flame_b200_ctx_destroy", "flame_b200_ctx_sync", "flame_b200_ctx_stream",
    "flame_b200_ctx_set_async_projection", "flame_b200_ctx_join", "flame_b200_ctx_set_pd_preference",
    "flame_b200_host_alloc", "flame_b200_host_free",
    "flame_b200_pd_default_params", "flame_b200_graph_upload", "flame_b200_graph_init",
    "flame_b200_graph_set_data", "flame_b200_graph_reset", "flame_b200_graph_download",
    "flame_b200_graph_project", "flame_b200_graph_project_marked", "flame_b200_detect_features",
    "flame_b200_graph_resync", "flame_b200_pd_run", "flame_b200_pd_wave",
    "flame_b200_pd_cost", "flame_b200_interpolate_mesh", "flame_b200_mesh_bind",
    "flame_b200_mesh_project", "flame_b200_mesh_download", "flame_b200_process_frame",
    "flame_b200_delaunay", "flame_b200_slot_device_ptrs", "flame_b200_mesh_redirect", "flame_b200_device_alloc",
    "flame_b200_device_free", "flame_b200_ipc_export", "flame_b200_ipc_open", "flame_b200_ipc_close",
    "flame_b200_process_frame_submit", "flame_b200_process_frame_wait", "flame_b200_process_frame_topology",
    "flame_b200_ctx_stats", "flame_b200_graph_plan",
    "flame_b200_tri_filter_default_params", "flame_b200_mesh_filter", "flame_b200_mesh_coverage",
    "flame_b200_mesh_normals", "flame_b200_mesh_filter_angles",
    "flame_b200_track_default_params",
    "flame_b200_set_camera", "flame_b200_frame_upload", "flame_b200_frame_upload_raw",
    "flame_b200_frame_download", "flame_b200_update_feature_idepths",
    "flame_b200_features_upload", "flame_b200_features_track", "flame_b200_features_download",
]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FlameB200Error(-2, "%s not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                                 "flame_b200 has no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.flame_b200_last_error.restype = C.c_char_p
    L.flame_b200_kernel_launches.restype = C.c_uint64
    L.flame_b200_ctx_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    L.flame_b200_ctx_destroy.argtypes = [vp]
    L.flame_b200_ctx_destroy.restype = None
    L.flame_b200_ctx_sync.argtypes = [vp]
    L.flame_b200_ctx_set_async_projection.argtypes = [vp, C.c_int]
    L.flame_b200_ctx_join.argtypes = [vp]
    L.flame_b200_ctx_set_pd_preference.argtypes = [vp, C.c_int]
    L.flame_b200_ctx_stream.argtypes = [vp]
    L.flame_b200_ctx_stream.restype = vp
    L.flame_b200_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.flame_b200_host_free.argtypes = [vp]
    L.flame_b200_pd_default_params.argtypes = [C.POINTER(PDParams)]
    L.flame_b200_pd_default_params.restype = None
    L.flame_b200_graph_upload.argtypes = [vp, C.c_int, C.POINTER(GraphIn)]
    L.flame_b200_graph_init.argtypes = [vp, C.c_int, C.c_int32, f32p, f32p, f32p, C.c_int32, i32p]
    L.flame_b200_graph_set_data.argtypes = [vp, C.c_int, f32p, f32p]
    L.flame_b200_graph_reset.argtypes = [vp, C.c_int, C.c_int]
    L.flame_b200_graph_download.argtypes = [vp, C.c_int, C.POINTER(GraphOut)]
    L.flame_b200_graph_project.argtypes = [vp, C.c_int, f32p, f32p, C.c_float, f32p, f32p]
    L.flame_b200_graph_resync.argtypes = [vp, C.c_int, C.POINTER(GraphSync)]
    L.flame_b200_graph_project_marked.argtypes = [vp, C.c_int, f32p, f32p, C.c_float, C.POINTER(ProjectRules), C.c_int,
                                                  f32p, f32p, u8p]
    L.flame_b200_detect_features.argtypes = [vp, C.POINTER(DetectParams), C.c_int, C.c_int, f32p, C.c_int32, f32p,
                                             C.c_int32, C.POINTER(C.c_int32)]
    L.flame_b200_pd_run.argtypes = [vp, C.c_int, C.c_int, C.POINTER(PDParams), C.c_int, C.c_int]
    L.flame_b200_pd_wave.argtypes = [vp, C.c_int, C.POINTER(C.c_int32)]
    L.flame_b200_pd_cost.argtypes = [vp, C.c_int, C.POINTER(PDParams), C.c_int, f32p, f32p]
    L.flame_b200_interpolate_mesh.argtypes = [vp, C.c_int, C.c_int32, i32p, C.c_int32, f32p, f32p,
                                              C.c_float, u8p, u8p, C.c_int, C.c_int, C.c_int, f32p]
    L.flame_b200_mesh_bind.argtypes = [vp, C.c_int, C.c_int32, i32p, u8p, u8p, C.c_int, C.c_int]
    L.flame_b200_mesh_project.argtypes = [vp, C.c_int, C.c_int, C.c_float]
    L.flame_b200_mesh_download.argtypes = [vp, C.c_int, vp]
    L.flame_b200_delaunay.argtypes = [C.c_int32, f32p, C.c_int32, C.POINTER(C.c_int32), i32p, C.c_int32,
                                      C.POINTER(C.c_int32), i32p]
    L.flame_b200_mesh_redirect.argtypes = [vp, C.c_int, vp, vp]
    L.flame_b200_device_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp)]
    L.flame_b200_device_free.argtypes = [vp]
    L.flame_b200_ipc_export.argtypes = [vp, C.c_char_p]
    L.flame_b200_ipc_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
    L.flame_b200_ipc_close.argtypes = [vp]
    L.flame_b200_slot_device_ptrs.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int32),
                                              C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.flame_b200_tri_filter_default_params.argtypes = [C.POINTER(TriFilterParams)]
    L.flame_b200_tri_filter_default_params.restype = None
    L.flame_b200_mesh_filter.argtypes = [vp, C.c_int, C.POINTER(TriFilterParams), f32p, C.c_float, u8p, C.c_int]
    L.flame_b200_mesh_coverage.argtypes = [vp, C.c_int, C.POINTER(C.c_int64)]
    L.flame_b200_mesh_normals.argtypes = [vp, C.c_int, f32p, C.c_float, f32p]
    L.flame_b200_mesh_filter_angles.argtypes = [vp, C.c_int, f32p, C.c_float, f32p]
    L.flame_b200_process_frame.argtypes = [vp, C.c_int, C.c_int32, f32p, f32p, f32p, C.c_int32, i32p, C.c_int32,
                                           i32p, C.c_int, C.c_int, C.POINTER(PDParams), C.c_int, C.c_float,
                                           vp, f32p, f32p, f32p]
    L.flame_b200_process_frame_submit.argtypes = [vp, C.c_int, C.c_int32, f32p, f32p, f32p, C.c_int32, i32p,
                                                  C.c_int32, i32p, C.c_int, C.c_int, C.POINTER(PDParams), C.c_int,
                                                  C.c_float, vp]
    L.flame_b200_process_frame_wait.argtypes = [vp, C.c_int, f32p, f32p, f32p]
    L.flame_b200_process_frame_topology.argtypes = [vp, C.c_int, C.c_int32, C.POINTER(C.c_int32), i32p, C.c_int32,
                                                    C.POINTER(C.c_int32), i32p]
    L.flame_b200_ctx_stats.argtypes = [vp, C.POINTER(Stats), C.c_int]
    L.flame_b200_graph_plan.argtypes = [C.c_int32, f32p, C.c_int32, i32p, C.c_int, i32p, C.POINTER(GraphLayout),
                                        i32p, C.c_int32, i32p, i32p]
    L.flame_b200_track_default_params.argtypes = [C.POINTER(TrackParams)]
    L.flame_b200_track_default_params.restype = None
    L.flame_b200_set_camera.argtypes = [vp, C.c_int, C.c_int, f32p, f32p]
    L.flame_b200_frame_upload.argtypes = [vp, C.c_int, u8p, f32p, f32p, C.c_int, f32p]
    L.flame_b200_frame_upload_raw.argtypes = [vp, C.c_int, u8p, C.c_int, f32p]
    L.flame_b200_frame_download.argtypes = [vp, C.c_int, u8p, f32p, f32p]
    L.flame_b200_update_feature_idepths.argtypes = [vp, C.POINTER(TrackParams), C.c_int, C.c_int, vp,
                                                    C.c_int32, vp, C.POINTER(C.c_int32)]
    L.flame_b200_features_upload.argtypes = [vp, vp, C.c_int32]
    L.flame_b200_features_track.argtypes = [vp, C.POINTER(TrackParams), C.c_int, C.c_int]
    L.flame_b200_features_download.argtypes = [vp, vp, vp, C.POINTER(C.c_int32), C.c_int]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise FlameB200Error(rc, lib().flame_b200_last_error().decode("utf-8", "replace"))


def fptr(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], "need contiguous float32"
    return a.ctypes.data_as(f32p)


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"], "need contiguous int32"
    return a.ctypes.data_as(i32p)


def bptr(a):
    if a is None:
        return None
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"], "need contiguous uint8"
    return a.ctypes.data_as(u8p)


def delaunay(pos):
    """flame_b200_delaunay: (tris [T,3] int32, edges [E,2] int32) of the points pos [V,2] (host-only, releases the GIL)."""
    pos = np.ascontiguousarray(pos, np.float32)
    V = pos.shape[0]
    tris = np.empty((max(2 * V, 1), 3), np.int32)
    edges = np.empty((max(3 * V, 1), 2), np.int32)
    nt, ne = C.c_int32(), C.c_int32()
    check(lib().flame_b200_delaunay(V, fptr(pos), tris.shape[0], C.byref(nt), iptr(tris), edges.shape[0],
                                    C.byref(ne), iptr(edges)))
    return tris[:nt.value].copy(), edges[:ne.value].copy()


# SMs kept busy by clusters of C CTAs on a B200 (GPCs of 20/20/20/20/18/18/18/14 SMs, one CTA per SM)
B200_BUSY_SMS = np.array([0] + [C_ * sum(n // C_ for n in (20, 20, 20, 20, 18, 18, 18, 14)) for C_ in range(1, 17)],
                         np.int32)


def graph_plan(pos, edges, preference=0, busy_sms=B200_BUSY_SMS):
    """flame_b200_graph_plan (host only): (layout dict, vperm, koff [n_chunks, max_degree + 1], chunk_base)."""
    pos = np.ascontiguousarray(pos, np.float32)
    edges = np.ascontiguousarray(edges, np.int32)
    lay = GraphLayout()
    bs = None if busy_sms is None else np.ascontiguousarray(busy_sms, np.int32)
    check(lib().flame_b200_graph_plan(pos.shape[0], fptr(pos), edges.shape[0], iptr(edges), preference, iptr(bs),
                                      C.byref(lay), None, 0, None, None))
    vperm = np.empty(pos.shape[0], np.int32)
    koff = np.empty((lay.n_chunks, lay.max_degree + 1), np.int32)
    chunk_base = np.empty(lay.n_chunks + 1, np.int32)
    check(lib().flame_b200_graph_plan(pos.shape[0], fptr(pos), edges.shape[0], iptr(edges), preference, iptr(bs),
                                      C.byref(lay), iptr(vperm), koff.size, iptr(koff), iptr(chunk_base)))
    return {k: getattr(lay, k) for k, _ in GraphLayout._fields_}, vperm, koff, chunk_base


def default_pd_params():
    p = PDParams()
    lib().flame_b200_pd_default_params(C.byref(p))
    return p


def default_track_params():
    p = TrackParams()
    lib().flame_b200_track_default_params(C.byref(p))
    return p


def kernel_launches():
    return int(lib().flame_b200_kernel_launches())


class DeviceBuffer:
    """Raw device allocation that can be shared with the other per-GPU processes of the box (CUDA IPC)."""

    def __init__(self, device, nbytes):
        self.device, self.nbytes, self._opened = device, nbytes, False
        self._p = C.c_void_p()
        check(lib().flame_b200_device_alloc(device, nbytes, C.byref(self._p)))

    @property
    def ptr(self):
        return self._p.value

    def export(self):
        h = C.create_string_buffer(64)
        check(lib().flame_b200_ipc_export(self._p, h))
        return h.raw

    @classmethod
    def open(cls, device, handle, nbytes):
        self = cls.__new__(cls)
        self.device, self.nbytes, self._opened = device, nbytes, True
        self._p = C.c_void_p()
        check(lib().flame_b200_ipc_open(device, handle, C.byref(self._p)))
        return self

    def tensor(self, shape, offset_bytes=0):
        """float32 torch view (no copy) of part of the buffer."""
        import torch

        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (self.ptr + offset_bytes, False),
                                      "version": 3, "strides": None}
        return torch.as_tensor(v, device=torch.device("cuda", self.device))

    def close(self):
        if self._p:
            (lib().flame_b200_ipc_close if self._opened else lib().flame_b200_device_free)(self._p)
            self._p = C.c_void_p()


class PinnedArray:
    """A numpy view over page-locked host memory from flame_b200_host_alloc."""

    def __init__(self, shape, dtype=np.float32):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = C.c_void_p()
        check(lib().flame_b200_host_alloc(self.nbytes, C.byref(self._p)))
        buf = (C.c_uint8 * self.nbytes).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            lib().flame_b200_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One CUDA stream + n_slots frame slots (flame_b200_ctx)."""

    def __init__(self, device=0, n_slots=1):
        self._h = C.c_void_p()
        check(lib().flame_b200_ctx_create(device, n_slots, C.byref(self._h)))
        self.device = device
        self.n_slots = n_slots

    def close(self):
        if self._h:
            lib().flame_b200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def stream(self):
        return lib().flame_b200_ctx_stream(self._h)

    def set_async_projection(self, on=True):
        check(lib().flame_b200_ctx_set_async_projection(self._h, int(on)))

    def graph_project_marked(self, slot, V, q, t, scale, rules, fnew=-1):
        """graph_project + the removal rules of Flame::projectGraph: (pos [V,2], idepth [V], remove u8[V])."""
        pos, idepth, rem = np.empty((V, 2), np.float32), np.empty(V, np.float32), np.empty(V, np.uint8)
        q, t = np.ascontiguousarray(q, np.float32), np.ascontiguousarray(t, np.float32)
        check(lib().flame_b200_graph_project_marked(self._h, slot, fptr(q), fptr(t), scale, C.byref(rules), fnew,
                                                    fptr(pos), fptr(idepth), bptr(rem)))
        return pos, idepth, rem

    def detect_features(self, params, fref, fprev, curr_xy, max_features=20000):
        """Grid detector of Flame::detectFeatures on frame slot fref: ((n, 2) float32, n)."""
        curr = np.ascontiguousarray(curr_xy, np.float32).reshape(-1, 2)
        out = np.empty((max_features, 2), np.float32)
        n = C.c_int32()
        check(lib().flame_b200_detect_features(self._h, C.byref(params), fref, fprev, fptr(curr) if len(curr) else None,
                                               curr.shape[0], fptr(out), max_features, C.byref(n)))
        return out[:min(n.value, max_features)].copy(), n.value

    def set_pd_preference(self, pref):
        """PREF_THROUGHPUT (smallest cluster per frame) or PREF_LATENCY; applies to later uploads."""
        check(lib().flame_b200_ctx_set_pd_preference(self._h, int(pref)))

    def join(self):
        """Device-side: order the pending (asynchronous) projection before later work on `stream`."""
        check(lib().flame_b200_ctx_join(self._h))

    def sync(self):
        check(lib().flame_b200_ctx_sync(self._h))

    # ---- graph ----
    def graph_upload(self, slot, g):
        """g: dict of contiguous numpy arrays named as flame_b200_graph_in's fields."""
        s = GraphIn()
        s.V = g["pos"].shape[0]
        s.E = g["edges"].shape[0]
        s.pos = fptr(g["pos"])
        s.edges = iptr(g["edges"])
        for k in ("data_term", "data_weight", "x", "w1", "w2", "x_bar", "w1_bar", "w2_bar", "x_prev",
                  "w1_prev", "w2_prev", "alpha", "beta", "q1", "q2", "q3"):
            setattr(s, k, fptr(g.get(k)))
        check(lib().flame_b200_graph_upload(self._h, slot, C.byref(s)))

    def graph_init(self, slot, pos, idepth, edges, weight=None):
        check(lib().flame_b200_graph_init(self._h, slot, pos.shape[0], fptr(pos), fptr(idepth),
                                          fptr(weight), edges.shape[0], iptr(edges)))

    def graph_set_data(self, slot, data_term=None, data_weight=None):
        check(lib().flame_b200_graph_set_data(self._h, slot, fptr(data_term), fptr(data_weight)))

    def graph_reset(self, slot0=0, n_slots=1):
        check(lib().flame_b200_graph_reset(self._h, slot0, n_slots))

    def graph_download(self, slot, V, E, fields=None):
        names = ("x", "w1", "w2", "x_bar", "w1_bar", "w2_bar", "x_prev", "w1_prev", "w2_prev",
                 "alpha", "beta", "q1", "q2", "q3")
        fields = names if fields is None else fields
        out = {}
        o = GraphOut()
        for k in fields:
            n = E if k in ("alpha", "beta", "q1", "q2", "q3") else V
            out[k] = np.empty(n, np.float32)
            setattr(o, k, fptr(out[k]))
        check(lib().flame_b200_graph_download(self._h, slot, C.byref(o)))
        return out

    def graph_project(self, slot, V, q_wxyz, t, graph_scale=1.0):
        """Flame::projectGraph's vertex loop on the device; returns (u_new [V,2], idepth_new [V])."""
        q = np.ascontiguousarray(q_wxyz, np.float32)
        tt = np.ascontiguousarray(t, np.float32)
        pos = np.empty((V, 2), np.float32)
        idepth = np.empty(V, np.float32)
        check(lib().flame_b200_graph_project(self._h, slot, fptr(q), fptr(tt), graph_scale, fptr(pos),
                                             fptr(idepth)))
        return pos, idepth

    def graph_resync(self, slot, pos, data_term, edges, vtx_from, edge_from, data_weight=None,
                     x_init=None, check_sticky_obstacles=False):
        """Flame::syncGraph's graph edits with the solver state carried over on the device."""
        s = GraphSync()
        s.V = pos.shape[0]
        s.E = edges.shape[0]
        s.pos = fptr(pos)
        s.data_term = fptr(data_term)
        s.data_weight = fptr(data_weight)
        s.vtx_from = iptr(vtx_from)
        s.x_init = fptr(x_init)
        s.edges = iptr(edges)
        s.edge_from = iptr(edge_from)
        s.check_sticky_obstacles = int(check_sticky_obstacles)
        check(lib().flame_b200_graph_resync(self._h, slot, C.byref(s)))

    def pd_run(self, params, n_iters, slot0=0, n_slots=1, mode=PD_AUTO):
        check(lib().flame_b200_pd_run(self._h, slot0, n_slots, C.byref(params), n_iters, mode))

    def pd_wave(self, slot=0):
        n = C.c_int32(0)
        check(lib().flame_b200_pd_wave(self._h, slot, C.byref(n)))
        return n.value

    def pd_cost(self, slot, params, exact_order=False):
        a, b = C.c_float(), C.c_float()
        check(lib().flame_b200_pd_cost(self._h, slot, C.byref(params), int(exact_order), C.byref(a),
                                       C.byref(b)))
        return a.value, b.value

    # ---- dense projection ----
    def interpolate_mesh(self, slot, tris, pos, values, rows, cols, vtx_valid=None, tri_valid=None,
                         img=None, scale=1.0):
        fill = img is None
        if img is None:
            img = np.empty((rows, cols), np.float32)
        check(lib().flame_b200_interpolate_mesh(self._h, slot, tris.shape[0], iptr(tris), pos.shape[0],
                                                fptr(pos), fptr(values), scale, bptr(vtx_valid),
                                                bptr(tri_valid), rows, cols, int(fill), fptr(img)))
        return img

    def mesh_bind(self, slot, tris, rows, cols, vtx_valid=None, tri_valid=None):
        check(lib().flame_b200_mesh_bind(self._h, slot, tris.shape[0], iptr(tris), bptr(vtx_valid),
                                         bptr(tri_valid), rows, cols))

    def mesh_project(self, slot0=0, n_slots=1, scale=1.0):
        check(lib().flame_b200_mesh_project(self._h, slot0, n_slots, scale))

    def mesh_download(self, slot, out):
        """Asynchronous; out: numpy array (ideally PinnedArray.array). Complete after sync()."""
        check(lib().flame_b200_mesh_download(self._h, slot, out.ctypes.data_as(C.c_void_p)))

    def mesh_filter(self, slot, T, Kinv, params=None, scale=1.0, apply=True, want=True):
        """Display filters of flame.cc:389-407 on the slot's mesh; returns tri_validity (u8[T]) if want."""
        if params is None:
            params = TriFilterParams()
            lib().flame_b200_tri_filter_default_params(C.byref(params))
        Kinv = np.ascontiguousarray(Kinv, np.float32).ravel()
        out = np.empty(T, np.uint8) if want else None
        check(lib().flame_b200_mesh_filter(self._h, slot, C.byref(params), fptr(Kinv), scale, bptr(out), int(apply)))
        return out

    def mesh_redirect(self, slot, map_ptr, values_ptr=None):
        """Project straight into (possibly peer) device memory; None restores the slot's own buffers."""
        check(lib().flame_b200_mesh_redirect(self._h, slot, C.c_void_p(map_ptr) if map_ptr else None,
                                             C.c_void_p(values_ptr) if values_ptr else None))

    def device_tensors(self, slot):
        """(dense map [rows, cols], vertex idepths [V]) of the slot as torch CUDA tensors that ALIAS the library's
        device memory (no copy): for NCCL collectives on the results (flame_b200.shard.gather_frames)."""
        import torch
        img, val = C.c_void_p(), C.c_void_p()
        rows, cols, V = C.c_int32(), C.c_int32(), C.c_int32()
        check(lib().flame_b200_slot_device_ptrs(self._h, slot, C.byref(img), C.byref(val), C.byref(rows),
                                                C.byref(cols), C.byref(V)))

        class _View:
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False),
                                                 "version": 3, "strides": None}

        dev = torch.device("cuda", self.device)
        m = torch.as_tensor(_View(img.value, (rows.value, cols.value)), device=dev)
        v = torch.as_tensor(_View(val.value, (max(V.value, 1),)), device=dev)[:V.value]
        return m, v

    def mesh_filter_angles(self, slot, T, Kinv, scale=1.0):
        """Oblique-filter angle of every triangle as computed on the device (diagnostic): f32[T]."""
        Kinv = np.ascontiguousarray(Kinv, np.float32).ravel()
        out = np.empty(T, np.float32)
        check(lib().flame_b200_mesh_filter_angles(self._h, slot, fptr(Kinv), scale, fptr(out)))
        return out

    def mesh_normals(self, slot, V, Kinv, scale=1.0):
        """Flame::getVertexNormals on the slot's mesh and current idepths: [V, 3] float32."""
        Kinv = np.ascontiguousarray(Kinv, np.float32).ravel()
        out = np.empty((V, 3), np.float32)
        check(lib().flame_b200_mesh_normals(self._h, slot, fptr(Kinv), scale, fptr(out)))
        return out

    def mesh_coverage(self, slot):
        n = C.c_int64(0)
        check(lib().flame_b200_mesh_coverage(self._h, slot, C.byref(n)))
        return n.value

    def process_frame(self, slot, pos, idepth, edges, tris, rows, cols, params, n_iters, out_map, out_x,
                      out_w1=None, out_w2=None, weight=None, scale=1.0):
        """graph_init + mesh_bind + pd_run + mesh_project + downloads + sync in ONE C call."""
        check(lib().flame_b200_process_frame(
            self._h, slot, pos.shape[0], fptr(pos), fptr(idepth), fptr(weight), edges.shape[0], iptr(edges),
            tris.shape[0], iptr(tris), rows, cols, C.byref(params), n_iters, scale,
            out_map.ctypes.data_as(C.c_void_p) if out_map is not None else None, fptr(out_x), fptr(out_w1),
            fptr(out_w2)))

    def process_frame_submit(self, slot, pos, idepth, edges, tris, rows, cols, params, n_iters, out_map,
                             weight=None, scale=1.0):
        """Enqueue a whole frame on the slot's own stream; inputs are consumed on return, out_map (pinned)
        is written asynchronously and is valid after process_frame_wait(slot). edges = tris = None: the call
        triangulates the vertices itself (process_frame_topology returns the arrays)."""
        check(lib().flame_b200_process_frame_submit(
            self._h, slot, pos.shape[0], fptr(pos), fptr(idepth), fptr(weight),
            edges.shape[0] if edges is not None else 0, iptr(edges),
            tris.shape[0] if tris is not None else 0, iptr(tris), rows, cols, C.byref(params), n_iters, scale,
            out_map.ctypes.data_as(C.c_void_p) if out_map is not None else None))

    def process_frame_topology(self, slot):
        """(tris [T,3], edges [E,2]) the slot's last self-triangulating process_frame_submit used."""
        nt, ne = C.c_int32(), C.c_int32()
        check(lib().flame_b200_process_frame_topology(self._h, slot, 0, C.byref(nt), None, 0, C.byref(ne), None))
        tris = np.empty((nt.value, 3), np.int32)
        edges = np.empty((ne.value, 2), np.int32)
        check(lib().flame_b200_process_frame_topology(self._h, slot, nt.value, C.byref(nt), iptr(tris), ne.value,
                                                      C.byref(ne), iptr(edges)))
        return tris, edges

    def process_frame_wait(self, slot, out_x=None, out_w1=None, out_w2=None):
        check(lib().flame_b200_process_frame_wait(self._h, slot, fptr(out_x), fptr(out_w1), fptr(out_w2)))

    def stats(self, reset=False):
        st = Stats()
        check(lib().flame_b200_ctx_stats(self._h, C.byref(st), int(reset)))
        return {k: getattr(st, k) for k, _ in Stats._fields_}

    # ---- epipolar search ----
    def set_camera(self, width, height, K, Kinv):
        K = np.ascontiguousarray(K, np.float32).ravel()
        Kinv = np.ascontiguousarray(Kinv, np.float32).ravel()
        check(lib().flame_b200_set_camera(self._h, width, height, fptr(K), fptr(Kinv)))

    def frame_upload(self, fslot, img_pad, gradx_pad, grady_pad, pad, pose):
        pose = np.ascontiguousarray(pose, np.float32)
        check(lib().flame_b200_frame_upload(self._h, fslot, bptr(np.ascontiguousarray(img_pad)),
                                            fptr(np.ascontiguousarray(gradx_pad)),
                                            fptr(np.ascontiguousarray(grady_pad)), pad, fptr(pose)))

    def frame_upload_raw(self, fslot, img, pad, pose):
        pose = np.ascontiguousarray(pose, np.float32)
        check(lib().flame_b200_frame_upload_raw(self._h, fslot, bptr(np.ascontiguousarray(img, np.uint8)),
                                                pad, fptr(pose)))

    def frame_download(self, fslot, height, width, pad):
        shape = (height + 2 * pad, width + 2 * pad)
        ip = np.empty(shape, np.uint8)
        gx = np.empty(shape, np.float32)
        gy = np.empty(shape, np.float32)
        check(lib().flame_b200_frame_download(self._h, fslot, bptr(ip), fptr(gx), fptr(gy)))
        return ip, gx, gy

    def update_feature_idepths(self, params, fnew, curr_pf, feats, want_diag=True):
        assert feats.dtype == FEATURE_DTYPE and feats.flags["C_CONTIGUOUS"]
        diag = np.zeros(feats.shape[0], DIAG_DTYPE) if want_diag else None
        n = C.c_int32()
        check(lib().flame_b200_update_feature_idepths(
            self._h, C.byref(params), fnew, curr_pf, feats.ctypes.data_as(C.c_void_p), feats.shape[0],
            diag.ctypes.data_as(C.c_void_p) if want_diag else None, C.byref(n)))
        return n.value, diag

    def features_upload(self, feats):
        """Replace the device-resident feature set."""
        assert feats.dtype == FEATURE_DTYPE and feats.flags["C_CONTIGUOUS"]
        check(lib().flame_b200_features_upload(self._h, feats.ctypes.data_as(C.c_void_p), feats.shape[0]))
        self._n_feats = feats.shape[0]

    def features_track(self, params, fnew, curr_pf):
        """Asynchronous, out-of-place search + measurement + fusion on the resident set."""
        check(lib().flame_b200_features_track(self._h, C.byref(params), fnew, curr_pf))

    def features_download(self, commit=True, want_diag=True):
        feats = np.zeros(self._n_feats, FEATURE_DTYPE)
        diag = np.zeros(self._n_feats, DIAG_DTYPE) if want_diag else None
        n = C.c_int32()
        check(lib().flame_b200_features_download(
            self._h, feats.ctypes.data_as(C.c_void_p), diag.ctypes.data_as(C.c_void_p) if want_diag else None,
            C.byref(n), int(commit)))
        return n.value, feats, diag
