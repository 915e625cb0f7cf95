"""Synthetic FLaME frames for tests and bench.py. Seeds fixed.

Produces the INPUTS of the hot path: camera, vertices, a Delaunay triangulation (index arrays),
noisy inverse-depth data terms, and — for the epipolar search — a textured reference / comparison
image pair with padded copies and central-difference gradients laid out as
flame::utils::Frame::create does (utils/frame.cc:33-71: border = 5, REFLECT_101 for the image,
zeros for the gradients).

Delaunay itself is outside the replaced path: index arrays come either from a
committed fixture produced by the reference's own Triangle (tests/golden/graph_*.npz, bit-exact by
construction) or, for arbitrary seeds, from scipy's Qhull.
"""
import os

import numpy as np

CONFIGS = {
    # name: (W, H, V, pd_iters)   — BASELINE.json configs A-D
    "A": (640, 480, 500, 50),
    "B": (640, 480, 2000, 100),
    "C": (1280, 720, 5000, 100),
    "D": (1920, 1080, 10000, 200),
}
BORDER = 4  # int(rescale_factor_max * win_size / 2 + 1), flame.cc:1576
PAD = 5     # Frame::create border = fparams.win_size, flame.cc:149-150
_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def camera(W, H):
    f = 525.0 * W / 640.0
    K = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1]], np.float32)
    Kinv = np.array([[1.0 / f, 0, -(W / 2.0) / f], [0, 1.0 / f, -(H / 2.0) / f], [0, 0, 1]], np.float32)
    return K, Kinv


def scene_idepth(u, v, W, H):
    """Ground-truth inverse depth: a tilted plane with a +0.3 step inside a rectangle."""
    rho = 0.5 + 0.2 * (u / W) + 0.1 * (v / H)
    inside = (u >= 0.55 * W) & (u <= 0.85 * W) & (v >= 0.3 * H) & (v <= 0.7 * H)
    return (rho + 0.3 * inside).astype(np.float32)


def make_vertices(W, H, V, seed=1):
    rng = np.random.default_rng(seed)
    out = np.empty((0, 2), np.float32)
    seen = set()
    while out.shape[0] < V:
        n = V - out.shape[0]
        p = np.stack([rng.uniform(BORDER, W - BORDER, 2 * n), rng.uniform(BORDER, H - BORDER, 2 * n)], 1)
        p = p.astype(np.float32)
        p = p[(p[:, 0] < W - BORDER) & (p[:, 1] < H - BORDER)]
        keep = []
        for i, (x, y) in enumerate(np.rint(p).astype(np.int64)):
            k = (int(x), int(y))
            if k not in seen:
                seen.add(k)
                keep.append(i)
        out = np.concatenate([out, p[keep][:n]], 0)
    return np.ascontiguousarray(out[:V], np.float32)


def _scipy_delaunay(pos):
    from scipy.spatial import Delaunay
    tri = Delaunay(pos.astype(np.float64)).simplices.astype(np.int32)
    a, b, c = pos[tri[:, 0]], pos[tri[:, 1]], pos[tri[:, 2]]
    area2 = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    # Triangle emits vertices so that the REVERSED order has positive edge functions in the
    # reference rasteriser (image_utils.cc:383-391): enforce the same winding (area2 > 0 in (x, y-down)).
    flip = area2 < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    e = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]], 0)
    e.sort(axis=1)
    e = np.unique(e, axis=0).astype(np.int32)
    return np.ascontiguousarray(tri), np.ascontiguousarray(e)


def make_graph(W, H, V, seed=1, noise_seed=2, triangulator="auto"):
    """Vertices + index arrays + PD inputs (x = x_bar = x_prev = data, w = q = 0; flame.cc:2039-2048)."""
    fixture = os.path.join(_GOLDEN, "graph_%dx%d_v%d_s%d.npz" % (W, H, V, seed))
    if triangulator in ("auto", "fixture") and os.path.exists(fixture):
        z = np.load(fixture)
        pos, tris, edges = z["pos"], z["tris"], z["edges"]
        source = "reference-triangle-fixture"
    elif triangulator == "fixture":
        raise FileNotFoundError(fixture)
    else:
        pos = make_vertices(W, H, V, seed)
        tris, edges = _scipy_delaunay(pos)
        source = "scipy-qhull"
    rng = np.random.default_rng(noise_seed)
    data = scene_idepth(pos[:, 0], pos[:, 1], W, H) + rng.normal(0, 0.02, pos.shape[0]).astype(np.float32)
    data = data.astype(np.float32)
    return {
        "W": W, "H": H, "source": source,
        "pos": np.ascontiguousarray(pos, np.float32),
        "tris": np.ascontiguousarray(tris, np.int32),
        "edges": np.ascontiguousarray(edges, np.int32),
        "data_term": data,
        "data_weight": np.ones(pos.shape[0], np.float32),
    }


def init_state(g, alpha, beta):
    """Full PD state dict (oracle / C-ABI layout) for graph inputs g with edge weights alpha, beta."""
    V, E = g["pos"].shape[0], g["edges"].shape[0]
    s = dict(g)
    for k in ("x", "x_bar", "x_prev"):
        s[k] = g["data_term"].copy()
    for k in ("w1", "w2", "w1_bar", "w2_bar", "w1_prev", "w2_prev"):
        s[k] = np.zeros(V, np.float32)
    for k in ("q1", "q2", "q3"):
        s[k] = np.zeros(E, np.float32)
    s["alpha"] = np.ascontiguousarray(alpha, np.float32)
    s["beta"] = np.ascontiguousarray(beta, np.float32)
    return s


def sync_topology(old_edges, keep, pos_kept, pos_new, delaunay=None):
    """Index bookkeeping of Flame::syncGraph (flame.cc:1984-2123) for flame_b200_graph_resync.

    old_edges [E0,2]; keep [V0] bool; pos_kept [sum(keep),2] new positions of the kept vertices (old
    order); pos_new [Vn,2] brand-new vertices. Vertices: kept ones in their old order, then the new ones
    (boost::add_vertex appends). Edges follow Boost's edge list: edges of the old graph that are still in
    the new triangulation keep their relative order AND orientation (boost::edge() finds the existing
    descriptor, :2092-2100), edges new to the triangulation are appended in triangulation order.
    Returns pos, tris, edges, vtx_from, edge_from.
    """
    keep = np.asarray(keep, bool)
    old_to_new = np.full(keep.shape[0], -1, np.int64)
    old_to_new[keep] = np.arange(int(keep.sum()))
    pos = np.ascontiguousarray(np.concatenate([pos_kept, pos_new], 0), np.float32)
    vtx_from = np.concatenate([np.flatnonzero(keep), np.full(pos_new.shape[0], -1)]).astype(np.int32)
    tris, tedges = (delaunay or _scipy_delaunay)(pos)
    want = {}
    for i, (a, b) in enumerate(tedges.tolist()):
        want[(min(a, b), max(a, b))] = i
    edges, edge_from, used = [], [], set()
    for e, (a, b) in enumerate(np.asarray(old_edges).tolist()):
        na, nb = old_to_new[a], old_to_new[b]
        if na < 0 or nb < 0:
            continue
        k = (min(na, nb), max(na, nb))
        if k in want:
            edges.append((na, nb))
            edge_from.append(e)
            used.add(k)
    for (a, b) in tedges.tolist():
        if (min(a, b), max(a, b)) not in used:
            edges.append((a, b))
            edge_from.append(-1)
    return (pos, tris, np.ascontiguousarray(np.array(edges, np.int32).reshape(-1, 2)),
            np.ascontiguousarray(vtx_from), np.array(edge_from, np.int32))


def sync_state_host(old, pos, edges, vtx_from, edge_from, data_term, data_weight=None, x_init=None,
                    check_sticky_obstacles=False):
    """The host-side edits graph_resync replaces: full PD state dict of the new graph from the old one
    (carried vertices / edges keep their state; defaults of flame.cc:2046-2048 and EdgeData otherwise).
    alpha / beta are left to the caller (1/len, 1)."""
    V, E = pos.shape[0], edges.shape[0]
    s = {"pos": pos, "edges": edges, "data_term": np.ascontiguousarray(data_term, np.float32),
         "data_weight": np.ones(V, np.float32) if data_weight is None else data_weight}
    cv, ce = vtx_from >= 0, edge_from >= 0
    init = s["data_term"] if x_init is None else x_init
    for k in ("x", "x_bar", "x_prev"):
        a = init.astype(np.float32).copy()
        a[cv] = old[k][vtx_from[cv]]
        s[k] = a
    for k in ("w1", "w2", "w1_bar", "w2_bar", "w1_prev", "w2_prev"):
        a = np.zeros(V, np.float32)
        a[cv] = old[k][vtx_from[cv]]
        s[k] = a
    for k in ("q1", "q2", "q3"):
        a = np.zeros(E, np.float32)
        a[ce] = old[k][edge_from[ce]]
        s[k] = a
    if check_sticky_obstacles:
        reset = cv & ((s["x"] - s["data_term"]) > np.float32(0.25))
        s["x"][reset] = s["data_term"][reset]
    return s


def triangle_soups(n=25, seed=123):
    """Random triangle lists outside the benchmark's shape for the dense projection: overlapping triangles of
    both windings, collinear / coincident vertices, horizontal lines reaching the right border (the 4-px sweep
    overshoot), validity masks and pre-initialised images. Yields dicts (W, H, pos, tris, vals, vv, tv, base)."""
    rng = np.random.default_rng(seed)
    for case in range(n):
        W, H = int(rng.integers(17, 90)), int(rng.integers(9, 70))
        V, T = int(rng.integers(3, 40)), int(rng.integers(1, 60))
        pos = np.stack([rng.uniform(0, W - 1, V), rng.uniform(0, H - 1, V)], 1).astype(np.float32)
        if case % 3 == 0:
            pos = np.rint(pos).astype(np.float32)          # many exactly collinear / coincident points
        if case % 4 == 1:
            pos[: V // 2, 1] = pos[0, 1]                   # horizontal lines
            pos[V // 2 - 1, 0] = W - 1                     # ... reaching the right border
        tris = rng.integers(0, V, (T, 3)).astype(np.int32)
        vals = rng.uniform(0.05, 3.0, V).astype(np.float32)
        vv = (rng.uniform(size=V) > 0.1).astype(np.uint8) if case % 2 else None
        tv = (rng.uniform(size=T) > 0.2).astype(np.uint8) if case % 5 else None
        base = rng.uniform(size=(H, W)).astype(np.float32) if case % 2 == 0 else None
        yield {"case": case, "W": W, "H": H, "pos": pos, "tris": tris, "vals": vals, "vv": vv, "tv": tv,
               "base": base}


def config_graph(name, frame=0):
    W, H, V, iters = CONFIGS[name]
    g = make_graph(W, H, V, seed=1 + 100 * frame, noise_seed=2 + 100 * frame)
    g["pd_iters"] = iters
    return g


# ------------------------------------------------------------------------------- images
def _texture(W, H, seed=3):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.zeros((H, W), np.float32)
    for _ in range(64):
        fx, fy = rng.uniform(-0.25, 0.25, 2)
        img += rng.uniform(0.3, 1.0) * np.sin(2 * np.pi * (fx * xx + fy * yy) + rng.uniform(0, 2 * np.pi))
    noise = rng.normal(0, 1, (H, W)).astype(np.float32)
    k = np.array([1, 4, 6, 4, 1], np.float32) / 16.0
    for _ in range(2):
        noise = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, noise)
        noise = np.apply_along_axis(lambda c: np.convolve(c, k, mode="same"), 0, noise)
    img += 6.0 * noise
    img = (img - img.min()) / (img.max() - img.min())
    return (16 + 224 * img).astype(np.float32)


def _bilinear(img, x, y):
    H, W = img.shape
    x = np.clip(x, 0, W - 1.001)
    y = np.clip(y, 0, H - 1.001)
    x0, y0 = np.floor(x).astype(np.int64), np.floor(y).astype(np.int64)
    dx, dy = x - x0, y - y0
    return ((1 - dx) * (1 - dy) * img[y0, x0] + dx * (1 - dy) * img[y0, x0 + 1] +
            (1 - dx) * dy * img[y0 + 1, x0] + dx * dy * img[y0 + 1, x0 + 1])


def quat_from_yaw(deg):
    h = np.deg2rad(deg) / 2.0
    return np.array([np.cos(h), 0.0, np.sin(h), 0.0], np.float32)  # rotation about the camera y axis


def _quat_to_rot(q):
    w, x, y, z = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def central_gradients(img):
    """utils/image_utils.h getCentralGradient: 0.5*(I[x+1]-I[x-1]); forward/backward at the borders."""
    f = img.astype(np.float32)
    gx = np.empty_like(f)
    gy = np.empty_like(f)
    gx[:, 1:-1] = 0.5 * (f[:, 2:] - f[:, :-2])
    gx[:, 0] = f[:, 1] - f[:, 0]
    gx[:, -1] = f[:, -1] - f[:, -2]
    gy[1:-1, :] = 0.5 * (f[2:, :] - f[:-2, :])
    gy[0, :] = f[1, :] - f[0, :]
    gy[-1, :] = f[-1, :] - f[-2, :]
    return gx, gy


def pad_frame(img_u8):
    """Frame::create (utils/frame.cc:54-63): image REFLECT_101, gradients zero-padded, border = 5."""
    gx, gy = central_gradients(img_u8)
    img_pad = np.pad(img_u8, PAD, mode="reflect")
    gx_pad = np.pad(gx, PAD, mode="constant")
    gy_pad = np.pad(gy, PAD, mode="constant")
    return (np.ascontiguousarray(img_pad, np.uint8), np.ascontiguousarray(gx_pad, np.float32),
            np.ascontiguousarray(gy_pad, np.float32))


def make_stereo_pair(W, H, seed=3, noise_seed=4, t_world=(0.10, 0.02, 0.01), yaw_deg=1.0):
    """Reference frame at identity, comparison camera at (t_world, yaw) in the world (camera-to-world
    poses, as flame::utils::Frame::pose). cmp is ref inverse-warped through the scene idepth."""
    K, Kinv = camera(W, H)
    ref = _texture(W, H, seed)
    q_cmp = quat_from_yaw(yaw_deg)
    R = _quat_to_rot(q_cmp)
    t = np.asarray(t_world, np.float64)
    # cmp pixel -> ray in world -> intersect scene (iterate: idepth is defined in the ref frame)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    f, cx, cy = float(K[0, 0]), float(K[0, 2]), float(K[1, 2])
    rays = np.stack([(xx - cx) / f, (yy - cy) / f, np.ones_like(xx)], -1) @ R.T
    depth = np.full((H, W), 1.0 / 0.6)
    for _ in range(8):
        P = t + rays * depth[..., None]          # point in ref(world) frame
        u = f * P[..., 0] / P[..., 2] + cx
        v = f * P[..., 1] / P[..., 2] + cy
        rho = scene_idepth(u.astype(np.float32), v.astype(np.float32), W, H).astype(np.float64)
        zref = 1.0 / rho
        depth = depth * (zref / P[..., 2])
    P = t + rays * depth[..., None]
    u = f * P[..., 0] / P[..., 2] + cx
    v = f * P[..., 1] / P[..., 2] + cy
    cmp_img = _bilinear(ref, u.astype(np.float32), v.astype(np.float32))
    rng = np.random.default_rng(noise_seed)
    cmp_img = cmp_img + rng.normal(0, 1, (H, W))
    ref_u8 = np.clip(np.rint(ref), 0, 255).astype(np.uint8)
    cmp_u8 = np.clip(np.rint(cmp_img), 0, 255).astype(np.uint8)
    pose_ref = np.array([1, 0, 0, 0, 0, 0, 0], np.float32)
    pose_cmp = np.concatenate([q_cmp, np.asarray(t_world, np.float32)]).astype(np.float32)
    frames = []
    for img, pose in ((ref_u8, pose_ref), (cmp_u8, pose_cmp)):
        ip, gx, gy = pad_frame(img)
        frames.append({"pose": pose, "img": img, "img_pad": ip, "gradx_pad": gx, "grady_pad": gy})
    return K, Kinv, frames


FEATURE_DTYPE = np.dtype([("id", "<u4"), ("frame_id", "<u4"), ("x", "<f4"), ("y", "<f4"),
                          ("idepth_mu", "<f4"), ("idepth_var", "<f4"), ("valid", "<i4"),
                          ("num_updates", "<u4"), ("num_dropouts", "<u4"), ("search_status", "<i4")])


def make_features(pos, W, H, seed=5, rel_sigma=0.05, var=0.01, frame_id=0):
    """Features at the vertices, prior mu = rho*(1+N(0, rel_sigma^2)), var = 0.01, parent = frame 0."""
    rng = np.random.default_rng(seed)
    n = pos.shape[0]
    f = np.zeros(n, FEATURE_DTYPE)
    f["id"] = np.arange(n)
    f["frame_id"] = frame_id
    f["x"], f["y"] = pos[:, 0], pos[:, 1]
    rho = scene_idepth(pos[:, 0], pos[:, 1], W, H)
    f["idepth_mu"] = rho * (1 + rng.normal(0, rel_sigma, n)).astype(np.float32)
    f["idepth_var"] = var
    f["valid"] = 1
    return f
