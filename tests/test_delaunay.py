"""CPU: flame_b200_delaunay (the frame pipeline's own host triangulator) produces a Delaunay triangulation in
Triangle's conventions - checked with exact integer arithmetic on the 1/256-px coordinates it works on:
positive orientation, locally Delaunay across every interior edge (hence globally), the triangles tile the
convex hull, edges = sides of the triangles, and the winding is the reference Triangle's."""
import numpy as np
import pytest

from flame_b200 import _lib, synth


def _q(pos):
    return np.rint(np.asarray(pos, np.float64) * 256.0).astype(np.int64)


def _orient(p, a, b, c):
    return int((p[b, 0] - p[a, 0]) * (p[c, 1] - p[a, 1]) - (p[b, 1] - p[a, 1]) * (p[c, 0] - p[a, 0]))


def _incircle(p, a, b, c, d):
    ax, ay = int(p[a, 0] - p[d, 0]), int(p[a, 1] - p[d, 1])
    bx, by = int(p[b, 0] - p[d, 0]), int(p[b, 1] - p[d, 1])
    cx, cy = int(p[c, 0] - p[d, 0]), int(p[c, 1] - p[d, 1])
    return ((ax * ax + ay * ay) * (bx * cy - by * cx) - (bx * bx + by * by) * (ax * cy - ay * cx) +
            (cx * cx + cy * cy) * (ax * by - ay * bx))


def _check(pos, tris, edges):
    p = _q(pos)
    area2 = 0
    sides = {}
    by_set = {tuple(sorted(t)): t for t in tris.tolist()}
    for t, (a, b, c) in enumerate(tris.tolist()):
        o = _orient(p, a, b, c)
        assert o > 0, "triangle %d not positively oriented" % t
        area2 += o
        for u, v, w in ((a, b, c), (b, c, a), (c, a, b)):
            sides.setdefault((min(u, v), max(u, v)), []).append(w)
    for (u, v), opp in sides.items():
        assert len(opp) <= 2
        if len(opp) == 2:  # locally Delaunay: the opposite vertex is not strictly inside the circumcircle
            a, b, c = by_set[tuple(sorted((u, v, opp[0])))]
            assert _incircle(p, a, b, c, opp[1]) <= 0
    eset = set((min(a, b), max(a, b)) for a, b in edges.tolist())
    assert len(eset) == len(edges)
    assert set(sides) <= eset
    return area2, eset, sides


def test_delaunay_of_config_meshes_matches_properties_and_triangle_winding(oracle):
    from scipy.spatial import ConvexHull
    for cfg in ("A", "C"):
        g = synth.config_graph(cfg)
        pos = g["pos"]
        tris, edges = _lib.delaunay(pos)
        area2, eset, sides = _check(pos, tris, edges)
        hull = ConvexHull(_q(pos).astype(np.float64))
        assert area2 == int(round(2 * hull.volume))          # the triangles tile the convex hull exactly
        assert eset == set(sides)                            # no stray edges for points in general position
        V = pos.shape[0]
        assert len(tris) == len(g["tris"]) and len(edges) == len(g["edges"])   # same counts as Triangle's
        # same winding as the reference's Triangle (which the rasteriser relies on), and the same triangle SET
        # wherever the Delaunay triangulation is unique (no exactly cocircular quadruple)
        p = _q(pos)
        assert all(_orient(p, a, b, c) > 0 for a, b, c in g["tris"].tolist())
        ours = set(tuple(sorted(t)) for t in tris.tolist())
        ref = set(tuple(sorted(t)) for t in g["tris"].tolist())
        assert len(ours & ref) >= 0.995 * len(ref)


def test_delaunay_degenerate_inputs():
    # pixel grid: every quadruple cocircular
    ys, xs = np.mgrid[0:12, 0:17]
    grid = np.stack([xs.ravel() * 8.0 + 4, ys.ravel() * 8.0 + 4], 1).astype(np.float32)
    tris, edges = _lib.delaunay(grid)
    area2, eset, sides = _check(grid, tris, edges)
    assert len(tris) == 2 * 11 * 16 and area2 == 2 * (16 * 8 * 256) * (11 * 8 * 256)
    # collinear points: a chain of edges, no triangle
    line = np.stack([np.arange(9) * 3.0, np.arange(9) * 1.5], 1).astype(np.float32)
    tris, edges = _lib.delaunay(line)
    assert len(tris) == 0 and len(edges) == 8
    # duplicates are left unconnected; tiny inputs
    dup = np.array([[5, 5], [20, 5], [5, 5], [12, 30], [20, 5]], np.float32)
    tris, edges = _lib.delaunay(dup)
    assert tris.tolist() and set(np.unique(tris).tolist()) == {0, 1, 3} and len(edges) == 3
    for n in (0, 1, 2):
        t, e = _lib.delaunay(np.array([[1, 1], [9, 4]], np.float32)[:n])
        assert len(t) == 0 and len(e) == max(0, n - 1)
    with pytest.raises(_lib.FlameB200Error):
        _lib.delaunay(np.array([[1, 1], [np.nan, 4], [3, 9]], np.float32))


def test_delaunay_random_clouds_including_clusters():
    rng = np.random.default_rng(4)
    for n in (3, 4, 7, 50, 400, 3000):
        pts = rng.uniform(0, 1900, (n, 2)).astype(np.float32)
        if n >= 50:
            pts[: n // 5] = np.rint(pts[: n // 5])               # pixel-grid points among sub-pixel ones
            pts[n // 5: n // 4] = pts[0] + rng.normal(0, 0.02, (n // 4 - n // 5, 2))  # a tight cluster
        tris, edges = _lib.delaunay(pts)
        _check(pts, tris, edges)
        q = _q(pts)
        uniq = len(set(map(tuple, q.tolist())))
        assert len(edges) <= 3 * uniq - 3 and (n < 3 or len(tris) >= 1)


def test_delaunay_alternating_cuts_on_lines_circles_and_every_small_size():
    """The divide step alternates between x and y cuts and moves the hull handles of the halves to the other
    axis' extreme vertices: exercise hulls that are segments, axis-parallel lines, cocircular rings and every
    small size (base cases of 2 and 3 points under both axes)."""
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(11)
    k = np.arange(13, dtype=np.float32)
    shapes = {
        "vertical line": np.stack([np.full(13, 40.0, np.float32), 5 + 7 * k], 1),
        "horizontal line": np.stack([5 + 7 * k, np.full(13, 40.0, np.float32)], 1),
        "two vertical lines": np.concatenate([np.stack([np.full(13, 40.0, np.float32), 5 + 7 * k], 1),
                                              np.stack([np.full(13, 90.0, np.float32), 5 + 7 * k], 1)]),
        "two horizontal lines": np.concatenate([np.stack([5 + 7 * k, np.full(13, 40.0, np.float32)], 1),
                                                np.stack([5 + 7 * k, np.full(13, 90.0, np.float32)], 1)]),
        "L": np.concatenate([np.stack([np.full(9, 10.0, np.float32), 10 + 5 * k[:9]], 1),
                             np.stack([15 + 5 * k[:9], np.full(9, 10.0, np.float32)], 1)]),
        # 3-4-5 multiples: exactly cocircular lattice points around (100, 100), radius 25 and 50
        "ring": np.array([[100 + s * a, 100 + t * b] for r in (1, 2) for a, b in ((15, 20), (20, 15), (7, 24), (24, 7), (0, 25), (25, 0))
                          for s in (r, -r) for t in (r, -r)], np.float32),
    }
    for name, pts in shapes.items():
        pts = np.unique(pts, axis=0)
        pts = pts[rng.permutation(len(pts))]
        tris, edges = _lib.delaunay(pts)
        area2, eset, sides = _check(pts, tris, edges)
        q = _q(pts).astype(np.float64)
        if "line" in name and "two" not in name:
            assert len(tris) == 0 and len(edges) == len(pts) - 1, name
        else:
            assert area2 == int(round(2 * ConvexHull(q).volume)), name
            assert eset == set(sides), name
    for n in range(2, 41):
        for rep in range(6):
            pts = rng.uniform(0, 300, (n, 2)).astype(np.float32)
            if rep % 2:
                pts = np.rint(pts / 16) * 16 + 3      # coarse lattice: many collinear / cocircular subsets
                pts = np.unique(pts, axis=0).astype(np.float32)
                if len(pts) < 2:
                    continue
            tris, edges = _lib.delaunay(pts)
            area2, eset, sides = _check(pts, tris, edges)
            q = _q(pts).astype(np.float64)
            if len(pts) >= 3 and np.linalg.matrix_rank(q[1:] - q[0]) == 2:
                assert area2 == int(round(2 * ConvexHull(q).volume)), (n, rep)
                assert eset == set(sides), (n, rep)
            else:
                assert len(tris) == 0 and len(edges) == len(pts) - 1, (n, rep)
