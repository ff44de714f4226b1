"""Lattice restatements of the shapely / GEOS / geopandas calls on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  shapely (>= 2.0.1 per
``setup.py``) and geopandas are un-vendored dependencies that are not installed in
this image; what the reference asks of them is restated here for geometries whose
vertices are integer lattice points (contour points are rounded to ints at
``indices/contour_index.py:110-112``, boxes are ints, split vertices are truncated
to ints at ``utils/index_utils.py:135``):

* ``LineString.touches`` / ``LineString.intersects`` (DE-9IM on segments),
  ``indices/streamer_index.py:197,234``
* ``geometry.buffer(r)`` + ``sjoin(points, predicate="contains")`` = the set of grid
  points strictly inside the buffered geometry, ``utils/index_utils.py:47-51`` and
  ``processing/events.py:75-79``
* the split of a polygon at the last meridian, ``utils/index_utils.py:148-173``
* ``intersection(...).area`` of two (multi)polygons, ``processing/events.py:205-214``
"""

from fractions import Fraction

import numpy as np


# --------------------------------------------------------------------------- predicates
def _orient(ax, ay, bx, by, cx, cy):
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)


def chord_touches_polyline(p, q, pts):
    """``LineString([p, q]).touches(LineString(pts))`` for lattice vertices.

    ``p`` and ``q`` are vertices of ``pts`` (all vertices of ``pts`` distinct), so the
    geometries always share a point; ``touches`` is then "the open chord meets the
    polyline nowhere except, possibly, at the location of its two end points
    ``pts[0]`` / ``pts[-1]``" (the boundary of the polyline).
    """
    pts = np.asarray(pts, dtype=np.int64)
    px, py = int(p[0]), int(p[1])
    qx, qy = int(q[0]), int(q[1])
    ax, ay = pts[:-1, 0], pts[:-1, 1]
    bx, by = pts[1:, 0], pts[1:, 1]
    ux, uy = qx - px, qy - py
    l2 = ux * ux + uy * uy
    d1 = _orient(px, py, qx, qy, ax, ay)
    d2 = _orient(px, py, qx, qy, bx, by)
    d3 = _orient(ax, ay, bx, by, px, py)
    d4 = _orient(ax, ay, bx, by, qx, qy)
    ta = (ax - px) * ux + (ay - py) * uy
    tb = (bx - px) * ux + (by - py) * uy

    e0x, e0y = pts[0]
    e1x, e1y = pts[-1]

    def is_end(x, y):
        return ((x == e0x) & (y == e0y)) | ((x == e1x) & (y == e1y))

    # proper crossing (interior of both); exempt if the crossing point is an end point location
    proper = (np.sign(d1) * np.sign(d2) < 0) & (np.sign(d3) * np.sign(d4) < 0)
    if proper.any():
        for ex, ey in ((e0x, e0y), (e1x, e1y)):
            on_chord = _orient(px, py, qx, qy, ex, ey) == 0
            on_seg = _orient(ax, ay, bx, by, ex, ey) == 0
            proper = proper & ~(on_chord & on_seg)
    # contour vertex strictly inside the chord
    a_in = (d1 == 0) & (ta > 0) & (ta < l2) & ~is_end(ax, ay)
    b_in = (d2 == 0) & (tb > 0) & (tb < l2) & ~is_end(bx, by)
    # collinear overlap of positive length
    lo = np.maximum(np.minimum(ta, tb), 0)
    hi = np.minimum(np.maximum(ta, tb), l2)
    overlap = (d1 == 0) & (d2 == 0) & (lo < hi)
    return not bool((proper | a_in | b_in | overlap).any())


def segments_intersect(p1, q1, p2, q2):
    """``LineString([p1,q1]).intersects(LineString([p2,q2]))`` (closed segments)."""
    p1x, p1y, q1x, q1y = int(p1[0]), int(p1[1]), int(q1[0]), int(q1[1])
    p2x, p2y, q2x, q2y = int(p2[0]), int(p2[1]), int(q2[0]), int(q2[1])
    d1 = _orient(p1x, p1y, q1x, q1y, p2x, p2y)
    d2 = _orient(p1x, p1y, q1x, q1y, q2x, q2y)
    d3 = _orient(p2x, p2y, q2x, q2y, p1x, p1y)
    d4 = _orient(p2x, p2y, q2x, q2y, q1x, q1y)
    if ((d1 > 0 and d2 < 0) or (d1 < 0 and d2 > 0)) and ((d3 > 0 and d4 < 0) or (d3 < 0 and d4 > 0)):
        return True

    def on_seg(ax, ay, bx, by, cx, cy):
        return min(ax, bx) <= cx <= max(ax, bx) and min(ay, by) <= cy <= max(ay, by)

    if d1 == 0 and on_seg(p1x, p1y, q1x, q1y, p2x, p2y):
        return True
    if d2 == 0 and on_seg(p1x, p1y, q1x, q1y, q2x, q2y):
        return True
    if d3 == 0 and on_seg(p2x, p2y, q2x, q2y, p1x, p1y):
        return True
    if d4 == 0 and on_seg(p2x, p2y, q2x, q2y, q1x, q1y):
        return True
    return False


# --------------------------------------------------------------------------- rasterisation
def buffered_contains(rings, r, px, py, edge_chunk=128):
    """Points (px, py) strictly inside ``Polygon/MultiPolygon(rings).buffer(r)``.

    Restated membership rule: inside any ring (non-zero winding) or
    on its boundary or closer than ``r`` to one of its edges (including the vertex
    discs).  ``rings`` is a list of (n, 2) arrays of open rings (the closing edge is
    implied); ``px``/``py`` are 1-D coordinate arrays in the same space.  Arithmetic
    is fp64, which is exact for lattice inputs.
    """
    px = np.asarray(px, dtype=np.float64)
    py = np.asarray(py, dtype=np.float64)
    out = np.zeros(px.shape, dtype=bool)
    r2 = float(r) * float(r)
    for ring in rings:
        ring = np.asarray(ring, dtype=np.float64)
        if len(ring) == 0:
            continue
        xa, ya = ring[:, 0], ring[:, 1]
        xb, yb = np.roll(xa, -1), np.roll(ya, -1)
        # candidate points: bbox +- r
        cand = np.nonzero(
            (px >= xa.min() - r) & (px <= xa.max() + r) & (py >= ya.min() - r) & (py <= ya.max() + r)
        )[0]
        if len(cand) == 0:
            continue
        cx, cy = px[cand][:, None], py[cand][:, None]
        wn = np.zeros(len(cand), dtype=np.int64)
        near = np.zeros(len(cand), dtype=bool)
        for s in range(0, len(xa), edge_chunk):
            x0, y0 = xa[None, s:s + edge_chunk], ya[None, s:s + edge_chunk]
            x1, y1 = xb[None, s:s + edge_chunk], yb[None, s:s + edge_chunk]
            dx, dy = x1 - x0, y1 - y0
            cross = dx * (cy - y0) - dy * (cx - x0)  # > 0: point left of the edge
            up = (y0 <= cy) & (cy < y1) & (cross > 0)
            dn = (y1 <= cy) & (cy < y0) & (cross < 0)
            wn += up.sum(axis=1) - dn.sum(axis=1)
            len2 = dx * dx + dy * dy
            dot = (cx - x0) * dx + (cy - y0) * dy
            within = (dot >= 0) & (dot <= len2) & (cross * cross < r2 * len2)
            # on the edge itself (also covers r == 0); a zero-length edge (repeated vertex) is only its end point,
            # which the last term below covers -- without the len2 test it would claim every candidate point
            on = (cross == 0) & (dot >= 0) & (dot <= len2) & (len2 > 0)
            disc = (cx - x0) ** 2 + (cy - y0) ** 2 < r2
            near |= (within | on | disc | ((cx == x0) & (cy == y0))).any(axis=1)
        out[cand] |= (wn != 0) | near
    return out


def ring_is_simple(ring):
    """True if the closed lattice ring neither touches nor crosses itself (vectorised over all edge pairs)."""
    ring = np.asarray(ring, dtype=np.int64).reshape(-1, 2)
    n = len(ring)
    if n < 3:
        return False
    a, b = ring, np.roll(ring, -1, axis=0)
    if np.any(np.all(a == b, axis=1)):
        return False  # repeated consecutive vertex
    # adjacent edges share one vertex; they must not fold back onto each other
    c = np.roll(ring, -2, axis=0)
    cr = (b[:, 0] - a[:, 0]) * (c[:, 1] - b[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - b[:, 0])
    dot = (a[:, 0] - b[:, 0]) * (c[:, 0] - b[:, 0]) + (a[:, 1] - b[:, 1]) * (c[:, 1] - b[:, 1])
    if np.any((cr == 0) & (dot > 0)):
        return False
    if n == 3:
        return True
    # non-adjacent edges must not share any point (closed segments)
    ax, ay, bx, by = a[:, 0][:, None], a[:, 1][:, None], b[:, 0][:, None], b[:, 1][:, None]
    cx, cy, dx, dy = a[:, 0][None, :], a[:, 1][None, :], b[:, 0][None, :], b[:, 1][None, :]
    d1 = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)
    d2 = (bx - ax) * (dy - ay) - (by - ay) * (dx - ax)
    d3 = (dx - cx) * (ay - cy) - (dy - cy) * (ax - cx)
    d4 = (dx - cx) * (by - cy) - (dy - cy) * (bx - cx)
    proper = (np.sign(d1) * np.sign(d2) < 0) & (np.sign(d3) * np.sign(d4) < 0)

    def on(px, py, qx, qy, rx, ry):  # r on the closed segment pq, given collinear
        return (np.minimum(px, qx) <= rx) & (rx <= np.maximum(px, qx)) & (np.minimum(py, qy) <= ry) & (ry <= np.maximum(py, qy))

    touch = ((d1 == 0) & on(ax, ay, bx, by, cx, cy)) | ((d2 == 0) & on(ax, ay, bx, by, dx, dy)) | \
            ((d3 == 0) & on(cx, cy, dx, dy, ax, ay)) | ((d4 == 0) & on(cx, cy, dx, dy, bx, by))
    hit = proper | touch
    i, j = np.indices((n, n))
    adjacent = (i == j) | ((i + 1) % n == j) | ((j + 1) % n == i)
    return not bool(np.any(hit & ~adjacent))


# --------------------------------------------------------------------------- meridian split
def _ring_area2(face):
    s = Fraction(0)
    m = len(face)
    for i in range(m):
        x0, y0 = face[i]
        x1, y1 = face[(i + 1) % m]
        s += x0 * y1 - x1 * y0
    return s


def _winding(px, py, ring):
    """Winding number of the closed ring around (px, py) (exact; the point must not lie on the ring)."""
    w = 0
    n = len(ring)
    for i in range(n):
        x0, y0 = ring[i]
        x1, y1 = ring[(i + 1) % n]
        cr = (x1 - x0) * (py - y0) - (y1 - y0) * (px - x0)
        if y0 <= py < y1 and cr > 0:
            w += 1
        elif y1 <= py < y0 and cr < 0:
            w -= 1
    return w


def _interior_point(face):
    """A point strictly inside the simple counter-clockwise polygon ``face`` (Fraction coordinates)."""
    n = len(face)
    k = min(range(n), key=lambda i: face[i])  # lexicographically smallest vertex: convex
    a, b, c = face[k - 1], face[k], face[(k + 1) % n]

    def inside_tri(p):
        d1 = (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
        d2 = (c[0] - b[0]) * (p[1] - b[1]) - (c[1] - b[1]) * (p[0] - b[0])
        d3 = (a[0] - c[0]) * (p[1] - c[1]) - (a[1] - c[1]) * (p[0] - c[0])
        return d1 >= 0 and d2 >= 0 and d3 >= 0

    best, best_d = None, None
    for i, p in enumerate(face):
        if p in (a, b, c):
            continue
        if inside_tri(p):
            # distance from the line ac towards b: the vertex closest to b wins
            d = abs((c[0] - a[0]) * (p[1] - a[1]) - (c[1] - a[1]) * (p[0] - a[0]))
            if best is None or d > best_d:
                best, best_d = p, d
    if best is None:
        return ((a[0] + b[0] + c[0]) / 3, (a[1] + b[1] + c[1]) / 3)
    return ((b[0] + best[0]) / 2, (b[1] + best[1]) / 2)


def split_ring_at_meridian(ring, nlon):
    """Pieces of an index-space ring after the split of ``utils/index_utils.py:148-173``.

    The reference overlays the polygon boundary with the boundary of the strip between x = nlon-1 and
    x = nlon, polygonises the union and keeps the faces that are not inside the strip and inside the
    polygon.  This is restated literally as a PLANAR-GRAPH FACE WALK with exact rational arithmetic
    (an independent formulation: the product clips chains against the two lines instead):

    1. the polygon edges are subdivided where they cross the two vertical lines; the lines themselves are cut
       at every node that lies on them (the strip's horizontal edges cannot separate faces of the polygon);
    2. at every node the outgoing half-edges are sorted by angle; a face is traced by always turning to the next
       half-edge clockwise from the one just arrived on (the face lies to the left);
    3. bounded faces (positive area) are kept if an interior point lies outside the strip and inside the polygon
       (non-zero winding).
    The kept faces' coordinates are truncated to ints (``:135``) and folded with ``x % nlon`` (``:139``).  Returns a
    list of (n, 2) int arrays (open rings, vertex order of the face walk); an empty list stands for ``Polygon()``.
    """
    ring = [(int(x), int(y)) for x, y in np.asarray(ring).reshape(-1, 2)]
    n = len(ring)
    c0, c1 = nlon - 1, nlon
    segs = set()
    on_line = {c0: set(), c1: set()}

    def node(x, y):
        p = (Fraction(x), Fraction(y))
        for c in (c0, c1):
            if p[0] == c:
                on_line[c].add(p)
        return p

    for i in range(n):
        (x0, y0), (x1, y1) = ring[i], ring[(i + 1) % n]
        if (x0, y0) == (x1, y1):
            continue
        pts = [node(x0, y0), node(x1, y1)]
        for c in (c0, c1):
            if min(x0, x1) < c < max(x0, x1):
                pts.append(node(c, Fraction(y0) + Fraction((c - x0) * (y1 - y0), x1 - x0)))
        # order along the edge
        pts.sort(key=lambda p: (p[0] - x0) * (x1 - x0) + (p[1] - y0) * (y1 - y0))
        for a, b in zip(pts[:-1], pts[1:]):
            if a != b:
                segs.add((a, b) if a < b else (b, a))
    for c in (c0, c1):
        line = sorted(on_line[c])
        for a, b in zip(line[:-1], line[1:]):
            segs.add((a, b))
    # half-edges sorted by angle around every node
    out = {}
    for a, b in segs:
        out.setdefault(a, []).append(b)
        out.setdefault(b, []).append(a)

    def angle_key(o):
        def key(p):
            dx, dy = p[0] - o[0], p[1] - o[1]
            half = 0 if (dy > 0 or (dy == 0 and dx > 0)) else 1
            return half, dx, dy
        return key

    import functools

    def cmp_at(o):
        def cmp(p, q):
            kp, kq = angle_key(o)(p), angle_key(o)(q)
            if kp[0] != kq[0]:
                return -1 if kp[0] < kq[0] else 1
            cr = kp[1] * kq[2] - kp[2] * kq[1]  # p before q (counter-clockwise) iff cross > 0
            return -1 if cr > 0 else (1 if cr < 0 else 0)
        return cmp

    for o in out:
        out[o].sort(key=functools.cmp_to_key(cmp_at(o)))
    visited = set()
    pieces = []
    for a in out:
        for b in out[a]:
            if (a, b) in visited:
                continue
            face = []
            u, v = a, b
            while (u, v) not in visited:
                visited.add((u, v))
                face.append(u)
                nb = out[v]
                k = nb.index(u)
                u, v = v, nb[k - 1]  # next half-edge: clockwise neighbour of the reverse edge (face on the left)
            if len(face) < 3 or _ring_area2(face) <= 0:
                continue  # the unbounded face (clockwise) and degenerate walks
            px, py = _interior_point(face)
            if c0 < px < c1:
                continue  # inside the strip
            if _winding(px, py, ring) == 0:
                continue  # a hole of the overlay, not part of the polygon
            xy = np.array([[int(x) % nlon, int(y)] for x, y in face], dtype=np.int64)
            keep = np.ones(len(xy), dtype=bool)
            keep[1:] = np.any(xy[1:] != xy[:-1], axis=1)
            if len(xy) > 1 and np.all(xy[0] == xy[-1]):
                keep[-1] = False
            # a sliver face may collapse to two points (or one) under the truncation: it is kept -- GEOS buffers a
            # degenerate ring like the segment / point it collapses to, so to_xarray still flags the cells next to it
            pieces.append(xy[keep])
    return pieces


def canonical_region(piece):
    """Canonical vertex list of a lattice ring as a REGION: collinear vertices removed, counter-clockwise, starting at
    the lexicographically smallest vertex (two rings that bound the same region compare equal)."""
    p = [(int(x), int(y)) for x, y in np.asarray(piece).reshape(-1, 2)]
    changed = True
    while changed and len(p) >= 3:
        changed = False
        for i in range(len(p)):
            a, b, c = p[i - 1], p[i], p[(i + 1) % len(p)]
            if (b[0] - a[0]) * (c[1] - b[1]) - (b[1] - a[1]) * (c[0] - b[0]) == 0:
                del p[i]
                changed = True
                break
    if len(p) < 3:
        return ()
    if _ring_area2(p) < 0:
        p = p[::-1]
    k = p.index(min(p))
    return tuple(p[k:] + p[:k])


# --------------------------------------------------------------------------- overlap area
def _ring_edges(rings):
    segs = []
    for ring in rings:
        ring = np.asarray(ring, dtype=np.float64)
        nxt = np.roll(ring, -1, axis=0)
        for a, b in zip(ring, nxt):
            if a[0] != b[0] or a[1] != b[1]:
                segs.append((a[0], a[1], b[0], b[1]))
    return np.array(segs, dtype=np.float64).reshape(-1, 4)


def overlap_areas(rings_a, rings_b):
    """(area(A), area(B), area(A n B)) for two (multi)polygons given as ring lists.

    Slab decomposition: between consecutive event abscissae (vertices and edge-edge
    intersections) no two edges cross, so the edges that span the slab can be sorted
    by their ordinate at the slab centre and the winding numbers of A and B swept
    upwards.  Regions are defined by non-zero winding.  fp64 throughout.
    """
    ea, eb = _ring_edges(rings_a), _ring_edges(rings_b)
    if len(ea) == 0 or len(eb) == 0:
        def _area(e):
            return abs(float(np.sum(e[:, 0] * e[:, 3] - e[:, 2] * e[:, 1])) / 2.0) if len(e) else 0.0
        return _area(ea), _area(eb), 0.0
    edges = np.concatenate([ea, eb])
    owner = np.concatenate([np.zeros(len(ea), dtype=np.int64), np.ones(len(eb), dtype=np.int64)])
    xs = set(edges[:, 0].tolist()) | set(edges[:, 2].tolist())
    # edge-edge intersections (all pairs; the oracle is for small cases)
    x1, y1, x2, y2 = edges[:, 0][:, None], edges[:, 1][:, None], edges[:, 2][:, None], edges[:, 3][:, None]
    x3, y3, x4, y4 = edges[:, 0][None, :], edges[:, 1][None, :], edges[:, 2][None, :], edges[:, 3][None, :]
    den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = ((x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)) / den
        u = ((x1 - x3) * (y1 - y2) - (y1 - y3) * (x1 - x2)) / den
    hit = (den != 0) & (t > 0) & (t < 1) & (u > 0) & (u < 1)
    ii, jj = np.nonzero(hit)
    for i, j in zip(ii, jj):
        xs.add(float(edges[i, 0] + t[i, j] * (edges[i, 2] - edges[i, 0])))
    xs = np.array(sorted(xs))
    area = np.zeros(3)  # A, B, A&B
    dxe = edges[:, 2] - edges[:, 0]
    for xl, xr in zip(xs[:-1], xs[1:]):
        if xr <= xl:
            continue
        xm = 0.5 * (xl + xr)
        span = (np.minimum(edges[:, 0], edges[:, 2]) <= xl) & (np.maximum(edges[:, 0], edges[:, 2]) >= xr) & (dxe != 0)
        idx = np.nonzero(span)[0]
        if len(idx) == 0:
            continue
        slope = (edges[idx, 3] - edges[idx, 1]) / dxe[idx]
        ym = edges[idx, 1] + slope * (xm - edges[idx, 0])
        order = np.argsort(ym, kind="stable")
        idx, ym, slope = idx[order], ym[order], slope[order]
        # crossing an edge upwards: winding changes by +1 if the edge runs left-to-right
        dw = np.where(dxe[idx] > 0, 1, -1)
        wa = np.cumsum(np.where(owner[idx] == 0, dw, 0))
        wb = np.cumsum(np.where(owner[idx] == 1, dw, 0))
        h = (ym[1:] - ym[:-1]) * (xr - xl)  # trapezoid area = mid-height * width
        ina, inb = wa[:-1] != 0, wb[:-1] != 0
        area[0] += h[ina].sum()
        area[1] += h[inb].sum()
        area[2] += h[ina & inb].sum()
    return float(area[0]), float(area[1]), float(area[2])


# --------------------------------------------------------------------------- exact overlap decision
def _sweep_exact(rings_a, rings_b, xl=None, xr=None, stop_at_first=False):
    """Exact slab sweep over two ring sets (non-zero winding each): Fractions (area A, area B, area A n B) of the
    part of the plane with xl <= x <= xr (default: everything).  With ``stop_at_first`` the sweep returns as soon
    as a piece of positive area lies in both."""
    def edges_of(rings):
        out = []
        for ring in rings:
            ring = [(int(x), int(y)) for x, y in np.asarray(ring).reshape(-1, 2)]
            for i in range(len(ring)):
                a, b = ring[i], ring[(i + 1) % len(ring)]
                if a != b:
                    out.append((a[0], a[1], b[0], b[1]))
        return out

    ea, eb = edges_of(rings_a), edges_of(rings_b)
    edges = [(e, 0) for e in ea] + [(e, 1) for e in eb]
    edges = [(e, o) for e, o in edges if e[0] != e[2]]  # vertical edges bound no area in a vertical-slab sweep
    if not edges:
        return Fraction(0), Fraction(0), Fraction(0)
    lo = min(min(e[0], e[2]) for e, _ in edges) if xl is None else xl
    hi = max(max(e[0], e[2]) for e, _ in edges) if xr is None else xr
    if lo >= hi:
        return Fraction(0), Fraction(0), Fraction(0)
    edges = [(e, o) for e, o in edges if min(e[0], e[2]) < hi and max(e[0], e[2]) > lo]
    E = np.array([e for e, _ in edges], dtype=np.int64).reshape(-1, 4)
    xs = {Fraction(lo), Fraction(hi)}
    for e, _ in edges:
        for x in (e[0], e[2]):
            if lo < x < hi:
                xs.add(Fraction(x))
    n = len(edges)
    if n > 1:  # intersection abscissae of every pair of edges (integer prefilter, exact rational points)
        x1, y1, x2, y2 = (E[:, k][:, None] for k in range(4))
        x3, y3, x4, y4 = (E[:, k][None, :] for k in range(4))
        den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
        tn = (x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)
        un = (x1 - x3) * (y1 - y2) - (y1 - y3) * (x1 - x2)
        sg = np.sign(den)
        hit = (den != 0) & (tn * sg > 0) & (tn * sg < den * sg) & (un * sg > 0) & (un * sg < den * sg)
        for i, j in zip(*np.nonzero(np.triu(hit, 1))):
            x = Fraction(int(E[i, 0])) + Fraction(int(tn[i, j]), int(den[i, j])) * int(E[i, 2] - E[i, 0])
            if lo < x < hi:
                xs.add(x)
    xs = sorted(xs)
    area = [Fraction(0), Fraction(0), Fraction(0)]
    for a, b in zip(xs[:-1], xs[1:]):
        xm = (a + b) / 2
        span = []
        for (x0, y0, x1_, y1_), owner in edges:
            l0, h0 = (x0, x1_) if x0 < x1_ else (x1_, x0)
            if l0 <= a and h0 >= b:
                ym = Fraction(y0) + Fraction(y1_ - y0, x1_ - x0) * (xm - x0)
                span.append((ym, owner, 1 if x1_ > x0 else -1))
        span.sort(key=lambda t: t[0])
        wa = wb = 0
        for k in range(len(span) - 1):
            ym, owner, dw = span[k]
            if owner == 0:
                wa += dw
            else:
                wb += dw
            h = (span[k + 1][0] - ym) * (b - a)  # mid-height x width = exact trapezoid area
            if h > 0:
                if wa != 0:
                    area[0] += h
                if wb != 0:
                    area[1] += h
                if wa != 0 and wb != 0:
                    area[2] += h
                    if stop_at_first:
                        return tuple(area)
    return tuple(area)


def regions_equal(rings_a, rings_b):
    """True iff the two ring sets cover the same region (exact: area(A) == area(B) == area(A n B))."""
    a, b, ab = _sweep_exact(rings_a, rings_b)
    return a == b == ab


def overlap_positive_exact(rings_a, rings_b):
    """True iff the (multi)polygons A and B (lattice rings, non-zero winding) share a region of positive area --
    what ``geom1.intersection(geom2).area > 0`` means in ``processing/events.py:205-214`` (GEOS returns an empty or
    lower-dimensional intersection for polygons that merely touch).

    Exact rational slab decomposition (:func:`_sweep_exact`), independent of the product's crossing / touching rules:
    between two consecutive abscissae of the vertices and edge-edge intersection points no two edges cross, so the
    edges spanning the slab are ordered by their ordinate at the slab centre and the winding numbers of A and B are
    swept upwards; a slab piece where both are non-zero and whose centre height is positive proves the overlap.
    Only the part of the plane where the bounding boxes intersect is swept.
    """
    def xrange(rings):
        xs = [int(x) for ring in rings for x in np.asarray(ring).reshape(-1, 2)[:, 0]]
        return (min(xs), max(xs)) if xs else None

    ra, rb = xrange(rings_a), xrange(rings_b)
    if ra is None or rb is None:
        return False
    xl, xr = max(ra[0], rb[0]), min(ra[1], rb[1])
    if xl >= xr:
        return False
    return _sweep_exact(rings_a, rings_b, xl, xr, stop_at_first=True)[2] > 0
