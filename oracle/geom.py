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
            # on the edge itself (also covers r == 0)
            on = (cross == 0) & (dot >= 0) & (dot <= len2)
            disc = (cx - x0) ** 2 + (cy - y0) ** 2 < r2
            near |= (within | on | disc | ((cx == x0) & (cy == y0))).any(axis=1)
        out[cand] |= (wn != 0) | near
    return out


def ring_is_simple(ring):
    """True if the closed lattice ring neither touches nor crosses itself."""
    ring = np.asarray(ring, dtype=np.int64)
    n = len(ring)
    if n < 3:
        return False
    nxt = np.roll(ring, -1, axis=0)
    for i in range(n):
        for j in range(i + 1, n):
            adjacent = (j == i + 1) or (i == 0 and j == n - 1)
            if adjacent:
                # adjacent edges share one vertex; they must not fold back onto each other
                if j == i + 1:
                    a, b, c = ring[i], ring[j], nxt[j]
                else:
                    a, b, c = ring[j], ring[i], nxt[i]
                if _orient(a[0], a[1], b[0], b[1], c[0], c[1]) == 0:
                    if (a[0] - b[0]) * (c[0] - b[0]) + (a[1] - b[1]) * (c[1] - b[1]) > 0:
                        return False
                continue
            if segments_intersect(ring[i], nxt[i], ring[j], nxt[j]):
                return False
    return True


# --------------------------------------------------------------------------- meridian split
def _clip_ring_vertical(ring, c, keep_le):
    """Faces of ``ring`` on one side of the vertical line x = c (kept side includes the line).

    Returns a list of rings whose vertices are (x, Fraction(y)) tuples.  Faces are
    separated (no bridge edges along the cut between different faces): crossings on
    the line are sorted by y and paired (even-odd along the line), then every chain of
    kept vertices is linked from its exit crossing to the paired entry crossing.
    """
    n = len(ring)
    xs = [int(v[0]) for v in ring]
    ys = [int(v[1]) for v in ring]
    inside = [(x <= c) if keep_le else (x >= c) for x in xs]
    if all(inside):
        return [[(x, Fraction(y)) for x, y in zip(xs, ys)]]
    if not any(inside):
        return []

    def cross_y(i, j):
        # intersection of edge i->j with x = c (exact)
        return Fraction(ys[i]) + Fraction((c - xs[i]) * (ys[j] - ys[i]), xs[j] - xs[i])

    # start at an entry: previous outside, current inside
    start = next(k for k in range(n) if inside[k] and not inside[k - 1])
    chains = []  # each: dict(points=[...], y_in=..., y_out=...)
    k = start
    visited = 0
    while visited < n:
        if inside[k] and not inside[k - 1]:
            pts = []
            prev = (k - 1) % n
            if xs[k] != c:
                y_in = cross_y(prev, k)
                pts.append((c, y_in))
            else:
                y_in = Fraction(ys[k])
            j = k
            while inside[j]:
                pts.append((xs[j], Fraction(ys[j])))
                j = (j + 1) % n
                visited += 1
            last = (j - 1) % n
            if xs[last] != c:
                y_out = cross_y(last, j)
                pts.append((c, y_out))
            else:
                y_out = Fraction(ys[last])
            chains.append({"pts": pts, "y_in": y_in, "y_out": y_out})
            k = j
        else:
            k = (k + 1) % n
            visited += 1

    # pair the crossings along the line
    crossings = []
    for ci, ch in enumerate(chains):
        crossings.append((ch["y_in"], 0, ci))  # 0 = entry
        crossings.append((ch["y_out"], 1, ci))  # 1 = exit
    crossings.sort(key=lambda t: (t[0], t[2], t[1]))
    partner_of_exit = {}
    for a in range(0, len(crossings) - 1, 2):
        c0, c1 = crossings[a], crossings[a + 1]
        if c0[1] == 1 and c1[1] == 0:
            partner_of_exit[c0[2]] = c1[2]
        elif c0[1] == 0 and c1[1] == 1:
            partner_of_exit[c1[2]] = c0[2]
        else:  # non-simple ring: fall back to closing every chain on itself
            partner_of_exit = {ci: ci for ci in range(len(chains))}
            break

    faces = []
    used = [False] * len(chains)
    for ci in range(len(chains)):
        if used[ci]:
            continue
        face = []
        cur = ci
        while not used[cur]:
            used[cur] = True
            face.extend(chains[cur]["pts"])
            cur = partner_of_exit.get(cur, cur)
        faces.append(face)
    return faces


def _ring_area2(face):
    s = Fraction(0)
    m = len(face)
    for i in range(m):
        x0, y0 = face[i]
        x1, y1 = face[(i + 1) % m]
        s += x0 * y1 - x1 * y0
    return s


def split_ring_at_meridian(ring, nlon):
    """Pieces of an index-space ring after the split of ``utils/index_utils.py:148-173``.

    The strip between x = nlon-1 and x = nlon is removed, the faces on either side
    are kept as separate pieces, their coordinates are truncated to ints
    (``:135``) and folded with ``x % nlon`` (``:139``).  Returns a list of (n, 2)
    int arrays (open rings); an empty list stands for the empty ``Polygon()``.
    """
    pieces = []
    for c, keep_le in ((nlon - 1, True), (nlon, False)):
        for face in _clip_ring_vertical(ring, c, keep_le):
            if len(face) < 3 or _ring_area2(face) == 0:
                continue
            xy = np.array([[int(x) % nlon, int(y)] for x, y in face], dtype=np.int64)
            # drop consecutive repeats created by the truncation
            keep = np.ones(len(xy), dtype=bool)
            keep[1:] = np.any(xy[1:] != xy[:-1], axis=1)
            if len(xy) > 1 and np.all(xy[0] == xy[-1]):
                keep[-1] = False
            pieces.append(xy[keep])
    return pieces


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
def overlap_positive_exact(rings_a, rings_b):
    """True iff the (multi)polygons A and B (lattice rings, non-zero winding) share a region of positive area --
    what ``geom1.intersection(geom2).area > 0`` means in ``processing/events.py:205-214`` (GEOS returns an empty or
    lower-dimensional intersection for polygons that merely touch).

    Exact rational slab decomposition, independent of the product's crossing / touching rules: between two
    consecutive abscissae of the vertices and edge-edge intersection points no two edges cross, so the edges
    spanning the slab are ordered by their ordinate at the slab centre and the winding numbers of A and B are swept
    upwards; a slab piece where both are non-zero and whose centre height is positive proves the overlap.
    ``fractions.Fraction`` throughout; only the part of the plane where the bounding boxes intersect is swept.
    """
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
    if not ea or not eb:
        return False
    xa0, xa1 = min(min(e[0], e[2]) for e in ea), max(max(e[0], e[2]) for e in ea)
    xb0, xb1 = min(min(e[0], e[2]) for e in eb), max(max(e[0], e[2]) for e in eb)
    xl, xr = max(xa0, xb0), min(xa1, xb1)
    if xl >= xr:
        return False
    edges = [(e, 0) for e in ea] + [(e, 1) for e in eb]
    # only non-vertical edges that reach into (xl, xr) matter
    edges = [(e, o) for e, o in edges if e[0] != e[2] and min(e[0], e[2]) < xr and max(e[0], e[2]) > xl]
    E = np.array([e for e, _ in edges], dtype=np.int64).reshape(-1, 4)
    xs = {Fraction(xl), Fraction(xr)}
    for e, _ in edges:
        for x in (e[0], e[2]):
            if xl < x < xr:
                xs.add(Fraction(x))
    # intersection abscissae of every pair of these edges (integer prefilter, exact rational points)
    n = len(edges)
    if n > 1:
        x1, y1, x2, y2 = (E[:, k][:, None] for k in range(4))
        x3, y3, x4, y4 = (E[:, k][None, :] for k in range(4))
        den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
        tn = (x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)
        un = (x1 - x3) * (y1 - y2) - (y1 - y3) * (x1 - x2)
        sg = np.sign(den)
        hit = (den != 0) & (tn * sg > 0) & (tn * sg < den * sg) & (un * sg > 0) & (un * sg < den * sg)
        for i, j in zip(*np.nonzero(np.triu(hit, 1))):
            x = Fraction(int(E[i, 0])) + Fraction(int(tn[i, j]), int(den[i, j])) * int(E[i, 2] - E[i, 0])
            if xl < x < xr:
                xs.add(x)
    xs = sorted(xs)
    for a, b in zip(xs[:-1], xs[1:]):
        xm = (a + b) / 2
        span = []
        for (x0, y0, x1_, y1_), owner in edges:
            lo, hi = (x0, x1_) if x0 < x1_ else (x1_, x0)
            if lo <= a and hi >= b:
                ym = Fraction(y0) + Fraction(y1_ - y0, x1_ - x0) * (xm - x0)
                span.append((ym, owner, 1 if x1_ > x0 else -1))
        span.sort(key=lambda s: s[0])
        wa = wb = 0
        for k in range(len(span) - 1):
            ym, owner, dw = span[k]
            if owner == 0:
                wa += dw
            else:
                wb += dw
            if wa != 0 and wb != 0 and span[k + 1][0] > ym:
                return True
    return False
