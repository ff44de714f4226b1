"""The oracle's restatement of shapely's LineString ``touches`` / ``intersects`` (GEOS DE-9IM), pinned
(a) on the examples the Shapely manual publishes for these predicates and (b) against an evaluator written straight
from the DE-9IM definitions (exact rational intersection of every segment pair; independent of the rule set in
``oracle.geom.chord_touches_polyline`` and of the CUDA chord test):

    A.touches(B)     <=>  A n B != {}  and  interior(A) n interior(B) == {}
    A.intersects(B)  <=>  A n B != {}
    interior(polyline) = the polyline without its two end points (an open, non-closed line)
"""

from fractions import Fraction

import numpy as np
import pytest

from oracle import geom as G


def _seg_intersection(a0, a1, b0, b1):
    """Intersection of two closed segments: None, a point (x, y) or a pair of points (an overlap), exact."""
    (x1, y1), (x2, y2), (x3, y3), (x4, y4) = a0, a1, b0, b1
    den = (x2 - x1) * (y4 - y3) - (y2 - y1) * (x4 - x3)
    if den != 0:
        t = Fraction((x3 - x1) * (y4 - y3) - (y3 - y1) * (x4 - x3), den)
        u = Fraction((x3 - x1) * (y2 - y1) - (y3 - y1) * (x2 - x1), den)
        if 0 <= t <= 1 and 0 <= u <= 1:
            return (x1 + t * (x2 - x1), y1 + t * (y2 - y1))
        return None
    if (x3 - x1) * (y2 - y1) - (y3 - y1) * (x2 - x1) != 0:
        return None  # parallel, not collinear
    # collinear: project on the direction of a
    dx, dy = x2 - x1, y2 - y1
    l2 = dx * dx + dy * dy
    tb = sorted((Fraction((x3 - x1) * dx + (y3 - y1) * dy, l2), Fraction((x4 - x1) * dx + (y4 - y1) * dy, l2)))
    lo, hi = max(tb[0], 0), min(tb[1], 1)
    if lo > hi:
        return None
    p, q = (x1 + lo * dx, y1 + lo * dy), (x1 + hi * dx, y1 + hi * dy)
    return p if lo == hi else (p, q)


def de9im_line_line(A, B):
    """(intersects, touches) of two open polylines from the definitions."""
    A = [tuple(map(int, p)) for p in A]
    B = [tuple(map(int, p)) for p in B]
    bndA = {A[0], A[-1]} if A[0] != A[-1] else set()
    bndB = {B[0], B[-1]} if B[0] != B[-1] else set()
    meets = interiors_meet = False
    for a0, a1 in zip(A[:-1], A[1:]):
        for b0, b1 in zip(B[:-1], B[1:]):
            r = _seg_intersection(a0, a1, b0, b1)
            if r is None:
                continue
            meets = True
            if isinstance(r[0], tuple):  # an overlap of positive length always contains interior points of both
                interiors_meet = True
            elif r not in bndA and r not in bndB:
                interiors_meet = True
    return meets, meets and not interiors_meet


def test_shapely_manual_examples():
    """Shapely manual, 'Binary Predicates': ``LineString([(0, 0), (1, 1)]).touches(LineString([(1, 1), (2, 2)]))`` is
    True; ``LineString([(0, 0), (1, 1)]).crosses(LineString([(0, 1), (1, 0)]))`` is True, hence they intersect and do
    NOT touch (crossing lines share interior points); disjoint lines neither intersect nor touch."""
    assert de9im_line_line([(0, 0), (1, 1)], [(1, 1), (2, 2)]) == (True, True)
    assert G.segments_intersect((0, 0), (1, 1), (1, 1), (2, 2))
    # crossing at (1/2, 1/2) -- scaled by 2 onto the lattice
    assert de9im_line_line([(0, 0), (2, 2)], [(0, 2), (2, 0)]) == (True, False)
    assert G.segments_intersect((0, 0), (2, 2), (0, 2), (2, 0))
    assert de9im_line_line([(0, 0), (1, 1)], [(3, 0), (4, 1)]) == (False, False)
    assert not G.segments_intersect((0, 0), (1, 1), (3, 0), (4, 1))


def test_chord_touches_polyline_known_cases():
    # polyline with a bulge; chord between two of its vertices
    pts = np.array([[0, 0], [2, 0], [2, 2], [4, 2], [4, 0], [6, 0]])
    assert G.chord_touches_polyline(pts[1], pts[4], pts)          # (2,0)-(4,0): only its end points lie on the line
    assert G.chord_touches_polyline(pts[0], pts[5], pts) is False  # (0,0)-(6,0) runs along two segments of the line
    assert G.chord_touches_polyline(pts[2], pts[4], pts)          # diagonal (2,2)-(4,0) touches only at its end points
    assert G.chord_touches_polyline(pts[0], pts[3], pts) is False  # (0,0)-(4,2) crosses the segment (2,0)-(2,2) at (2,1)
    for i, j in ((1, 4), (0, 5), (2, 4), (0, 3), (0, 2), (1, 3)):
        assert G.chord_touches_polyline(pts[i], pts[j], pts) == de9im_line_line([pts[i], pts[j]], pts)[1], (i, j)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_chord_touches_polyline_vs_definition_random(seed):
    """random lattice polylines with distinct vertices (contours are deduplicated) on a small lattice -- collinear
    runs, vertices inside chords and chords through the polyline's end points are frequent"""
    rng = np.random.default_rng(seed)
    checked = touching = 0
    for _ in range(120):
        n = int(rng.integers(4, 12))
        pts, seen = [], set()
        p = (int(rng.integers(0, 8)), int(rng.integers(0, 8)))
        while len(pts) < n:
            if p not in seen:
                pts.append(p)
                seen.add(p)
            p = (int(np.clip(p[0] + rng.integers(-2, 3), 0, 8)), int(np.clip(p[1] + rng.integers(-2, 3), 0, 8)))
        pts = np.array(pts)
        for _ in range(6):
            i, j = sorted(rng.choice(n, 2, replace=False))
            want = de9im_line_line([pts[i], pts[j]], pts)[1]
            assert G.chord_touches_polyline(pts[i], pts[j], pts) == want, (pts.tolist(), i, j)
            checked += 1
            touching += want
    assert checked > 500 and 20 < touching < checked - 20


@pytest.mark.parametrize("seed", [0, 1])
def test_segments_intersect_vs_definition_random(seed):
    rng = np.random.default_rng(seed)
    for _ in range(2000):
        a0, a1, b0, b1 = (tuple(int(v) for v in rng.integers(0, 6, 2)) for _ in range(4))
        if a0 == a1 or b0 == b1:
            continue
        assert G.segments_intersect(a0, a1, b0, b1) == (_seg_intersection(a0, a1, b0, b1) is not None)
