"""Restatement of ``skimage.measure.find_contours`` (test infrastructure only).

The reference calls ``measure.find_contours(ds.values, level)`` with defaults
(``wavebreaking/indices/contour_index.py:103``).  scikit-image is an unpinned,
un-vendored dependency (``setup.py:13-24``) and is not installed in this image,
so its published algorithm is restated here:

* ``_get_contour_segments`` (``skimage/measure/_find_contours_cy.pyx``): raster
  scan over the (R-1)x(C-1) squares, strict ``>`` against the level, squares
  with a NaN corner skipped, linear interpolation
  ``(level - from) / (to - from)`` along the crossed edges (0 when the two
  values are equal), directed segments with the low values on the left,
  ``fully_connected='low'`` for the two saddle cases.
* ``_assemble_contours`` (``skimage/measure/_find_contours.py``): sequential
  joining through two dicts keyed by the float point tuples; when two partial
  contours join the one that was created first survives; the result is sorted
  by creation number; ``positive_orientation='low'`` keeps the direction.

Points are (row, col) float64 pairs, exactly as skimage returns them.
"""

from collections import deque

import numpy as np

# (from_edge, to_edge) per marching-squares case; edges: 0=top 1=bottom 2=left 3=right
_T, _B, _L, _R = 0, 1, 2, 3
CASE_SEGMENTS = {
    1: [(_T, _L)],
    2: [(_R, _T)],
    3: [(_R, _L)],
    4: [(_L, _B)],
    5: [(_T, _B)],
    6: [(_R, _T), (_L, _B)],  # fully_connected == 'low'
    7: [(_R, _B)],
    8: [(_B, _R)],
    9: [(_T, _L), (_B, _R)],  # fully_connected == 'low'
    10: [(_B, _T)],
    11: [(_B, _L)],
    12: [(_L, _R)],
    13: [(_T, _R)],
    14: [(_L, _T)],
}


def _fraction(from_value, to_value, level):
    """``_get_fraction``: 0 if the values are equal else (level-from)/(to-from)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        frac = (level - from_value) / (to_value - from_value)
    return np.where(to_value == from_value, 0.0, frac)


def get_contour_segments(array, level):
    """Vectorised restatement of ``_get_contour_segments``.

    Returns ``(seg_from, seg_to, seg_rid)``: float64 (n, 2) arrays of (row, col)
    points in skimage's raster order and the raster id ``2*square + sub`` of
    every segment (``square = r0 * (C-1) + c0``).
    """
    a = np.asarray(array, dtype=np.float64)
    level = float(level)
    ul, ur = a[:-1, :-1], a[:-1, 1:]
    ll, lr = a[1:, :-1], a[1:, 1:]
    nan = np.isnan(ul) | np.isnan(ur) | np.isnan(ll) | np.isnan(lr)
    with np.errstate(invalid="ignore"):
        case = (
            (ul > level).astype(np.int8)
            + 2 * (ur > level).astype(np.int8)
            + 4 * (ll > level).astype(np.int8)
            + 8 * (lr > level).astype(np.int8)
        )
    active = (~nan) & (case != 0) & (case != 15)
    r0, c0 = np.nonzero(active)  # row-major == raster order
    if len(r0) == 0:
        z = np.zeros((0, 2))
        return z, z.copy(), np.zeros(0, dtype=np.int64)
    cs = case[r0, c0]
    vul, vur, vll, vlr = ul[r0, c0], ur[r0, c0], ll[r0, c0], lr[r0, c0]
    r0f, c0f = r0.astype(np.float64), c0.astype(np.float64)
    # edge points (row, col)
    pts = np.empty((len(r0), 4, 2))
    pts[:, _T, 0] = r0f
    pts[:, _T, 1] = c0f + _fraction(vul, vur, level)
    pts[:, _B, 0] = r0f + 1.0
    pts[:, _B, 1] = c0f + _fraction(vll, vlr, level)
    pts[:, _L, 0] = r0f + _fraction(vul, vll, level)
    pts[:, _L, 1] = c0f
    pts[:, _R, 0] = r0f + _fraction(vur, vlr, level)
    pts[:, _R, 1] = c0f + 1.0

    first_from = np.zeros(16, dtype=np.int64)
    first_to = np.zeros(16, dtype=np.int64)
    second_from = np.zeros(16, dtype=np.int64)
    second_to = np.zeros(16, dtype=np.int64)
    nseg = np.zeros(16, dtype=np.int64)
    for k, segs in CASE_SEGMENTS.items():
        nseg[k] = len(segs)
        first_from[k], first_to[k] = segs[0]
        if len(segs) == 2:
            second_from[k], second_to[k] = segs[1]

    n_per = nseg[cs]
    sq = np.repeat(np.arange(len(cs)), n_per)
    # sub index 0/1 inside a square
    starts = np.cumsum(n_per) - n_per
    sub = np.arange(len(sq)) - np.repeat(starts, n_per)
    e_from = np.where(sub == 0, first_from[cs[sq]], second_from[cs[sq]])
    e_to = np.where(sub == 0, first_to[cs[sq]], second_to[cs[sq]])
    seg_from = pts[sq, e_from]
    seg_to = pts[sq, e_to]
    ncols_sq = a.shape[1] - 1
    rid = 2 * (r0[sq].astype(np.int64) * ncols_sq + c0[sq]) + sub
    return seg_from, seg_to, rid


def assemble_contours(seg_from, seg_to):
    """Restatement of ``_assemble_contours`` (sequential dict/deque joining)."""
    current_index = 0
    contours = {}
    starts = {}
    ends = {}
    f_list = list(map(tuple, seg_from.tolist()))
    t_list = list(map(tuple, seg_to.tolist()))
    for from_point, to_point in zip(f_list, t_list):
        if from_point == to_point:  # degenerate segment
            continue
        tail, tail_num = starts.pop(to_point, (None, None))
        head, head_num = ends.pop(from_point, (None, None))

        if tail is not None and head is not None:
            if tail is head:
                head.append(to_point)  # close the ring
            else:
                if tail_num > head_num:
                    head.extend(tail)
                    contours.pop(tail_num, None)
                    starts[head[0]] = (head, head_num)
                    ends[head[-1]] = (head, head_num)
                else:
                    tail.extendleft(reversed(head))
                    starts.pop(head[0], None)
                    contours.pop(head_num, None)
                    starts[tail[0]] = (tail, tail_num)
                    ends[tail[-1]] = (tail, tail_num)
        elif tail is None and head is None:
            new_contour = deque((from_point, to_point))
            contours[current_index] = new_contour
            starts[from_point] = (new_contour, current_index)
            ends[to_point] = (new_contour, current_index)
            current_index += 1
        elif head is None:
            tail.appendleft(from_point)
            starts[from_point] = (tail, tail_num)
        else:
            head.append(to_point)
            ends[to_point] = (head, head_num)

    return [np.array(contour) for _, contour in sorted(contours.items())]


def find_contours(image, level):
    """``skimage.measure.find_contours(image, level)`` with default arguments."""
    image = np.asarray(image)
    if image.ndim != 2:
        raise ValueError("Only 2D arrays are supported.")
    seg_from, seg_to, _ = get_contour_segments(image, level)
    return assemble_contours(seg_from, seg_to)


def has_lattice_vertex_points(image, level):
    """True if a contour vertex of ``image`` at ``level`` falls exactly on a grid vertex.

    That is the one situation in which skimage's float-tuple joining differs from
    joining by edge identity (used on the GPU); it needs a grid value exactly equal
    to the level (or an interpolated coordinate that rounds to an integer).
    """
    seg_from, seg_to, _ = get_contour_segments(image, level)
    pts = np.concatenate([seg_from, seg_to])
    if len(pts) == 0:
        return False
    return bool(np.any((pts[:, 0] == np.rint(pts[:, 0])) & (pts[:, 1] == np.rint(pts[:, 1]))))
