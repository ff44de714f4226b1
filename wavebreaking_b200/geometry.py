"""Host-side lattice geometry of the event tables: fold to the real grid and the meridian split.

Reference: wavebreaking/utils/index_utils.py:129-184 (``transform_polygons``).  Events live in
index coordinates of the periodically extended grid; their rings are folded with ``x % nlon`` and,
when a ring has vertices on both sides of the last meridian, split into the faces left of
``x = nlon-1`` and right of ``x = nlon`` (the strip between the two columns is removed, the cut
vertices are truncated to ints).  Only those few straddling events are touched here; everything
else of the path runs in the CUDA kernels.  Integer arithmetic throughout.
"""

import numpy as np


def _clip_vertical(xs, ys, c, keep_le):
    """Faces of the closed ring (xs, ys) strictly on one side of x = c; list of (x, y_num, y_den) vertex lists.

    A chain is a maximal run of vertices STRICTLY beyond the line; it is entered and left through a crossing point
    on the line: the neighbouring ring vertex if that one lies on the line, else the exact intersection of the edge.
    Ring edges that run along the line belong to no chain -- whether they bound a face is decided by the pairing of
    the crossings (even-odd along the line), exactly as in the overlay + polygonize of the reference
    (index_utils.py:158-163): a vertex on the line never sticks out of a face as a zero-area antenna, and two parts
    of the polygon that only meet along the line stay separate faces."""
    n = len(xs)
    inside = (xs < c) if keep_le else (xs > c)
    if not inside.any():
        return []
    if inside.all():
        return [[(int(x), int(y), 1) for x, y in zip(xs, ys)]]

    def cross(i, j):
        # y of edge i->j at x = c as an exact fraction (num, den), den > 0 (the end point itself if it is on the line)
        if xs[i] == c:
            return int(ys[i]), 1
        if xs[j] == c:
            return int(ys[j]), 1
        den = int(xs[j] - xs[i])
        num = int(ys[i]) * den + (c - int(xs[i])) * int(ys[j] - ys[i])
        if den < 0:
            num, den = -num, -den
        return num, den

    start = next(k for k in range(n) if inside[k] and not inside[k - 1])
    chains = []
    k, visited = start, 0
    while visited < n:
        if inside[k] and not inside[k - 1]:
            prev = (k - 1) % n
            y_in = cross(prev, k)
            pts = [(c, y_in[0], y_in[1])]
            j = k
            while inside[j]:
                pts.append((int(xs[j]), int(ys[j]), 1))
                j = (j + 1) % n
                visited += 1
            last = (j - 1) % n
            y_out = cross(last, j)
            pts.append((c, y_out[0], y_out[1]))
            chains.append((pts, y_in, y_out))
            k = j
        else:
            k = (k + 1) % n
            visited += 1

    import functools

    def cmp(a, b):
        # compare fractions a[0] = (num, den), ties by chain id then entry-before-exit
        l, r = a[0][0] * b[0][1], b[0][0] * a[0][1]
        if l != r:
            return -1 if l < r else 1
        return (a[2] > b[2]) - (a[2] < b[2]) or (a[1] > b[1]) - (a[1] < b[1])

    crossings = []
    for ci, (_, y_in, y_out) in enumerate(chains):
        crossings.append((y_in, 0, ci))
        crossings.append((y_out, 1, ci))
    crossings.sort(key=functools.cmp_to_key(cmp))
    partner = {}
    for a in range(0, len(crossings) - 1, 2):
        c0, c1 = crossings[a], crossings[a + 1]
        if c0[1] == 1 and c1[1] == 0:
            partner[c0[2]] = c1[2]
        elif c0[1] == 0 and c1[1] == 1:
            partner[c1[2]] = c0[2]
        else:  # non-simple ring: close every chain on itself
            partner = {ci: ci for ci in range(len(chains))}
            break
    faces, used = [], [False] * len(chains)
    for ci in range(len(chains)):
        if used[ci]:
            continue
        face, cur = [], ci
        while not used[cur]:
            used[cur] = True
            face.extend(chains[cur][0])
            cur = partner.get(cur, cur)
        faces.append(face)
    return faces


def _area2_sign(face):
    """Sign-insensitive zero test of the exact doubled area of a face with rational y."""
    # sum over edges of (x0*y1 - x1*y0) with y = num/den: accumulate as a fraction
    num, den = 0, 1
    m = len(face)
    for i in range(m):
        x0, n0, d0 = face[i]
        x1, n1, d1 = face[(i + 1) % m]
        # x0*n1/d1 - x1*n0/d0
        tn = x0 * n1 * d0 - x1 * n0 * d1
        td = d0 * d1
        num = num * td + tn * den
        den = den * td
    return num != 0


def split_ring(ring, nlon):
    """Pieces (list of (n, 2) int arrays, folded) of an index-space ring split at the last meridian."""
    ring = np.asarray(ring, dtype=np.int64)
    xs, ys = ring[:, 0], ring[:, 1]
    pieces = []
    for c, keep_le in ((nlon - 1, True), (nlon, False)):
        for face in _clip_vertical(xs, ys, c, keep_le):
            if len(face) < 3 or not _area2_sign(face):
                continue
            xy = np.array([[x % nlon, (num // den) if num >= 0 else -((-num) // den)] for x, num, den in face],
                          dtype=np.int64)
            keep = np.ones(len(xy), dtype=bool)
            keep[1:] = np.any(xy[1:] != xy[:-1], axis=1)
            if len(xy) > 1 and np.all(xy[0] == xy[-1]):
                keep[-1] = False
            pieces.append(xy[keep])
    return pieces


def transform_ring(ring, nlon):
    """``transform_polygons`` for one event ring: list of folded integer pieces."""
    ring = np.asarray(ring, dtype=np.int64)
    x = ring[:, 0]
    if not (x >= nlon).any():
        return [ring.copy()]
    if (x >= nlon).all():
        out = ring.copy()
        out[:, 0] -= nlon
        out[:, 0] %= nlon
        return [out]
    return split_ring(ring, nlon)


def interleave_pieces(xy, off, split, pieces, first, nlon):
    """Ragged rings of a table of events after ``transform_polygons`` (utils/index_utils.py:129-184), vectorised:
    ordinary events keep their single ring folded with ``x % nlon``; events that straddle the last meridian
    (``split == 1``) are replaced by the pieces the device clipper produced (``wbk_split_fetch``: dict with ``ev`` =
    event row in the batch's kind-major order, ``off``, ``xy``).  ``first`` = row of the table's first event in that
    order.  Returns ``(xy int64 [N, 2], ring_off, poly_off)``: polygon e = rings [poly_off[e], poly_off[e + 1])."""
    from .tracking import _ranges

    n = len(off) - 1
    split = np.asarray(split)
    pev = np.asarray(pieces["ev"]).astype(np.int64) - first
    mine = np.nonzero((pev >= 0) & (pev < n))[0]
    order = mine[np.argsort(pev[mine], kind="stable")]
    poff = np.asarray(pieces["off"], dtype=np.int64)
    plen = np.diff(poff)[order]
    pxy = np.asarray(pieces["xy"])[_ranges(poff[:-1][order], plen)].astype(np.int64)
    npieces = np.bincount(pev[order], minlength=n).astype(np.int64)
    keep_ev = split != 1
    lens = np.diff(off)
    nrings = np.where(keep_ev, 1, npieces)
    poly_off = np.r_[0, np.cumsum(nrings)].astype(np.int64)
    ring_len = np.zeros(int(poly_off[-1]), dtype=np.int64)
    ring_len[poly_off[:-1][keep_ev]] = lens[keep_ev]
    piece_slots = _ranges(poly_off[:-1][~keep_ev], npieces[~keep_ev])
    ring_len[piece_slots] = plen
    ring_off = np.r_[0, np.cumsum(ring_len)].astype(np.int64)
    allxy = np.zeros((int(ring_off[-1]), 2), dtype=np.int64)
    src = _ranges(np.asarray(off[:-1])[keep_ev], lens[keep_ev])
    dst = _ranges(ring_off[:-1][poly_off[:-1][keep_ev]], lens[keep_ev])
    folded = np.asarray(xy)[src].astype(np.int64)
    folded[:, 0] %= nlon
    allxy[dst] = folded
    allxy[_ranges(ring_off[:-1][piece_slots], plen)] = pxy
    return allxy, ring_off, poly_off
