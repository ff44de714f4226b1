"""Temporal tracking of events (host logic + the overlap-area kernel).

Reference: wavebreaking/processing/events.py:113-241 (``track_events``) and
wavebreaking/utils/index_utils.py:187-214 (``combine_shared``).
"""

import numpy as np
import pandas as pd
import torch

from . import _lib, compat


def combine_shared(lst):
    """Connected components of index lists that share elements.

    Components come out in order of first appearance (like the reference's sweep over ``lst``);
    the members of a component are returned in ascending order (the reference returns them in
    Python-set iteration order, which is ascending for the small non-negative ints it is used with).
    """
    parent = {}

    def find(a):
        root = a
        while parent[root] != root:
            root = parent[root]
        while parent[a] != root:
            parent[a], a = root, parent[a]
        return root

    order = []
    for item in lst:
        item = [int(v) if isinstance(v, (np.integer,)) else v for v in item]
        for v in item:
            if v not in parent:
                parent[v] = v
                order.append(v)
        for v in item[1:]:
            ra, rb = find(item[0]), find(v)
            if ra != rb:
                parent[rb] = ra
    groups, seen = {}, []
    for v in order:
        r = find(v)
        if r not in groups:
            groups[r] = []
            seen.append(r)
        groups[r].append(v)
    return [sorted(groups[r]) for r in seen]


def _haversine_pairs(p1, p2):
    """sklearn DistanceMetric('haversine') for paired rows of (lat, lon) radians."""
    s0 = np.sin(0.5 * (p1[:, 0] - p2[:, 0]))
    s1 = np.sin(0.5 * (p1[:, 1] - p2[:, 1]))
    return 2 * np.arcsin(np.sqrt(s0 * s0 + np.cos(p1[:, 0]) * np.cos(p2[:, 0]) * s1 * s1))


def overlap_areas(geoms, pairs):
    """(area A, area B, area A n B) for every pair of (Multi)Polygons, computed by wbk_track_overlap."""
    lib = _lib.get()
    rings_xy, ring_off, poly_off = [], [0], [0]
    for g in geoms:
        for ring in compat.geometry_rings(g):
            rings_xy.append(np.asarray(ring, dtype=np.float64))
            ring_off.append(ring_off[-1] + len(ring))
        poly_off.append(len(ring_off) - 1)
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    if len(pairs) == 0:
        return np.zeros((0, 3))
    xy = np.ascontiguousarray(np.concatenate(rings_xy) if rings_xy else np.zeros((0, 2)))
    dev = lib.device
    d_xy = torch.from_numpy(xy).to(dev)
    d_ro = torch.from_numpy(np.asarray(ring_off, dtype=np.int32)).to(dev)
    d_po = torch.from_numpy(np.asarray(poly_off, dtype=np.int32)).to(dev)
    d_pairs = torch.from_numpy(pairs).to(dev)
    out = torch.empty((len(pairs), 3), dtype=torch.float64, device=dev)
    lib.call("wbk_track_overlap", _lib.ptr(d_xy), _lib.ptr(d_ro), _lib.ptr(d_po), _lib.ptr(d_pairs), len(pairs),
             _lib.ptr(out), lib.stream())
    return out.cpu().numpy()


def range_combinations(dates, time_range):
    """Pairs (i, j) with 0 < date_j - date_i <= time_range (events.py:160-181), as an (n, 2) array."""
    dates = np.asarray(dates)
    if np.issubdtype(dates.dtype, np.datetime64):
        hours = (dates - dates.min()).astype("timedelta64[ns]").astype(np.int64) / 3.6e12
        order = np.argsort(hours, kind="stable")
        hs = hours[order]
        lo = np.searchsorted(hs, hs, side="right")
        hi = np.searchsorted(hs, hs + time_range, side="right")
        ii = np.repeat(np.arange(len(hs)), np.maximum(hi - lo, 0))
        jj = np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)]) if len(hs) else np.zeros(0, dtype=int)
        return np.c_[order[ii], order[jj.astype(int)]] if len(ii) else np.zeros((0, 2), dtype=int)
    vals = dates.astype(np.float64)
    diffs = np.abs(vals[None, :] - vals[:, None])
    ii, jj = np.nonzero((diffs > 0) & (diffs <= time_range))
    return np.c_[ii, jj]


def track_events(events, time_range=None, method="by_overlap", buffer=0, overlap=0, distance=1000):
    """events.py:151-241; ``buffer`` other than 0 is not supported (GEOS buffering is out of scope)."""
    events = events.reset_index(drop=True)
    if time_range is None:
        date_dif = events.date.diff()
        time_range = date_dif[date_dif > pd.Timedelta(0)].min().total_seconds() / 3600
    range_comb = range_combinations(events.date.values, time_range)
    if len(range_comb) == 0:
        raise ValueError("No events detected in the time range: {}".format(time_range))

    if method == "by_distance":
        com1 = np.asarray(list(events.iloc[range_comb[:, 0]].com), dtype=np.float64)
        com2 = np.asarray(list(events.iloc[range_comb[:, 1]].com), dtype=np.float64)
        # the reference feeds (lon, lat) where haversine expects (lat, lon) (events.py:189-195); kept
        dist_com = _haversine_pairs(np.radians(com1), np.radians(com2))
        combine = range_comb[dist_com * 6371 < distance]
    elif method == "by_overlap":
        if buffer != 0:
            raise NotImplementedError("track_events(by_overlap): only buffer=0 is supported")
        geoms = list(events.geometry)
        # cheap bounding-box prefilter on the host, exact areas on the device
        boxes = np.array([_bounds(g) for g in geoms])
        a, b = range_comb[:, 0], range_comb[:, 1]
        cand = ~((boxes[a, 2] < boxes[b, 0]) | (boxes[b, 2] < boxes[a, 0])
                 | (boxes[a, 3] < boxes[b, 1]) | (boxes[b, 3] < boxes[a, 1]))
        check = np.zeros(len(range_comb), dtype=bool)
        if cand.any():
            areas = overlap_areas(geoms, range_comb[cand])
            inter = np.where(areas[:, 2] > 1e-12 * (areas[:, 0] + areas[:, 1]), areas[:, 2], 0.0)
            with np.errstate(divide="ignore", invalid="ignore"):
                check[cand] = inter / (areas[:, 1] + areas[:, 0] - inter) > overlap
        combine = range_comb[check]
    else:
        raise ValueError("'{}' not supported as method! Supported methods are 'by_overlap' and 'by_distance'".format(method))

    groups = combine_shared([list(map(int, c)) for c in combine])
    label = np.arange(len(events))
    for item in groups:
        label[item] = min(item)
    _, dense = np.unique(label, return_inverse=True)  # smallest possible label numbers (events.py:233-238)
    events["label"] = dense
    return events.sort_values(by=["label", "date"], kind="stable")


def _bounds(g):
    rings = compat.geometry_rings(g)
    if not rings:
        return (np.inf, np.inf, -np.inf, -np.inf)
    xy = np.concatenate(rings)
    return (xy[:, 0].min(), xy[:, 1].min(), xy[:, 0].max(), xy[:, 1].max())
