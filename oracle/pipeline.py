"""CPU restatement of the reference detection pipeline (test infrastructure only).

Every function cites the reference lines it follows (paths relative to the
reference checkout).  Data model: fields are numpy arrays ``[ntime, nlat, nlon]``
on a :class:`Grid` with ascending latitude / longitude (the reference re-sorts the
data that way first, ``utils/data_utils.py:196-213``); tables are pandas DataFrames
with the reference's column names; geometries are numpy arrays (a LineString is an
(n, 2) array, a (Multi)Polygon a list of open (n, 2) rings, the empty polygon ``[]``).

The loop structure, the N x N pair formulation and the third-party calls that are
available here (``scipy.ndimage.convolve``, sklearn haversine ``pairwise``, pandas
groupby) are kept as in the reference so that timing this module is a fair stand-in
for timing the reference on the host CPU.
"""

import itertools
from dataclasses import dataclass

import numpy as np
import pandas as pd
from scipy import ndimage
from sklearn.metrics import DistanceMetric

from . import geom
from .skimage_contours import find_contours

dist = DistanceMetric.get_metric("haversine")

DEFAULT_WEIGHTS = np.array([[0, 1, 0], [1, 2, 1], [0, 1, 0]])


@dataclass
class Grid:
    """Coordinates of a regular lat/lon grid (ascending) and its time axis."""

    lon: np.ndarray
    lat: np.ndarray
    time: np.ndarray

    @property
    def nlon(self):
        return len(self.lon)

    @property
    def nlat(self):
        return len(self.lat)

    @property
    def ntime(self):
        return len(self.time)

    @property
    def dlon(self):
        return spatial_resolution(self.lon, "lon")

    @property
    def dlat(self):
        return spatial_resolution(self.lat, "lat")


def spatial_resolution(coord, dim):
    """``get_spatial_resolution`` (utils/data_utils.py:129-144)."""
    coord = np.asarray(coord)
    delta = abs(np.unique(coord[1:] - coord[:-1]))
    if len(delta) > 1:
        raise ValueError("No regular grid found for dimension {}.".format(dim))
    elif delta[0] == 0:
        raise ValueError("Two equivalent coordinates found for dimension {}.".format(dim))
    return delta[0]


# --------------------------------------------------------------------------- spatial.py
def smooth_field(data, passes, weights=DEFAULT_WEIGHTS, mode="wrap"):
    """``calculate_smoothed_field`` body (processing/spatial.py:99-109)."""
    weights = np.asarray(weights)
    smoothed = []
    for step in range(data.shape[0]):
        temp = data[step]
        for _ in range(passes):
            temp = ndimage.convolve(temp, weights=weights, mode=mode) / np.sum(weights)
        border_size = int(weights.shape[0] / 2 + 0.5)
        temp = np.array(temp, copy=True)
        if not np.issubdtype(temp.dtype, np.floating):
            raise ValueError("cannot convert float NaN to integer")
        temp[np.arange(-border_size, border_size), :] = np.nan
        smoothed.append(temp)
    return np.asarray(smoothed)


def momentum_flux(u, v):
    """``calculate_momentum_flux`` body (processing/spatial.py:50-54); lon is the last axis."""
    with np.errstate(invalid="ignore"):
        u_prime = u - np.nanmean(u, axis=-1, keepdims=True)
        v_prime = v - np.nanmean(v, axis=-1, keepdims=True)
    return u_prime * v_prime


# --------------------------------------------------------------------------- contour_index.py
def contours_one(field2d, level, grid, periodic_add=120, original_coordinates=True):
    """One (step, level) of ``calculate_contours`` (indices/contour_index.py:93-194).

    Returns a list of dicts ``closed, exp_lon, mean_lat, geometry`` (geometry = (n, 2)
    array of x, y).
    """
    nlon, dlon = grid.nlon, grid.dlon
    add = int(periodic_add / dlon)
    ext = np.concatenate([field2d, field2d[:, :add]], axis=1)  # :94-100

    contours_from_measure = find_contours(ext, level)  # :103

    contours_index_expanded, closed = [], []
    for item in contours_from_measure:  # :106-117
        check_closed = all(item[0] == item[-1])
        indices = np.asarray(list(dict.fromkeys(map(tuple, np.round(item).astype("int")))))[:, ::-1]
        if len(indices) >= 4:
            contours_index_expanded.append(indices)
            closed.append(check_closed)

    def check_duplicates(list_of_arrays):  # :119-140
        temp = [np.c_[item[:, 0] % nlon, item[:, 1]] for item in list_of_arrays]
        sets = [set(map(tuple, e)) for e in temp]
        check = [
            (i1, i2)
            for i1, i2 in itertools.permutations(range(len(temp)), r=2)
            if sets[i1].issubset(sets[i2])
        ]
        drop = []
        lens = np.array([len(item) for item in temp])
        for indices in check:
            if lens[indices[0]] == lens[indices[1]]:
                drop.append(max(indices))
            else:
                drop.append(indices[np.argmin(lens[[indices[0], indices[1]]])])
        return list(set(drop))

    drop = check_duplicates(contours_index_expanded)  # :143-147
    contours_index_expanded = [it for i, it in enumerate(contours_index_expanded) if i not in drop]
    closed = [it for i, it in enumerate(closed) if i not in drop]

    def rows(list_of_arrays):  # :149-169
        return [
            {
                "closed": bool(c),
                "exp_lon": len(set(item[:, 0])) * dlon,
                "mean_lat": np.round(item[:, 1].mean(), 2),
                "geometry": item,
            }
            for item, c in zip(list_of_arrays, closed)
        ]

    if original_coordinates is False:
        return rows(contours_index_expanded)

    original = [  # :179-191
        np.c_[grid.lon[item[:, 0] % nlon], grid.lat[item[:, 1]]] for item in contours_index_expanded
    ]
    original = [np.asarray(list(dict.fromkeys(map(tuple, item)))) for item in original]
    return rows(original)


CONTOUR_COLUMNS = ["date", "level", "closed", "exp_lon", "mean_lat", "geometry"]


def _as_levels(contour_levels):
    try:
        iter(contour_levels)
    except Exception:
        contour_levels = [contour_levels]
    return list(contour_levels)


def calculate_contours(data, contour_levels, grid, periodic_add=120, original_coordinates=True):
    """``calculate_contours`` with its time / level loops (utils/index_utils.py:217-258)."""
    out = []
    for t in range(grid.ntime):
        for level in _as_levels(contour_levels):
            for row in contours_one(data[t], level, grid, periodic_add, original_coordinates):
                out.append({"date": grid.time[t], "level": level, **row})
    return pd.DataFrame(out, columns=CONTOUR_COLUMNS)


# --------------------------------------------------------------------------- index_utils.py
def combine_shared(lst):
    """``combine_shared`` (utils/index_utils.py:187-214)."""
    elements = list(lst)
    output = []
    while len(elements) > 0:
        first, *rest = elements
        first = set(first)
        lf = -1
        while len(first) > lf:
            lf = len(first)
            rest2 = []
            for r in rest:
                if len(first.intersection(set(r))) > 0:
                    first |= set(r)
                else:
                    rest2.append(r)
            rest = rest2
        output.append(list(first))
        elements = rest
    return output


def cell_areas(grid):
    """Per-latitude cell area in km^2 (utils/index_utils.py:58-63)."""
    area_cell = np.round(6371 * 2 * np.pi / (360 / ((grid.dlon + grid.dlat) / 2))) ** 2
    return np.cos(np.radians(grid.lat)) * area_cell


def calculate_properties(events, data, intensity, periodic_add, grid):
    """``calculate_properties`` (utils/index_utils.py:35-126).

    ``events`` has columns id, date, level, geometry (index-space rings) and optionally
    orientation.  Returns the property DataFrame plus, under ``_members``, the list of
    (x, y) member arrays of every event (extended-grid indices) for the parity tests.
    """
    nlon, nlat = grid.nlon, grid.nlat
    width = nlon + int(periodic_add / grid.dlon)
    x, y = np.meshgrid(np.arange(0, width), np.arange(0, nlat))
    x, y = x.flatten(), y.flatten()
    r = ((grid.dlon + grid.dlat) / 2) / 2  # buffer in index units (:47-50)

    weight_lat = cell_areas(grid)
    time_index = {t: i for i, t in enumerate(grid.time.tolist())}

    recs = []
    members = []
    for ev_id, date, rings in zip(events["id"], events["date"], events["geometry"]):
        sel = geom.buffered_contains(rings, r, x, y)
        mx, my = x[sel], y[sel]
        members.append(np.c_[mx, my])
        t = time_index[pd.Timestamp(date).to_datetime64().astype(grid.time.dtype).item()
                       if np.issubdtype(grid.time.dtype, np.datetime64) else date]
        areas = weight_lat[my]
        vals = data[t, my, mx % nlon]
        recs.append(
            pd.DataFrame(
                {
                    "id": ev_id,
                    "areas": areas,
                    "mean_var": areas * vals,
                    "intensity": areas * intensity[t, my, mx % nlon] if intensity is not None else 0,
                    "x_com": mx * areas,
                    "y_com": my * areas,
                }
            )
        )
    merged = pd.concat(recs) if recs else pd.DataFrame(
        columns=["id", "areas", "mean_var", "intensity", "x_com", "y_com"])
    agg = merged.groupby("id").agg(
        {"areas": "sum", "mean_var": "sum", "intensity": "sum", "x_com": "sum", "y_com": "sum"}
    )
    agg = agg.reindex(events["id"])  # events without any member would be missing in the reference

    com_x = grid.lon[(agg.x_com / agg.areas).astype("int") % nlon]
    com_y = grid.lat[(agg.y_com / agg.areas).astype("int")]
    com = list(map(tuple, np.c_[com_x, com_y]))

    prop = {
        "date": events["date"].values,
        "level": events["level"].values,
        "com": com,
        "mean_var": (agg.mean_var / agg.areas).round(2).values,
        "intensity": (agg.intensity / agg.areas).round(2).values,
        "event_area": agg.areas.round(2).values,
    }
    if "orientation" in events.columns:
        prop["orientation"] = events["orientation"].values
    df = pd.DataFrame(prop)
    df.attrs["_members"] = members
    df.attrs["_sums"] = agg.reset_index(drop=True)
    return df


def transform_polygons(events, grid):
    """``transform_polygons`` (utils/index_utils.py:129-184).

    Returns ``(geometry, index_pieces)``: per event the list of rings in lon/lat
    coordinates and the same rings as folded integer index coordinates.
    """
    nlon = grid.nlon
    geoms, pieces_all = [], []
    for rings in events["geometry"]:
        ring = np.asarray(rings[0]).astype("int")
        split = (ring[:, 0] >= nlon).any()
        if not split:
            pieces = [np.c_[ring[:, 0] % nlon, ring[:, 1]]]
        else:
            pieces = geom.split_ring_at_meridian(ring, nlon)
        pieces_all.append(pieces)
        geoms.append([np.c_[grid.lon[p[:, 0]], grid.lat[p[:, 1]]] for p in pieces])
    return geoms, pieces_all


def _finish_events(gdf, data, intensity, periodic_add, grid):
    props = calculate_properties(gdf, data, intensity, periodic_add, grid)
    geoms, pieces = transform_polygons(gdf, grid)
    out = props.copy()
    out["geometry"] = geoms
    out.attrs = dict(props.attrs)
    out.attrs["_index_rings"] = [g[0] for g in gdf["geometry"]]
    out.attrs["_index_pieces"] = pieces
    return out


def _empty_events():
    return pd.DataFrame()


# --------------------------------------------------------------------------- streamer_index.py
def streamer_basepoints(points, grid, geo_dis=800, cont_dis=1500, diagnostics=None):
    """Per-contour part of ``calculate_streamers`` (indices/streamer_index.py:119-264).

    ``points`` is the (n, 2) integer (x, y) array of one contour.  Returns the final
    base-point table ``ind1, x1, y1, ind2, x2, y2`` (int arrays, reference row order).
    """
    nlon = grid.nlon
    contour_index = pd.DataFrame(np.asarray(points), columns=["x", "y"]).astype("int")
    contour_coords = np.c_[
        grid.lat[contour_index.y.values], grid.lon[contour_index.x.values % nlon]
    ]
    geo_matrix = dist.pairwise(np.radians(contour_coords)) * 6371  # :130
    on = np.insert(np.diagonal(geo_matrix, 1), 0, 0)  # :133
    on_mat = np.triu(np.tile(on, (len(on), 1)), k=1)
    cont_matrix = np.cumsum(on_mat, axis=1)
    check = (geo_matrix < geo_dis) * (cont_matrix > cont_dis)  # :138
    check = np.transpose(check.nonzero())
    if diagnostics is not None:
        diagnostics["on"] = on
        diagnostics["geo_margin"] = np.abs(geo_matrix - geo_dis)
        diagnostics["cont_margin"] = np.abs(cont_matrix - cont_dis)
        diagnostics["raw_pairs"] = check

    df1 = (
        contour_index[["x", "y"]].iloc[check[:, 0]].reset_index()
        .rename(columns={"x": "x1", "y": "y1", "index": "ind1"})
    )
    df2 = (
        contour_index[["x", "y"]].iloc[check[:, 1]].reset_index()
        .rename(columns={"x": "x2", "y": "y2", "index": "ind2"})
    )
    df_bp = pd.concat([df1, df2], axis=1)
    df_bp = df_bp.drop(df_bp[np.abs(df_bp.x1 - df_bp.x2) > 120].index)  # :157
    pts = contour_index[["x", "y"]].values

    def check_duplicates(df):  # :160-183
        temp = pd.concat([df.x1 % nlon, df.y1, df.x2 % nlon, df.y2], axis=1)
        temp = temp[temp.duplicated(keep=False)]
        if len(temp) == 0:
            check = []
        else:
            groups = temp.groupby(list(temp)).indices  # key -> positional indices
            index = temp.index.values
            check = [index[v][1] for v in groups.values()]
        return df.drop(check)

    def check_intersections(df):  # :185-200
        keep = [
            geom.chord_touches_polyline((r.x1, r.y1), (r.x2, r.y2), pts)
            for r in df.itertuples()
        ]
        return df[keep]

    def check_overlapping(df):  # :202-222
        # the reference walks itertools.permutations(df.index, 2) in pure Python (O(P^2), ~5e8 iterations for
        # the 2e4 candidates of a 0.25-degree contour); the same predicate is evaluated here with chunked numpy
        # broadcasting so that the oracle stays usable as a CPU baseline at that size (identical result)
        i1 = df.ind1.values.astype(np.int64)
        i2 = df.ind2.values.astype(np.int64)
        idx = np.arange(len(df))
        drop = np.zeros(len(df), dtype=bool)
        for s0 in range(0, len(df), 512):
            a1, a2, ai = i1[s0:s0 + 512, None], i2[s0:s0 + 512, None], idx[s0:s0 + 512, None]
            inside = (i1[None, :] <= a1) & (a1 <= i2[None, :]) & (i1[None, :] <= a2) & (a2 <= i2[None, :])
            inside &= ai != idx[None, :]
            drop[s0:s0 + 512] = inside.any(axis=1)
        return df.drop(df.index[drop])

    def check_groups(df):  # :224-251
        index_combinations = np.asarray(
            list(itertools.combinations_with_replacement(df.index, r=2))
        )
        rows = df[["x1", "y1", "x2", "y2"]].values
        check_crossing = [
            geom.segments_intersect(rows[a][:2], rows[a][2:], rows[b][:2], rows[b][2:])
            for a, b in index_combinations
        ]
        groups = combine_shared(index_combinations[check_crossing])
        keep_index = [
            item[
                np.argmax(
                    [on[int(r.ind1): int(r.ind2) + 1].sum() for r in df.iloc[item].itertuples()]
                )
            ]
            for item in groups
        ]
        return df.iloc[keep_index]

    routines = [check_duplicates, check_intersections, check_overlapping, check_groups]
    i = 0
    while len(df_bp.index) > 1 and i <= 3:  # :261-264
        df_bp = routines[i](df_bp).reset_index(drop=True)
        i += 1
    return df_bp[["ind1", "x1", "y1", "ind2", "x2", "y2"]].astype("int64").reset_index(drop=True)


def _select_full_contours(contours):
    return contours[contours.exp_lon == contours.exp_lon.max()].reset_index(drop=True)


def calculate_streamers(data, grid, contours, geo_dis=800, cont_dis=1500, intensity=None,
                        periodic_add=120, diagnostics=None):
    """``calculate_streamers`` (indices/streamer_index.py:100-289); ``contours`` in index coordinates."""
    contours = _select_full_contours(contours)  # :106
    rows = []
    for k, series in enumerate(contours.itertuples()):
        pts = np.asarray(series.geometry).astype("int")
        diag = {} if diagnostics is not None else None
        df_bp = streamer_basepoints(pts, grid, geo_dis, cont_dis, diag)
        if diagnostics is not None:
            diagnostics.setdefault("contours", []).append(diag)
        for r in df_bp.itertuples():  # :269-272
            rows.append({"date": series.date, "level": series.level,
                         "geometry": [pts[int(r.ind1): int(r.ind2) + 1]],
                         "_contour": k, "_ind1": int(r.ind1), "_ind2": int(r.ind2)})
    if len(rows) == 0:
        return _empty_events()
    gdf = pd.DataFrame(rows).reset_index().rename(columns={"index": "id"})
    out = _finish_events(gdf, data, intensity, periodic_add, grid)
    out.attrs["_basepoints"] = gdf[["_contour", "_ind1", "_ind2"]]
    return out


# --------------------------------------------------------------------------- overturning_index.py
def overturning_boxes(points, grid, range_group=5, min_exp=5):
    """Per-contour part of ``calculate_overturnings`` (indices/overturning_index.py:119-213)."""
    nlon, dlon = grid.nlon, grid.dlon
    contour_index = pd.DataFrame(np.asarray(points), columns=["x", "y"]).astype("int")
    lons, counts = np.unique(contour_index.x, return_counts=True)
    ot_lons = pd.DataFrame({"lon": lons[counts >= 3]})
    ot_lons["label"] = (ot_lons.diff() > range_group / dlon).cumsum()
    groups = ot_lons.groupby("label")
    df_ot = groups.agg(["min", "max"]).astype("int").reset_index(drop=True)
    df_ot.columns = ["min_lon", "max_lon"]

    def check_duplicates(df):  # :136-162
        temp = [np.array(range(r.min_lon, r.max_lon + 1)) % nlon for r in df.itertuples()]
        sets = [set(t.tolist()) for t in temp]
        check = [it for it in itertools.permutations(df.index, r=2) if sets[it[0]].issubset(sets[it[1]])]
        drop = []
        for item in check:
            lens = [len(temp[i]) for i in item]
            if lens[0] == lens[1]:
                drop.append(max(item))
            else:
                drop.append(item[np.argmin(lens)])
        return df[~df.reset_index(drop=True).index.isin(drop)]

    def check_expansion(df):  # :164-166
        exp_lon = df.max_lon - df.min_lon
        return df[exp_lon >= min_exp / dlon]

    def find_lat_expansion(df):  # :168-180
        lats = [
            contour_index[contour_index.x.isin(range(r.min_lon, r.max_lon + 1))].y
            for r in df.itertuples()
        ]
        ot_lats = pd.DataFrame(
            [(item.min(), item.max()) for item in lats], columns=["min_lat", "max_lat"]
        ).astype("int")
        return pd.concat([df, ot_lats], axis=1)

    routines = [check_duplicates, check_expansion, find_lat_expansion]
    i = 0
    while len(df_ot.index) > 0 and i <= 2:  # :185-188
        df_ot = routines[i](df_ot).reset_index(drop=True)
        i += 1

    def check_orientation(r):  # :191-202
        lat_west = grid.lat[contour_index[contour_index.x.eq(r.min_lon)].y.values[0]]
        lat_east = grid.lat[contour_index[contour_index.x.eq(r.max_lon)].y.values[-1]]
        return "cyclonic" if abs(lat_west) <= abs(lat_east) else "anticyclonic"

    df_ot = df_ot.reset_index(drop=True)
    df_ot["orientation"] = [check_orientation(r) for r in df_ot.itertuples()]
    return df_ot


def calculate_overturnings(data, grid, contours, range_group=5, min_exp=5, intensity=None,
                           periodic_add=120):
    """``calculate_overturnings`` (indices/overturning_index.py:99-232)."""
    contours = _select_full_contours(contours)  # :105
    rows = []
    for series in contours.itertuples():
        pts = np.asarray(series.geometry).astype("int")
        df_ot = overturning_boxes(pts, grid, range_group, min_exp)
        for r in df_ot.itertuples():  # box(minx, miny, maxx, maxy): ccw from (maxx, miny)
            ring = np.array([[r.max_lon, r.min_lat], [r.max_lon, r.max_lat],
                             [r.min_lon, r.max_lat], [r.min_lon, r.min_lat]])
            rows.append({"date": series.date, "level": series.level,
                         "orientation": r.orientation, "geometry": [ring]})
    if len(rows) == 0:
        return _empty_events()
    gdf = pd.DataFrame(rows).reset_index().rename(columns={"index": "id"})
    return _finish_events(gdf, data, intensity, periodic_add, grid)


# --------------------------------------------------------------------------- cutoff_index.py
def calculate_cutoffs(data, grid, contours, min_exp=5, intensity=None, periodic_add=120):
    """``calculate_cutoffs`` (indices/cutoff_index.py:83-104)."""
    sel = contours[
        (contours.exp_lon < contours.exp_lon.max()) & (contours.exp_lon >= min_exp) & contours.closed
    ].reset_index(drop=True)
    if len(sel) == 0:
        return _empty_events()
    gdf = pd.DataFrame({
        "date": sel.date.values,
        "level": sel.level.values,
        "geometry": [[np.asarray(g).astype("int")] for g in sel.geometry],
    }).reset_index().rename(columns={"index": "id"})
    return _finish_events(gdf, data, intensity, periodic_add, grid)


# --------------------------------------------------------------------------- events.py
def to_xarray(data, events, grid, flag="ones"):
    """``to_xarray`` (processing/events.py:66-108); returns the flag array [ntime, nlat, nlon]."""
    if len(events) == 0:
        raise ValueError("geopandas.GeoDataFrame is empty!")
    lon, lat = np.meshgrid(grid.lon, grid.lat)
    lonf, latf = lon.flatten(), lat.flatten()
    r = ((grid.dlon + grid.dlat) / 2) / 2  # degrees (:75-78)
    if flag != "ones" and flag not in events.columns:
        raise KeyError("{} is not a column of the events geopandas.GeoDataFrame.".format(flag))
    flagged = np.zeros_like(data)
    time_index = {t: i for i, t in enumerate(grid.time.tolist())}
    for k, (date, rings) in enumerate(zip(events["date"], events["geometry"])):
        sel = geom.buffered_contains(rings, r, lonf, latf)
        key = (pd.Timestamp(date).to_datetime64().astype(grid.time.dtype).item()
               if np.issubdtype(grid.time.dtype, np.datetime64) else date)
        t = time_index[key]
        val = 1 if flag == "ones" else events[flag].iloc[k]
        flagged[t].reshape(-1)[sel] = val
    if flag == "ones":
        flagged = flagged.astype("int8")
    return flagged


def track_events(events, time_range=None, method="by_overlap", buffer=0, overlap=0, distance=1000, _memo=None):
    """``track_events`` (processing/events.py:151-241); ``buffer`` must be 0 (no GEOS here).
    ``_memo`` (tests): dict that caches the per-pair areas between calls on the same table."""
    events = events.reset_index(drop=True)
    if len(events) == 0:
        raise ValueError("geopandas.GeoDataFrame is empty!")
    if time_range is None:
        date_dif = events.date.diff()
        time_range = date_dif[date_dif > pd.Timedelta(0)].min().total_seconds() / 3600

    if np.issubdtype(events.date.dtype, np.datetime64):
        hours = (events.date - events.date.iloc[0]).dt.total_seconds().values / 3600
        diffs = hours[None, :] - hours[:, None]
    else:
        vals = events.date.values
        diffs = abs(vals[None, :] - vals[:, None])
    ii, jj = np.nonzero((diffs > 0) & (diffs <= time_range))
    range_comb = np.c_[ii, jj]
    if len(range_comb) == 0:
        raise ValueError("No events detected in the time range: {}".format(time_range))

    if method == "by_distance":
        com1 = np.asarray(list(events.iloc[range_comb[:, 0]].com))
        com2 = np.asarray(list(events.iloc[range_comb[:, 1]].com))
        dist_com = np.asarray([dist.pairwise(np.radians([p1, p2]))[0, 1] for p1, p2 in zip(com1, com2)])
        combine = range_comb[dist_com * 6371 < distance]
    elif method == "by_overlap":
        if buffer != 0:
            raise NotImplementedError("oracle restates by_overlap for buffer=0 only")
        check = []
        geoms = [[np.asarray(r).reshape(-1, 2) for r in g] for g in events.geometry]
        boxes = [(min(r[:, 0].min() for r in g), min(r[:, 1].min() for r in g), max(r[:, 0].max() for r in g),
                  max(r[:, 1].max() for r in g)) if len(g) else None for g in geoms]
        lattice = all(np.array_equal(r, np.rint(r)) for g in geoms for r in g)
        for a, b in range_comb:
            ba, bb = boxes[a], boxes[b]
            if ba is None or bb is None:
                check.append(False if overlap >= 0 else (ba is not None or bb is not None))
                continue
            if ba[2] < bb[0] or bb[2] < ba[0] or ba[3] < bb[1] or bb[3] < ba[1]:
                # disjoint boxes: the intersection is empty, 0 / union
                check.append(0.0 > overlap)
                continue
            key = (int(a), int(b))
            if _memo is not None and key in _memo:
                a1, a2, inter = _memo[key]
            else:
                if lattice:
                    # GEOS' overlay is robust: polygons that merely touch give area 0.0, overlapping ones a positive
                    # area.  A float64 sum is not, so the sign of the intersection area is decided exactly and the
                    # float64 area is only evaluated where it is positive.
                    a1 = sum(abs(geom._ring_area2([(int(x), int(y)) for x, y in r])) for r in geoms[a]) / 2.0
                    a2 = sum(abs(geom._ring_area2([(int(x), int(y)) for x, y in r])) for r in geoms[b]) / 2.0
                    a1, a2 = float(a1), float(a2)
                    if geom.overlap_positive_exact(geoms[a], geoms[b]):
                        inter = max(geom.overlap_areas(geoms[a], geoms[b])[2], np.finfo(np.float64).tiny)
                    else:
                        inter = 0.0
                else:
                    a1, a2, inter = geom.overlap_areas(geoms[a], geoms[b])
                if _memo is not None:
                    _memo[key] = (a1, a2, inter)
            with np.errstate(divide="ignore", invalid="ignore"):
                check.append(np.float64(inter) / np.float64(a2 + a1 - inter) > overlap)
        combine = range_comb[np.asarray(check, dtype=bool)]
    else:
        raise ValueError("'{}' not supported as method! Supported methods are 'by_overlap' and 'by_distance'".format(method))

    combine = combine_shared([list(map(int, c)) for c in combine])
    label = np.arange(len(events))
    for item in combine:
        label[item] = min(item)
    # dense rank of the labels (:233-238)
    _, dense = np.unique(label, return_inverse=True)
    events["label"] = dense
    return events.sort_values(by=["label", "date"], kind="stable")


# --------------------------------------------------------------------------- whole path (bench / tests)
def detect_steps(raw, grid, levels=(2,), passes=5, periodic_add=120, intensity=None):
    """The benchmarked path on the CPU, call for call as a user of the reference would write it:
    smoothing -> contours -> streamers / overturnings / cutoffs -> to_xarray of each."""
    sm = smooth_field(raw, passes) if passes > 0 else raw
    contours = calculate_contours(sm, list(levels), grid, periodic_add, original_coordinates=False)
    events = dict(
        streamers=calculate_streamers(sm, grid, contours, intensity=intensity, periodic_add=periodic_add),
        overturnings=calculate_overturnings(sm, grid, contours, intensity=intensity, periodic_add=periodic_add),
        cutoffs=calculate_cutoffs(sm, grid, contours, intensity=intensity, periodic_add=periodic_add),
    )
    flags = {}
    for kind, ev in events.items():
        flags[kind] = to_xarray(np.zeros_like(sm), ev, grid) if len(ev) else np.zeros(sm.shape, dtype="int8")
    return dict(smoothed=sm, contours=contours, events=events, flags=flags)
