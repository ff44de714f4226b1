"""Reference-held parity fixtures: run the REAL ``wavebreaking`` package on the frozen synthetic fields.

This image has no xarray / geopandas / shapely / scikit-image, so the reference cannot be imported here and the
oracle's restatement of the skimage / GEOS semantics stays "parity unpinned" (DESIGN.md 2).  This script closes that
gap wherever the reference IS installable (``pip install wavebreaking==0.3.8`` or ``pip install /path/to/reference``):

    python tools/make_reference_fixtures.py [tests/golden/reference_v038.npz]

It needs nothing from this repository except ``wavebreaking_b200/synthetic.py`` (pure numpy, loaded by path so
that neither torch nor a GPU is required).  For each frozen configuration it stores what the reference returns for
``calculate_smoothed_field -> calculate_contours(original_coordinates=False) -> calculate_streamers /
overturnings / cutoffs(contours=...) -> to_xarray -> track_events``: the smoothed field, every contour (date,
level, closed, exp_lon, mean_lat, vertices), every event (date, level, com, mean_var, intensity, event_area,
orientation, polygon parts) in the reference's row order, the three int8 flag grids and the track labels.
``tests/test_reference_fixtures.py`` compares the CUDA path (and the oracle) with the file when it exists and
reports "fixture missing" otherwise.  The generated .npz is small (a few hundred KB) and is meant to be committed.
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# frozen configurations: name -> (nlat, nlon, number of steps, hours between steps, latitude descending)
CONFIGS = {
    "demo_like": (179, 360, 3, 6.0, False),     # shape of the reference's tests/data/demo_data.nc
    "one_degree": (181, 360, 6, 6.0, False),    # BASELINE.json configs[1]
    "era5_quarter": (721, 1440, 2, 1.0, True),  # configs[2], stored with descending latitude like ERA5 files
}


def load_synthetic():
    spec = importlib.util.spec_from_file_location("wbk_synthetic", os.path.join(ROOT, "wavebreaking_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def field_for(syn, name):
    nlat, nlon, nt, step, desc = CONFIGS[name]
    if name == "demo_like":
        raw = syn.pv_field(181, 360, np.arange(nt) * step, dtype=np.float64)[:, 1:-1, :]
        raw = np.roll(raw, 180, axis=2).astype(np.float32)
        lat, lon = np.arange(-89.0, 90.0), np.arange(-180.0, 180.0)
    else:
        raw = syn.pv_field(nlat, nlon, np.arange(nt) * step)
        lat, lon = syn.grid_coords(nlat, nlon)
    time = np.datetime64("2000-01-01T00", "ns") + (np.arange(nt) * step * 3600e9).astype("timedelta64[ns]")
    if desc:
        raw, lat = raw[:, ::-1, :].copy(), lat[::-1].copy()
    return raw, lat, lon, time


def ragged(list_of_arrays):
    off = np.zeros(len(list_of_arrays) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(a) for a in list_of_arrays])
    flat = np.concatenate(list_of_arrays) if list_of_arrays else np.zeros((0, 2))
    return off, flat


def polygon_parts(geom):
    if geom is None or geom.is_empty:
        return []
    parts = list(geom.geoms) if hasattr(geom, "geoms") else [geom]
    return [np.asarray(p.exterior.coords)[:-1] for p in parts]


def main(out_path):
    import xarray as xr

    import wavebreaking as wb

    syn = load_synthetic()
    out = {"reference_version": np.array(getattr(wb, "__version__", "unknown"))}
    for name in CONFIGS:
        raw, lat, lon, time = field_for(syn, name)
        pv = xr.DataArray(raw, dims=("time", "lat", "lon"), coords={"time": time, "lat": lat, "lon": lon}, name="PV")
        sm = wb.calculate_smoothed_field(data=pv, passes=5)
        contours = wb.calculate_contours(data=sm, contour_levels=2, periodic_add=120, original_coordinates=False)
        out[name + "/smoothed"] = np.asarray(sm.transpose("time", "lat", "lon").values)
        out[name + "/lat"], out[name + "/lon"] = lat, lon
        out[name + "/contour_date"] = contours.date.values.astype("datetime64[ns]")
        for col in ("level", "closed", "exp_lon", "mean_lat"):
            out[name + "/contour_" + col] = contours[col].values
        off, flat = ragged([np.asarray(g.coords) for g in contours.geometry])
        out[name + "/contour_off"], out[name + "/contour_xy"] = off, flat
        calls = {"streamers": wb.calculate_streamers, "overturnings": wb.calculate_overturnings, "cutoffs": wb.calculate_cutoffs}
        for kind, fn in calls.items():
            ev = fn(data=sm, contour_levels=2, contours=contours)
            key = "{}/{}_".format(name, kind)
            n = len(ev)
            out[key + "n"] = np.array(n)
            if n == 0:
                out[key + "flags"] = np.zeros(raw.shape, dtype=np.int8)
                continue
            out[key + "date"] = ev.date.values.astype("datetime64[ns]")
            out[key + "level"] = ev.level.values
            out[key + "com"] = np.asarray(list(ev.com), dtype=np.float64)
            for col in ("mean_var", "intensity", "event_area"):
                out[key + col] = ev[col].values.astype(np.float64)
            if "orientation" in ev:
                out[key + "orientation"] = np.array([str(v) for v in ev.orientation])
            parts = [polygon_parts(g) for g in ev.geometry]
            out[key + "nparts"] = np.array([len(p) for p in parts])
            off, flat = ragged([r for p in parts for r in p])
            out[key + "part_off"], out[key + "part_xy"] = off, flat
            flags = wb.to_xarray(data=sm, events=ev)
            out[key + "flags"] = np.asarray(flags.transpose("time", "lat", "lon").values).astype(np.int8)
            if kind == "streamers" and len(time) > 1:
                tracked = wb.track_events(events=ev, time_range=CONFIGS[name][3], method="by_overlap")
                out[key + "label"] = tracked.label.sort_index().values
                tracked = wb.track_events(events=ev, time_range=CONFIGS[name][3], method="by_distance", distance=1000)
                out[key + "label_by_distance"] = tracked.label.sort_index().values
        print(name, "contours", len(contours), {k: int(out["{}/{}_n".format(name, k)]) for k in calls})
    np.savez_compressed(out_path, **out)
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "reference_v038.npz"))
