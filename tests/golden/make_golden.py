"""Generate the golden fixtures under tests/golden/ (run once, outputs committed).

  python tests/golden/make_golden.py

* third_party.npz -- live outputs of the third-party routines the reference calls and that ARE installed
  here: scipy.ndimage.convolve (spatial.py:103) and sklearn DistanceMetric("haversine").pairwise
  (streamer_index.py:130), on small seeded inputs.
* pipeline_91x180.npz -- outputs of the oracle's whole path (oracle.pipeline.detect_steps) on a small
  seeded synthetic field: contour table, event tables, flag grids.  The reference itself cannot be imported
  in this image (xarray / geopandas / shapely / scikit-image are not installed), so these vectors pin the
  oracle + kernels against regressions, not against the reference.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from scipy import ndimage  # noqa: E402
from sklearn.metrics import DistanceMetric  # noqa: E402

from oracle import pipeline as P  # noqa: E402
from wavebreaking_b200 import synthetic  # noqa: E402


def third_party():
    rng = np.random.default_rng(123)
    w = np.array([[0, 1, 0], [1, 2, 1], [0, 1, 0]])
    f32 = rng.standard_normal((9, 14)).astype(np.float32)
    f64 = rng.standard_normal((9, 14))
    conv32 = ndimage.convolve(f32, weights=w, mode="wrap")
    conv64 = ndimage.convolve(f64, weights=w, mode="wrap")
    latlon = np.c_[rng.uniform(-80, 80, 40), rng.uniform(-180, 180, 40)]
    hav = DistanceMetric.get_metric("haversine").pairwise(np.radians(latlon)) * 6371
    np.savez_compressed(os.path.join(HERE, "third_party.npz"), f32=f32, f64=f64, conv32=conv32, conv64=conv64,
                        div32=conv32 / np.sum(w), latlon=latlon, hav=hav)


def pipeline_case():
    nlat, nlon, ntime = 91, 180, 2
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = synthetic.pv_field(nlat, nlon, np.arange(ntime) * 6.0)
    grid = P.Grid(lon, lat, synthetic.time_axis(ntime, 6))
    out = P.detect_steps(raw, grid, levels=[2.0, -2.0])
    c = out["contours"]
    save = dict(raw=raw, smoothed=out["smoothed"],
                c_level=c.level.values.astype(float), c_closed=c.closed.values, c_exp_lon=c.exp_lon.values.astype(float),
                c_mean_lat=c.mean_lat.values.astype(float), c_npts=np.array([len(g) for g in c.geometry]),
                c_pts=np.concatenate([np.asarray(g) for g in c.geometry]))
    for kind, ev in out["events"].items():
        save[kind + "_flags"] = out["flags"][kind]
        save[kind + "_level"] = ev.level.values.astype(float)
        save[kind + "_com"] = np.array([list(v) for v in ev.com])
        save[kind + "_mean_var"] = ev.mean_var.values
        save[kind + "_event_area"] = ev.event_area.values
        save[kind + "_ring_n"] = np.array([len(r) for r in ev.attrs["_index_rings"]])
        save[kind + "_rings"] = np.concatenate([np.asarray(r) for r in ev.attrs["_index_rings"]])
    np.savez_compressed(os.path.join(HERE, "pipeline_91x180.npz"), **save)
    print({k: len(v) for k, v in out["events"].items()}, "contours", len(c))


if __name__ == "__main__":
    third_party()
    pipeline_case()
