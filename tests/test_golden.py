"""The oracle against golden vectors: skimage's own published test vectors, the reference's fixture-free
known-answer tests, live third-party outputs (scipy / sklearn) and the committed pipeline fixture; then the
CUDA path against the same fixture."""

import math
import os

import numpy as np
import pandas as pd
import pytest

from oracle import geom as G
from oracle import pipeline as P
from oracle.skimage_contours import find_contours
from wavebreaking_b200 import detect, pipeline, spatial, synthetic

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------ scikit-image vectors
def test_skimage_binary_vector():
    """skimage/measure/tests/test_find_contours.py::test_binary: ring, start point and direction."""
    a = np.ones((8, 8), dtype=np.float32)
    a[1:-1, 1] = 0
    a[1, 1:-1] = 0
    ref = [[6., 1.5], [5., 1.5], [4., 1.5], [3., 1.5], [2., 1.5], [1.5, 2.], [1.5, 3.], [1.5, 4.], [1.5, 5.], [1.5, 6.],
           [1., 6.5], [0.5, 6.], [0.5, 5.], [0.5, 4.], [0.5, 3.], [0.5, 2.], [0.5, 1.], [1., 0.5], [2., 0.5], [3., 0.5],
           [4., 0.5], [5., 0.5], [6., 0.5], [6.5, 1.], [6., 1.5]]
    contours = find_contours(a, 0.5)  # positive_orientation='low' == reversed 'high' result of the skimage test
    assert len(contours) == 1
    assert np.array_equal(contours[0], ref)


def test_skimage_float_vector_through_lattice_vertices():
    """::test_float: the level passes exactly through grid vertices (float-tuple joining)."""
    x, y = np.mgrid[-1:1:5j, -1:1:5j]
    r = np.sqrt(x ** 2 + y ** 2)
    contours = find_contours(r, 0.5)
    assert len(contours) == 1
    assert np.array_equal(contours[0], [[2., 3.], [1., 2.], [2., 1.], [3., 2.], [2., 3.]])


def test_skimage_docstring_vector():
    a = np.zeros((3, 3))
    a[0, 0] = 1
    c = find_contours(a, 0.5)
    assert len(c) == 1 and np.array_equal(c[0], [[0., 0.5], [0.5, 0.]])


# ------------------------------------------------------------------ reference KATs (tests/test_wavebreaking.py)
def test_reference_kat_combine_shared():
    assert P.combine_shared([[1, 2, 3], [2, 3, 4], [5, 6]]) == [[1, 2, 3, 4], [5, 6]]  # :101-104


def _demo_grid(ntime=3):
    lat = np.arange(-89.0, 90.0)
    lon = np.arange(-180.0, 180.0)
    time = np.datetime64("1959-06-03T12", "ns") + np.arange(ntime) * np.timedelta64(6, "h")
    return P.Grid(lon, lat, time)


def test_reference_kat_to_xarray():
    grid = _demo_grid()  # :116-128
    sq = np.array([[0, 0], [10, 0], [10, 10], [0, 10]], dtype=float)
    ev = pd.DataFrame({"date": [grid.time[0]], "geometry": [[sq]]})
    flags = P.to_xarray(np.zeros((3, 179, 360), dtype=np.float32), ev, grid)
    assert flags.dtype == np.int8
    assert flags[0, list(grid.lat).index(5.0), list(grid.lon).index(5.0)] == 1
    assert flags.sum() == 121


def test_reference_kat_track_events():
    grid = _demo_grid()  # :130-144
    sq = np.array([[0, 0], [10, 0], [10, 10], [0, 10]], dtype=float)
    ev = pd.DataFrame({"date": pd.to_datetime(grid.time[:2]), "geometry": [[sq], [sq]]})
    tracked = P.track_events(ev, method="by_overlap")
    assert tracked.iloc[0].label == 0 and tracked.iloc[1].label == 0


# ------------------------------------------------------------------ third-party fixtures
def test_third_party_fixture():
    z = np.load(os.path.join(GOLD, "third_party.npz"))
    w = np.array([[0, 1, 0], [1, 2, 1], [0, 1, 0]])
    from scipy import ndimage

    assert np.array_equal(ndimage.convolve(z["f32"], weights=w, mode="wrap"), z["conv32"])
    assert np.array_equal(ndimage.convolve(z["f64"], weights=w, mode="wrap"), z["conv64"])
    # tap order of A.7: ((((N + W) + 2C) + E) + S) in double, cast to the input dtype
    f = z["f64"]
    man = ((((np.roll(f, 1, 0) + np.roll(f, 1, 1)) + 2 * f) + np.roll(f, -1, 1)) + np.roll(f, -1, 0))
    assert np.array_equal(man, z["conv64"])
    assert z["div32"].dtype == (np.float64 if int(np.__version__.split(".")[0]) >= 2 else np.float32)
    # haversine: scalar libm restatement of sklearn's expression is bit-identical (A.3)
    X = np.radians(z["latlon"])
    hav = np.zeros((len(X), len(X)))
    for i in range(len(X)):
        for j in range(i, len(X)):
            s0 = math.sin(0.5 * (X[i, 0] - X[j, 0]))
            s1 = math.sin(0.5 * (X[i, 1] - X[j, 1]))
            hav[i, j] = hav[j, i] = 2 * math.asin(math.sqrt(s0 * s0 + math.cos(X[i, 0]) * math.cos(X[j, 0]) * s1 * s1))
    assert np.array_equal(hav * 6371, z["hav"])


# ------------------------------------------------------------------ pipeline fixture
def _fixture():
    z = np.load(os.path.join(GOLD, "pipeline_91x180.npz"))
    lat, lon = synthetic.grid_coords(91, 180)
    grid = P.Grid(lon, lat, synthetic.time_axis(2, 6))
    return z, grid


def test_oracle_matches_pipeline_fixture():
    z, grid = _fixture()
    out = P.detect_steps(z["raw"], grid, levels=[2.0, -2.0])
    assert np.array_equal(np.nan_to_num(out["smoothed"]), np.nan_to_num(z["smoothed"]))
    c = out["contours"]
    assert np.array_equal(np.concatenate([np.asarray(g) for g in c.geometry]), z["c_pts"])
    assert np.array_equal(c.exp_lon.values.astype(float), z["c_exp_lon"])
    for kind, ev in out["events"].items():
        assert np.array_equal(out["flags"][kind], z[kind + "_flags"])
        assert np.array_equal(ev.event_area.values, z[kind + "_event_area"])
        assert np.array_equal(np.concatenate([np.asarray(r) for r in ev.attrs["_index_rings"]]), z[kind + "_rings"])


def _check_cuda_against_fixture():
    z, grid = _fixture()
    det = pipeline.Detector(grid.lat, grid.lon, levels=[2.0, -2.0])
    res = det.run_batch(spatial.to_device(z["raw"]))
    h = res.contours.host()
    assert np.array_equal(np.c_[h["x"], h["y"]], z["c_pts"])
    assert np.array_equal(h["nx"] * grid.dlon, z["c_exp_lon"])
    assert np.array_equal(h["closed"], z["c_closed"])
    for k, kind in enumerate(detect.KINDS):
        tab = res.tables[kind]
        assert np.array_equal(res.flags[k].cpu().numpy(), z[kind + "_flags"])
        props = detect.finish_properties(tab, grid.lon, grid.lat, grid.nlon)
        assert np.array_equal(props["event_area"], z[kind + "_event_area"])
        assert np.array_equal(props["mean_var"], z[kind + "_mean_var"])
        assert np.array_equal(np.array([list(v) for v in props["com"]]), z[kind + "_com"])
        if kind != "overturnings":
            rings = detect.event_rings(res.contours, tab)
            assert np.array_equal(np.concatenate(rings), z[kind + "_rings"])


def test_cuda_sources_match_pipeline_fixture_emu(emu):
    _check_cuda_against_fixture()


@pytest.mark.gpu
def test_cuda_matches_pipeline_fixture_gpu(gpu):
    _check_cuda_against_fixture()
