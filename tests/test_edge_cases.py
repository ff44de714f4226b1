"""Edge cases of the path: empty / degenerate inputs, capacity regrowth, multi-level + intensity batches."""

import numpy as np
import pandas as pd
import pytest
import torch

import wavebreaking_b200 as wb
from oracle import pipeline as P
from wavebreaking_b200 import _lib, compat, detect, pipeline, spatial, synthetic


def _field(values, nlat, nlon, ntime):
    lat, lon = synthetic.grid_coords(nlat, nlon)
    return compat.Field(values, ("time", "lat", "lon"),
                        {"time": synthetic.time_axis(ntime, 6), "lat": lat, "lon": lon}, name="PV")


def test_constant_field_has_no_contours_emu(emu):
    data = _field(np.full((2, 31, 60), 1.0, dtype=np.float32), 31, 60, 2)
    c = wb.calculate_contours(data, 2, original_coordinates=False)
    assert len(c) == 0
    det = pipeline.Detector(*synthetic.grid_coords(31, 60), levels=[2.0], passes=0)
    res = det.run_batch(spatial.to_device(np.full((2, 31, 60), 1.0, dtype=np.float32)))
    assert pipeline.summarize(res)["contours"] == 0 and int(res.flags.sum()) == 0
    assert all(len(t) == 0 for t in res.tables.values())


def test_all_nan_field_emu(emu):
    det = pipeline.Detector(*synthetic.grid_coords(31, 60), levels=[2.0], passes=2)
    res = det.run_batch(spatial.to_device(np.full((1, 31, 60), np.nan, dtype=np.float32)))
    assert pipeline.summarize(res)["contours"] == 0


def test_single_time_step_and_tiny_batches_emu(emu):
    lat, lon = synthetic.grid_coords(46, 90)
    raw = synthetic.pv_field(46, 90, np.arange(3) * 6.0)
    det = pipeline.Detector(lat, lon, levels=[2.0])
    whole = det.run_batch(spatial.to_device(raw))
    parts = [det.run_batch(spatial.to_device(raw[t:t + 1])) for t in range(3)]
    for t, one in enumerate(parts):
        assert np.array_equal(one.flags[:, 0].numpy(), whole.flags[:, t].numpy())
    streamed = list(det.stream([spatial.to_device(raw[t:t + 1]) for t in range(3)], depth=2))
    assert [pipeline.summarize(r) for r in streamed] == [pipeline.summarize(r) for r in parts]


def test_capacity_regrow_in_pipeline_emu(emu, monkeypatch):
    lat, lon = synthetic.grid_coords(46, 90)
    raw = synthetic.pv_field(46, 90, np.arange(2) * 6.0)
    want = pipeline.summarize(pipeline.Detector(lat, lon, levels=[2.0, -2.0]).run_batch(spatial.to_device(raw)))
    tiny = dict(max_jobs=4, seg_cap=64, contour_cap=4, sel_cap=1, pair_cap=16, event_cap=1)
    monkeypatch.setattr(detect, "default_caps", lambda nlat, nlon, add, njobs: dict(tiny, max_jobs=int(njobs)))
    det = pipeline.Detector(lat, lon, levels=[2.0, -2.0])
    got = pipeline.summarize(det.run_batch(spatial.to_device(raw)))
    assert got == want
    assert det._grow  # the arenas were regrown at least once


def test_multi_level_with_intensity_matches_oracle_emu(emu):
    nlat, nlon, ntime = 91, 180, 2
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = synthetic.pv_field(nlat, nlon, np.arange(ntime) * 6.0)
    grid = P.Grid(lon, lat, synthetic.time_axis(ntime, 6))
    levels = [1.5, 2.0, 3.0]
    sm = P.smooth_field(raw, 5)
    inten = np.random.default_rng(1).standard_normal(sm.shape)
    det = pipeline.Detector(lat, lon, levels=levels)
    res = det.run_batch(spatial.to_device(raw), intensity=spatial.to_device(inten))
    c = P.calculate_contours(sm, levels, grid, 120, original_coordinates=False)
    for kind, fn in (("streamers", P.calculate_streamers), ("overturnings", P.calculate_overturnings),
                     ("cutoffs", P.calculate_cutoffs)):
        want = fn(sm, grid, c, intensity=inten)
        tab = res.tables[kind]
        assert len(tab) == len(want)
        if len(want):
            props = detect.finish_properties(tab, lon, lat, nlon)
            assert np.array_equal(props["intensity"], want.intensity.values)
            assert np.array_equal(props["mean_var"], want.mean_var.values)
            assert [levels[j % 3] for j in tab.job] == list(want.level)


def test_to_xarray_rejects_off_grid_vertices_emu(emu):
    data = _field(np.zeros((1, 31, 60), dtype=np.float32), 31, 60, 1)
    ev = compat.make_frame({"date": [data["time"].values[0]]}, [compat.Polygon([(0.5, 0.3), (10, 0), (10, 10)])])
    with pytest.raises(ValueError, match="grid points"):
        wb.to_xarray(data, ev)


def test_smoothing_of_extreme_values_is_exact_emu(emu):
    """Tiles with non-finite / tiny / huge values take the plain-division path and stay bit-exact."""
    rng = np.random.default_rng(0)
    f = rng.standard_normal((1, 70, 140))
    f[0, 10, 10] = np.inf
    f[0, 40, 100] = 1e-300
    f[0, 60, 20] = 1e300
    f[0, 5, 70] = 0.0
    want = P.smooth_field(f, 3)
    got = spatial.smooth(f, 3).numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(np.nan_to_num(got, posinf=1e308, neginf=-1e308), np.nan_to_num(want, posinf=1e308, neginf=-1e308))


def test_smoothing_of_non_finite_float32_is_exact_emu(emu):
    """float32 input: tiles holding NaN / Inf leave the fast exact-division path (one fused x*0 check per tile)."""
    rng = np.random.default_rng(3)
    f = rng.standard_normal((2, 70, 140)).astype(np.float32)
    f[0, 12, 17] = np.inf
    f[0, 50, 120] = -np.inf
    f[1, 33, 3] = np.nan
    want = P.smooth_field(f, 5)
    got = spatial.smooth(f, 5).numpy()
    assert got.dtype == want.dtype
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(np.nan_to_num(got, posinf=1e308, neginf=-1e308), np.nan_to_num(want, posinf=1e308, neginf=-1e308))


def test_detector_descending_latitude_emu(emu):
    lat, lon = synthetic.grid_coords(46, 90)
    raw = synthetic.pv_field(46, 90, np.arange(2) * 6.0)
    a = pipeline.Detector(lat, lon, levels=[2.0]).run_batch(spatial.to_device(raw))
    d = pipeline.Detector(lat[::-1].copy(), lon, levels=[2.0]).run_batch(spatial.to_device(raw[:, ::-1, :].copy()))
    assert pipeline.summarize(a) == pipeline.summarize(d)
    assert np.array_equal(a.flags.numpy(), d.flags.numpy())
