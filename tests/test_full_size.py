"""Parity and size-independent properties at BASELINE.json's shapes (GPU only).

configs[1]: 1.0 degree (181 x 360), 6-hourly;  configs[2]: 0.25 degree (721 x 1440), hourly, three indices +
to_xarray;  configs[3]: several contour levels with an intensity field.  The oracle is too slow for whole years, so
the long batches are checked through properties (sharding invariance, two independent rasterisers agreeing,
uniqueness / bounds of the contour points) and a spread of single steps is compared event by event.
"""

import numpy as np
import pytest

import wavebreaking_b200 as wb
from oracle import pipeline as P
from wavebreaking_b200 import compat, detect, pipeline, spatial, synthetic

pytestmark = pytest.mark.gpu


def test_c1_year_sample_matches_oracle(gpu):
    """configs[1]: 16 six-hourly steps spread over the year at 181 x 360, levels +-2, event by event."""
    nlat, nlon = 181, 360
    lat, lon = synthetic.grid_coords(nlat, nlon)
    hours = np.sort(np.random.default_rng(5).choice(1460, 16, replace=False)) * 6.0
    raw = synthetic.pv_field(nlat, nlon, hours)
    grid = P.Grid(lon, lat, synthetic.time_axis(len(hours), 6))
    res = pipeline.Detector(lat, lon, levels=[2.0, -2.0]).run_batch(spatial.to_device(raw))
    want = P.detect_steps(raw, grid, levels=[2.0, -2.0])
    assert res.contours.ncontours == len(want["contours"])
    for k, kind in enumerate(detect.KINDS):
        w = want["events"][kind]
        tab = res.tables[kind]
        assert len(tab) == len(w) > 0, kind
        assert np.array_equal(res.flags[k].cpu().numpy(), want["flags"][kind]), kind
        props = detect.finish_properties(tab, lon, lat, nlon)
        assert [tuple(c) for c in props["com"]] == [tuple(c) for c in w.com]
        assert np.array_equal(props["event_area"], w.event_area.values)
        assert np.array_equal(props["mean_var"], w.mean_var.values)
    assert int(sum((t.near & 1).sum() for t in res.tables.values())) == 0  # no decision within 1e-9 of a threshold


def test_c25_sharding_invariance_and_contour_lattice(gpu):
    """configs[2]: 64 hourly steps == two shards of 32 (what `sharding.run_sharded` relies on); every contour has
    at least four unique lattice points inside the extended grid."""
    nlat, nlon, T = 721, 1440, 64
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = spatial.synth_pv(T, nlat, nlon, hour0=4000.0, hour_step=1.0)
    det = pipeline.Detector(lat, lon, levels=[2.0])
    whole = det.run_batch(raw)
    gmax = int(whole.contours.host()["nx"].max())
    lo, hi = det.run_batch(raw[:T // 2], gmax_nx=gmax), det.run_batch(raw[T // 2:], gmax_nx=gmax)
    assert np.array_equal(whole.flags[:, :T // 2].cpu().numpy(), lo.flags.cpu().numpy())
    assert np.array_equal(whole.flags[:, T // 2:].cpu().numpy(), hi.flags.cpu().numpy())
    for kind in detect.KINDS:
        w = whole.tables[kind]
        first = w.job < T // 2
        assert int(first.sum()) == len(lo.tables[kind]) and int((~first).sum()) == len(hi.tables[kind])
        assert np.array_equal(w.sums[first], lo.tables[kind].sums) and np.array_equal(w.sums[~first], hi.tables[kind].sums)
    h = whole.contours.host()
    W = nlon + int(120 / 0.25)
    assert h["x"].min() >= 0 and h["x"].max() <= W - 1 and h["y"].min() >= 0 and h["y"].max() <= nlat - 1
    for c in range(whole.contours.ncontours):
        a, b = h["pt_off"][c], h["pt_off"][c + 1]
        x, y = h["x"][a:b].astype(np.int64), h["y"][a:b].astype(np.int64)
        assert b - a >= 4
        assert len(np.unique(x * 4096 + y)) == b - a           # keep-first dedupe (contour_index.py:110-112)


def test_c25_two_rasterisers_agree(gpu):
    """configs[2]: flag grids written by the event rasteriser (device meridian split) == `to_xarray` of the same
    events through the API (host integer split + the ring rasteriser), three steps, all three indices."""
    nlat, nlon, T = 721, 1440, 3
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = spatial.synth_pv(T, nlat, nlon, hour0=2500.0, hour_step=11.0)
    res = pipeline.Detector(lat, lon, levels=[2.0]).run_batch(raw)
    time = synthetic.time_axis(T, 11)
    data = compat.Field(raw.cpu().numpy(), ("time", "lat", "lon"), {"time": time, "lat": lat, "lon": lon}, name="PV")
    sm = wb.calculate_smoothed_field(data, 5)
    index = wb.calculate_contours(sm, 2, original_coordinates=False)
    for k, fn in enumerate((wb.calculate_streamers, wb.calculate_overturnings, wb.calculate_cutoffs)):
        ev = fn(sm, 2, contours=index)
        assert len(ev) == len(res.tables[detect.KINDS[k]]) > 0
        assert np.array_equal(np.asarray(wb.to_xarray(sm, ev).values), res.flags[k].cpu().numpy()), detect.KINDS[k]


def test_c25_levels_and_intensity_match_oracle(gpu):
    """configs[3]: levels 1.5 / 2 / 3 PVU with an intensity field, one 721 x 1440 step, event by event."""
    nlat, nlon = 721, 1440
    lat, lon = synthetic.grid_coords(nlat, nlon)
    levels = [1.5, 2.0, 3.0]
    raw = spatial.synth_pv(1, nlat, nlon, hour0=6100.0).cpu().numpy()
    grid = P.Grid(lon, lat, synthetic.time_axis(1, 1))
    sm = P.smooth_field(raw, 5)
    inten = np.cos(np.radians(lat))[None, :, None] * np.nan_to_num(sm) * 0.25 + 1.0
    res = pipeline.Detector(lat, lon, levels=levels).run_batch(spatial.to_device(raw), intensity=spatial.to_device(inten))
    c = P.calculate_contours(sm, levels, grid, 120, original_coordinates=False)
    assert res.contours.ncontours == len(c)
    for kind, fn in (("streamers", P.calculate_streamers), ("overturnings", P.calculate_overturnings),
                     ("cutoffs", P.calculate_cutoffs)):
        want = fn(sm, grid, c, intensity=inten)
        tab = res.tables[kind]
        assert len(tab) == len(want) > 0, kind
        props = detect.finish_properties(tab, lon, lat, nlon)
        assert [levels[j % 3] for j in tab.job] == list(want.level)
        assert np.array_equal(props["intensity"], want.intensity.values)
        assert np.array_equal(props["mean_var"], want.mean_var.values)
        assert np.array_equal(props["event_area"], want.event_area.values)
        assert [tuple(c) for c in props["com"]] == [tuple(c) for c in want.com]
