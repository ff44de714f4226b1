"""Whole detection path (Detector) vs the oracle's call-for-call CPU path."""

import numpy as np
import pytest

from oracle import pipeline as P
from wavebreaking_b200 import detect, pipeline, spatial, synthetic


def _run(nlat, nlon, ntime, levels, step_hours=6.0):
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = synthetic.pv_field(nlat, nlon, np.arange(ntime) * step_hours)
    grid = P.Grid(lon, lat, synthetic.time_axis(ntime, step_hours))
    det = pipeline.Detector(lat, lon, levels=levels)
    res = det.run_batch(spatial.to_device(raw))
    want = P.detect_steps(raw, grid, levels=levels)
    return res, want


def _compare(res, want):
    for k, kind in enumerate(detect.KINDS):
        assert len(res.tables[kind]) == len(want["events"][kind]), kind
        assert np.array_equal(res.flags[k].cpu().numpy(), want["flags"][kind]), kind
    assert res.contours.ncontours == len(want["contours"])


def test_detector_emu(emu):
    res, want = _run(91, 180, 2, [2.0, -2.0])
    _compare(res, want)
    assert sum(len(t) for t in res.tables.values()) > 0


def test_detector_host_path_emu(emu):
    import torch

    lat, lon = synthetic.grid_coords(46, 90)
    raw = synthetic.pv_field(46, 90, np.arange(2) * 6.0)
    det = pipeline.Detector(lat, lon, levels=[2.0])
    res = det.run_batch_host(torch.from_numpy(raw))
    res2 = det.run_batch(spatial.to_device(raw))
    assert np.array_equal(res.flags.numpy(), res2.flags.numpy())
    assert pipeline.summarize(res) == pipeline.summarize(res2)


@pytest.mark.gpu
def test_detector_gpu_one_degree(gpu):
    res, want = _run(181, 360, 4, [2.0, -2.0])
    _compare(res, want)


@pytest.mark.gpu
def test_detector_gpu_batch_invariance(gpu):
    """Size-independent property at the benchmark shape: a batch of T steps == T batches of one step."""
    lat, lon = synthetic.grid_coords(721, 1440)
    raw = spatial.synth_pv(6, 721, 1440, hour0=0.0, hour_step=1.0)
    det = pipeline.Detector(lat, lon, levels=[2.0])
    whole = det.run_batch(raw)
    for t in range(6):
        one = det.run_batch(raw[t:t + 1])
        assert np.array_equal(one.flags[:, 0].cpu().numpy(), whole.flags[:, t].cpu().numpy())
        for kind in detect.KINDS:
            sel = whole.tables[kind].job == t
            assert int(sel.sum()) == len(one.tables[kind])
            assert np.array_equal(whole.tables[kind].sums[sel], one.tables[kind].sums)


def _same_result(a, b):
    assert pipeline.summarize(a) == pipeline.summarize(b)
    assert np.array_equal(a.flags.cpu().numpy(), b.flags.cpu().numpy())
    ha, hb = a.contours.host(), b.contours.host()
    for k in ("x", "y", "closed", "nx", "job"):
        assert np.array_equal(ha[k], hb[k]), k
    for kind in detect.KINDS:
        assert np.array_equal(a.tables[kind].sums, b.tables[kind].sums)


@pytest.mark.parametrize("shape,dtype,passes", [((46, 90), np.float32, 5), ((57, 112), np.float64, 3), ((71, 140), np.float32, 8)])
def test_fused_smoothing_contours_equals_separate_emu(emu, shape, dtype, passes):
    """wbk_smooth_contours == wbk_smooth + wbk_contours (tile seams, odd shapes, both dtypes)."""
    nlat, nlon = shape
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = synthetic.pv_field(nlat, nlon, np.arange(2) * 6.0, dtype=np.float64)
    raw = np.ascontiguousarray(raw).astype(dtype)
    a = pipeline.Detector(lat, lon, levels=[2.0, -2.0], passes=passes, fuse=True).run_batch(spatial.to_device(raw))
    b = pipeline.Detector(lat, lon, levels=[2.0, -2.0], passes=passes, fuse=False).run_batch(spatial.to_device(raw))
    _same_result(a, b)


@pytest.mark.gpu
def test_fused_smoothing_contours_equals_separate_gpu(gpu):
    lat, lon = synthetic.grid_coords(721, 1440)
    raw = spatial.synth_pv(3, 721, 1440, hour0=100.0, hour_step=50.0)
    a = pipeline.Detector(lat, lon, levels=[2.0, 1.5], fuse=True).run_batch(raw)
    b = pipeline.Detector(lat, lon, levels=[2.0, 1.5], fuse=False).run_batch(raw)
    _same_result(a, b)


def test_stream_uses_the_global_exp_lon_maximum_emu(emu, caplog):
    """exp_lon.max() is global over all dates (streamer_index.py:106): a batch WITHOUT a circumglobal contour must not
    promote its widest contour to 'full width'.  Stream of two batches vs. one call on the concatenated record."""
    import logging

    nlat, nlon = 46, 90
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = synthetic.pv_field(nlat, nlon, np.arange(2) * 6.0)
    # second batch: the field is capped below the contour level over a band of longitudes, so no contour is circumglobal
    broken = raw.copy()
    broken[:, :, 30:40] = np.minimum(broken[:, :, 30:40], 0.5)
    record = np.concatenate([raw, broken])
    det = pipeline.Detector(lat, lon, levels=[2.0])
    whole = det.run_batch(spatial.to_device(record))
    assert whole.contours.max_nx == nlon + det.add
    parts = [(pipeline.summarize(r), r.flags.cpu().numpy().copy()) for r in
             det.stream([spatial.to_device(raw), spatial.to_device(broken)], depth=1)]
    assert parts[1][0]["streamers"] == 0 and parts[1][0]["overturnings"] == 0
    flags = np.concatenate([p[1] for p in parts], axis=1)
    assert np.array_equal(flags, whole.flags.cpu().numpy())
    for k in ("streamers", "overturnings", "cutoffs"):
        assert parts[0][0][k] + parts[1][0][k] == len(whole.tables[k])
    # per-batch maxima (the old default) give a different, wrong answer for the second batch
    own = list(det.stream([spatial.to_device(broken)], depth=1, gmax_nx=None))[0]
    assert pipeline.summarize(own)["cutoffs"] != parts[1][0]["cutoffs"] or pipeline.summarize(own)["streamers"] > 0
    # a record without any circumglobal contour is reported
    with caplog.at_level(logging.WARNING, logger="wavebreaking_b200.pipeline"):
        list(det.stream([spatial.to_device(broken)], depth=1))
    assert "re-run with gmax_nx" in caplog.text


def test_packed_flags_emu(emu):
    import torch

    nlat, nlon, nt = 46, 90, 3
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = synthetic.pv_field(nlat, nlon, np.arange(nt) * 6.0)
    det = pipeline.Detector(lat, lon, levels=[2.0])
    packed = torch.zeros(pipeline.packed_nbytes(3 * nt * nlat * nlon), dtype=torch.uint8)
    slot = det._slot(nt)
    det.submit(slot, spatial.to_device(raw), packed_host=packed)
    res = det.collect(slot)
    assert res.flags_packed is packed
    assert np.array_equal(pipeline.unpack_flags(packed, nt, nlat, nlon), res.flags.cpu().numpy())
    assert res.flags.cpu().numpy().any()


@pytest.mark.gpu
def test_cuda_graph_replay_equals_eager_gpu(gpu):
    """Detector(graphs=True): the third submission of the same buffers replays a captured graph; results, packed
    flags and host tables equal the eager run, for different data in the same buffers"""
    import torch

    nlat, nlon, T = 181, 360, 8
    lat, lon = synthetic.grid_coords(nlat, nlon)
    eager = pipeline.Detector(lat, lon, levels=[2.0])
    det = pipeline.Detector(lat, lon, levels=[2.0], graphs=True)
    host = torch.empty((T, nlat, nlon), dtype=torch.float32).pin_memory()
    packed = torch.zeros(pipeline.packed_nbytes(3 * T * nlat * nlon), dtype=torch.uint8).pin_memory()
    for rep in range(4):
        raw = synthetic.pv_field(nlat, nlon, (np.arange(T) + 10 * rep) * 6.0).astype(np.float32)
        host.copy_(torch.from_numpy(raw))
        got = list(det.stream([host], depth=1, packed_host=[packed]))[0]
        want = eager.run_batch(spatial.to_device(raw))
        assert pipeline.summarize(got) == pipeline.summarize(want), rep
        assert np.array_equal(pipeline.unpack_flags(packed, T, nlat, nlon), want.flags.cpu().numpy())
        for kind in detect.KINDS:
            assert np.array_equal(got.tables[kind].sums, want.tables[kind].sums)
    assert det.graph_replays >= 2 and det.graph_kernel_launches > 20


def _nan_patch_case(seed, fuse):
    """missing values in the input: a NaN blob and a NaN at the seam column.  The smoothing spreads them by `passes`
    cells (scipy semantics), the contour stage must skip every square that touches one (skimage), and the events
    whose members would include a NaN cell carry NaN properties as in the reference.  (Infinite values are not a
    parity target: skimage's interpolation fraction is NaN beside them and the reference casts that to int.)"""
    from test_contours import compare_contours

    rng = np.random.default_rng(900 + seed)
    nlat, nlon, nt = 91, 180, 2
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = synthetic.pv_field(nlat, nlon, np.arange(nt) * 6.0).astype(np.float32)
    raw += 0.2 * rng.standard_normal(raw.shape).astype(np.float32)
    for t in range(nt):
        cy, cx = rng.integers(25, 65), rng.integers(20, 160)
        raw[t, cy:cy + rng.integers(1, 4), cx:cx + rng.integers(1, 5)] = np.nan
        raw[t, rng.integers(20, 70), nlon - 1] = np.nan  # its spread wraps around the seam
    grid = P.Grid(lon, lat, synthetic.time_axis(nt, 6.0))
    with np.errstate(all="ignore"):
        want = P.detect_steps(raw, grid, levels=[2.0])
    det = pipeline.Detector(lat, lon, levels=[2.0], fuse=fuse)
    res = det.run_batch(spatial.to_device(raw))
    compare_contours(res.contours, want["contours"], grid, [2.0])
    _compare(res, want)
    assert np.isnan(want["smoothed"]).sum() > 100


@pytest.mark.parametrize("fuse", [True, False])
def test_missing_values_in_the_input_emu(emu, fuse):
    _nan_patch_case(0, fuse)


@pytest.mark.gpu
def test_missing_values_in_the_input_gpu(gpu):
    _nan_patch_case(1, True)
    _nan_patch_case(2, False)


def _noisy_detector_case(seed, amp, levels):
    """the fused path (smoothing with bit planes -> marching squares on planes -> ... -> flags) on a rough field:
    contours point by point, every event with its properties, and the flag grids against the oracle"""
    from test_contours import compare_contours
    from test_indices import compare_events

    rng = np.random.default_rng(1100 + seed)
    nlat, nlon, nt = 91, 180, 2
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = synthetic.pv_field(nlat, nlon, np.arange(nt) * 6.0).astype(np.float32)
    raw += amp * rng.standard_normal(raw.shape).astype(np.float32)
    grid = P.Grid(lon, lat, synthetic.time_axis(nt, 6.0))
    want = P.detect_steps(raw, grid, levels=levels)
    res = pipeline.Detector(lat, lon, levels=levels).run_batch(spatial.to_device(raw))
    compare_contours(res.contours, want["contours"], grid, levels)
    compare_events(res.contours, res.tables, res.flags, want["events"], grid, levels, want["smoothed"], simple_pieces_only=True)
    return res


def test_detector_on_rough_fields_emu(emu):
    res = _noisy_detector_case(0, 6.0, [2.0])
    assert res.contours.ncontours > 60
    _noisy_detector_case(1, 3.0, [2.0, -2.0])


@pytest.mark.gpu
def test_detector_on_rough_fields_gpu(gpu):
    _noisy_detector_case(2, 6.0, [2.0])
    _noisy_detector_case(3, 3.0, [2.0, -2.0])
