"""The register-streamed smoothing kernel (wbk_smooth.cu): orientation folded into the loads, packed int16
input, the plain-division fallback for values outside the fast division's range, strips narrower / wider than
the grid, and the bit planes it leaves for the marching-squares stage."""

import numpy as np
import pytest
import torch

from oracle import pipeline as P
from wavebreaking_b200 import _lib, detect, pipeline, spatial, synthetic


def _smooth_raw(x, passes, rmode, out_dtype, opts=None):
    lib = _lib.get()
    x = x.to(lib.device).contiguous()
    nt, nlat, nlon = x.shape
    out = torch.empty(x.shape, dtype=out_dtype, device=lib.device)
    tmp = torch.empty_like(out) if passes > _lib.SMOOTH_MAX_FUSED else None
    lib.call("wbk_smooth", _lib.ptr(x), _lib.dtype_code(x.dtype), _lib.ptr(out), _lib.dtype_code(out_dtype),
             _lib.ptr(tmp), nt, nlat, nlon, passes, rmode, opts, lib.stream())
    return out.cpu().numpy()


def _eq(got, want):
    assert got.dtype == want.dtype
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(np.nan_to_num(got, posinf=1e308, neginf=-1e308), np.nan_to_num(want, posinf=1e308, neginf=-1e308))


def _field(nlat, nlon, ntime, dtype, seed=0):
    rng = np.random.default_rng(seed)
    f = synthetic.pv_field(nlat, nlon, np.arange(ntime) * 6.0, dtype=np.float64)
    return (f + 0.05 * rng.standard_normal(f.shape)).astype(dtype)


@pytest.mark.parametrize("shape", [(9, 20), (33, 54), (40, 55), (37, 130)])
@pytest.mark.parametrize("flips", [(True, False), (False, True), (True, True)])
def test_orientation_folded_into_loads_emu(emu, shape, flips):
    """descending latitude / longitude (utils/data_utils.py:196-213): same result as smoothing the re-sorted field"""
    nlat, nlon = shape
    fl, fo = flips
    f = _field(nlat, nlon, 2, np.float32, 3)
    stored = f[:, ::-1] if fl else f
    stored = stored[:, :, ::-1] if fo else stored
    want = P.smooth_field(f, 5)
    got = _smooth_raw(torch.from_numpy(np.ascontiguousarray(stored)), 5, _lib.ROUND_FIRST, torch.float64,
                      _lib.smooth_opts(fl, fo))
    _eq(got, want)


@pytest.mark.parametrize("passes", [1, 2, 8, 11])
def test_numpy1_rounding_and_f64_emu(emu, passes):
    f = _field(21, 70, 1, np.float32, 5)
    # NumPy 1.x: every pass stays float32
    want32 = f.copy()
    from scipy import ndimage

    w = spatial.DEFAULT_WEIGHTS
    for _ in range(passes):
        want32[0] = (ndimage.convolve(want32[0], w, mode="wrap") / np.float32(6)).astype(np.float32)
    want32[:, [0, 1, -2, -1]] = np.nan
    _eq(_smooth_raw(torch.from_numpy(f), passes, _lib.ROUND_ALL, torch.float32), want32)
    f64 = f.astype(np.float64) * 1.000000123
    _eq(_smooth_raw(torch.from_numpy(f64), passes, _lib.ROUND_NONE, torch.float64), P.smooth_field(f64, passes))


def test_values_outside_fast_division_range_emu(emu):
    """Inf / NaN / huge / tiny values switch the affected rows to the plain division; results stay bit-exact"""
    f = _field(30, 70, 1, np.float64, 7)
    f[0, 5, 10] = np.inf
    f[0, 12, 60] = 1e300
    f[0, 13, 3] = -1e299
    f[0, 20, 33] = 1e-200
    f[0, 21, 34] = -3e-310  # subnormal
    f[0, 25, 1] = np.nan
    f[0, 27, :] = 0.0
    with np.errstate(all="ignore"):
        want = P.smooth_field(f, 5)
        got = _smooth_raw(torch.from_numpy(f), 5, _lib.ROUND_NONE, torch.float64)
    _eq(got, want)
    f32 = _field(30, 70, 1, np.float32, 8)
    f32[0, 9, 9] = np.inf
    f32[0, 15, 40] = 3e38
    f32[0, 16, 40] = 1e-44
    with np.errstate(all="ignore"):
        _eq(_smooth_raw(torch.from_numpy(f32), 5, _lib.ROUND_FIRST, torch.float64), P.smooth_field(f32, 5))


def test_packed_int16_input_emu(emu):
    """CF-packed shorts are decoded in the load exactly as xarray does (float64(v) * scale + offset, fill -> NaN)"""
    rng = np.random.default_rng(11)
    packed = rng.integers(-32000, 32000, size=(2, 25, 80)).astype(np.int16)
    packed[1, 10, 10] = -32767
    scale, offset, fill = 2.3283064365386963e-04 * 3.1, 1.234567, -32767
    decoded = packed.astype(np.float64) * scale + offset
    decoded[packed == fill] = np.nan
    for fl in (False, True):
        stored = packed[:, ::-1] if fl else packed
        got = _smooth_raw(torch.from_numpy(np.ascontiguousarray(stored)), 4, _lib.ROUND_NONE, torch.float64,
                          _lib.smooth_opts(fl, False, scale, offset, fill))
        _eq(got, P.smooth_field(decoded, 4))
    # passes == 0: decode + orientation + NaN border only
    got0 = _smooth_raw(torch.from_numpy(packed), 0, _lib.ROUND_NONE, torch.float64, _lib.smooth_opts(False, False, scale, offset, fill))
    want0 = decoded.copy()
    want0[:, [0, 1, -2, -1]] = np.nan
    _eq(got0, want0)


def test_detector_descending_latitude_with_regrow_emu(emu):
    """ERA5-style descending latitude through the fused path, with arenas so small that the batch is re-run with
    larger ones (the re-run must see the caller's buffer, not an already re-oriented one), device and host input"""
    nlat, nlon = 46, 90
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = synthetic.pv_field(nlat, nlon, np.arange(3) * 6.0)
    ref = pipeline.Detector(lat, lon, levels=[2.0]).run_batch(spatial.to_device(raw))
    for host in (False, True):
        det = pipeline.Detector(lat[::-1].copy(), lon, levels=[2.0])
        tiny = dict(seg_cap=32, contour_cap=4, pair_cap=16, event_cap=1, sel_cap=1)
        orig = detect.default_caps

        def small(*a, **k):
            caps = orig(*a, **k)
            caps.update(tiny)
            return caps

        detect.default_caps = small
        try:
            stored = np.ascontiguousarray(raw[:, ::-1])
            res = det.run_batch_host(torch.from_numpy(stored)) if host else det.run_batch(spatial.to_device(stored))
        finally:
            detect.default_caps = orig
        assert det._grow, "the tiny arenas were expected to overflow"
        assert pipeline.summarize(res) == pipeline.summarize(ref)
        assert np.array_equal(res.flags.cpu().numpy(), ref.flags.cpu().numpy())
    # the streaming entry (explicit global exp_lon maximum) must notice overflowing arenas too: the status bits of the
    # contour / packing stages survive the index stage
    det = pipeline.Detector(lat, lon, levels=[2.0])
    tiny = dict(seg_cap=32, contour_cap=4, pair_cap=16, event_cap=1, sel_cap=1)
    orig = detect.default_caps
    detect.default_caps = lambda *a, **k: {**orig(*a, **k), **tiny}
    try:
        res = list(det.stream([spatial.to_device(raw)], depth=1))[0]
        res_flags = res.flags.cpu().numpy().copy()
    finally:
        detect.default_caps = orig
    assert det._grow and pipeline.summarize(res) == pipeline.summarize(ref)
    assert np.array_equal(res_flags, ref.flags.cpu().numpy())


def test_detector_int16_input_emu(emu):
    nlat, nlon = 46, 90
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = synthetic.pv_field(nlat, nlon, np.arange(2) * 6.0, dtype=np.float64)
    lo, hi = raw.min(), raw.max()
    scale, offset = (hi - lo) / 65000.0, (hi + lo) / 2
    packed = np.round((raw - offset) / scale).astype(np.int16)
    decoded = packed.astype(np.float64) * scale + offset
    a = pipeline.Detector(lat, lon, levels=[2.0], packing=(scale, offset, None)).run_batch(torch.from_numpy(packed))
    b = pipeline.Detector(lat, lon, levels=[2.0]).run_batch(spatial.to_device(decoded))
    assert pipeline.summarize(a) == pipeline.summarize(b)
    assert np.array_equal(a.flags.cpu().numpy(), b.flags.cpu().numpy())
    for kind in detect.KINDS:
        assert np.array_equal(a.tables[kind].sums, b.tables[kind].sums)


@pytest.fixture
def two_halves():
    """force the 128-column strips (two 64-column halves per warp) that wide grids use, whatever the grid width"""
    lib = _lib.get()
    lib.cdll.wbk_tune_smooth_halves(2)
    yield
    lib.cdll.wbk_tune_smooth_halves(0)


@pytest.mark.parametrize("shape,passes", [((9, 20), 5), ((33, 118), 4), ((30, 119), 5), ((26, 131), 1), ((21, 260), 5)])
def test_two_halves_per_warp_emu(emu, two_halves, shape, passes):
    """strips of 128 columns: the seam between the halves, strips narrower / wider than the grid, flips, all dtypes"""
    nlat, nlon = shape
    f = _field(nlat, nlon, 2, np.float32, 13)
    f[1, nlat // 2, nlon // 3] = np.inf  # one slow-path stretch
    with np.errstate(all="ignore"):
        want = P.smooth_field(f, passes)
        stored = np.ascontiguousarray(f[:, ::-1, ::-1])
        _eq(_smooth_raw(torch.from_numpy(stored), passes, _lib.ROUND_FIRST, torch.float64, _lib.smooth_opts(True, True)), want)
        f64 = f.astype(np.float64) * 1.0000001
        _eq(_smooth_raw(torch.from_numpy(f64), passes, _lib.ROUND_NONE, torch.float64), P.smooth_field(f64, passes))
    packed = np.round(_field(nlat, nlon, 1, np.float64, 14) * 3000).astype(np.int16)
    got = _smooth_raw(torch.from_numpy(packed), passes, _lib.ROUND_NONE, torch.float64, _lib.smooth_opts(False, False, 1 / 3000.0, 0.25, None))
    _eq(got, P.smooth_field(packed.astype(np.float64) * (1 / 3000.0) + 0.25, passes))


@pytest.mark.parametrize("nlon,levels", [(90, [2.0]), (118, [2.0, -2.0, 1.5]), (120, [2.0]), (128, [2.0, -2.0, 1.5]),
                                         (236, [2.0]), (238, [2.0, -2.0, 1.5])])
def test_two_halves_bit_planes_detector_emu(emu, two_halves, nlon, levels):
    """the marching-squares stage reads the planes of both halves: same contours / events as the unfused path"""
    nlat = nlon // 2 + 1  # dlon == dlat
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = synthetic.pv_field(nlat, nlon, np.arange(2) * 6.0)
    a = pipeline.Detector(lat, lon, levels=levels).run_batch(spatial.to_device(raw))
    b = pipeline.Detector(lat, lon, levels=levels, fuse=False).run_batch(spatial.to_device(raw))
    assert pipeline.summarize(a) == pipeline.summarize(b)
    ha, hb = a.contours.host(), b.contours.host()
    for k in ("job", "pt_off", "closed", "x", "y"):
        assert np.array_equal(ha[k], hb[k]), k
    assert np.array_equal(a.flags.cpu().numpy(), b.flags.cpu().numpy())


@pytest.mark.gpu
def test_smooth_stream_gpu_full_size_special_cases(gpu):
    f = _field(721, 1440, 1, np.float32, 3)
    want = P.smooth_field(f, 5)
    stored = np.ascontiguousarray(f[:, ::-1, ::-1])
    _eq(_smooth_raw(torch.from_numpy(stored), 5, _lib.ROUND_FIRST, torch.float64, _lib.smooth_opts(True, True)), want)
    f64 = f.astype(np.float64)
    f64[0, 100, 100] = np.inf
    f64[0, 300, 700] = 1e-250
    f64[0, 500, 1439] = np.nan
    with np.errstate(all="ignore"):
        _eq(_smooth_raw(torch.from_numpy(f64), 5, _lib.ROUND_NONE, torch.float64), P.smooth_field(f64, 5))


@pytest.mark.gpu
def test_two_halves_full_size_gpu(gpu, two_halves):
    """the opt-in 128-column strips at the benchmark shape: smoothed field bit-exact, detection results unchanged"""
    f = _field(721, 1440, 2, np.float32, 4)
    _eq(_smooth_raw(torch.from_numpy(f), 5, _lib.ROUND_FIRST, torch.float64), P.smooth_field(f, 5))
    lat, lon = synthetic.grid_coords(721, 1440)
    raw = spatial.synth_pv(3, 721, 1440, hour0=500.0)
    a = pipeline.Detector(lat, lon, levels=[2.0]).run_batch(raw)
    _lib.get().cdll.wbk_tune_smooth_halves(1)
    b = pipeline.Detector(lat, lon, levels=[2.0]).run_batch(raw)
    assert pipeline.summarize(a) == pipeline.summarize(b)
    assert torch.equal(a.flags, b.flags)
    for kind in detect.KINDS:
        assert np.array_equal(a.tables[kind].sums, b.tables[kind].sums)
