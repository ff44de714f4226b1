"""Smoothing / momentum flux: CUDA kernels vs the real scipy routine the reference calls."""

import numpy as np
import pytest
import torch

from oracle import pipeline as P
from wavebreaking_b200 import spatial, synthetic


def _field(nlat, nlon, ntime, dtype, seed=0):
    rng = np.random.default_rng(seed)
    f = synthetic.pv_field(nlat, nlon, np.arange(ntime) * 6.0, dtype=np.float64)
    f = f + 0.05 * rng.standard_normal(f.shape)
    return f.astype(dtype)


def _check_smooth(dtype, passes, nlat=37, nlon=72, ntime=2, **kw):
    f = _field(nlat, nlon, ntime, dtype)
    want = P.smooth_field(f, passes, **kw)
    got = spatial.smooth(f, passes, **kw).cpu().numpy()
    assert got.dtype == want.dtype
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(np.nan_to_num(got), np.nan_to_num(want))  # bit-exact


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("passes", [0, 1, 5, 10])
def test_smooth_emu_bit_exact(emu, dtype, passes):
    _check_smooth(dtype, passes)


def test_smooth_emu_nan_and_odd_shapes(emu):
    f = _field(19, 50, 1, np.float32)
    f[0, 7, 11] = np.nan
    want = P.smooth_field(f, 3)
    got = spatial.smooth(f, 3).cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(np.nan_to_num(got), np.nan_to_num(want))


@pytest.mark.parametrize("mode", ["wrap", "reflect", "mirror", "nearest", "constant"])
def test_convolve_generic_emu(emu, mode):
    w = np.array([[1, 2, 0], [0, 3, 1], [2, 0, 1], [1, 1, 1]])
    _check_smooth(np.float32, 2, nlat=12, nlon=17, ntime=1, weights=w, mode=mode)
    _check_smooth(np.float64, 1, nlat=12, nlon=17, ntime=1, weights=w.T.copy(), mode=mode)


def test_mflux_emu(emu):
    u = _field(19, 36, 2, np.float32, 1)
    v = _field(19, 36, 2, np.float32, 2)
    u[0, 3, 5] = np.nan
    want = P.momentum_flux(u, v)
    got = spatial.momentum_flux(u, v).cpu().numpy()
    assert got.dtype == np.float32
    # the zonal means follow numpy's pairwise float32 summation: bit-exact
    assert np.array_equal(got, want, equal_nan=True)
    for nlon, dtype in ((7, np.float32), (128, np.float32), (129, np.float64), (1440, np.float32), (1447, np.float64)):
        a, b = _field(5, nlon, 1, dtype, 3), _field(5, nlon, 1, dtype, 4)
        a[0, 1, nlon // 2] = np.nan
        assert np.array_equal(spatial.momentum_flux(a, b).cpu().numpy(), P.momentum_flux(a, b), equal_nan=True)
    # longitude not the contiguous axis of the user's array (dims time, lon, lat): numpy sums sequentially
    a, b = _field(9, 300, 2, np.float32, 5), _field(9, 300, 2, np.float32, 6)
    at, bt = np.ascontiguousarray(a.transpose(0, 2, 1)), np.ascontiguousarray(b.transpose(0, 2, 1))
    want = ((at - np.nanmean(at, axis=1, keepdims=True)) * (bt - np.nanmean(bt, axis=1, keepdims=True))).transpose(0, 2, 1)
    assert np.array_equal(spatial.momentum_flux(a, b, lon_contiguous=False).cpu().numpy(), want)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(181, 360), (721, 1440)])
def test_smooth_gpu_bit_exact(gpu, dtype, shape):
    _check_smooth(dtype, 5, nlat=shape[0], nlon=shape[1], ntime=2)


@pytest.mark.gpu
def test_smooth_gpu_many_passes_and_generic(gpu):
    _check_smooth(np.float32, 10, nlat=91, nlon=180, ntime=1)
    _check_smooth(np.float64, 19, nlat=91, nlon=180, ntime=1)
    w = np.array([[1, 2, 0], [0, 3, 1], [2, 0, 1]])
    for mode in ["wrap", "reflect", "mirror", "nearest", "constant"]:
        _check_smooth(np.float32, 2, nlat=45, nlon=90, ntime=1, weights=w, mode=mode)


@pytest.mark.gpu
def test_mflux_gpu(gpu):
    u = _field(181, 360, 2, np.float32, 1)
    v = _field(181, 360, 2, np.float32, 2)
    want = P.momentum_flux(u, v)
    got = spatial.momentum_flux(u, v).cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True)  # numpy's pairwise float32 order reproduced
    u64, v64 = _field(721, 1440, 1, np.float64, 3), _field(721, 1440, 1, np.float64, 4)
    assert np.array_equal(spatial.momentum_flux(u64, v64).cpu().numpy(), P.momentum_flux(u64, v64))


@pytest.mark.gpu
def test_synth_gpu_matches_host_recipe(gpu):
    import torch

    got = spatial.synth_pv(2, 91, 180, hour0=0.0, hour_step=6.0, dtype=torch.float64).cpu().numpy()
    want = synthetic.pv_field(91, 180, np.array([0.0, 6.0]), dtype=np.float64)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)


@pytest.mark.gpu
def test_staged_upload_of_pageable_arrays_gpu(gpu):
    """arrays of >= 128 MiB in ordinary host memory go up through the pinned staging ring: every byte arrives, for
    sizes that are not a multiple of the chunk, twice in a row (buffer reuse) and from a read-only source"""
    rng = np.random.default_rng(5)
    for dtype, n in ((np.float32, (40 << 20) + 12345), (np.float64, (17 << 20) + 7)):
        a = rng.standard_normal(n).astype(dtype)
        a.setflags(write=False)
        for _ in range(2):
            t = spatial.to_device(a)
            assert t.dtype == torch.from_numpy(a[:1].copy()).dtype and tuple(t.shape) == a.shape
            assert np.array_equal(t.cpu().numpy(), a)
    b = rng.standard_normal((37, 721, 1440)).astype(np.float32)
    assert np.array_equal(spatial.to_device(b).cpu().numpy(), b)
