"""Contour stage: CUDA kernels vs the skimage / contour_index restatement of the oracle."""

import numpy as np
import pytest
import torch

from oracle import pipeline as P
from oracle import skimage_contours as SK
from wavebreaking_b200 import detect, spatial, synthetic


def make_case(nlat, nlon, ntime, step_hours=6.0, passes=5, seed_noise=None):
    lat, lon = synthetic.grid_coords(nlat, nlon)
    pv = synthetic.pv_field(nlat, nlon, np.arange(ntime) * step_hours)
    if seed_noise is not None:
        rng = np.random.default_rng(seed_noise)
        pv = (pv + 0.3 * rng.standard_normal(pv.shape)).astype(np.float32)
    grid = P.Grid(lon, lat, synthetic.time_axis(ntime, step_hours))
    sm = P.smooth_field(pv, passes)
    return grid, pv, sm


def compare_contours(cs, oracle_df, grid, levels):
    """Row-by-row equality (order, closed flag, nx, mean_lat ingredients, every point)."""
    h = cs.host()
    assert cs.ncontours == len(oracle_df)
    nlev = len(levels)
    for c, row in enumerate(oracle_df.itertuples()):
        job = h["job"][c]
        t, l = divmod(int(job), nlev)
        assert grid.time[t] == row.date and levels[l] == row.level
        pts = cs.contour_points(c)
        want = np.asarray(row.geometry)
        assert pts.shape == want.shape, (c, pts.shape, want.shape)
        assert np.array_equal(pts, want), c
        assert bool(h["closed"][c]) == bool(row.closed)
        assert h["nx"][c] * grid.dlon == row.exp_lon
        assert np.round(h["sum_y"][c] / len(want), 2) == row.mean_lat
    if len(oracle_df):
        assert cs.max_nx * grid.dlon == oracle_df.exp_lon.max()


def run_contours(sm, levels, grid, periodic_add):
    add = int(periodic_add / grid.dlon)
    field = spatial.to_device(sm)
    return detect.contours(field, levels, add)


@pytest.mark.parametrize("levels", [[2], [2, -2], [1.5, 2, 3]])
def test_contours_emu_small(emu, levels):
    grid, pv, sm = make_case(46, 90, 2)
    cs = run_contours(sm, levels, grid, 120)
    want = P.calculate_contours(sm, levels, grid, 120, original_coordinates=False)
    compare_contours(cs, want, grid, levels)
    assert not np.any(cs.status)


def test_contours_emu_noisy_many_small_contours(emu):
    grid, pv, sm = make_case(46, 90, 1, passes=1, seed_noise=3)
    cs = run_contours(sm, [2, -2], grid, 120)
    want = P.calculate_contours(sm, [2, -2], grid, 120, original_coordinates=False)
    assert len(want) >= 8
    compare_contours(cs, want, grid, [2, -2])


def test_contours_emu_no_periodic_add_and_raw_f32(emu):
    grid, pv, sm = make_case(31, 60, 1)
    cs = run_contours(pv, [2], grid, 0)  # unsmoothed float32 field, no extension
    want = P.calculate_contours(pv, [2], grid, 0, original_coordinates=False)
    compare_contours(cs, want, grid, [2])


def test_contours_emu_capacity_regrow(emu):
    grid, pv, sm = make_case(31, 60, 1, passes=0, seed_noise=5)
    detect.clear_contexts()
    field = spatial.to_device(np.nan_to_num(sm))
    cs = detect.contours(field, [2.0], 20, caps=None)
    want = P.calculate_contours(np.nan_to_num(sm), [2.0], grid, 120, original_coordinates=False)
    compare_contours(cs, want, grid, [2.0])
    detect.clear_contexts()


def test_lattice_vertex_flag_emu(emu):
    grid, pv, sm = make_case(31, 60, 1)
    f = np.nan_to_num(sm, nan=0.0)
    f[0, 10, 10] = 2.0  # a grid value exactly at the level
    f[0, 10, 11] = 3.0
    cs = run_contours(f, [2.0], grid, 120)
    assert SK.has_lattice_vertex_points(np.concatenate([f[0], f[0][:, :20]], axis=1), 2.0)
    assert cs.status[0] & 4


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("shape,levels", [((181, 360), [2, -2]), ((721, 1440), [2])])
def test_contours_gpu(gpu, shape, levels):
    grid, pv, sm = make_case(shape[0], shape[1], 2 if shape[0] < 700 else 1)
    cs = run_contours(sm, levels, grid, 120)
    want = P.calculate_contours(sm, levels, grid, 120, original_coordinates=False)
    compare_contours(cs, want, grid, levels)


@pytest.mark.gpu
def test_contours_gpu_noisy(gpu):
    grid, pv, sm = make_case(181, 360, 3, passes=1, seed_noise=7)
    cs = run_contours(sm, [2, -2], grid, 120)
    want = P.calculate_contours(sm, [2, -2], grid, 120, original_coordinates=False)
    assert len(want) > 50
    compare_contours(cs, want, grid, [2, -2])


# ------------------------------------------------------------------ contour vertices exactly on grid vertices
def _int_grid(nlat, nlon):
    lat = np.linspace(-60.0, 60.0, nlat)
    lon = np.arange(nlon) * (360.0 / nlon)
    return P.Grid(lon, lat, synthetic.time_axis(1, 6))


def test_lattice_vertices_skimage_vector_emu(emu):
    """skimage's test_float field: the level passes exactly through four grid vertices."""
    x, y = np.mgrid[-1:1:5j, -1:1:5j]
    r = np.sqrt(x ** 2 + y ** 2)[None]
    grid = _int_grid(5, 5)
    cs = detect.contours(spatial.to_device(r), [0.5], 0)
    want = P.calculate_contours(r, [0.5], grid, 0, original_coordinates=False)
    assert cs.status[0] & 4
    compare_contours(cs, want, grid, [0.5])
    assert len(want) == 1 and cs.contour_points(0).tolist() == [[3, 2], [2, 1], [1, 2], [2, 3]]


@pytest.mark.parametrize("seed", range(6))
def test_lattice_vertices_random_integer_fields_emu(emu, seed):
    """Integer-valued fields with an integer level: many vertices on grid points, saddles and touching rings."""
    rng = np.random.default_rng(seed)
    nlat, nlon = 14, 24
    f = rng.integers(0, 4, size=(1, nlat, nlon)).astype(np.float64)
    # smooth a little so that longer contours exist, keep many exact hits of the level
    f = np.round((f + np.roll(f, 1, 2) + np.roll(f, 1, 1)) / 3.0 * 2.0) / 2.0
    grid = _int_grid(nlat, nlon)
    for level, add_deg in ((1.0, 0), (1.5, 0), (2.0, 120)):
        cs = detect.contours(spatial.to_device(f), [level], int(add_deg / grid.dlon))
        want = P.calculate_contours(f, [level], grid, add_deg, original_coordinates=False)
        compare_contours(cs, want, grid, [level])


@pytest.mark.gpu
def test_lattice_vertices_gpu(gpu):
    rng = np.random.default_rng(11)
    f = rng.integers(0, 4, size=(2, 40, 72)).astype(np.float64)
    f = np.round((f + np.roll(f, 1, 2) + np.roll(f, 1, 1)) / 3.0 * 2.0) / 2.0
    lat = np.linspace(-60.0, 60.0, 40)
    lon = np.arange(72) * 5.0
    grid = P.Grid(lon, lat, synthetic.time_axis(2, 6))
    cs = detect.contours(spatial.to_device(f), [1.0, 2.0], 24)
    want = P.calculate_contours(f, [1.0, 2.0], grid, 120, original_coordinates=False)
    assert np.all(cs.status & 4)
    compare_contours(cs, want, grid, [1.0, 2.0])
