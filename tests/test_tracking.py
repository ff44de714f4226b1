"""track_events: candidate pairs, the exact lattice-polygon overlap test, distances, labels, and the sharded run
with an event-table halo (processing/events.py:113-241)."""

import os
import socket
import sys

import numpy as np
import pandas as pd
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import geom
from oracle import pipeline as P
from wavebreaking_b200 import compat, tracking

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ generators
def _rect(rng, span=8):
    x0, y0 = rng.integers(0, span, 2)
    w, h = rng.integers(1, 5, 2)
    r = np.array([[x0, y0], [x0 + w, y0], [x0 + w, y0 + h], [x0, y0 + h]])
    return r if rng.random() < 0.5 else r[::-1].copy()


def _star(rng, span=10):
    cx, cy = rng.integers(2, span - 2, 2)
    k = rng.integers(3, 9)
    ang = np.sort(rng.uniform(0, 2 * np.pi, k))
    rad = rng.uniform(1, 4, k)
    pts = np.clip(np.c_[np.rint(cx + rad * np.cos(ang)), np.rint(cy + rad * np.sin(ang))].astype(int), 0, span)
    keep = [0]
    for i in range(1, len(pts)):
        if (pts[i] != pts[keep[-1]]).any():
            keep.append(i)
    pts = pts[keep]
    return pts[:-1] if len(pts) > 1 and (pts[0] == pts[-1]).all() else pts


def random_lattice_polygons(seed, n=90, multi=8):
    """Simple lattice polygons on a small lattice, so that shared vertices, shared edges, nesting, identical
    shapes and collinear contacts are frequent; plus a few two-part multipolygons."""
    rng = np.random.default_rng(seed)
    polys = []
    while len(polys) < n:
        p = _rect(rng) if rng.random() < 0.5 else _star(rng)
        if len(p) < 3 or geom._ring_area2([(int(x), int(y)) for x, y in p]) == 0 or not geom.ring_is_simple(p):
            continue
        polys.append([p])
    for _ in range(multi):
        polys.append([_rect(rng), _rect(rng) + [20, 0]])
    return polys


def soup_of(polys):
    xy = np.concatenate([r for p in polys for r in p]).astype(np.int32)
    ring_len = [len(r) for p in polys for r in p]
    return tracking.PolygonSoup(xy, np.r_[0, np.cumsum(ring_len)], np.r_[0, np.cumsum([len(p) for p in polys])], True)


def _check_exact(seed):
    polys = random_lattice_polygons(seed)
    n = len(polys)
    pairs = np.array([(i, j) for i in range(n) for j in range(i + 1, n)], dtype=np.int32)
    got = tracking.overlap_exact(soup_of(polys), pairs)
    want = np.array([geom.overlap_positive_exact(polys[i], polys[j]) for i, j in pairs])
    assert np.array_equal((got & 1).astype(bool), want)
    return int(((got & 2) > 0).sum()), int(want.sum()), len(pairs)


@pytest.mark.parametrize("seed", [0, 1])
def test_overlap_exact_matches_rational_oracle_emu(emu, seed):
    touching, positive, total = _check_exact(seed)
    assert touching > 200 and positive > 500  # the degenerate classes are well represented


def test_overlap_exact_known_cases_emu(emu):
    sq = lambda x, y, s: np.array([[x, y], [x + s, y], [x + s, y + s], [x, y + s]])
    tri = np.array([[0, 0], [4, 0], [0, 4]])
    polys = [[sq(0, 0, 4)], [sq(2, 2, 4)], [sq(4, 0, 4)], [sq(4, 4, 4)], [sq(1, 1, 2)], [sq(0, 0, 4)[::-1].copy()],
             [sq(5, 0, 4)], [tri], [np.array([[4, 4], [4, 0], [0, 4]])], [sq(0, 0, 2), sq(10, 10, 2)], [sq(11, 11, 3)]]
    soup = soup_of(polys)
    cases = {(0, 1): 1, (0, 2): 0, (0, 3): 0, (0, 4): 1, (0, 5): 1, (0, 6): 0, (7, 8): 0, (0, 7): 1, (9, 10): 1, (9, 3): 0,
             (2, 6): 1, (4, 7): 1}
    pairs = np.array(list(cases), dtype=np.int32)
    got = tracking.overlap_exact(soup, pairs) & 1
    assert got.tolist() == list(cases.values())


def _event_table(seed, nsteps=10, per_step=7, hours=1):
    rng = np.random.default_rng(seed)
    polys = random_lattice_polygons(seed + 100, n=nsteps * per_step, multi=0)
    t0 = np.datetime64("2000-01-01T00", "ns")
    dates = np.array([t0 + (k // per_step) * np.timedelta64(hours, "h") for k in range(len(polys))])
    com = np.array([[float(p[0][:, 0].mean()) * 3.0, float(p[0][:, 1].mean()) * 2.0] for p in polys])
    return dates, polys, com


def _oracle_labels(dates, polys, com, **kw):
    ev = pd.DataFrame({"date": pd.to_datetime(dates), "geometry": polys, "com": list(map(tuple, com))})
    out = P.track_events(ev, **kw)
    return out.label.sort_index().values


def test_track_columnar_matches_oracle_emu(emu):
    dates, polys, com = _event_table(3)
    memo = {}
    for overlap in (0, 0.2):
        got, _ = tracking.track_columnar(dates, "by_overlap", soup=soup_of(polys), overlap=overlap)
        assert np.array_equal(got, _oracle_labels(dates, polys, com, method="by_overlap", overlap=overlap, _memo=memo))
    got, _ = tracking.track_columnar(dates, "by_distance", com=com, distance=700)
    assert np.array_equal(got, _oracle_labels(dates, polys, com, method="by_distance", distance=700))
    # longer range, unsorted input
    perm = np.random.default_rng(0).permutation(len(dates))
    got, _ = tracking.track_columnar(dates[perm], "by_overlap", soup=soup_of([polys[i] for i in perm]), time_range=3)
    want = _oracle_labels(dates[perm], [polys[i] for i in perm], com[perm], method="by_overlap", time_range=3)
    assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        tracking.track_columnar(dates, "by_overlap", soup=soup_of(polys), time_range=0.5)


def test_track_events_frame_and_errors_emu(emu):
    import wavebreaking_b200 as wb

    dates, polys, com = _event_table(5, nsteps=4, per_step=5, hours=6)
    geoms = [compat.Polygon(p[0] * 0.25 - 30.0) for p in polys]  # lon / lat values on a 0.25 degree lattice
    ev = compat.make_frame({"date": list(dates), "com": list(map(tuple, com))}, geoms)
    got = wb.track_events(ev)
    want = _oracle_labels(dates, polys, com, method="by_overlap")
    assert np.array_equal(got.label.sort_index().values, want)
    assert list(got.label) == sorted(got.label)
    with pytest.raises(ValueError, match="not supported as method"):
        wb.track_events(ev, method="nope")
    with pytest.raises(ValueError, match="No events detected"):
        wb.track_events(ev, time_range=1)


# ------------------------------------------------------------------ sharded (gloo)
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_worker(rank, world, port, emu_path, out_dir, nsteps, time_range, method):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wavebreaking_b200 import _lib, sharding

    _lib.use_library(emu_path, "cpu")
    dates, polys, com = _event_table(7, nsteps=nsteps, per_step=6)
    step = ((dates - dates[0]) / np.timedelta64(1, "h")).astype(int)
    t0, t1 = sharding.shard_range(nsteps, rank, world)
    sel = np.nonzero((step >= t0) & (step < t1))[0]
    stats = {}
    labels = sharding.track_sharded(dates[sel], method, soup=soup_of([polys[i] for i in sel]) if len(sel) else
                                    tracking.PolygonSoup(np.zeros((0, 2), np.int32), [0], [0], True),
                                    com=com[sel], time_range=time_range, distance=700, stats=stats)
    np.save(os.path.join(out_dir, "labels_{}.npy".format(rank)), labels)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nsteps,time_range,method", [(2, 8, None, "by_overlap"), (4, 10, 3, "by_overlap"),
                                                             (3, 7, 2, "by_distance")])
def test_track_sharded_equals_single_process(emu, emu_lib, tmp_path, world, nsteps, time_range, method):
    """labels of 2 / 3 / 4 time-sharded ranks (halo of the event table; with time_range = 3 hours and shards of 2-3
    steps the halo reaches two shards ahead) == labels of the single-process run"""
    port = _free_port()
    mp.spawn(_shard_worker, args=(world, port, emu_lib, str(tmp_path), nsteps, time_range, method), nprocs=world, join=True)
    dates, polys, com = _event_table(7, nsteps=nsteps, per_step=6)
    want, _ = tracking.track_columnar(dates, method, soup=soup_of(polys), com=com, time_range=time_range, distance=700)
    got = np.concatenate([np.load(tmp_path / "labels_{}.npy".format(r)) for r in range(world)])
    assert np.array_equal(got, want)
    assert len(np.unique(want)) < len(want)  # some events were linked across steps


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_overlap_exact_gpu(gpu):
    for seed in (0, 1, 2):
        _check_exact(seed)


@pytest.mark.gpu
def test_track_real_streamers_gpu(gpu):
    """48 consecutive hourly steps of the 721 x 1440 workload: streamers of the CUDA pipeline, tracked by the CUDA
    path (exact overlap test, areas, distances) and by the oracle on the same polygons"""
    from wavebreaking_b200 import detect, geometry, pipeline, spatial, synthetic

    nlat, nlon, nt = 721, 1440, 48
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = spatial.synth_pv(nt, nlat, nlon, hour0=500.0, hour_step=1.0)
    det = pipeline.Detector(lat, lon, levels=[2.0])
    res = det.run_batch(raw)
    tab = res.tables["streamers"]
    soup, com = pipeline.events_soup(res, "streamers", det)
    props = detect.finish_properties(tab, lon, lat, nlon)
    dates = synthetic.time_axis(nt, 1)[tab.job]
    polys = [[soup.xy[soup.ring_off[r]:soup.ring_off[r + 1]].astype(np.int64) for r in range(soup.poly_off[e], soup.poly_off[e + 1])]
             for e in range(len(soup))]
    assert len(polys) == len(tab) > 400
    memo, stats = {}, {}
    for overlap in (0, 0.2):
        got, near = tracking.track_columnar(dates, "by_overlap", soup=soup, overlap=overlap, stats=stats)
        want = _oracle_labels(dates, polys, np.asarray(props["com"]), method="by_overlap", overlap=overlap, _memo=memo)
        assert np.array_equal(got, want), overlap
        assert len(near) == 0
    assert len(np.unique(got)) < len(got) / 3  # events persist over many hours
    com = np.asarray(props["com"], dtype=np.float64)
    got, near = tracking.track_columnar(dates, "by_distance", com=com, distance=400)
    assert np.array_equal(got, _oracle_labels(dates, polys, com, method="by_distance", distance=400))
