"""Multi-rank host logic with gloo (world_size 2, CPU): sharding, global exp_lon max, id offsets, gathers."""

import os
import socket
import sys

import numpy as np
import pandas as pd
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, emu_path, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wavebreaking_b200 import _lib, detect, pipeline, sharding, spatial, synthetic

    _lib.use_library(emu_path, "cpu")
    nlat, nlon, ntime = 46, 90, 5
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = synthetic.pv_field(nlat, nlon, np.arange(ntime) * 6.0)
    det = pipeline.Detector(lat, lon, levels=[2.0])
    t0, t1, res = sharding.run_sharded(det, spatial.to_device(raw))
    assert (t0, t1) == sharding.shard_range(ntime, rank, world)
    # global ids by exclusive scan of the per-rank counts
    tab = res.tables["cutoffs"]
    off, total = sharding.exclusive_offset(len(tab))
    frame = pd.DataFrame({"id": off + np.arange(len(tab)), "t": t0 + tab.job, "area": tab.sums[:, 0]})
    merged = sharding.gather_frames(frame)
    flags = res.flags.numpy()
    np.save(os.path.join(out_dir, "flags_{}.npy".format(rank)), flags)
    if rank == 0:
        merged.to_pickle(os.path.join(out_dir, "merged.pkl"))
        assert list(merged.id) == list(range(total))
    assert sharding.global_max(res.gmax_nx) >= res.gmax_nx
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_axis():
    from wavebreaking_b200 import sharding

    for n in (1, 5, 8, 8760):
        for w in (1, 2, 3, 8):
            blocks = [sharding.shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))


def test_two_rank_gloo_matches_single_process(emu_lib, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, emu_lib, str(tmp_path)), nprocs=2, join=True)
    # single-process reference run
    from wavebreaking_b200 import _lib, pipeline, spatial, synthetic

    prev = _lib._LIB
    _lib.use_library(emu_lib, "cpu")
    try:
        lat, lon = synthetic.grid_coords(46, 90)
        raw = synthetic.pv_field(46, 90, np.arange(5) * 6.0)
        res = pipeline.Detector(lat, lon, levels=[2.0]).run_batch(spatial.to_device(raw))
    finally:
        _lib._LIB = prev
    flags = np.concatenate([np.load(tmp_path / "flags_0.npy"), np.load(tmp_path / "flags_1.npy")], axis=1)
    assert np.array_equal(flags, res.flags.numpy())
    merged = pd.read_pickle(tmp_path / "merged.pkl")
    tab = res.tables["cutoffs"]
    assert len(merged) == len(tab)
    assert np.array_equal(merged.t.values, tab.job) and np.array_equal(merged.area.values, tab.sums[:, 0])
