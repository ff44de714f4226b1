"""File -> pinned host -> device streaming gives the same results as one resident batch."""

import numpy as np
import pytest

from wavebreaking_b200 import detect, io, pipeline, spatial, synthetic


def _check(tmp_path, nlat, nlon, ntime, batch, raw_binary):
    lat, lon = synthetic.grid_coords(nlat, nlon)
    raw = synthetic.pv_field(nlat, nlon, np.arange(ntime) * 6.0)
    if raw_binary:
        path = tmp_path / "pv.bin"
        raw.tofile(path)
        kw = dict(shape=raw.shape, dtype=np.float32)
    else:
        path = tmp_path / "pv.npy"
        np.save(path, raw)
        kw = {}
    whole = pipeline.Detector(lat, lon, levels=[2.0]).run_batch(spatial.to_device(raw))
    gmax = int(whole.contours.host()["nx"].max())
    seen = 0
    for t0, res in io.detect_file(str(path), lat, lon, batch=batch, depth=2, levels=[2.0], **kw):
        assert t0 == seen
        nt = res.ntime
        assert int(res.contours.host()["nx"].max()) <= gmax
        assert np.array_equal(res.flags.cpu().numpy(), whole.flags[:, t0:t0 + nt].cpu().numpy())
        for kind in detect.KINDS:
            sel = (whole.tables[kind].job >= t0) & (whole.tables[kind].job < t0 + nt)
            assert np.array_equal(whole.tables[kind].sums[sel], res.tables[kind].sums)
        seen += nt
    assert seen == ntime


def test_detect_file_npy_emu(emu, tmp_path):
    _check(tmp_path, 46, 90, 5, 2, raw_binary=False)


def test_detect_file_raw_binary_emu(emu, tmp_path):
    _check(tmp_path, 46, 90, 3, 2, raw_binary=True)


def test_open_field_rejects_bad_input(tmp_path):
    np.save(tmp_path / "bad.npy", np.zeros((4, 5), dtype=np.float32))
    with pytest.raises(ValueError, match="time, lat, lon"):
        io.open_field(str(tmp_path / "bad.npy"))
    np.save(tmp_path / "int.npy", np.zeros((2, 4, 5), dtype=np.int32))
    with pytest.raises(TypeError, match="float32 or float64"):
        io.open_field(str(tmp_path / "int.npy"))
    with pytest.raises(ValueError, match="shape is required"):
        io.open_field(str(tmp_path / "x.bin"))


@pytest.mark.gpu
def test_detect_file_gpu(gpu, tmp_path):
    _check(tmp_path, 181, 360, 7, 3, raw_binary=False)
