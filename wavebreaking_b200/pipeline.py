"""Batched per-time-step detection pipeline (the benchmarked hot path).

One ``Detector.run_batch`` call takes ``T`` time steps of a raw field and produces what the
reference produces with
``calculate_smoothed_field`` -> ``calculate_contours`` -> ``calculate_streamers`` /
``calculate_overturnings`` / ``calculate_cutoffs`` -> ``to_xarray`` (x3):
columnar event tables (host) and the three int8 flag grids.  All arithmetic runs in libwbk's
CUDA kernels (including the meridian split of the events that straddle the date line); the host only
sizes buffers (two small device->host count reads per batch) and receives the tables.
"""

from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib, detect, spatial


@dataclass
class BatchResult:
    ntime: int
    contours: detect.ContourSet
    tables: dict              # kind -> detect.EventTable
    flags: torch.Tensor       # int8 [3, ntime, nlat, nlon] on the device (or pinned host if fetched)
    gmax_nx: int
    n_split: int = 0


class Detector:
    """Holds the grid, thresholds and reusable device buffers of one detection configuration."""

    def __init__(self, lat, lon, levels=(2.0,), periodic_add=120, passes=5, which=detect.KINDS, geo_dis=800.0,
                 cont_dis=1500.0, range_group=5.0, ot_min_exp=5.0, co_min_exp=5.0, want_flags=True):
        self.lib = _lib.get()
        self.lat = np.asarray(lat, dtype=np.float64)
        self.lon = np.asarray(lon, dtype=np.float64)
        self.nlat, self.nlon = len(self.lat), len(self.lon)
        self.dlon = float(abs(self.lon[1] - self.lon[0]))
        self.dlat = float(abs(self.lat[1] - self.lat[0]))
        self.add = int(periodic_add / self.dlon)
        self.levels = np.atleast_1d(np.asarray(levels, dtype=np.float64))
        self.passes = int(passes)
        self.which = tuple(which)
        self.params = dict(geo_dis=geo_dis, cont_dis=cont_dis, range_group=range_group, ot_min_exp=ot_min_exp,
                           co_min_exp=co_min_exp)
        self.want_flags = want_flags
        self.coords = detect.coord_tables(self.lat, self.lon, self.dlon, self.dlat, self.lib)

    # ------------------------------------------------------------------ device-resident input
    def run_batch(self, raw, gmax_nx=None, smoothed=None):
        """raw: device tensor [T, nlat, nlon] (float32 / float64).  Returns a BatchResult."""
        sm = smoothed if smoothed is not None else (spatial.smooth(raw, self.passes) if self.passes > 0 else raw)
        cs = detect.contours(sm, self.levels, self.add)
        g = cs.max_nx if gmax_nx is None else max(int(gmax_nx), cs.max_nx)
        tables, flags = detect.run_indices(cs, sm, self.coords, self.dlon, self.dlat, which=self.which, gmax_nx=g,
                                           want_flags=self.want_flags, **self.params)
        n_split = int(sum(int((t.split == 1).sum()) for t in tables.values()))
        return BatchResult(ntime=int(sm.shape[0]), contours=cs, tables=tables, flags=flags, gmax_nx=g, n_split=n_split)

    # ------------------------------------------------------------------ host input (end to end)
    def run_batch_host(self, raw_host, flags_host=None):
        """raw_host: pinned host tensor [T, nlat, nlon]; copies in, runs, copies the flag grids out."""
        raw = raw_host.to(self.lib.device, non_blocking=True)
        res = self.run_batch(raw)
        if self.want_flags:
            if flags_host is None:
                flags_host = torch.empty(res.flags.shape, dtype=torch.int8, pin_memory=self.lib.is_cuda)
            flags_host.copy_(res.flags, non_blocking=True)
            if self.lib.is_cuda:
                torch.cuda.current_stream().synchronize()
            res.flags = flags_host
        return res


def summarize(res):
    """Small dict of counts (used by the benchmark's result read-back and the tests)."""
    return dict(
        ntime=res.ntime, contours=res.contours.ncontours, points=res.contours.npoints,
        streamers=len(res.tables["streamers"]), overturnings=len(res.tables["overturnings"]),
        cutoffs=len(res.tables["cutoffs"]), split=res.n_split,
    )
