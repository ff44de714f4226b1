"""Batched per-time-step detection pipeline (the benchmarked hot path).

One batch = ``T`` time steps of a raw field through what the reference does with
``calculate_smoothed_field`` -> ``calculate_contours`` -> ``calculate_streamers`` /
``calculate_overturnings`` / ``calculate_cutoffs`` -> ``to_xarray`` (x3):
columnar event tables (host) and the three int8 flag grids.  All arithmetic runs in libwbk's
CUDA kernels (including the meridian split of the events that straddle the date line).

A batch is *enqueued without any host synchronisation* (``submit``): smoothing, marching squares, linking,
device-side packing, the three indices, rasterisation and one gather of all results, then the asynchronous
device->host copies.  ``collect`` waits once and slices the tables.  Several slots (each with its own
context, buffers and CUDA stream) keep the GPU busy while the host handles the previous batch and let
host<->device copies overlap with kernels (``Detector.stream``).
"""

import ctypes
import logging
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, detect, spatial

logger = logging.getLogger(__name__)


@dataclass
class BatchResult:
    ntime: int
    contours: detect.ContourSet
    tables: dict              # kind -> detect.EventTable
    flags: torch.Tensor       # int8 [3, ntime, nlat, nlon] (device, or pinned host in the end-to-end path)
    gmax_nx: int
    n_split: int = 0
    # streamer pairs decided within 1e-9 of geo_dis / cont_dis, kept AND rejected ones (wbk_near_list):
    # int32 [m, 6] = time step, level index, contour, i, j, flags (1 kept, 2 geo in band, 4 cont in band)
    near: np.ndarray = None
    near_total: int = 0
    pieces: dict = None  # device-clipped pieces of the straddling events: ev (row in kind-major order), off, xy
    counts: tuple = None  # events per kind (streamers, overturnings, cutoffs)
    work: dict = None  # device work counters of the batch: marching-squares segments, candidate pairs, scan tiles
    flags_packed: torch.Tensor = None  # pinned uint8: the three grids, one bit per cell (unpack_flags restores them)


def packed_nbytes(ncells):
    """bytes of the bit-packed form of ``ncells`` flag cells (whole 32-bit words)."""
    return 4 * ((int(ncells) + 31) // 32)


def unpack_flags(packed, nt, nlat, nlon):
    """int8 [3, nt, nlat, nlon] flag grids from the bit-packed download (host numpy)."""
    ncells = 3 * nt * nlat * nlon
    buf = packed.numpy() if isinstance(packed, torch.Tensor) else np.asarray(packed)
    bits = np.unpackbits(buf[:packed_nbytes(ncells)], bitorder="little")[:ncells]
    return bits.view(np.int8).reshape(3, nt, nlat, nlon)


def _pinned(shape, dtype, lib):
    return torch.empty(shape, dtype=dtype, pin_memory=lib.is_cuda)


class _Slot:
    """Context + device / pinned buffers + stream for one batch in flight."""

    def __init__(self, det, T, caps=None):
        lib = det.lib
        self.det, self.T = det, int(T)
        dev = lib.device
        L = len(det.levels)
        J = self.T * L
        self.J = J
        want = detect.default_caps(det.nlat, det.nlon, det.add, J)
        if caps:
            for k, v in caps.items():
                if k in want:  # (cap_c / cap_p / cap_e / cap_r size the slot's own buffers, below)
                    want[k] = max(want[k], int(v))
        self.ctx = detect.Context(det.nlat, det.nlon, det.add, want)
        self.cap_c = int(max(64 * J, 1024))
        self.cap_p = int(max(min(want["seg_cap"], 16384) * J, 1 << 16))
        self.cap_e = int(max(48 * J, 1024))
        self.cap_r = int(max(8192 * J, 1 << 16))
        if caps:
            self.cap_c = max(self.cap_c, int(caps.get("cap_c", 0)))
            self.cap_p = max(self.cap_p, int(caps.get("cap_p", 0)))
            self.cap_e = max(self.cap_e, int(caps.get("cap_e", 0)))
            self.cap_r = max(self.cap_r, int(caps.get("cap_r", 0)))
        shape = (self.T, det.nlat, det.nlon)
        i32, f64 = torch.int32, torch.float64
        self.sm = torch.empty(shape, dtype=f64, device=dev) if det.passes > 0 else None
        self.sm_f32 = None
        self.sm_tmp = torch.empty(shape, dtype=f64, device=dev) if det.passes > _lib.SMOOTH_MAX_FUSED else None
        self.raw_dev = None  # allocated on first host submit
        self.flipped = None
        self.job_off = torch.empty(J + 1, dtype=i32, device=dev)
        self.pt_off = torch.empty(self.cap_c + 1, dtype=i32, device=dev)
        self.meta = torch.empty((self.cap_c, 4), dtype=i32, device=dev)
        self.pts = torch.empty(self.cap_p, dtype=i32, device=dev)
        self.work = torch.empty(2 * self.cap_p, dtype=f64, device=dev)
        self.flags = None
        self.ev_int = torch.empty((self.cap_e, _lib.EV_INTS), dtype=i32, device=dev)
        self.ev_f64 = torch.empty((self.cap_e, _lib.EV_F64), dtype=f64, device=dev)
        self.ev_job = torch.empty(self.cap_e, dtype=i32, device=dev)
        self.ring_off = torch.empty(self.cap_e + 1, dtype=i32, device=dev)
        self.ring_pts = torch.empty(self.cap_r, dtype=i32, device=dev)
        self.summary = torch.zeros(16, dtype=i32, device=dev)
        self.h_summary = _pinned(16, i32, lib)
        # pieces of the events that straddle the last meridian (device clipper output, wbk_split_fetch)
        self.cap_sr = int(max(16 * J, 1024))
        self.cap_sv = int(max(4096 * J, 1 << 16))
        if caps:
            self.cap_sr = max(self.cap_sr, int(caps.get("cap_sr", 0)))
            self.cap_sv = max(self.cap_sv, int(caps.get("cap_sv", 0)))
        self.sp_ev = torch.zeros(self.cap_sr, dtype=i32, device=dev)
        self.sp_off = torch.zeros(self.cap_sr + 1, dtype=i32, device=dev)
        self.sp_xy = torch.zeros((self.cap_sv, 2), dtype=i32, device=dev)
        self.sp_cnt = torch.zeros(4, dtype=i32, device=dev)
        self.h_sp_ev = _pinned(self.cap_sr, i32, lib)
        self.h_sp_off = _pinned(self.cap_sr + 1, i32, lib)
        self.h_sp_xy = _pinned((self.cap_sv, 2), i32, lib)
        self.h_sp_cnt = _pinned(4, i32, lib)
        self.near = torch.zeros((_lib.NEAR_CAP, 4), dtype=i32, device=dev)
        self.near_cnt = torch.zeros(1, dtype=i32, device=dev)
        self.h_near = _pinned((_lib.NEAR_CAP, 4), i32, lib)
        self.h_near_cnt = _pinned(1, i32, lib)
        self.h_ev_int = _pinned((self.cap_e, _lib.EV_INTS), i32, lib)
        self.h_ev_f64 = _pinned((self.cap_e, _lib.EV_F64), f64, lib)
        self.h_ev_job = _pinned(self.cap_e, i32, lib)
        self.h_ring_off = _pinned(self.cap_e + 1, i32, lib)
        self.h_ring_pts = _pinned(self.cap_r, i32, lib)
        self.stream = torch.cuda.Stream(device=dev) if lib.is_cuda else None
        self.done = torch.cuda.Event() if lib.is_cuda else None
        self.pending = None   # (raw, ntime, flags_dev, flags_host, gmax)
        self.grow_seen = 0
        self.packed = None    # bit-packed flag grids (device), allocated on first use
        self.graphs = {}      # buffers key -> (torch.cuda.CUDAGraph, pending record)
        self.graph_seen = set()

    def close(self):
        self.graphs.clear()
        self.ctx.close()


class Detector:
    """Holds the grid, thresholds and the reusable slots of one detection configuration."""

    def __init__(self, lat, lon, levels=(2.0,), periodic_add=120, passes=5, which=detect.KINDS, geo_dis=800.0,
                 cont_dis=1500.0, range_group=5.0, ot_min_exp=5.0, co_min_exp=5.0, want_flags=True, fuse=True,
                 packing=None, graphs=False, nvtx=True, want_pieces=False):
        self.lib = _lib.get()
        self.lat = np.asarray(lat, dtype=np.float64)
        self.lon = np.asarray(lon, dtype=np.float64)
        # descending coordinates (ERA5 latitudes): batches are flipped on the device after the upload
        # (utils/data_utils.py:196-213); all results refer to the ascending orientation
        self.flip_lat = bool(len(self.lat) > 1 and self.lat[0] > self.lat[-1])
        self.flip_lon = bool(len(self.lon) > 1 and self.lon[0] > self.lon[-1])
        if self.flip_lat:
            self.lat = self.lat[::-1].copy()
        if self.flip_lon:
            self.lon = self.lon[::-1].copy()
        self.nlat, self.nlon = len(self.lat), len(self.lon)
        self.dlon = float(abs(self.lon[1] - self.lon[0]))
        self.dlat = float(abs(self.lat[1] - self.lat[0]))
        self.add = int(periodic_add / self.dlon)
        self.levels = np.ascontiguousarray(np.atleast_1d(np.asarray(levels, dtype=np.float64)))
        self.passes = int(passes)
        self.which = tuple(which)
        self.params = dict(geo_dis=geo_dis, cont_dis=cont_dis, range_group=range_group, ot_min_exp=ot_min_exp,
                           co_min_exp=co_min_exp)
        self.want_flags = want_flags
        # smoothing that also leaves the marching-squares bit planes (wbk_smooth_contours): the contour stage then
        # does not re-read the smoothed field.  fuse=False runs wbk_smooth + wbk_contours (same results).
        self.fuse = bool(fuse)
        # CF packing of int16 input (scale_factor, add_offset, _FillValue or None), decoded in the smoothing loads
        self.packing = tuple(packing) if packing is not None else (1.0, 0.0, None)
        # also download the device clipper's pieces of the events that straddle the last meridian (events_soup then
        # needs no host-side polygon clipping; used by the tracking leg)
        self.want_pieces = bool(want_pieces)
        self.graphs = bool(graphs)  # replay one captured CUDA graph per batch (see submit)
        self.graph_replays = 0
        self.graph_kernel_launches = 0
        self.nvtx = bool(nvtx)      # NVTX ranges per stage (upload / smooth+contours / indices / properties / download)
        self.coords = detect.coord_tables(self.lat, self.lon, self.dlon, self.dlat, self.lib)
        self._slots = {}
        self._grow = {}
        self._grow_version = 0

    # ------------------------------------------------------------------ slots
    def _slot(self, T, index=0):
        key = (int(T), int(index))
        s = self._slots.get(key)
        if s is not None and s.pending is None and s.grow_seen != self._grow_version:
            # another slot had to regrow its arenas: adopt the larger capacities before the next batch overflows too
            if self.lib.is_cuda:
                torch.cuda.synchronize()
            s.close()
            s = None
        if s is None:
            s = _Slot(self, T, self._grow)
            s.grow_seen = self._grow_version
            self._slots[key] = s
        return s

    def close(self):
        for s in self._slots.values():
            s.close()
        self._slots.clear()

    def _prm(self, gmax):
        p = self.params
        return _lib.IndexParams(
            do_streamers=int("streamers" in self.which), do_overturnings=int("overturnings" in self.which),
            do_cutoffs=int("cutoffs" in self.which), gmax_nx=int(gmax), dlon=self.dlon, dlat=self.dlat,
            geo_dis=float(p["geo_dis"]), cont_dis=float(p["cont_dis"]), range_group=float(p["range_group"]),
            ot_min_exp=float(p["ot_min_exp"]), co_min_exp=float(p["co_min_exp"]))

    # ------------------------------------------------------------------ enqueue / collect
    def _enqueue(self, slot, raw, flags_out, flags_host, gmax_nx, smoothed, intensity, packed_host):
        """All device work of one batch on the CURRENT stream; returns what ``collect`` needs.  Allocation-free in
        the steady state (so it can be captured into a CUDA graph)."""
        lib = self.lib
        nvtx = torch.cuda.nvtx if (lib.is_cuda and self.nvtx) else None
        nt = int(raw.shape[0])
        st = lib.stream()
        if nvtx:
            nvtx.range_push("wbk.upload")
        if raw.device != lib.device:  # host input: H2D on this slot's stream
            if slot.raw_dev is None or slot.raw_dev.dtype != raw.dtype:
                slot.raw_dev = torch.empty((slot.T, self.nlat, self.nlon), dtype=raw.dtype, device=lib.device)
            slot.raw_dev[:nt].copy_(raw, non_blocking=True)
            raw = slot.raw_dev[:nt]
        raw = raw.contiguous()
        raw_in = raw  # what the caller handed over (the regrow path re-submits exactly this)
        if intensity is not None:
            intensity = intensity.to(lib.device).contiguous()
        if nvtx:
            nvtx.range_pop()
            nvtx.range_push("wbk.smooth+contours")
        h = slot.ctx.handle
        L = len(self.levels)
        lv = self.levels.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        # the stored orientation (descending ERA5 latitudes, utils/data_utils.py:196-213) and the CF packing of
        # int16 input are resolved inside the smoothing loads: no re-oriented / decoded copy of the batch
        opts = _lib.smooth_opts(self.flip_lat, self.flip_lon, *self.packing) if (
            self.flip_lat or self.flip_lon or raw.dtype == torch.int16) else None
        fused = False
        if smoothed is not None:
            sm = smoothed
        elif self.passes > 0:
            f32 = raw.dtype == torch.float32
            if f32 and not spatial._numpy2():
                if slot.sm_f32 is None:
                    slot.sm_f32 = torch.empty((slot.T, self.nlat, self.nlon), dtype=torch.float32, device=lib.device)
                sm, rmode = slot.sm_f32[:nt], _lib.ROUND_ALL
            else:
                sm, rmode = slot.sm[:nt], (_lib.ROUND_FIRST if f32 else _lib.ROUND_NONE)
            if self.fuse and sm.dtype == torch.float64 and self.passes <= _lib.SMOOTH_MAX_FUSED and self.nlat >= 4:
                lib.call("wbk_smooth_contours", h, _lib.ptr(raw), _lib.dtype_code(raw.dtype), _lib.ptr(sm), nt,
                         self.passes, lv, L, opts, st)
                fused = True
            else:
                lib.call("wbk_smooth", _lib.ptr(raw), _lib.dtype_code(raw.dtype), _lib.ptr(sm),
                         _lib.dtype_code(sm.dtype), _lib.ptr(slot.sm_tmp), nt, self.nlat, self.nlon, self.passes,
                         rmode, opts, st)
        else:
            if opts is not None:  # no smoothing: orientation / decode only
                want = torch.float64 if raw.dtype == torch.int16 else raw.dtype
                if slot.flipped is None or slot.flipped.dtype != want:
                    slot.flipped = torch.empty((slot.T, self.nlat, self.nlon), dtype=want, device=lib.device)
                lib.call("wbk_orient", _lib.ptr(raw), _lib.dtype_code(raw.dtype), _lib.ptr(slot.flipped), nt,
                         self.nlat, self.nlon, opts, st)
                raw = slot.flipped[:nt]
            sm = raw
        if not fused:
            lib.call("wbk_contours", h, _lib.ptr(sm), _lib.dtype_code(sm.dtype), nt, lv, L, st)
        lib.call("wbk_contours_pack_auto", h, _lib.ptr(slot.job_off), _lib.ptr(slot.pt_off), _lib.ptr(slot.meta),
                 _lib.ptr(slot.pts), slot.cap_c, slot.cap_p, st)
        if nvtx:
            nvtx.range_pop()
            nvtx.range_push("wbk.indices")
        if intensity is not None and intensity.dtype != sm.dtype:
            intensity = intensity.to(sm.dtype)
        prm = self._prm(-1 if gmax_nx is None else gmax_nx)
        lib.call("wbk_index_run", h, nt * L, L, _lib.ptr(slot.job_off), _lib.ptr(slot.pt_off), _lib.ptr(slot.meta),
                 _lib.ptr(slot.pts), slot.cap_c, slot.cap_p, _lib.ptr(self.coords), _lib.ptr(slot.work),
                 ctypes.byref(prm), st)
        lib.call("wbk_near_list", h, _lib.ptr(slot.near), _lib.NEAR_CAP, _lib.ptr(slot.near_cnt), st)
        if nvtx:
            nvtx.range_pop()
            nvtx.range_push("wbk.properties+to_xarray")
        flags = None
        if self.want_flags:
            if flags_out is not None:
                flags = flags_out
            else:
                if slot.flags is None:
                    slot.flags = torch.empty((3, slot.T, self.nlat, self.nlon), dtype=torch.int8, device=lib.device)
                flags = slot.flags if nt == slot.T else torch.empty((3, nt, self.nlat, self.nlon), dtype=torch.int8,
                                                                    device=lib.device)
        lib.call("wbk_events_raster", h, _lib.ptr(slot.job_off), _lib.ptr(slot.pt_off), _lib.ptr(slot.pts),
                 _lib.ptr(self.coords), _lib.ptr(sm), _lib.dtype_code(sm.dtype),
                 None if intensity is None else _lib.ptr(intensity), nt, _lib.ptr(flags), ctypes.byref(prm), st)
        lib.call("wbk_batch_fetch", h, _lib.ptr(slot.pt_off), _lib.ptr(slot.pts), _lib.ptr(slot.ev_int),
                 _lib.ptr(slot.ev_f64), _lib.ptr(slot.ev_job), _lib.ptr(slot.ring_off), _lib.ptr(slot.ring_pts),
                 slot.cap_e, slot.cap_r, _lib.ptr(slot.summary), st)
        if flags is not None and self.want_pieces:
            lib.call("wbk_split_fetch", h, _lib.ptr(slot.sp_ev), _lib.ptr(slot.sp_off), _lib.ptr(slot.sp_xy), slot.cap_sr,
                     slot.cap_sv, _lib.ptr(slot.sp_cnt), st)
            slot.h_sp_cnt.copy_(slot.sp_cnt, non_blocking=True)
            slot.h_sp_ev.copy_(slot.sp_ev, non_blocking=True)
            slot.h_sp_off.copy_(slot.sp_off, non_blocking=True)
            slot.h_sp_xy.copy_(slot.sp_xy, non_blocking=True)
        if nvtx:
            nvtx.range_pop()
            nvtx.range_push("wbk.download")
        # asynchronous read-back of everything the host needs (tables are small; flags only on request)
        slot.h_summary.copy_(slot.summary, non_blocking=True)
        slot.h_near_cnt.copy_(slot.near_cnt, non_blocking=True)
        slot.h_near.copy_(slot.near, non_blocking=True)
        slot.h_ev_int.copy_(slot.ev_int, non_blocking=True)
        slot.h_ev_f64.copy_(slot.ev_f64, non_blocking=True)
        slot.h_ev_job.copy_(slot.ev_job, non_blocking=True)
        slot.h_ring_off.copy_(slot.ring_off, non_blocking=True)
        slot.h_ring_pts.copy_(slot.ring_pts, non_blocking=True)
        if flags_host is not None and flags is not None:
            flags_host[:, :nt].copy_(flags[:, :nt] if flags.shape[1] != nt else flags, non_blocking=True)
        if packed_host is not None and flags is not None:
            # one bit per cell on the wire (wbk_pack_flags); unpack_flags() restores the int8 grids on the host
            ncells = 3 * nt * self.nlat * self.nlon
            if slot.packed is None:
                slot.packed = torch.empty(packed_nbytes(3 * slot.T * self.nlat * self.nlon), dtype=torch.uint8,
                                          device=lib.device)
            fl = flags if flags.shape[1] == nt and flags.is_contiguous() else flags[:, :nt].contiguous()
            lib.call("wbk_pack_flags", _lib.ptr(fl), _lib.ptr(slot.packed), ncells, st)
            nb = packed_nbytes(ncells)
            packed_host[:nb].copy_(slot.packed[:nb], non_blocking=True)
        if nvtx:
            nvtx.range_pop()
        return dict(raw=raw_in, nt=nt, flags=flags, flags_host=flags_host, gmax=gmax_nx, smoothed=smoothed, sm=sm,
                    intensity=intensity, packed_host=packed_host)

    def submit(self, slot, raw, flags_out=None, flags_host=None, gmax_nx=None, smoothed=None, stream=None,
               intensity=None, packed_host=None):
        """Enqueue one batch on the slot's stream (or ``stream``).  ``raw``: device tensor or (pinned) host
        tensor [T' <= T, nlat, nlon].  No host synchronisation happens here.

        With ``Detector(graphs=True)`` the whole batch (uploads, ~25 kernels and memsets, downloads) is captured into
        ONE CUDA graph per (slot, buffers) the second time the same buffers are submitted and replayed afterwards: one
        launch per batch instead of ~40 driver calls, which is what limits several ranks sharing one host."""
        lib = self.lib
        nt = int(raw.shape[0])
        if nt > slot.T:
            raise ValueError("batch of {} time steps exceeds the slot size {}".format(nt, slot.T))
        use_stream = stream if stream is not None else slot.stream
        args = (slot, raw, flags_out, flags_host, gmax_nx, smoothed, intensity, packed_host)
        if use_stream is None:  # emulator
            slot.pending = self._enqueue(*args)
            return slot
        graphable = self.graphs and smoothed is None and intensity is None
        key = None
        if graphable:
            key = (raw.data_ptr(), raw.dtype, nt, None if flags_out is None else flags_out.data_ptr(),
                   None if flags_host is None else flags_host.data_ptr(),
                   None if packed_host is None else packed_host.data_ptr(), gmax_nx, use_stream.cuda_stream)
        with torch.cuda.stream(use_stream):
            use_stream.wait_stream(torch.cuda.default_stream(lib.device))
            if key is not None and key in slot.graphs:
                g, pend, nk = slot.graphs[key]
                g.replay()
                self.graph_replays += 1
                self.graph_kernel_launches += nk  # kernels inside the graph (the library only counts direct launches)
                slot.pending = pend
            elif key is not None and key in slot.graph_seen:
                # second submission of these buffers: everything is allocated, capture the batch
                use_stream.synchronize()
                g = torch.cuda.CUDAGraph()
                n0 = lib.cdll.wbk_launch_count()
                with torch.cuda.graph(g, stream=use_stream, capture_error_mode="relaxed"):
                    pend = self._enqueue(*args)
                nk = int(lib.cdll.wbk_launch_count() - n0)
                slot.graphs[key] = (g, pend, nk)
                g.replay()
                self.graph_replays += 1
                self.graph_kernel_launches += nk
                slot.pending = pend
            else:
                slot.pending = self._enqueue(*args)
                if key is not None:
                    slot.graph_seen.add(key)
            slot.done.record()
        return slot

    def collect(self, slot):
        """Wait for the slot's batch and return its BatchResult (re-runs it with larger arenas on overflow)."""
        pend = slot.pending
        if pend is None:
            raise RuntimeError("nothing was submitted on this slot")
        if slot.done is not None:
            slot.done.synchronize()
        s = slot.h_summary.numpy()
        C, P, ns, no, nc, status, max_nx, n_split = (int(v) for v in s[:8])
        work = dict(segments=int(s[8]), pairs=int(s[9]), tiles=int(s[11]))
        bad = status & (_lib.ST_SEG_OVERFLOW | _lib.ST_CONTOUR_OVERFLOW | _lib.ST_PAIR_OVERFLOW | _lib.ST_EVENT_OVERFLOW
                        | _lib.ST_SEL_OVERFLOW | _lib.ST_PACK_OVERFLOW | _lib.ST_FETCH_OVERFLOW)
        if status & _lib.ST_SPLIT_CHAINS:
            raise _lib.WbkError(_lib.ERR_INVALID, "an event crosses the last meridian too often for the device clipper; "
                                                  "its cells would be missing from the flag grids")
        if bad:
            return self._regrow_and_rerun(slot, status)
        slot.pending = None
        nt = pend["nt"]
        L = len(self.levels)
        counts = (ns, no, nc)
        tables, o = {}, 0
        ring_off = slot.h_ring_off.numpy()
        ring_pts = slot.h_ring_pts.numpy().view(np.uint32)
        for kind, n in zip(detect.KINDS, counts):
            ints = slot.h_ev_int.numpy()[o:o + n].copy()
            f64 = slot.h_ev_f64.numpy()[o:o + n].copy()
            job = slot.h_ev_job.numpy()[o:o + n].copy()
            rings = None
            if kind != "overturnings":
                ro = ring_off[o:o + n + 1].copy()
                rings = detect.LazyRings(ro, ring_pts[ro[0]:ro[-1]].copy())
            tab = detect.EventTable(kind=kind, job=job, contour=ints[:, 0].copy(), ind1=ints[:, 1].copy(),
                                    ind2=ints[:, 2].copy(), box=ints[:, 3:7].copy(), orientation=ints[:, 7].copy(),
                                    split=ints[:, 8].copy(), near=ints[:, 9].copy(), sums=f64)
            tab.rings = rings
            tables[kind] = tab
            o += n
        cs = detect.ContourSet(
            njobs=nt * L, nlevels=L, nlat=self.nlat, nlon=self.nlon, add=self.add, levels=self.levels,
            job_off=slot.job_off[:nt * L + 1], pt_off=slot.pt_off[:C + 1], meta=slot.meta[:C], pts=slot.pts[:P],
            status=np.full(nt * L, status, dtype=np.int32), max_nx=max_nx, h_ncontours=None, h_npoints=None)
        flags = pend["flags_host"] if pend["flags_host"] is not None else pend["flags"]
        packed = pend.get("packed_host")
        n_sp = int(sum(int((t.split == 1).sum()) for t in tables.values()))
        pieces = None
        if self.want_pieces and pend["flags"] is not None:
            npc, nvx, clip_over = (int(v) for v in slot.h_sp_cnt.numpy()[:3])
            if npc > slot.cap_sr or nvx > slot.cap_sv:
                grow = dict(self._grow)
                grow["cap_sr"], grow["cap_sv"] = max(2 * slot.cap_sr, 2 * npc), max(2 * slot.cap_sv, 2 * nvx)
                self._grow = grow
                return self._regrow_and_rerun(slot, 0)
            off = slot.h_sp_off.numpy()[:npc + 1].astype(np.int64)
            pieces = dict(ev=slot.h_sp_ev.numpy()[:npc].copy(), off=off, xy=slot.h_sp_xy.numpy()[:nvx].copy(),
                          overflow=bool(clip_over))
        n_near = int(slot.h_near_cnt[0]) if "streamers" in self.which else 0
        rec = slot.h_near.numpy()[:min(n_near, _lib.NEAR_CAP)]
        near = np.c_[rec[:, 0] // L, rec[:, 0] % L, rec[:, 1], rec[:, 2], rec[:, 3] & 0x0FFFFFFF, (rec[:, 3] >> 28) & 7]
        return BatchResult(ntime=nt, contours=cs, tables=tables, flags=flags,
                           gmax_nx=max_nx if pend["gmax"] is None else int(pend["gmax"]), n_split=n_sp,
                           near=near.astype(np.int32), near_total=n_near, flags_packed=packed, work=work, pieces=pieces,
                           counts=(ns, no, nc))

    def _regrow_and_rerun(self, slot, status):
        pend = slot.pending
        caps = dict(slot.ctx.caps)
        grow = dict(self._grow)
        if status & _lib.ST_SEG_OVERFLOW:
            grow["seg_cap"] = caps["seg_cap"] * 2
        if status & _lib.ST_CONTOUR_OVERFLOW:
            grow["contour_cap"] = caps["contour_cap"] * 2
        if status & _lib.ST_PAIR_OVERFLOW:
            grow["pair_cap"] = caps["pair_cap"] * 4
        if status & _lib.ST_EVENT_OVERFLOW:
            grow["event_cap"] = caps["event_cap"] * 4
        if status & _lib.ST_SEL_OVERFLOW:
            grow["sel_cap"] = caps["sel_cap"] * 4
        # the summary record holds the true totals of the batch: size the buffers for them (with headroom) in one go
        C, P, ns, no, nc = (int(v) for v in slot.h_summary.numpy()[:5])
        if status & _lib.ST_PACK_OVERFLOW:
            grow["cap_c"] = max(slot.cap_c * 2, int(1.3 * C))
            grow["cap_p"] = max(slot.cap_p * 2, int(1.3 * P))
        if status & _lib.ST_FETCH_OVERFLOW:
            grow["cap_e"] = max(slot.cap_e * 2, int(1.3 * (ns + no + nc)))
            grow["cap_r"] = slot.cap_r * 2
        self._grow = grow
        self._grow_version += 1
        if len(grow) and max(grow.values()) > (1 << 28):
            raise _lib.WbkError(_lib.ERR_CAPACITY, "arenas still overflow after repeated regrowth (status {})".format(status))
        key = next(k for k, v in self._slots.items() if v is slot)
        if self.lib.is_cuda:
            torch.cuda.synchronize()
        slot.close()
        new = _Slot(self, slot.T, grow)
        new.grow_seen = self._grow_version
        self._slots[key] = new
        self.submit(new, pend["raw"], flags_out=pend["flags"], flags_host=pend["flags_host"], gmax_nx=pend["gmax"],
                    smoothed=pend["smoothed"], intensity=pend["intensity"], packed_host=pend.get("packed_host"))
        return self.collect(new)

    # ------------------------------------------------------------------ convenience
    def run_batch(self, raw, gmax_nx=None, smoothed=None, intensity=None):
        """Synchronous: raw device tensor [T, nlat, nlon] -> BatchResult (flags and contour set are private
        copies, valid after further calls)."""
        nt = int(raw.shape[0])
        slot = self._slot(nt)
        flags = None
        if self.want_flags:
            flags = torch.empty((3, nt, self.nlat, self.nlon), dtype=torch.int8, device=self.lib.device)
        self.submit(slot, raw, flags_out=flags, gmax_nx=gmax_nx, smoothed=smoothed, intensity=intensity)
        res = self.collect(slot)
        cs = res.contours
        cs.job_off, cs.pt_off, cs.meta, cs.pts = cs.job_off.clone(), cs.pt_off.clone(), cs.meta.clone(), cs.pts.clone()
        return res

    def run_batch_host(self, raw_host, flags_host=None):
        """Synchronous end-to-end: (pinned) host input -> tables + flag grids in (pinned) host memory."""
        nt = int(raw_host.shape[0])
        slot = self._slot(nt)
        if self.want_flags and flags_host is None:
            flags_host = _pinned((3, nt, self.nlat, self.nlon), torch.int8, self.lib)
        self.submit(slot, raw_host, flags_host=flags_host)
        return self.collect(slot)

    def stream(self, batches, depth=3, flags_host=None, shared_stream=False, gmax_nx="full", packed_host=None):
        """Pipelined execution: yields one BatchResult per input batch, in order.

        ``batches``: iterable of device or pinned-host tensors of (at most) equal length.  ``depth`` batches are
        in flight at once, each on its own CUDA stream, so uploads, kernels, downloads and the host-side table
        handling overlap.  With ``shared_stream`` all slots enqueue on ONE side stream: kernels of different
        batches do not overlap each other, the host merely runs ahead (device-resident inputs).
        ``flags_host``: optional list of pinned int8 buffers, one per slot.
        ``packed_host``: optional list of pinned uint8 buffers (``packed_nbytes(3 * T * nlat * nlon)`` bytes), one per
        slot: the flag grids come back bit-packed (8x less device -> host traffic; ``unpack_flags`` restores them).
        ``gmax_nx``: the global ``exp_lon.max()`` in columns.  The reference takes it over ALL dates of a call
        (streamer_index.py:106, overturning_index.py:105, cutoff_index.py:90), a stream only sees one batch at a
        time, so the default ``"full"`` uses the width of the extended grid, ``nlon + periodic_add / dlon``: the value
        the maximum has as soon as ANY date of the record holds a circumglobal contour (a batch without one then yields
        no streamers / overturnings, exactly like those dates do in the reference).  If no batch at all reaches the full
        width a warning is logged with the maximum that was seen -- re-run with ``gmax_nx=<that value>`` to get the
        reference's result for such a record.  ``None`` = every batch uses its own maximum; an int = that value.
        The flag grids / contour set of a result are only valid until its slot is reused (``depth`` batches later).
        """
        common = None
        if shared_stream and self.lib.is_cuda:
            if getattr(self, "_shared", None) is None:
                self._shared = torch.cuda.Stream(device=self.lib.device)
            common = self._shared
        full = self.nlon + self.add
        assumed = gmax_nx == "full"
        if assumed:
            gmax_nx = full
        seen_max = 0
        inflight = []
        i = 0
        for raw in batches:
            idx = i % depth
            if len(inflight) == depth:
                res = self.collect(inflight.pop(0))
                seen_max = max(seen_max, res.contours.max_nx)
                yield res
            slot = self._slot(int(raw.shape[0]), idx)
            fh = flags_host[idx] if flags_host is not None else None
            ph = packed_host[idx] if packed_host is not None else None
            inflight.append(self.submit(slot, raw, flags_host=fh, stream=common, gmax_nx=gmax_nx, packed_host=ph))
            i += 1
        while inflight:
            res = self.collect(inflight.pop(0))
            seen_max = max(seen_max, res.contours.max_nx)
            yield res
        if assumed and i > 0 and seen_max < full:
            logger.warning("no contour of the %d batches spans the extended grid (%d columns; widest: %d): the "
                           "reference's exp_lon.max() would be %d columns, re-run with gmax_nx=%d", i, full, seen_max,
                           seen_max, seen_max)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def summarize(res):
    """Small dict of counts (used by the benchmark's result read-back and the tests)."""
    return dict(
        ntime=res.ntime, contours=res.contours.ncontours, points=res.contours.npoints,
        streamers=len(res.tables["streamers"]), overturnings=len(res.tables["overturnings"]),
        cutoffs=len(res.tables["cutoffs"]), split=res.n_split, near=res.near_total,
        segments=(res.work or {}).get("segments", 0), pairs=(res.work or {}).get("pairs", 0),
    )


def events_soup(res, kind, det):
    """Lattice polygons of the events of one kind of a BatchResult, as ``track_events`` sees them after
    ``transform_polygons`` (utils/index_utils.py:129-184): index coordinates of the REAL grid, folded with
    ``x % nlon``, events that straddle the last meridian split into their pieces (a multipolygon).
    Returns ``(tracking.PolygonSoup, com)`` with ``com`` the (lon, lat) centre-of-mass column."""
    from . import geometry, tracking

    tab = res.tables[kind]
    n = len(tab)
    nlon = det.nlon
    if kind == "overturnings":
        x0, y0, x1, y1 = (tab.box[:, k].astype(np.int64) for k in range(4))
        xy = np.stack([np.c_[x1, y0], np.c_[x1, y1], np.c_[x0, y1], np.c_[x0, y0]], axis=1).reshape(-1, 2)
        off = np.arange(n + 1, dtype=np.int64) * 4
    else:
        off = tab.rings.off - tab.rings.off[0]
        p = tab.rings.packed
        xy = np.c_[(p & 0xFFFF).astype(np.int64), (p >> 16).astype(np.int64)]
    straddle = np.nonzero(tab.split == 1)[0]
    if len(straddle) == 0:
        xy[:, 0] %= nlon
        soup = tracking.PolygonSoup(xy.astype(np.int32), off, np.arange(n + 1), True)
    elif res.pieces is not None and not res.pieces["overflow"]:
        # pieces from the device clipper (wbk_split_fetch): everything stays vectorised.  Rows of this kind start at
        # `first` in the kind-major gather order of the batch
        first = int(sum(res.counts[:detect.KINDS.index(kind)]))
        allxy, ring_off, poly_off = geometry.interleave_pieces(xy, off, tab.split, res.pieces, first, nlon)
        soup = tracking.PolygonSoup(allxy.astype(np.int32), ring_off, poly_off, True)
    else:
        # only the few events that straddle the meridian are rebuilt one by one
        ring_xy, ring_len, poly_nr = [], [], np.ones(n, dtype=np.int64)
        prev = 0
        for e in straddle:
            a, b = off[prev], off[e]
            if b > a:
                seg = xy[a:b].copy()
                seg[:, 0] %= nlon
                ring_xy.append(seg)
                ring_len.extend(np.diff(off[prev:e + 1]).tolist())
            pieces = geometry.transform_ring(xy[off[e]:off[e + 1]], nlon)
            poly_nr[e] = len(pieces)
            for pc in pieces:
                ring_xy.append(np.asarray(pc, dtype=np.int64).reshape(-1, 2))
                ring_len.append(len(pc))
            prev = e + 1
        if off[n] > off[prev]:
            seg = xy[off[prev]:off[n]].copy()
            seg[:, 0] %= nlon
            ring_xy.append(seg)
            ring_len.extend(np.diff(off[prev:n + 1]).tolist())
        allxy = np.concatenate(ring_xy) if ring_xy else np.zeros((0, 2), dtype=np.int64)
        soup = tracking.PolygonSoup(allxy.astype(np.int32), np.r_[0, np.cumsum(ring_len)], np.r_[0, np.cumsum(poly_nr)],
                                    True)
    props = detect.finish_properties(tab, det.lon, det.lat, nlon)
    return soup, np.asarray(props["com"], dtype=np.float64).reshape(-1, 2)
