"""Batched detection core: device contexts, contour sets and the index kernels' host wrappers.

Columnar layer under the drop-in API (``wavebreaking_b200.api``).  A *job* is one
(time step, contour level); jobs are ordered time-outer, level-inner like the reference's
``iterate_time_dimension`` / ``iterate_contour_levels`` loops (utils/index_utils.py:217-258).
"""

import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


def _iptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def default_caps(nlat, nlon, add, njobs):
    """Arena capacities that fit smooth geophysical fields with headroom (regrown on overflow)."""
    W = nlon + add
    p2 = lambda v: 1 << max(4, int(np.ceil(np.log2(max(v, 2)))))
    return dict(max_jobs=int(njobs), seg_cap=p2(8 * (W + nlat)), contour_cap=p2(max(256, W // 2)), sel_cap=4,
                pair_cap=p2(16 * W), event_cap=256)


class Context:
    """Owns a wbk_ctx and its torch-allocated device workspace."""

    def __init__(self, nlat, nlon, add, caps):
        self.lib = _lib.get()
        self.nlat, self.nlon, self.add = int(nlat), int(nlon), int(add)
        self.caps = dict(caps)
        c = _lib.Caps(**self.caps)
        need = self.lib.cdll.wbk_workspace_bytes(ctypes.byref(c), self.nlat, self.nlon, self.add)
        if need == 0:
            raise ValueError("invalid capacities / grid for wbk_create: {}".format(self.caps))
        self.workspace = torch.empty(need + 256, dtype=torch.uint8, device=self.lib.device)
        base = self.workspace.data_ptr()
        aligned = (base + 255) & ~255
        self.handle = ctypes.c_void_p()
        self.lib.call("wbk_create", ctypes.byref(self.handle), ctypes.byref(c), self.nlat, self.nlon, self.add,
                      ctypes.c_void_p(aligned), need)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.cdll.wbk_destroy(self.handle)
            self.handle = None
            self.workspace = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_CTX_CACHE = {}


def get_context(nlat, nlon, add, njobs, min_caps=None):
    """A cached context for this grid whose capacities cover ``njobs`` (and ``min_caps``)."""
    key = (int(nlat), int(nlon), int(add), _lib.get().path)
    want = default_caps(nlat, nlon, add, njobs)
    if min_caps:
        for k, v in min_caps.items():
            want[k] = max(want[k], int(v))
    ctx = _CTX_CACHE.get(key)
    if ctx is not None and all(ctx.caps[k] >= want[k] for k in want):
        return ctx
    if ctx is not None:
        for k in want:
            want[k] = max(want[k], ctx.caps[k])
        ctx.close()
    ctx = Context(nlat, nlon, add, want)
    _CTX_CACHE[key] = ctx
    return ctx


def clear_contexts():
    for ctx in _CTX_CACHE.values():
        ctx.close()
    _CTX_CACHE.clear()


@dataclass
class ContourSet:
    """Packed contours of a batch (device tensors) in index coordinates of the extended grid."""

    njobs: int
    nlevels: int
    nlat: int
    nlon: int
    add: int
    levels: np.ndarray
    job_off: torch.Tensor   # int32 [njobs + 1]
    pt_off: torch.Tensor    # int32 [C + 1]
    meta: torch.Tensor      # int32 [C, 4]: closed, nx, sum_y, job
    pts: torch.Tensor       # uint32-as-int32 [P]: x | y << 16
    status: np.ndarray      # int32 [njobs]
    max_nx: int
    h_ncontours: np.ndarray
    h_npoints: np.ndarray
    _host: dict = None

    @property
    def ncontours(self):
        return int(self.pt_off.shape[0]) - 1

    @property
    def npoints(self):
        return int(self.pts.shape[0])

    def host(self):
        """Host copies: dict(job_off, pt_off, closed, nx, sum_y, job, x, y)."""
        if self._host is None:
            meta = self.meta.cpu().numpy().reshape(-1, 4)
            pts = self.pts.cpu().numpy().view(np.uint32)
            self._host = dict(
                job_off=self.job_off.cpu().numpy(), pt_off=self.pt_off.cpu().numpy(),
                closed=meta[:, 0].astype(bool), nx=meta[:, 1].copy(), sum_y=meta[:, 2].copy(), job=meta[:, 3].copy(),
                x=(pts & 0xFFFF).astype(np.int64), y=(pts >> 16).astype(np.int64),
            )
        return self._host

    def contour_points(self, c):
        h = self.host()
        a, b = h["pt_off"][c], h["pt_off"][c + 1]
        return np.c_[h["x"][a:b], h["y"][a:b]]


def contours(field, levels, add, caps=None):
    """Contours of every (time step, level) of ``field`` [ntime, nlat, nlon] (device tensor).

    contour_index.py:86-175 with ``original_coordinates=False``.  Retries with larger arenas
    when a job overflows its capacity.
    """
    lib = _lib.get()
    if field.dim() != 3:
        raise ValueError("expected [ntime, nlat, nlon]")
    field = field.contiguous()
    ntime, nlat, nlon = (int(v) for v in field.shape)
    levels = np.ascontiguousarray(np.atleast_1d(np.asarray(levels, dtype=np.float64)))
    nlevels = len(levels)
    njobs = ntime * nlevels
    grow = dict(caps or {})
    for _attempt in range(8):
        ctx = get_context(nlat, nlon, add, max(njobs, 1), grow)
        lib.call("wbk_contours", ctx.handle, _lib.ptr(field), _lib.dtype_code(field.dtype), ntime,
                 levels.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), nlevels, lib.stream())
        nc = np.zeros(max(njobs, 1), dtype=np.int32)
        npnt = np.zeros(max(njobs, 1), dtype=np.int32)
        status = np.zeros(max(njobs, 1), dtype=np.int32)
        max_nx = ctypes.c_int(0)
        try:
            lib.call("wbk_contours_counts", ctx.handle, _iptr(nc), _iptr(npnt), _iptr(status), ctypes.byref(max_nx),
                     lib.stream())
            break
        except _lib.CapacityError:
            if status.max() & _lib.ST_SEG_OVERFLOW or np.any(status & _lib.ST_SEG_OVERFLOW):
                grow["seg_cap"] = ctx.caps["seg_cap"] * 2
            if np.any(status & _lib.ST_CONTOUR_OVERFLOW):
                grow["contour_cap"] = ctx.caps["contour_cap"] * 2
    else:
        raise _lib.WbkError(_lib.ERR_CAPACITY, "contour arenas still overflow after 8 regrowths")
    nc, npnt, status = nc[:njobs], npnt[:njobs], status[:njobs]
    C, Pn = int(nc.sum()), int(npnt.sum())
    dev = lib.device
    job_off = torch.empty(njobs + 1, dtype=torch.int32, device=dev)
    pt_off = torch.zeros(C + 1, dtype=torch.int32, device=dev)
    meta = torch.empty((C, 4), dtype=torch.int32, device=dev)
    pts = torch.empty(Pn, dtype=torch.int32, device=dev)
    if njobs > 0:
        lib.call("wbk_contours_pack", ctx.handle, _iptr(nc), _iptr(npnt), _lib.ptr(job_off), _lib.ptr(pt_off),
                 _lib.ptr(meta) if C else _lib.ptr(job_off), _lib.ptr(pts) if Pn else _lib.ptr(job_off), lib.stream())
    else:
        job_off.zero_()
    return ContourSet(njobs=njobs, nlevels=nlevels, nlat=nlat, nlon=nlon, add=int(add), levels=levels,
                      job_off=job_off, pt_off=pt_off, meta=meta, pts=pts, status=status.copy(),
                      max_nx=int(max_nx.value), h_ncontours=nc.copy(), h_npoints=npnt.copy())
