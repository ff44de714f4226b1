"""Parity sweep at the benchmark shape: N hourly 721x1440 time steps, CUDA pipeline vs the oracle, event by event.

  python tools/validate_c25.py [nsteps] [hour0] [noise]   (needs a GPU; the oracle takes ~8 s per step on the box)
`noise` (PVU, default 0): white noise per cell added before the smoothing, the recipe of bench.py's stress line
(1.5 gives ~3x the contour points and hundreds of contours per step; the oracle then needs ~1 min per step).
Writes a one-line-per-step report and a summary (profiles/r*_parity_c25*.txt).
"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from oracle import geom as G
from oracle import pipeline as P
from wavebreaking_b200 import detect, pipeline, spatial, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
hour0 = float(sys.argv[2]) if len(sys.argv) > 2 else 1000.0
noise = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
nlat, nlon = 721, 1440
lat, lon = synthetic.grid_coords(nlat, nlon)
raw = spatial.synth_pv(n, nlat, nlon, hour0=hour0, hour_step=37.0)  # spread over the year
if noise:
    g = torch.Generator(device="cuda").manual_seed(20260101)
    raw = raw + noise * torch.randn(raw.shape, generator=g, device="cuda", dtype=torch.float32)
det = pipeline.Detector(lat, lon, levels=[2.0])
res = det.run_batch(raw)
raw_h = raw.cpu().numpy()
tot = dict(events=0, mismatched_events=0, flag_cells_diff=0, near_listed=int(res.near_total), contours=0, points=0,
           non_simple_rings=0, near_tie_group_winners=int(((res.tables["streamers"].near & 2) != 0).sum()))
non_simple = []  # (step, kind, event) whose ring touches / crosses itself: invalid for GEOS ("GEOS-invalid class")
for t in range(n):
    t0 = time.time()
    grid = P.Grid(lon, lat, synthetic.time_axis(1, 1))
    want = P.detect_steps(raw_h[t:t + 1], grid, levels=[2.0])
    line = ["step {:2d}".format(t)]
    c = want["contours"]
    h = res.contours.host()
    sel = h["job"] == t
    npts = int(sum(h["pt_off"][k + 1] - h["pt_off"][k] for k in np.nonzero(sel)[0]))
    ok_c = int(sel.sum()) == len(c) and npts == sum(len(g) for g in c.geometry)
    tot["contours"] += len(c)
    tot["points"] += npts
    line.append("contours {} pts {} {}".format(len(c), npts, "ok" if ok_c else "MISMATCH"))
    for k, kind in enumerate(detect.KINDS):
        tab = res.tables[kind]
        idx = np.nonzero(tab.job == t)[0]
        w = want["events"][kind]
        bad = 0
        if len(idx) != len(w):
            bad = max(len(idx), len(w))
        else:
            props = detect.finish_properties(tab, lon, lat, nlon)
            tot["com_near_integer"] = tot.get("com_near_integer", 0) + int(props["com_near_integer"][idx].sum())
            for j, e in enumerate(idx):
                row = w.iloc[j]
                same = (tuple(props["com"][e]) == tuple(row.com) and props["mean_var"][e] == row.mean_var
                        and props["event_area"][e] == row.event_area)
                if kind != "overturnings":
                    same = same and np.array_equal(tab.rings[e], np.asarray(w.attrs["_index_rings"][j]))
                bad += 0 if same else 1
        fd = int(np.count_nonzero(res.flags[k, t].cpu().numpy() != want["flags"][kind][0]))
        tot["events"] += len(w)
        tot["mismatched_events"] += bad
        tot["flag_cells_diff"] += fd
        ns = 0
        if kind != "overturnings":
            for j, e in enumerate(idx):
                if not G.ring_is_simple(np.asarray(tab.rings[e])):
                    ns += 1
                    non_simple.append((t, kind, j))
        tot["non_simple_rings"] += ns
        line.append("{} {} bad {} flagdiff {} non-simple {}".format(kind, len(w), bad, fd, ns))
    line.append("({:.1f} s)".format(time.time() - t0))
    print(" | ".join(line), flush=True)
print("NEAR-THRESHOLD PAIRS (step, level, contour, i, j, flags 1 kept 2 geo 4 cont):", res.near.tolist())
print("NON-SIMPLE EVENT RINGS (step, kind, event):", non_simple)
print("SUMMARY", tot)
