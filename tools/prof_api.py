"""cProfile of one pass through the drop-in API at the benchmark shape (python tools/prof_api.py [T])."""
import cProfile
import pstats
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

import wavebreaking_b200 as wb
from wavebreaking_b200 import compat, spatial, synthetic

T = int(sys.argv[1]) if len(sys.argv) > 1 else 148
lat, lon = synthetic.grid_coords(721, 1440)
host_nps = [spatial.synth_pv(T, 721, 1440, hour0=float(k * T)).cpu().numpy() for k in range(3)]  # no reuse across passes
tt = np.datetime64("2000-01-01T00", "ns") + np.arange(T) * np.timedelta64(3600 * 10**9, "ns")


def api_pass(k=0):
    t00 = time.perf_counter()
    pv = compat.Field(host_nps[k], ("time", "lat", "lon"), {"time": tt, "lat": lat, "lon": lon}, name="PV")
    t = [time.perf_counter()]

    def lap(name):
        torch.cuda.synchronize()
        now = time.perf_counter()
        print("  {:28s} {:8.1f} ms".format(name, 1000 * (now - t[0])))
        t[0] = now

    sm = wb.calculate_smoothed_field(pv, 5)
    lap("calculate_smoothed_field")
    contours = wb.calculate_contours(sm, 2, original_coordinates=False)
    lap("calculate_contours")
    evs = []
    for fn in (wb.calculate_streamers, wb.calculate_overturnings, wb.calculate_cutoffs):
        evs.append(fn(sm, 2, contours=contours))
        lap(fn.__name__)
    for ev in evs:
        g = np.asarray(wb.to_xarray(sm, ev).values)
        lap("to_xarray + values")
    print("  pass total {:8.1f} ms".format(1000 * (time.perf_counter() - t00)))
    return evs


api_pass(0)
print("second pass (fresh host array)")
api_pass(1)
print("third pass (fresh host array, under cProfile)")
pr = cProfile.Profile()
pr.enable()
api_pass(2)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
