"""A/B of the smoothing strip width (one or two 64-column halves per warp) on the benchmark batch: per-kernel
times of smooth_fused / ms_segments over 296 time steps, and equality of every result.

  python tools/ab_halves.py [nsteps]      (needs a GPU)
"""
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from wavebreaking_b200 import _lib, pipeline, spatial, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
nlat, nlon = 721, 1440
lat, lon = synthetic.grid_coords(nlat, nlon)
raw = spatial.synth_pv(n, nlat, nlon, hour0=0.0, hour_step=1.0)
lib = _lib.get()
ref = None
for halves in (1, 2, 1, 2):
    lib.cdll.wbk_tune_smooth_halves(halves)
    det = pipeline.Detector(lat, lon, levels=[2.0], nvtx=False)
    for _ in range(2):
        res = det.run_batch(raw)
    torch.cuda.synchronize()
    lib.cdll.wbk_prof_reset()
    lib.cdll.wbk_prof_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        res = det.run_batch(raw)
    e1.record()
    torch.cuda.synchronize()
    lib.cdll.wbk_prof_enable(0)
    prof = _lib.prof_read(lib)
    line = {k: round(v[1] / v[0], 4) for k, v in prof.items() if k in ("smooth_fused", "ms_segments", "contour_link")}
    summ = pipeline.summarize(res)
    flags = res.flags.cpu().numpy()
    if ref is None:
        ref = (summ, flags)
    same = summ == ref[0] and np.array_equal(flags, ref[1])
    print("halves", halves, "batch ms", round(e0.elapsed_time(e1) / 5, 3), line, "same results" if same else "DIFFERENT", flush=True)
lib.cdll.wbk_tune_smooth_halves(0)
