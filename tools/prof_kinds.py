"""Per-kernel times of one batch with each index on its own (python tools/prof_kinds.py [batch])."""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
from wavebreaking_b200 import _lib, detect, pipeline, spatial, synthetic
T = int(sys.argv[1]) if len(sys.argv) > 1 else 296
lat, lon = synthetic.grid_coords(721, 1440)
raw = spatial.synth_pv(T, 721, 1440, hour0=0.0)
lib = _lib.get()
for which in (("streamers",), ("overturnings",), ("cutoffs",), detect.KINDS):
    det = pipeline.Detector(lat, lon, levels=[2.0], which=which)
    for _ in range(2):
        res = det.run_batch(raw)
    torch.cuda.synchronize()
    lib.cdll.wbk_prof_reset(); lib.cdll.wbk_prof_enable(1)
    for _ in range(3):
        res = det.run_batch(raw)
    torch.cuda.synchronize()
    prof = _lib.prof_read(lib)
    lib.cdll.wbk_prof_enable(0)
    n = {k: len(res.tables[k]) for k in detect.KINDS}
    members = {k: float(res.tables[k].sums[:, 5].sum()) for k in detect.KINDS}
    print(which, n, members)
    print("   ", {k: round(v[1] / 3, 3) for k, v in prof.items() if v[1] > 0.01})
