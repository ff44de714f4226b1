"""Small driver for ncu captures: one batch through the pipeline (python tools/ncu_kernels.py [batch])."""
import sys
sys.path.insert(0, '.')
import torch
from wavebreaking_b200 import pipeline, spatial, synthetic
T = int(sys.argv[1]) if len(sys.argv) > 1 else 148
lat, lon = synthetic.grid_coords(721, 1440)
det = pipeline.Detector(lat, lon, levels=[2.0])
raw = spatial.synth_pv(T, 721, 1440, hour0=0.0)
for _ in range(2):
    res = det.run_batch(raw)
torch.cuda.synchronize()
print(pipeline.summarize(res))
