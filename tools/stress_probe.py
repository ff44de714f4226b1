"""Probe of the noisy ("stress") workload on the GPU: python tools/stress_probe.py [T] [amp] [graphs] [nbatches]"""
import faulthandler
import sys
import time

sys.path.insert(0, ".")
import torch

from wavebreaking_b200 import _lib, pipeline, spatial, synthetic

faulthandler.dump_traceback_later(100, exit=True)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
amp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
graphs = len(sys.argv) > 3 and sys.argv[3] == "1"
nb = int(sys.argv[4]) if len(sys.argv) > 4 else 6
lat, lon = synthetic.grid_coords(721, 1440)
det = pipeline.Detector(lat, lon, levels=[2.0], graphs=graphs)
g = torch.Generator(device="cuda").manual_seed(1)
slabs = []
for s in range(2):
    raw = spatial.synth_pv(T, 721, 1440, hour0=float(s * T))
    slabs.append(raw + amp * torch.randn(raw.shape, generator=g, device="cuda", dtype=torch.float32))
lib = _lib.get()
t0 = time.time()
for i, res in enumerate(det.stream((slabs[k % 2] for k in range(nb)), depth=3)):
    torch.cuda.synchronize()
    print("batch", i, round(time.time() - t0, 2), "s", pipeline.summarize(res), "grow", det._grow,
          "mem GB", round(torch.cuda.memory_allocated() / 1e9, 1), flush=True)
print("max mem GB", torch.cuda.max_memory_allocated() / 1e9)
