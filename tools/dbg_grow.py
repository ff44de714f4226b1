import sys, time
sys.path.insert(0, '.')
import torch
from wavebreaking_b200 import pipeline, spatial, synthetic, _lib
lat, lon = synthetic.grid_coords(721, 1440)
det = pipeline.Detector(lat, lon, levels=[2.0])
raw = [spatial.synth_pv(296, 721, 1440, hour0=296.0 * i) for i in range(3)]
for i in range(6):
    torch.cuda.synchronize(); t = time.perf_counter()
    res = list(det.stream([raw[i % 3]], depth=1, shared_stream=True))[0]
    torch.cuda.synchronize(); print(i, round((time.perf_counter() - t) * 1e3, 2), 'ms', det._grow, pipeline.summarize(res))
lib = _lib.get()
lib.cdll.wbk_prof_reset(); lib.cdll.wbk_prof_enable(1)
t = time.perf_counter()
for r in det.stream([raw[i % 3] for i in range(6)], depth=3, shared_stream=True): pass
torch.cuda.synchronize(); print('6 batches', round((time.perf_counter() - t) * 1e3, 2), 'ms')
print(_lib.prof_read(lib))
