import cProfile, pstats, sys, numpy as np, torch, time
sys.path.insert(0,'.')
from wavebreaking_b200 import pipeline, spatial, synthetic
lat,lon=synthetic.grid_coords(721,1440)
det=pipeline.Detector(lat,lon,levels=[2.0])
raw=[spatial.synth_pv(296,721,1440,hour0=296.0*i) for i in range(3)]
for r in raw[:2]: det.run_batch(r)
torch.cuda.synchronize()
pr=cProfile.Profile(); pr.enable()
t=time.perf_counter(); res=det.run_batch(raw[2]); torch.cuda.synchronize(); print('batch s',time.perf_counter()-t)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
