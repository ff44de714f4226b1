import sys; sys.path.insert(0,".")
import numpy as np
from wavebreaking_b200 import pipeline, spatial, synthetic, detect
nlat,nlon=181,360
lat,lon=synthetic.grid_coords(nlat,nlon)
hours=np.sort(np.random.default_rng(5).choice(1460,16,replace=False))*6.0
raw=synthetic.pv_field(nlat,nlon,hours)
a=pipeline.Detector(lat,lon,levels=[2.0,-2.0],fuse=True).run_batch(spatial.to_device(raw))
b=pipeline.Detector(lat,lon,levels=[2.0,-2.0],fuse=False).run_batch(spatial.to_device(raw))
print(pipeline.summarize(a)); print(pipeline.summarize(b))
ha,hb=a.contours.host(),b.contours.host()
print(np.bincount(ha["job"],minlength=32)); print(np.bincount(hb["job"],minlength=32))
