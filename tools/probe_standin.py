import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import tinyraytracing_b200 as trt
from tinyraytracing_b200 import scenes, workloads
a=scenes.cornell_with_standin()
host=trt.HostScene.from_arrays(a['v9'],a['mtl'],a['materials'],a['lights'],a['eye'],a['lookat'],a['up'],a['fovy'],a['width'],a['height'],vn9=a['vn9'],vt6=a['vt6'])
dev=trt.DeviceScene(host,0)
print(dev.stats())
rng=np.random.default_rng(0)
rays=np.concatenate([workloads.camera_rays(host.camera(), 1<<18, rng), workloads.box_rays(*host.root_box(), 1<<18, rng)])
print("counters", dev.trace_counters(rays))
b,l=host.nodes()
leaf=l[:,3]>0
ext=b[leaf,3:]-b[leaf,:3]
print("leaves", leaf.sum(), "leaf extent max quantiles", np.quantile(ext.max(1),[0.5,0.9,0.99,0.999,1.0]))
big=np.flatnonzero(ext.max(1)>300); print("n big leaves", len(big), "tris in them", l[leaf][big][:,3].sum())
import time
for flags in (0, 8, 4):
    t0=time.time(); ids,t=dev.trace_closest(rays, flags); print(flags, "%.3fs"%(time.time()-t0), (ids>=0).mean())
