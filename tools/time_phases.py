"""where does a device-resident step spend its wall time? (upload / build / export_device)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import PARAMS, WEIGHTS, make_instances
from ctrm_b200 import RoadmapBuilder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = RoadmapBuilder(WEIGHTS, 0)
ins = make_instances(B, 0)
packed = b.pack_instances(ins)
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b.upload(None, packed=packed); torch.cuda.synchronize(); t1 = time.perf_counter()
    b.run(seed=1, export=False, **PARAMS); torch.cuda.synchronize(); t2 = time.perf_counter()
    r = b.export_device(); torch.cuda.synchronize(); t3 = time.perf_counter()
    c = b.export(); t4 = time.perf_counter()
    print(f"iter {it}: upload {1e3*(t1-t0):.1f} ms  build {1e3*(t2-t1):.1f} ms  export_device {1e3*(t3-t2):.1f} ms  export_host {1e3*(t4-t3):.1f} ms  {b.timing()}")
