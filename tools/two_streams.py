"""does running two half-batches concurrently (two contexts = two streams/graphs, two host threads) beat one batch?"""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import PARAMS, WEIGHTS, make_instances
from ctrm_b200 import RoadmapBuilder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ins = make_instances(B, 0)
one = RoadmapBuilder(WEIGHTS, 0)
p = one.pack_instances(ins)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    one.upload(None, packed=p); one.run(seed=1, export=False, **PARAMS); torch.cuda.synchronize()
    print(f"single ctx B={B}: {1e3*(time.perf_counter()-t0):.1f} ms", one.result_sizes()[:2])
one.close()
bs = [RoadmapBuilder(WEIGHTS, 0) for _ in range(NS)]
ps = [b.pack_instances(ins[i * B // NS:(i + 1) * B // NS]) for i, b in enumerate(bs)]
def work(i):
    bs[i].upload(None, packed=ps[i]); bs[i].run(seed=1, instance_id_base=i * B // NS, export=False, **PARAMS)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(NS)]
    [t.start() for t in th]; [t.join() for t in th]; torch.cuda.synchronize()
    print(f"{NS} ctx x B={B//NS}: {1e3*(time.perf_counter()-t0):.1f} ms", [b.result_sizes()[:2] for b in bs])
