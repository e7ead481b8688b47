"""diagnostic: where does the CVAE output error against the oracle come from? (GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import ctrm_oracle as O
from helpers import *
from rng_tables import PhiloxTables
from ctrm_b200 import RoadmapBuilder, _lib

W = os.path.join(ROOT, "tests/golden/aamas22_main_best.npz")
w = O.OWeights.load(W)
z = np.load(os.path.join(ROOT, "tests/golden/ref_trace_demo.npz"))
oi = oins_from_npz(z)
NT, MT, K, seed = 2, 64, 3, 20221019
recs = []
O.build_timed_roadmaps(oi, w, O.OParams(N_traj=NT, max_T=MT, merge_distance_rate=0.1), PhiloxTables(seed, 7), recorder=recs)
b = RoadmapBuilder(W, 0)
b.upload([instance_from_oins(oi)])
v = b.device_view()
lib = _lib.load()
A, F, M = v.A, v.F, v.M
st = torch.cuda.current_stream().cuda_stream
ex, es_given_X, es_own = [], [], []
for r in recs:
    cur, prev = torch.from_numpy(r["cur"]).cuda(), torch.from_numpy(r["prev"]).cuda()
    fov = torch.zeros(A * 2 * F * F + 8, dtype=torch.float32, device="cuda")
    X = torch.zeros(A * 72, dtype=torch.float32, device="cuda")
    lib.ctrm_fov_raster(v.d_occ, v.d_ctg, v.ctg_layout, v.d_agent_inst, cur.data_ptr(), A, M, F, fov.data_ptr(), st)
    lib.ctrm_encode_features(b.ctx, fov.data_ptr(), cur.data_ptr(), prev.data_ptr(), X.data_ptr(), st)
    Xg = X.cpu().numpy().reshape(A, 72)
    ex.append([rel_err(Xg[:, lo:hi], r["X"][:, lo:hi]) for lo, hi in ((0, 8), (8, 40), (40, 72))])
    U = torch.from_numpy(np.ascontiguousarray(r["U"].reshape(K, A, 64).transpose(1, 0, 2))).cuda()
    ref = r["SAMPLES_nn"].reshape(K, A, 3).transpose(1, 0, 2)
    for src, acc in ((torch.from_numpy(r["X"]).cuda(), es_given_X), (X, es_own)):
        out = torch.zeros(A * K * 3, dtype=torch.float32, device="cuda")
        lib.ctrm_cvae_sample(b.ctx, src.data_ptr(), U.data_ptr(), K, 0, 0, 0, out.data_ptr(), None, st)
        g = out.cpu().numpy().reshape(A, K, 3)
        acc.append([rel_err(g[..., c], ref[..., c]) for c in range(3)] + [rel_err(g, ref)])
ex, es_given_X, es_own = np.array(ex), np.array(es_given_X), np.array(es_own)
print("steps", len(recs))
print("X err (self_info, fov_self, comm): max", ex.max(0), "median", np.median(ex, 0))
print("SAMPLES err given oracle X (dx,dy,mag,all): max", es_given_X.max(0), "median", np.median(es_given_X, 0))
print("SAMPLES err own X                       : max", es_own.max(0), "median", np.median(es_own, 0))
