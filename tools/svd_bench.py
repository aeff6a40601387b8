#!/usr/bin/env python
"""SURVEY §8(f)-3 row at the config-C shape (Yelp2018-shaped interactions, dim 64): compute_svd_embeddings on the GPU
(block Krylov on the CSR SpMM kernel) against scipy's svds on the host cores, with the singular-value parity."""
import json, os, sys, time
import numpy as np, scipy.sparse as sp, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gdr
from gdr import synth
cfg = synth.BIPARTITE[sys.argv[1] if len(sys.argv) > 1 else "C"]
u, i = synth.bipartite_interactions(cfg["users"], cfg["items"], cfg["inter"], seed=1236)
R = sp.csr_matrix((np.ones(u.shape[0], np.float32), (u, i)), shape=(cfg["users"], cfg["items"]))
dim = cfg["d"]
A = gdr.CSR.from_scipy(R, device="cuda:0")
for _ in range(2):
    ue, ie = gdr.compute_svd_embeddings(A, dim, seed=42)
torch.cuda.synchronize()
l0 = gdr.launch_count(); t0 = time.perf_counter()
ue, ie = gdr.compute_svd_embeddings(A, dim, seed=42)
torch.cuda.synchronize()
t_gpu = time.perf_counter() - t0
launches = gdr.launch_count() - l0
from scipy.sparse.linalg import svds
t0 = time.perf_counter()
U, S, VT = svds(R.astype(np.float64), k=dim)
t_cpu = time.perf_counter() - t0
S = np.sort(S)[::-1]
sig = (ue.astype(np.float64) ** 2).sum(0)
print(json.dumps({"what": "compute_svd_embeddings", "config": f"{cfg['name']}-shaped: {cfg['users']} x {cfg['items']}, nnz={R.nnz}, dim={dim}",
                  "gpu_ms": t_gpu * 1e3, "gpu_launches": int(launches), "cpu_svds_ms": t_cpu * 1e3, "speedup": t_cpu / t_gpu,
                  "sigma_max_rel_err": float(np.max(np.abs(sig - S) / S)), "sigma_first_last": [float(S[0]), float(S[-1])]}))
