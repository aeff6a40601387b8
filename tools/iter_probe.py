#!/usr/bin/env python
"""One rank's share of a config-E Lloyd fit on ONE GPU (rows / world rows, all K centres): what an iteration costs
besides the tensor-core screen.  Run under `ncu --metrics gpu__time_duration.sum` for the launch list, or alone for the
per-iteration time.    python tools/iter_probe.py [world=8] [iters=6]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdr
from gdr import synth

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = synth.CONFIGS["E"]
n, f, K = (cfg["n"] + world - 1) // world, cfg["f"], cfg["k"]
dev = torch.device("cuda:0")
X = torch.from_numpy(synth.features(n, f, 1338)).to(dev)
C0 = X[torch.from_numpy(np.random.RandomState(3).permutation(n)[:K].astype(np.int64)).to(dev)].clone()
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    km = gdr.KMeans(n_clusters=K, init=C0, n_init=1, max_iter=iters, tol=0, precision="tc").fit(X)
    e1.record()
    torch.cuda.synchronize()
    print(f"rep {rep}: fit({iters} iterations) {e0.elapsed_time(e1):.3f} ms -> {e0.elapsed_time(e1) / (iters + 1):.3f} ms per E-step-equivalent, n_iter {km.n_iter_}")
