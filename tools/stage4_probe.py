#!/usr/bin/env python
"""Where the multi-GPU stage 4 spends its time (config E, random labels).  torchrun --nproc-per-node N tools/stage4_probe.py"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdr
from gdr import synth, parallel as par

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
cfg = synth.CONFIGS["E"]
n = cfg["n"]
u, v = synth.uniform_graph(n, cfg["pairs"], 1238)
part, comm, ops = par.RowPartition(n, world, rank), par.Comm(dist), par.CudaOps()
per = (u.shape[0] + world - 1) // world
u_sl, v_sl = torch.from_numpy(u[rank * per:(rank + 1) * per].copy()).to(dev), torch.from_numpy(v[rank * per:(rank + 1) * per].copy()).to(dev)
A = par.dist_build_adjacency(comm, part, u_sl, v_sl, n, ops=ops)
import time
for symm in (False, True):
    comm.use_symm_exchange = symm
    for rep in range(4):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        A = par.dist_build_adjacency(comm, part, u_sl, v_sl, n, ops=ops)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0:
        print("stage 1", "symm" if symm else "nccl", round((t1 - t0) * 1e3, 3), "ms", flush=True)
labels = torch.from_numpy(np.random.RandomState(0).randint(0, cfg["k"], n).astype(np.int32)[part.lo:part.hi].copy()).to(dev)
for merge, symm in (("route", False), ("route", True), ("records", True)):
    comm.use_symm_exchange = symm
    for rep in range(3):
        t = {}
        par.dist_graph_compress(comm, part, labels, A, ops=ops, merge=merge, timing=t)
    if rank == 0:
        print(merge, "symm" if symm else "nccl", {k: round(x, 3) for k, x in t.items()}, "total", round(sum(t.values()), 3), flush=True)
for rep in range(2):
    t = {}
    par.dist_graph_compress(comm, part, labels, A, ops=ops, replicate=False, timing=t)
if rank == 0:
    print("route, not replicated", {k: round(x, 3) for k, x in t.items()}, "total", round(sum(t.values()), 3), flush=True)
comm.close()
dist.destroy_process_group()
