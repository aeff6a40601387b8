#!/usr/bin/env python
"""Stage 4 at a BASELINE shape on one GPU: shared-memory accumulation (one CTA per coarse row) against the sort path."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdr
from gdr import synth, _lib

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "E"
cfg = synth.CONFIGS[name]
n, k = cfg["n"], cfg["k"]
u, v = synth.uniform_graph(n, cfg["pairs"], 1238)
A = gdr.sym_normalize(gdr.coo_to_csr(torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev), None, (n, n), symmetrize=True,
                                     binarize=True), 2)
labels = torch.from_numpy(np.random.RandomState(0).randint(0, k, n).astype(np.int32)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for dense in (0, 1):
    _lib.call("gdr_debug_set", b"coarsen_dense", dense)
    for what, fn in (("coarsen_edges (cells, counts, sums)", lambda: gdr.coarsen_edges(labels, labels, k, k, csr=A, weights=A.vals, drop_diag=True)),
                     ("graph_compress (whole stage)", lambda: gdr.graph_compress(labels, A, []))):
        ts = []
        for _ in range(4):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"{name} {'shared-memory rows' if dense else 'sort path'}: {what} {min(ts[1:]):.3f} ms", flush=True)
_lib.call("gdr_debug_set", b"coarsen_dense", 1)

# owner-side merge of routed pairs on one GPU (world 1: every pair is "received"): sort form vs shared-memory form
from gdr import parallel as par
ops = par.CudaOps()
keys, w, counts = ops.coarsen_route(A, labels, labels, k, 1)
m = counts[0]
stats = ops.cluster_stats(A, labels, k)
for dense in (0, 1):
    _lib.call("gdr_debug_set", b"coarsen_dense", dense)
    ts = []
    for _ in range(4):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = ops.coarse_merge_edges(keys[:m], w[:m], 0, k, k, stats=stats); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{name} merge of {m} routed pairs, {'shared-memory rows' if dense else 'sort'}: {min(ts[1:]):.3f} ms", flush=True)
_lib.call("gdr_debug_set", b"coarsen_dense", 1)

# the same for the pairs one of TWO owners would receive (coarse rows [0, k/2))
keys2, w2, counts2 = ops.coarsen_route(A, labels, labels, k, 2)
m0 = counts2[0]
cr = (k + 1) // 2
for dense in (0, 1):
    _lib.call("gdr_debug_set", b"coarsen_dense", dense)
    ts = []
    for _ in range(4):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = ops.coarse_merge_edges(keys2[:m0], w2[:m0], 0, cr, k, stats=stats); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{name} merge of {m0} routed pairs into {cr} coarse rows, {'shared-memory rows' if dense else 'sort'}: {min(ts[1:]):.3f} ms", flush=True)
_lib.call("gdr_debug_set", b"coarsen_dense", 1)
