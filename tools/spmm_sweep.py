#!/usr/bin/env python
"""Sweep the SpMM tuning knobs on a BASELINE-shaped graph (run on the GPU box).

    python tools/spmm_sweep.py [B|E|skew] > gpurun_out/spmm_sweep.txt
"""
import ctypes, itertools, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdr
from gdr import synth, _lib

which = sys.argv[1] if len(sys.argv) > 1 else "B"
dev = torch.device("cuda:0")
cfg = synth.CONFIGS["E" if which == "E" else "B"]
n, f = cfg["n"], cfg["f"]
gen = synth.skewed_graph if which == "skew" else synth.uniform_graph
u, v = gen(n, cfg["pairs"], 1235)
A = gdr.sym_normalize(gdr.coo_to_csr(torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev), None, (n, n), symmetrize=True, binarize=True), 2)
X = torch.randn(n, f, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
nnz = A.nnz
b_min = nnz * 8 + (n + 1) * 4 + 4 * n * f * 4
b_gather = nnz * (8 + 4 * f) + (n + 1) * 4 + 3 * n * f * 4
ref = None
print(f"# {which}: N={n} nnz={nnz} F={f}  B_min={b_min/1e6:.0f} MB  B_gather={b_gather/1e6:.0f} MB")
for unroll, hints, split in itertools.product([4, 2, 8], [0, 1], [1, 2]):
    _lib.call("gdr_debug_set", b"spmm_unroll", unroll)
    _lib.call("gdr_debug_set", b"spmm_hints", hints)
    _lib.call("gdr_debug_set", b"spmm_split", split)
    T = torch.zeros_like(X)
    for _ in range(2):
        y = gdr.spmm(A, X, alpha=0.8, accumulate_into=T, beta=0.2)
    if ref is None:
        ref = y.clone()
    assert torch.equal(y, ref), "variant changed the result"
    ts = []
    for _ in range(5):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gdr.spmm(A, X, alpha=0.8, accumulate_into=T, beta=0.2); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"unroll={unroll} hints={hints} split={split}: {ms*1e3:8.1f} us  B_min {b_min/ms/1e6:7.0f} GB/s  B_gather {b_gather/ms/1e6:7.0f} GB/s")
