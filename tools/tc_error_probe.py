#!/usr/bin/env python
"""How big is the re-score set of the tensor-core E-step, and how large is the real 3xTF32 error?
Run on the GPU box:  python tools/tc_error_probe.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdr
from gdr import synth
from gdr._dev import padded_rows, new_padded
from gdr.kmeans import TcOperand, assign_labels, segment_sum

dev = torch.device("cuda:0")
cfg = synth.CONFIGS["B"]
n, f, K = cfg["n"], cfg["f"], cfg["k"]
u, v = synth.uniform_graph(n, cfg["pairs"], 1235)
A = gdr.sym_normalize(gdr.coo_to_csr(torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev), None, (n, n), symmetrize=True, binarize=True), 2)
X = torch.from_numpy(synth.features(n, f, 1335)).to(dev)
_, tgt = gdr.propagate(A, X, 3, 0.8)
Xc = padded_rows((tgt - tgt.mean(0)).contiguous())
perm = torch.from_numpy(np.random.RandomState(1235).permutation(n)[:K].astype(np.int64)).to(dev)
C = padded_rows(Xc[perm].clone())
op = TcOperand(Xc)
xn = Xc.norm(dim=1)
for it in range(6):
    lab = torch.empty(n, dtype=torch.int32, device=dev)
    best = torch.empty(n, dtype=torch.float32, device=dev)
    nref = torch.zeros(1, dtype=torch.int32, device=dev)
    assign_labels(Xc, C, lab, best=best, tc_operand=op, n_refined=nref)
    # exact distances in fp64 for a sample of rows
    idx = torch.arange(0, n, 37, device=dev)
    d64 = (C.double() ** 2).sum(1)[None, :] - 2.0 * Xc[idx].double() @ C.double().t()
    srt, _ = torch.sort(d64, dim=1)
    gap = (srt[:, 1] - srt[:, 0])
    cmax = C.norm(dim=1).max()
    scale = (xn[idx].double() * cmax.double())
    print(f"iter {it}: n_refined={int(nref.item())} ({100*int(nref.item())/n:.2f}%)  |x| med={xn.median().item():.3f} cmax={cmax.item():.3f}  "
          f"band/scale=2^-15  gap/scale quantiles 1%={torch.quantile(gap/scale,0.01).item():.2e} 10%={torch.quantile(gap/scale,0.1).item():.2e} 50%={torch.quantile(gap/scale,0.5).item():.2e}")
    sums, counts = segment_sum(Xc, lab, K)
    Cn = sums / counts.clamp_min(1).unsqueeze(1)
    C = padded_rows(Cn.contiguous())
# real error of the tensor-core distances: compare its best value with the fp64 minimum on unambiguous rows
import ctypes
from gdr import _lib
lab = torch.empty(n, dtype=torch.int32, device=dev); best = torch.empty(n, dtype=torch.float32, device=dev)
assign_labels(Xc, C, lab, best=best, tc_operand=op)
idx = torch.arange(0, n, 11, device=dev)
d64 = (C.double() ** 2).sum(1)[None, :] - 2.0 * Xc[idx].double() @ C.double().t()
err = (best[idx].double() - d64.min(dim=1).values).abs() / (xn[idx].double() * C.norm(dim=1).max().double())
print(f"|d_tc - d_fp64| / (|x| cmax): max={err.max().item():.3e} p99={torch.quantile(err,0.99).item():.3e} median={err.median().item():.3e}   (2^-15={2**-15:.3e}, 2^-18={2**-18:.3e})")
