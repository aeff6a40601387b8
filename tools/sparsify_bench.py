#!/usr/bin/env python
"""Measurement of the SURVEY §8(f)-1 row (edge scoring + per-class top-k sparsification) at the config-B
shape: graph_sparse(sp_type='attaw') on the GPU against the CPU port of the reference lines, with the
HBM byte model of the per-class pass.  Run on the GPU box:  python tools/sparsify_bench.py [B]
Prints one JSON line (kept under profiles/)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gdr
from gdr import synth
from oracle import ref_port as rp

name = sys.argv[1] if len(sys.argv) > 1 else "B"
cfg = synth.CONFIGS[name]
n, C, ratio = cfg["n"], cfg["d_logit"], 0.1
dev = torch.device("cuda:0")
u, v = synth.uniform_graph(n, cfg["pairs"], 1235)
A = gdr.sym_normalize(gdr.coo_to_csr(torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev), None, (n, n), symmetrize=True, binarize=True), 2)
ebd_h = (np.random.RandomState(7).randn(n, C) * 2).astype(np.float32)
ebd = torch.from_numpy(ebd_h).to(dev)
nnz, k = A.nnz, int(A.nnz * ratio)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
adj_coo = A.to_torch_coo()
adj_coo._gdr_csr = A if hasattr(adj_coo, "__dict__") else None

def run():
    return gdr.graph_sparse(A, ratio, ebd=ebd, sp_type="attaw")

for _ in range(3):
    out = run()
torch.cuda.synchronize()
ts = []
l0 = gdr.launch_count()
for _ in range(5):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = run(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
launches = (gdr.launch_count() - l0) // 5
ms = float(np.median(ts))
# per class: class weight (colidx, er, w: 12 B/edge) + 4 histogram passes (4 B/edge each) + equal flags (8) + scan (~12)
# + selection flags (16) + scan (~12) + compaction (pos 8 + colidx/vals 8 + k * 8 out)
bytes_class = nnz * (12 + 16 + 8 + 12 + 16 + 12 + 16) + k * 8
bytes_total = C * bytes_class + nnz * 40
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
# CPU port of the reference lines, bounded sample: 2 classes (the per-class loop is what scales), scaled to C
adj_cpu = rp.to_tensor_sparse(A.to_scipy())
ebd_cpu = torch.from_numpy(ebd_h)
t0 = time.perf_counter(); rp.graph_sparse_attaw(adj_cpu, ratio, ebd_cpu, max_classes=1); t1 = time.perf_counter()
rp.graph_sparse_attaw(adj_cpu, ratio, ebd_cpu, max_classes=3); t2 = time.perf_counter()
per_class_cpu = ((t2 - t1) - (t1 - t0)) / 2
cpu_total = (t1 - t0) + per_class_cpu * (C - 1)
# parity on the same inputs: class-0 edge set against the CPU port
ref0 = rp.graph_sparse_attaw(adj_cpu, ratio, ebd_cpu, max_classes=1)[0].coalesce()._indices().numpy()
got0 = out[0].coalesce()._indices().cpu().numpy()
sref, sgot = set(zip(ref0[0].tolist(), ref0[1].tolist())), set(zip(got0[0].tolist(), got0[1].tolist()))
print(json.dumps({"what": "graph_sparse(sp_type='attaw')", "config": f"config {name}: N={n}, nnz={nnz}, classes={C}, ratio={ratio}, k={k}",
                  "gpu_ms": ms, "gpu_ms_per_class": ms / C, "gpu_launches": int(launches), "edges_per_s": nnz * C / (ms / 1e3),
                  "roofline": {"bound": "hbm", "achieved": bytes_total / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": bytes_total / (ms / 1e3) / 1e9 / peak, "bytes_model": "per class nnz*92 + k*8, setup nnz*40"},
                  "cpu_baseline": {"value_ms": cpu_total * 1e3, "kind": "port", "cores": rp.host_info()["torch_threads"],
                                   "sample": "1 and 3 classes timed, per-class cost extrapolated to all classes"},
                  "speedup": cpu_total * 1e3 / ms, "parity_class0": {"k": k, "sym_diff_edges": len(sref ^ sgot)}}))
