#!/usr/bin/env python
"""Per-fit time series of KMeans.fit under different surroundings (run on the GPU box)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdr
from gdr import synth, _lib

dev = torch.device("cuda:0")
cfg = synth.CONFIGS["B"]
n, f, K = cfg["n"], cfg["f"], cfg["k"]
u, v = synth.uniform_graph(n, cfg["pairs"], 1235)
u_d, v_d = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
X = torch.from_numpy(synth.features(n, f, 1335)).to(dev)
A = gdr.sym_normalize(gdr.coo_to_csr(u_d, v_d, None, (n, n), symmetrize=True, binarize=True), 2)
_, tgt = gdr.propagate(A, X, 3, 0.8)
perm = torch.from_numpy(np.random.RandomState(1235).permutation(n)[:K].astype(np.int64)).to(dev)
C0 = tgt[perm].clone()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def fit(x):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    km = gdr.KMeans(n_clusters=K, init=C0, n_init=1, max_iter=20, tol=0, precision="tc").fit(x)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3

def series(name, body, reps=12):
    ts = [body() for _ in range(reps)]
    print(f"{name:46s} " + " ".join(f"{t:6.2f}" for t in ts))

series("fit only (graph)", lambda: fit(tgt))
_lib.call("gdr_debug_set", b"lloyd_graph", 0)
series("fit only (no graph)", lambda: fit(tgt))
_lib.call("gdr_debug_set", b"lloyd_graph", 1)
def with_flush():
    flush.fill_(1); return fit(tgt)
series("flush + fit (graph)", with_flush)
def with_stage12():
    A2 = gdr.sym_normalize(gdr.coo_to_csr(u_d, v_d, None, (n, n), symmetrize=True, binarize=True), 2)
    _, t2 = gdr.propagate(A2, X, 3, 0.8)
    return fit(t2)
series("stage1+2 then fit (graph)", with_stage12)
def with_all():
    flush.fill_(1)
    A2 = gdr.sym_normalize(gdr.coo_to_csr(u_d, v_d, None, (n, n), symmetrize=True, binarize=True), 2)
    _, t2 = gdr.propagate(A2, X, 3, 0.8)
    t = fit(t2)
    gdr.graph_compress(torch.zeros(n, dtype=torch.int32, device=dev), A2, [])
    return t
series("flush+stage1+2, fit, stage4 (graph)", with_all)
_lib.call("gdr_debug_set", b"lloyd_graph", 0)
series("flush+stage1+2, fit, stage4 (no graph)", with_all)
print("mem allocated MB", torch.cuda.memory_allocated() / 1e6, "reserved MB", torch.cuda.memory_reserved() / 1e6)

# does an NVML query between fits perturb the next fit?
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
_lib.call("gdr_debug_set", b"lloyd_graph", 1)
def with_nvml(kind):
    def body():
        t = with_all()
        if kind >= 1: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        if kind >= 2: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        if kind >= 3: pynvml.nvmlDeviceGetPowerUsage(h)
        return t
    return body
for kind, name in [(0, "no NVML"), (1, "+clock query"), (2, "+reasons query"), (3, "+power query"), (0, "no NVML again")]:
    series(f"step + NVML between steps: {name}", with_nvml(kind))

# bench.py-like step: events on the current stream around each stage + host wall time of fit
def bench_like():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    flush.fill_(1); torch.cuda.synchronize()
    ev[0].record()
    A2 = gdr.sym_normalize(gdr.coo_to_csr(u_d, v_d, None, (n, n), symmetrize=True, binarize=True), 2)
    ev[1].record()
    _, t2 = gdr.propagate(A2, X, 3, 0.8)
    ev[2].record()
    h0 = time.perf_counter()
    km = gdr.KMeans(n_clusters=K, init=C0, n_init=1, max_iter=20, tol=0, precision="tc")
    km.fit(t2)
    h1 = time.perf_counter()
    ev[3].record()
    gdr.graph_compress(km.labels_, A2, [])
    ev[4].record()
    torch.cuda.synchronize()
    return ev[2].elapsed_time(ev[3]), (h1 - h0) * 1e3
rows = [bench_like() for _ in range(12)]
print("bench-like  event ms:", " ".join(f"{a:6.2f}" for a, _ in rows))
print("bench-like  host  ms:", " ".join(f"{b:6.2f}" for _, b in rows))
