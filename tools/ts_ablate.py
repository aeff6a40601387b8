#!/usr/bin/env python
"""Role ablations of the A-in-TMEM first-level kernel at config E (tc_ablate bit mask: 1 no epilogue arithmetic, 2 no
tcgen05.ld, 4 no MMAs, 8 no centre stream), epilogue gate off so that garbage accumulators cost the same."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdr
from gdr import synth, _lib
from gdr._dev import padded_rows
from gdr.kmeans import TcOperand, assign_labels

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts[1:])


cfg = synth.CONFIGS["E"]
n, K = cfg["n"], cfg["k"]
for D in []:
    X = padded_rows(torch.from_numpy(synth.features(n, D, 1338)).to(dev))
    X = padded_rows((X - X.mean(0)).contiguous())
    perm = torch.from_numpy(np.random.RandomState(1235).permutation(n)[:K].astype(np.int64)).to(dev)
    C = padded_rows(X[perm].clone())
    op = TcOperand(X)
    ws = torch.empty(_lib.query("gdr_kmeans_assign_tc_ws_bytes", n, K, D), dtype=torch.uint8, device=dev)
    out = torch.empty(n, dtype=torch.int32, device=dev)
    for gate in (1, 0):
        _lib.call("gdr_debug_set", b"tc_gate", gate)
        for ab in ((-1,) if gate else (-1, 1, 2, 4, 8, 2 | 4, 2 | 8, 4 | 8, 2 | 4 | 8)):
            _lib.call("gdr_debug_set", b"tc_ablate", ab)
            t = timed(lambda: assign_labels(X, C, out, tc_operand=op, ws=ws))
            names = [nm for bit, nm in ((1, "no epi math"), (2, "no ld"), (4, "no MMA"), (8, "no stream")) if ab > 0 and ab & bit]
            print(f"D={D} gate {gate} level 1 [{', '.join(names) or 'full'}]: {t*1e3:.0f} us", flush=True)
    _lib.call("gdr_debug_set", b"tc_ablate", 0)
    _lib.call("gdr_debug_set", b"tc_gate", 1)

# in-kernel cycle accounts of CTA 0 (tc_ablate bit 16)
import ctypes
NAMES = ["mma total", "mma wait stage", "mma wait acc", "mma wait A", "mma issue+commit", "prod total", "prod wait free stage",
         "epi total", "epi wait acc", "epi ld+math", "epi arrive", "epi A store"]
D = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["f"]
X = padded_rows(torch.from_numpy(synth.features(n, D, 1338)).to(dev))
X = padded_rows((X - X.mean(0)).contiguous())
perm = torch.from_numpy(np.random.RandomState(1235).permutation(n)[:K].astype(np.int64)).to(dev)
C = padded_rows(X[perm].clone())
op = TcOperand(X)
ws = torch.empty(_lib.query("gdr_kmeans_assign_tc_ws_bytes", n, K, D), dtype=torch.uint8, device=dev)
out = torch.empty(n, dtype=torch.int32, device=dev)
for ab in (16, 16 | 2 | 4 | 8, 16 | 8, 16 | 4, 16 | 2):
    _lib.call("gdr_debug_set", b"tc_ablate", ab)
    t = timed(lambda: assign_labels(X, C, out, tc_operand=op, ws=ws))
    v = ctypes.c_int64()
    vals = []
    for i in range(12):
        _lib.call("gdr_debug_get", f"ts_probe_{i}".encode(), ctypes.addressof(v))
        vals.append(v.value)
    print(f"D={D} ablate {ab}: {t*1e3:.0f} us; kcycles: " + ", ".join(f"{nm} {x/1e3:.0f}" for nm, x in zip(NAMES, vals)), flush=True)
_lib.call("gdr_debug_set", b"tc_ablate", 0)
