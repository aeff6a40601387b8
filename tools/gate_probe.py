#!/usr/bin/env python
"""First-level E-step kernel at a BASELINE shape: epilogue gate modes (tc_gate 0 / 1 / 2, with the converged labels as
hint) and role ablations (tc_ablate 1: tcgen05.ld but no epilogue math, 2: no tcgen05.ld either, 3: no MMA) for two
widths.  Run on the GPU box:  python tools/gate_probe.py [E|B]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdr
from gdr import synth, _lib
from gdr._dev import padded_rows
from gdr.kmeans import TcOperand, assign_labels, segment_sum

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "E"
cfg = synth.CONFIGS[name]
n, K = cfg["n"], cfg["k"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=4):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts[1:])


for D in (cfg["f"], cfg["d_logit"]):
    X = padded_rows(torch.from_numpy(synth.features(n, D, 1338)).to(dev))
    X = padded_rows((X - X.mean(0)).contiguous())
    perm = torch.from_numpy(np.random.RandomState(1235).permutation(n)[:K].astype(np.int64)).to(dev)
    C = padded_rows(X[perm].clone())
    op = TcOperand(X)
    ws = torch.empty(_lib.query("gdr_kmeans_assign_tc_ws_bytes", n, K, D), dtype=torch.uint8, device=dev)
    lab = torch.empty(n, dtype=torch.int32, device=dev)
    for it in range(3):      # a few Lloyd updates: centres become means
        assign_labels(X, C, lab, tc_operand=op, ws=ws)
        sums, counts = segment_sum(X, lab, K)
        C = padded_rows((sums / counts.clamp_min(1).unsqueeze(1)).contiguous())
    truth = lab.clone()
    assign_labels(X, C, truth, tc_operand=op, ws=ws)
    out = torch.empty(n, dtype=torch.int32, device=dev)
    nch = torch.zeros(1, dtype=torch.int32, device=dev)
    for gate in (0, 1, 2):
        _lib.call("gdr_debug_set", b"tc_gate", gate)
        for hint, hname in ((None, "no hint"), (truth, "true labels"), (lab, "previous labels")):
            if gate < 2 and hint is not None:
                continue
            t = timed(lambda: assign_labels(X, C, out, labels_prev=hint, n_changed=nch if hint is not None else None, tc_operand=op, ws=ws))
            print(f"D={D} gate {gate} ({hname}): whole E-step {t*1e3:.0f} us, labels equal {bool(torch.equal(out, truth))}", flush=True)
    for gate in (0, 2):
        _lib.call("gdr_debug_set", b"tc_gate", gate)
        for ab in (0, 1, 2, 3):
            _lib.call("gdr_debug_set", b"tc_ablate", ab if ab else -1)    # -1: no ablation, but return after level 1
            t = timed(lambda: assign_labels(X, C, out, labels_prev=truth, n_changed=nch, tc_operand=op, ws=ws))
            print(f"D={D} gate {gate} level-1 kernel only, ablate {ab}: {t*1e3:.0f} us", flush=True)
        _lib.call("gdr_debug_set", b"tc_ablate", 0)
    _lib.call("gdr_debug_set", b"tc_gate", 1)
    del X, C, op, ws
