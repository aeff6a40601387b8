#!/usr/bin/env python
"""First-level E-step kernel with the row tile in tensor memory (k_assign_tc_ts) against the shared-memory-operand form
(tc_screen 3): whole E-step and level-1-only times, labels against the exact fp32 kernel.
Run on the GPU box:  python tools/ts_probe.py [B|E] ..."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdr
from gdr import synth, _lib
from gdr._dev import padded_rows
from gdr.kmeans import TcOperand, assign_labels, segment_sum

dev = torch.device("cuda:0")
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


for name in (sys.argv[1:] or ["B", "E"]):
    cfg = synth.CONFIGS[name]
    n, K = cfg["n"], cfg["k"]
    for D in (cfg["f"], cfg["d_logit"], 64):
        X = padded_rows(torch.from_numpy(synth.features(n, D, 1338)).to(dev))
        X = padded_rows((X - X.mean(0)).contiguous())
        perm = torch.from_numpy(np.random.RandomState(1235).permutation(n)[:K].astype(np.int64)).to(dev)
        C = padded_rows(X[perm].clone())
        op = TcOperand(X)
        ws = torch.empty(_lib.query("gdr_kmeans_assign_tc_ws_bytes", n, K, D), dtype=torch.uint8, device=dev)
        lab = torch.empty(n, dtype=torch.int32, device=dev)
        _lib.call("gdr_debug_set", b"tc_screen", 3)
        for it in range(2):
            assign_labels(X, C, lab, tc_operand=op, ws=ws)
            sums, counts = segment_sum(X, lab, K)
            C = padded_rows((sums / counts.clamp_min(1).unsqueeze(1)).contiguous())
        truth = torch.empty(n, dtype=torch.int32, device=dev)
        assign_labels(X, C, truth)
        out = torch.empty(n, dtype=torch.int32, device=dev)
        for screen, sname in ((3, "SS 128x256"), (6, "TS 128x192")):
            _lib.call("gdr_debug_set", b"tc_screen", screen)
            out.fill_(-1)
            t = timed(lambda: assign_labels(X, C, out, tc_operand=op, ws=ws))
            eq = bool(torch.equal(out, truth))
            _lib.call("gdr_debug_set", b"tc_ablate", -1)
            t1 = timed(lambda: assign_labels(X, C, out, tc_operand=op, ws=ws))
            _lib.call("gdr_debug_set", b"tc_ablate", 3)
            t3 = timed(lambda: assign_labels(X, C, out, tc_operand=op, ws=ws))
            _lib.call("gdr_debug_set", b"tc_ablate", 0)
            print(f"{name} D={D} {sname}: E-step {t*1e3:.0f} us, level 1 {t1*1e3:.0f} us (no MMA {t3*1e3:.0f}), labels == exact fp32: {eq}"
                  f"{'' if eq else ' differ ' + str(int((out != truth).sum()))}", flush=True)
        _lib.call("gdr_debug_set", b"tc_screen", 0)
