#!/usr/bin/env python
"""Is the first-level kernel's MMA rate a per-SM limit or a chip-wide (power / memory system) one?
Times the first-level kernel alone (role ablations of gdr_debug_set "tc_ablate") on 1, 2, 8, 37, 74 and 148
row tiles — one tile per SM, so the number of tiles is the number of busy SMs — at the config-E width."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdr
from gdr import synth, _lib
from gdr._dev import padded_rows
from gdr.kmeans import TcOperand, assign_labels
dev = torch.device("cuda:0")
K, D = 10000, 100
X = torch.from_numpy(synth.features(148 * 128, D, 7)).to(dev)
X = padded_rows((X - X.mean(0)).contiguous())
C = padded_rows(X[torch.randperm(X.shape[0], device=dev)[:K] % X.shape[0]].clone()) if X.shape[0] >= K else None
C = padded_rows(torch.randn(K, D, device=dev))
_lib.call("gdr_debug_set", b"tc_screen", 3)
if True:
  for ab, what in ((2, "TMA + MMA, no epilogue"), (5, "MMA only (no centre stream)"), (6, "MMA only, no stage barriers / commits"), (8, "MMA only, no barriers at all, no epilogue warps"), (7, "same, issued as kind::f16 BF16 (garbage numerics)"), (3, "TMA only (no MMA)")):
      if ab in (7,): continue
      _lib.call("gdr_debug_set", b"tc_ablate", ab)
      line = []
      for tiles in (1, 148):
          n = tiles * 128
          Xs = X[:n]
          op = TcOperand(Xs)
          lab = torch.empty(n, dtype=torch.int32, device=dev)
          ws = torch.empty(_lib.query("gdr_kmeans_assign_tc_ws_bytes", n, K, D), dtype=torch.uint8, device=dev)
          ts = []
          for rep in range(5):
              a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
              a.record(); assign_labels(Xs, C, lab, tc_operand=op, ws=ws); b.record(); torch.cuda.synchronize()
              ts.append(a.elapsed_time(b))
          t = min(ts) * 1e3
          import ctypes
          cyc, ns, mm = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
          _lib.call("gdr_debug_get", b"tc_probe_cycles", ctypes.addressof(cyc))
          _lib.call("gdr_debug_get", b"tc_probe_ns", ctypes.addressof(ns))
          _lib.call("gdr_debug_get", b"tc_probe_mmas", ctypes.addressof(mm))
          line.append(f"{tiles} SMs: call {t:.0f} us; MMA thread of CTA 0: {ns.value / 1e3:.0f} us, {cyc.value / max(1, mm.value):.0f} clk/MMA over {mm.value} MMAs, SM clock {cyc.value / max(1, ns.value) * 1e3:.0f} MHz")
      print(f"ablate {ab} [{what}]: " + "  ".join(line), flush=True)
_lib.call("gdr_debug_set", b"tc_ablate", 0)
_lib.call("gdr_debug_set", b"tc_screen", 0)
