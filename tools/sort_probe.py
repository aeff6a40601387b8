"""Stage-1 build at config E once (for `ncu -k regex:k_rs_` captures of the radix-sort passes) or timed with CUDA events."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import gdr
from gdr import synth

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "E"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = cfg["n"]
u, v = synth.uniform_graph(n, cfg["pairs"], 1238)
dev = torch.device("cuda:0")
u_d, v_d = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
from gdr import _lib
for mode in ((9, 10, 11) if reps > 1 else (9,)):
  _lib.call("gdr_debug_set", b"rs_max_bits", mode)
  print("rs_max_bits", mode)
  for r in range(reps):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    A = gdr.coo_to_csr(u_d, v_d, None, (n, n), symmetrize=True, binarize=True)
    e1.record()
    An = gdr.sym_normalize(A, 2)
    e2.record()
    torch.cuda.synchronize()
    print(f"  rep {r}: coo_to_csr {e0.elapsed_time(e1):.3f} ms, sym_normalize {e1.elapsed_time(e2):.3f} ms, nnz {An.nnz}")
lab = torch.from_numpy(np.random.RandomState(0).randint(0, cfg["k"], n).astype(np.int32)).to(dev)
for mode in ((9, 10, 11) if reps > 1 else (9,)):
  _lib.call("gdr_debug_set", b"rs_max_bits", mode)
  for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gdr.graph_compress(lab, An, [])
    e1.record()
    torch.cuda.synchronize()
    print(f"  rs_max_bits {mode} rep {r}: graph_compress {e0.elapsed_time(e1):.3f} ms")
_lib.call("gdr_debug_set", b"rs_max_bits", 0)
