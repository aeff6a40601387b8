#!/usr/bin/env python
"""Raw tcgen05.mma rate of ONE SM: cycles per instruction for 128 x N x 32B shapes (gdr_debug_mma_probe)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, gdr
from gdr import _lib
torch.zeros(1, device="cuda")
iters = 4000
for variant, what in ((8, "tf32, same two operand tiles"), (8 + 256, "tf32, production pattern 4-4-4-1"), (8 + 256 + 512, "tf32, production pattern 4-4-4-4"),
                      (8 + 256 + 16, "tf32, production pattern 4-4-4-1 + commits")):
    out = []
    for N in (64, 128, 256):
        c = ctypes.c_int64(0)
        _lib.call("gdr_debug_mma_probe", N, iters, variant, ctypes.addressof(c))
        _lib.call("gdr_debug_mma_probe", N, iters, variant, ctypes.addressof(c))
        out.append(f"N={N}: {c.value / iters:.1f} clk/MMA")
    print(f"{what:40s}" + "   ".join(out), flush=True)

c = ctypes.c_int64(0)
_lib.call("gdr_debug_mma_probe", 256, iters, 1024, ctypes.addressof(c))
_lib.call("gdr_debug_mma_probe", 256, iters, 1024, ctypes.addressof(c))
print(f"CTA pair, cta_group::2, 256 x 256 x 8 tf32, production pattern: {c.value / iters:.1f} clk/MMA (two SMs)")
