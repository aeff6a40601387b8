"""Stage 2 — K-hop feature propagation  target = (1-a) * sum_{t<T} (a*A_hat)^t X.

The reference has no function for this: it is the inline loop at
``clustgdd_agent_transduct.py:55-65`` (and three copies at
``clustgdd_agent_induct.py:67-94``).  ``propagate`` keeps those tensors-in /
tensors-out semantics; each hop is one launch of the sm_100a CSR SpMM with the
``target +=`` axpy fused into its epilogue.
"""
from __future__ import annotations

from typing import Tuple, Union

import torch

from . import _lib
from ._dev import need_cuda, new_padded, padded_rows, ptr, stream
from .graph import CSR


def _as_csr(adj: Union[CSR, torch.Tensor]) -> CSR:
    if isinstance(adj, CSR):
        return adj
    if isinstance(adj, torch.Tensor) and adj.layout == torch.sparse_coo:
        return CSR.from_torch_coo(adj)
    raise TypeError("adj must be a gdr CSR or a torch sparse COO tensor on a CUDA device")


def spmm(adj: Union[CSR, torch.Tensor], x: torch.Tensor, alpha: float = 1.0, out: torch.Tensor = None,
         accumulate_into: torch.Tensor = None, beta: float = 0.0) -> torch.Tensor:
    """Y = (alpha*A) @ X  [; accumulate_into += beta * Y].  X: dense f32 [n_cols, F]."""
    A = _as_csr(adj)
    need_cuda(x, "x")
    if x.dim() != 2 or x.shape[0] != A.shape[1]:
        raise ValueError(f"x must be [{A.shape[1]}, F], got {tuple(x.shape)}")
    xs = padded_rows(x.to(torch.float32))
    n, f = A.shape[0], x.shape[1]
    y = out if out is not None else new_padded(n, f, x.device)
    t = accumulate_into
    plan = A.spmm_plan()
    _lib.call("gdr_spmm_prop_planned", n, f, ptr(A.rowptr), ptr(A.colidx), ptr(A.vals), float(alpha),
              ptr(xs), xs.stride(0), ptr(y), y.stride(0), ptr(t), 0 if t is None else t.stride(0),
              float(beta), ptr(plan), plan.numel() - 1, stream())
    return y


def propagate(adj_norm: Union[CSR, torch.Tensor], features: torch.Tensor, prop_num: int,
              alpha: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """The loop of clustgdd_agent_transduct.py:59-65.

    for t in range(T): t == 0: prop = X, target = (1-alpha)*prop
                       else:   prop = alpha*adj_norm @ prop ; target = target + (1-alpha)*prop
    Returns ``(prop_feat, target_feat)`` — dense f32 [N, F] on the input's device.
    ``prop_num = T`` means T-1 hops, exactly as in the reference.
    """
    A = _as_csr(adj_norm)
    need_cuda(features, "features")
    T = int(prop_num)
    if T < 1:
        raise ValueError("prop_num must be >= 1")  # the reference would leave prop_feat undefined
    x = padded_rows(features.to(torch.float32))
    n, f = x.shape
    if A.shape[0] != A.shape[1] or A.shape[1] != n:
        raise ValueError("adj_norm must be [N, N] with N = features.shape[0]")
    one_minus = float(1.0 - alpha)
    target = new_padded(n, f, x.device)
    _lib.call("gdr_scale_rows", n, f, one_minus, ptr(x), x.stride(0), ptr(target), target.stride(0),
              stream())
    prop = x
    bufs = [new_padded(n, f, x.device) for _ in range(min(2, T - 1))]
    plan = A.spmm_plan() if T > 1 else None
    for t in range(1, T):
        y = bufs[(t - 1) % len(bufs)]
        _lib.call("gdr_spmm_prop_planned", n, f, ptr(A.rowptr), ptr(A.colidx), ptr(A.vals), float(alpha),
                  ptr(prop), prop.stride(0), ptr(y), y.stride(0), ptr(target), target.stride(0),
                  one_minus, ptr(plan), plan.numel() - 1, stream())
        prop = y
    if T == 1:
        prop = features
    return prop, target
