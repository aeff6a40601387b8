"""Bipartite LightGCN propagation on the GPU kernels (stages 1-2, bipartite form).

Forward semantics of ``LightGCNCondensed.propagate`` (distill_recsys.py:319-353):
    deg_u = sum_e w ; deg_i = sum_e w
    norm  = w / (sqrt(deg_u[cu] + 1e-8) * sqrt(deg_i[ci] + 1e-8))
    L layers of  u <- sum norm * i ,  i <- sum norm * u  (simultaneous), output = layer mean.
The two scatter-adds per layer become two CSR SpMMs (the cu x ci matrix and its
transpose); autograd stays with torch (``BipartitePropagate`` implements the backward
with the transposed SpMMs for the embedding inputs).
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib
from ._dev import need_cuda, new_padded, padded_rows, ptr, stream
from .graph import CSR, coo_to_csr
from .propagation import spmm


class BipartiteGraph:
    """Device CSR of the cu x ci weight matrix, its transpose, and the normalised values."""

    def __init__(self, edge_index: torch.Tensor, edge_weight: torch.Tensor, num_u: int, num_i: int,
                 eps: float = 1e-8):
        need_cuda(edge_index, "edge_index")
        self.num_u, self.num_i = int(num_u), int(num_i)
        W = coo_to_csr(edge_index[0], edge_index[1], edge_weight, (self.num_u, self.num_i),
                       device=edge_index.device)
        WT, t_perm = W.transpose()
        dev = W.device
        self.deg_u = torch.empty(self.num_u, dtype=torch.float32, device=dev)
        self.deg_i = torch.empty(self.num_i, dtype=torch.float32, device=dev)
        norm = torch.empty(W.nnz, dtype=torch.float32, device=dev)
        t_norm = torch.empty(W.nnz, dtype=torch.float32, device=dev)
        _lib.call("gdr_bipartite_normalize", self.num_u, self.num_i, W.nnz, ptr(W.rowptr), ptr(W.colidx),
                  ptr(W.vals), ptr(WT.rowptr), ptr(t_perm), float(eps), ptr(norm), ptr(t_norm),
                  ptr(self.deg_u), ptr(self.deg_i), stream())
        self.A = CSR(W.rowptr, W.colidx, norm, W.shape)        # users <- items
        self.AT = CSR(WT.rowptr, WT.colidx, t_norm, WT.shape)  # items <- users
        self.W = W


def lightgcn_propagate(graph: BipartiteGraph, u0: torch.Tensor, i0: torch.Tensor,
                       num_layers: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Forward of distill_recsys.py:337-353 (no autograd)."""
    u = padded_rows(u0.detach().to(torch.float32))
    it = padded_rows(i0.detach().to(torch.float32))
    d = u.shape[1]
    dev = u.device
    u_acc = new_padded(graph.num_u, d, dev)
    i_acc = new_padded(graph.num_i, d, dev)
    u_acc.copy_(u)
    i_acc.copy_(it)
    for _ in range(int(num_layers)):
        u_next = spmm(graph.A, it, accumulate_into=u_acc, beta=1.0)
        i_next = spmm(graph.AT, u, accumulate_into=i_acc, beta=1.0)
        u, it = u_next, i_next
    inv = 1.0 / float(num_layers + 1)
    u_out = new_padded(graph.num_u, d, dev)
    i_out = new_padded(graph.num_i, d, dev)
    _lib.call("gdr_scale_rows", graph.num_u, d, inv, ptr(u_acc), u_acc.stride(0), ptr(u_out), u_out.stride(0), stream())
    _lib.call("gdr_scale_rows", graph.num_i, d, inv, ptr(i_acc), i_acc.stride(0), ptr(i_out), i_out.stride(0), stream())
    return u_out, i_out


class RankformerGCNGraph:
    """CSR of the raw user-item interactions with the Rankformer GCN weights
    (Rankformer/code/rec.py:92-140): du / di = interaction counts clamped to >= 1,
    w1 = 1/du^alpha/di^beta (items -> users), w2 = 1/du^beta/di^alpha (users -> items).
    Duplicate (u, i) lines are summed into the CSR value, which is what scattering every line
    separately adds up to."""

    def __init__(self, u: torch.Tensor, i: torch.Tensor, num_users: int, num_items: int, alpha: float = 1.0,
                 beta: float = 0.0):
        need_cuda(u, "u")
        self.n, self.m = int(num_users), int(num_items)
        R = coo_to_csr(u, i, None, (self.n, self.m), device=u.device)
        RT, t_perm = R.transpose()
        dev = R.device
        self.deg_u = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.deg_i = torch.empty(self.m, dtype=torch.float32, device=dev)
        w1 = torch.empty(R.nnz, dtype=torch.float32, device=dev)
        w2t = torch.empty(R.nnz, dtype=torch.float32, device=dev)
        scratch = torch.empty(max(R.nnz, 1), dtype=torch.float32, device=dev)
        _lib.call("gdr_bipartite_pow_normalize", self.n, self.m, R.nnz, ptr(R.rowptr), ptr(R.colidx), ptr(R.vals),
                  ptr(RT.rowptr), ptr(t_perm), float(alpha), float(beta), ptr(w1), ptr(w2t), ptr(self.deg_u),
                  ptr(self.deg_i), ptr(scratch), stream())
        self.A = CSR(R.rowptr, R.colidx, w1, R.shape)         # zu = A @ xi
        self.AT = CSR(RT.rowptr, RT.colidx, w2t, RT.shape)    # zi = AT @ xu


def rankformer_gcn_forward(graph: RankformerGCNGraph, x: torch.Tensor) -> torch.Tensor:
    """GCN.forward of Rankformer/code/rec.py:92-140: x = [users; items] -> [zu; zi]."""
    xu, xi = x[: graph.n], x[graph.n:]
    zu = spmm(graph.A, xi.contiguous())
    zi = spmm(graph.AT, xu.contiguous())
    return torch.cat([zu, zi], dim=0)


class BipartitePropagate(torch.autograd.Function):
    """Differentiable (w.r.t. the embeddings) wrapper: the backward of a layer-mean of
    alternating SpMMs is the same propagation with A and A^T swapped."""

    @staticmethod
    def forward(ctx, u0, i0, graph: BipartiteGraph, num_layers: int):
        ctx.graph, ctx.num_layers = graph, int(num_layers)
        return lightgcn_propagate(graph, u0, i0, num_layers)

    @staticmethod
    def backward(ctx, gu, gi):
        # out_u = mean_l U_l with U_{l+1} = A I_l, I_{l+1} = A^T U_l.  The adjoint recursion is
        # the same alternating propagation applied to (gu, gi) because [[0, A], [A^T, 0]] is symmetric.
        du, di = lightgcn_propagate(ctx.graph, gu.contiguous(), gi.contiguous(), ctx.num_layers)
        return du, di, None, None
