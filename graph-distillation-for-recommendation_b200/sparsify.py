"""Edge scoring + top-k sparsification — the step between the k-means
stage and the coarsened graph.

Mirrors, with the reference's names and argument meaning,
  ``ER_estimator(adj, src, dst)``                  utils_clustgdd.py:151-162
  ``attaw_ER_estimator(adj, ebd, src, dst)``       utils_clustgdd.py:165-184
  ``ClustGDD.graph_sparse(adj, ratio, ebd, sp_type)``   clustgdd_agent_transduct.py:131-232
on the device: degrees, cosine re-weighting, class probabilities, edge weights, the top-k
threshold (radix select) and the rebuild of the kept edges all run in libgdr_b200 without a
host round trip; the reference copies to the host and rebuilds a scipy COO per class.

Differences a caller can observe: the returned sparse tensors are in row-major (coalesced)
order instead of top-k order, and among weights EQUAL to the k-th largest the first ones in
stored order are kept (torch.topk leaves that choice unspecified).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib
from ._dev import need_cuda, ptr, stream, workspace
from .graph import CSR


def _csr_of(adj) -> CSR:
    return adj if isinstance(adj, CSR) else CSR.from_torch_coo(adj)


def row_degree(A: CSR) -> torch.Tensor:
    """``adj @ ones`` in fp32, stored order."""
    deg = torch.empty(A.shape[0], dtype=torch.float32, device=A.device)
    _lib.call("gdr_row_sums_f32", A.shape[0], ptr(A.rowptr), ptr(A.vals), ptr(deg), stream())
    return deg


def er_lower(A: CSR) -> torch.Tensor:
    """Lower bound of the effective resistance of every stored edge: v/deg[src] + v/deg[dst]."""
    deg = torch.empty(A.shape[0], dtype=torch.float32, device=A.device)
    er = torch.empty(max(A.nnz, 1), dtype=torch.float32, device=A.device)
    _lib.call("gdr_er_lower", A.shape[0], A.nnz, ptr(A.rowptr), ptr(A.colidx), ptr(A.vals), ptr(deg), ptr(er), stream())
    return er[: A.nnz]


def cosine_reweight(A: CSR, ebd: torch.Tensor, eps: float = 1e-8) -> CSR:
    """values * F.cosine_similarity(ebd[src], ebd[dst], dim=-1): the re-weighted graph (same pattern)."""
    e = need_cuda(ebd, "ebd").to(torch.float32).contiguous()
    if e.dim() != 2 or e.shape[0] != A.shape[0]:
        raise ValueError("ebd must be [n_nodes, n_classes]")
    inv = torch.empty(A.shape[0], dtype=torch.float32, device=A.device)
    out = torch.empty(max(A.nnz, 1), dtype=torch.float32, device=A.device)
    _lib.call("gdr_edge_cosine_scale", A.shape[0], A.nnz, e.shape[1], ptr(A.rowptr), ptr(A.colidx), ptr(A.vals),
              ptr(e), e.stride(0), float(eps), ptr(inv), ptr(out), stream())
    return CSR(A.rowptr, A.colidx, out[: A.nnz], A.shape)


def softmax_rows(x: torch.Tensor) -> torch.Tensor:
    x = need_cuda(x, "ebd").to(torch.float32).contiguous()
    out = torch.empty_like(x)
    _lib.call("gdr_softmax_rows", x.shape[0], x.shape[1], ptr(x), x.stride(0), ptr(out), out.stride(0), stream())
    return out


def class_edge_weight(A: CSR, er: torch.Tensor, prob: torch.Tensor, cls: int) -> torch.Tensor:
    """(prob[src, cls] * prob[dst, cls]) * er."""
    w = torch.empty(max(A.nnz, 1), dtype=torch.float32, device=A.device)
    _lib.call("gdr_class_edge_weight", A.shape[0], A.nnz, ptr(A.rowptr), ptr(A.colidx), ptr(er), ptr(prob),
              prob.stride(0), int(cls), ptr(w), stream())
    return w[: A.nnz]


def topk_filter(A: CSR, weight: torch.Tensor, k: int, values: Optional[torch.Tensor] = None) -> CSR:
    """The k stored entries of A with the largest ``weight`` as a sorted CSR (values from ``values`` or A)."""
    k = int(k)
    if not 0 <= k <= A.nnz:
        raise ValueError("k out of range")   # torch.topk raises for k > nedges as well
    n, dev = A.shape[0], A.device
    vals = A.vals if values is None else values
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    colidx = torch.empty(max(k, 1), dtype=torch.int32, device=dev)
    vout = torch.empty(max(k, 1), dtype=torch.float32, device=dev)
    nnz_out = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = workspace(_lib.query("gdr_topk_filter_ws_bytes", n, A.nnz), dev)
    _lib.call("gdr_topk_filter_csr", n, A.nnz, ptr(A.rowptr), ptr(A.colidx), ptr(vals), ptr(weight.contiguous()), k,
              ptr(rowptr), ptr(colidx), ptr(vout), ptr(nnz_out), ptr(ws), ws.numel(), stream())
    return CSR(rowptr, colidx[:k], vout[:k], A.shape)   # exactly k entries by construction


def sparsify_classes(R: CSR, er: torch.Tensor, prob: torch.Tensor, k: int) -> List[CSR]:
    """One sparsified graph per class (column of ``prob``): top-k of (prob[src,c] * prob[dst,c]) * er, all classes
    in a single library call."""
    n, dev, C, k = R.shape[0], R.device, int(prob.shape[1]), int(k)
    if not 0 <= k <= R.nnz:
        raise ValueError("k out of range")
    rowptr = torch.empty((C, n + 1), dtype=torch.int32, device=dev)
    colidx = torch.empty((C, max(k, 1)), dtype=torch.int32, device=dev)
    vals = torch.empty((C, max(k, 1)), dtype=torch.float32, device=dev)
    nnz_out = torch.zeros(C, dtype=torch.int64, device=dev)
    ws = workspace(_lib.query("gdr_sparsify_classes_ws_bytes", n, R.nnz, C), dev)
    _lib.call("gdr_sparsify_classes", n, R.nnz, C, ptr(R.rowptr), ptr(R.colidx), ptr(R.vals), ptr(er), ptr(prob),
              prob.stride(0), k, ptr(rowptr), ptr(colidx), ptr(vals), ptr(nnz_out), ptr(ws), ws.numel(), stream())
    return [CSR(rowptr[c], colidx[c, :k], vals[c, :k], R.shape) for c in range(C)]


# ---------------------------------------------------------------------------------
# reference-signature functions
# ---------------------------------------------------------------------------------
def ER_estimator(adj, src=None, dst=None) -> torch.Tensor:
    """utils_clustgdd.py:151-162.  ``src`` / ``dst`` are accepted for signature compatibility: they are the
    coalesced indices of ``adj`` in the reference and are implied by the CSR here."""
    return er_lower(_csr_of(adj))


def attaw_ER_estimator(adj, ebd, src=None, dst=None):
    """utils_clustgdd.py:165-184.  Returns (ER_lower, reweighted_graph as a torch sparse COO tensor)."""
    R = cosine_reweight(_csr_of(adj), ebd)
    return er_lower(R), R.to_torch_coo()


def graph_sparse(adj, ratio: float, ebd: Optional[torch.Tensor] = None, sp_type: str = "vanilla") -> List[torch.Tensor]:
    """ClustGDD.graph_sparse (clustgdd_agent_transduct.py:131-232): list of sparsified graphs (torch sparse COO,
    row-major order) — one for 'vanilla' / 'single' / 'no_sp', one per class for 'attaw' / 'rand'."""
    if sp_type == "no_sp":
        return [adj]
    A = _csr_of(adj)
    k = int(A.nnz * ratio)
    if sp_type == "vanilla":
        return [topk_filter(A, er_lower(A), k).to_torch_coo()]
    if sp_type == "rand":
        # clustgdd_agent_transduct.py:208-230: one random edge subset per class, torch.randperm on the default
        # (CPU) generator exactly as the reference draws it; the kept entries are rebuilt on the device
        out = []
        for _ in range(ebd.shape[-1]):
            pick = torch.randperm(A.nnz)[:k].to(A.device)
            w = torch.zeros(A.nnz, dtype=torch.float32, device=A.device)
            w[pick] = 1.0
            out.append(topk_filter(A, w, k).to_torch_coo())
        return out
    if ebd is None:
        raise ValueError(f"sp_type={sp_type!r} needs ebd")
    R = cosine_reweight(A, ebd)
    er = er_lower(R)
    if sp_type == "single":
        return [topk_filter(R, er, k).to_torch_coo()]
    if sp_type == "attaw":
        return [g.to_torch_coo() for g in sparsify_classes(R, er, softmax_rows(ebd), k)]
    raise ValueError(f"unknown sp_type {sp_type!r}")
