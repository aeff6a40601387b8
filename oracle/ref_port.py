"""CPU port of the reference's call sites for this path — TEST / BASELINE INFRASTRUCTURE.

The reference (/root/reference, pure Python) cannot travel to the GPU box, so the CPU arm of
``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg) times THIS port instead:
the same expressions the reference evaluates, on the same third-party libraries that own
its arithmetic (scipy sparsetools, torch CPU sparse, scikit-learn), with every host thread
those libraries choose to use.  cpu_baseline.kind = "port".

Each function restates (not copies) the call site it cites; ``tests/test_ref_port.py``
checks them against the golden vectors generated from the real reference.
Only tests/ and bench.py may import this module.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch


def to_tensor_sparse(adj_csr: sp.csr_matrix) -> torch.Tensor:
    """deep_robust_utils.py:389-396: scipy -> torch sparse COO f32 (CPU)."""
    coo = adj_csr.tocoo().astype(np.float32)
    idx = torch.from_numpy(np.vstack((coo.row, coo.col)).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(coo.data), torch.Size(coo.shape))


def build_adjacency(u: np.ndarray, v: np.ndarray, n: int) -> sp.csr_matrix:
    """utils.py:66-67 + utils_graphsaint.py:20-22: CSR of ones, A + A^T, clip to 1."""
    A = sp.csr_matrix((np.ones(u.shape[0]), (u, v)), shape=(n, n))
    A.data[:] = 1.0
    A = A + A.T
    A[A > 1] = 1
    return sp.csr_matrix(A)


def normalize_adj_tensor_sparse(adj: torch.Tensor) -> torch.Tensor:
    """deep_robust_utils.py:245-256 (sparse=True): to_scipy -> normalize_adj -> back."""
    vals, idx = adj._values().numpy(), adj._indices().numpy()
    mx = sp.csr_matrix((vals, idx), shape=tuple(adj.shape))          # :408-417
    mx = mx.tolil()                                                    # :197-198
    if mx[0, 0] == 0:                                                  # :199-200
        mx = mx + sp.eye(mx.shape[0])
    rowsum = np.array(mx.sum(1))                                       # :201
    with np.errstate(divide="ignore"):
        r_inv = np.power(rowsum, -1 / 2).flatten()                     # :202
    r_inv[np.isinf(r_inv)] = 0.0                                       # :203
    r_mat_inv = sp.diags(r_inv)                                        # :204
    mx = r_mat_inv.dot(mx).dot(r_mat_inv)                              # :205-206
    return to_tensor_sparse(mx)


def propagate(adj_norm: torch.Tensor, features: torch.Tensor, prop_num: int, alpha: float):
    """clustgdd_agent_transduct.py:59-65 on torch CPU sparse."""
    prop_feat = target_feat = None
    for t in range(prop_num):
        if t == 0:
            prop_feat = features
            target_feat = (1 - alpha) * prop_feat
        else:
            prop_feat = alpha * adj_norm @ prop_feat
            target_feat = target_feat + (1 - alpha) * prop_feat
    return prop_feat, target_feat


def kmeans_fit(X: np.ndarray, C0: np.ndarray, max_iter: int, tol: float = 0.0):
    """clustgdd_agent_transduct.py:105 with the initialisation pinned:
    KMeans(n_clusters=K, init=C0, n_init=1, algorithm='lloyd').fit(X)."""
    from sklearn.cluster import KMeans
    return KMeans(n_clusters=C0.shape[0], init=C0, n_init=1, max_iter=max_iter, tol=tol, algorithm="lloyd").fit(X)


def cluster_means(target_feat: torch.Tensor, labels: np.ndarray, n: int) -> torch.Tensor:
    """clustgdd_agent_transduct.py:116-125 (the O(n*N) loop, as written)."""
    lab = torch.FloatTensor(labels)
    return torch.stack([target_feat[torch.where(lab == i)[0]].mean(dim=0) for i in range(n)], dim=0)


def graph_compress_sparse(labels: np.ndarray, adj_norm: torch.Tensor) -> sp.csr_matrix:
    """clustgdd_agent_transduct.py:234-250 with a SPARSE one-hot P (the dense N x n one-hot of
    the reference is 1.35 GB at config B and 196 GB at config E, so this stand-in is what
    BASELINE.md §4 prescribes): S = P^T A P, P = onehot / colsum, diagonal removed."""
    n = int(labels.max()) + 1
    N = labels.shape[0]
    sizes = np.bincount(labels, minlength=n).astype(np.float32)
    P = sp.csr_matrix((1.0 / sizes[labels], (np.arange(N), labels)), shape=(N, n), dtype=np.float32)
    A = sp.csr_matrix((adj_norm._values().numpy(), adj_norm._indices().numpy()), shape=tuple(adj_norm.shape))
    S = (P.T @ A @ P).tolil()
    S.setdiag(0)
    S = S.tocsr()
    S.eliminate_zeros()
    return S


def build_condensed_bipartite(train_u, train_i, u2cu, i2ci, num_cu, num_ci) -> sp.csr_matrix:
    """distill_recsys.py:184-201."""
    cu, ci = u2cu[train_u], i2ci[train_i]
    C = sp.coo_matrix((np.ones_like(cu, dtype=np.float32), (cu, ci)), shape=(num_cu, num_ci))
    C.sum_duplicates()
    return C.tocsr()


def build_interaction_matrix(num_users: int, num_items: int, u: np.ndarray, i: np.ndarray) -> sp.csr_matrix:
    """distill_recsys.py:110-117: COO of ones -> CSR (duplicate lines summed by tocsr)."""
    vals = np.ones_like(u, dtype=np.float32)
    return sp.coo_matrix((vals, (u, i)), shape=(num_users, num_items)).tocsr()


def lightgcn_propagate(cu: torch.Tensor, ci: torch.Tensor, w: torch.Tensor, u0: torch.Tensor, i0: torch.Tensor,
                       num_layers: int):
    """distill_recsys.py:329-353 on torch CPU: index_add_ degrees, edge norm, L simultaneous layers, layer mean."""
    deg_u = torch.zeros(u0.shape[0]).index_add_(0, cu, w)
    deg_i = torch.zeros(i0.shape[0]).index_add_(0, ci, w)
    norm = w / (torch.sqrt(deg_u[cu] + 1e-8) * torch.sqrt(deg_i[ci] + 1e-8))
    u, it = u0, i0
    u_layers, i_layers = [u], [it]
    for _ in range(num_layers):
        u_msg = torch.zeros_like(u).index_add_(0, cu, it[ci] * norm.unsqueeze(1))
        i_msg = torch.zeros_like(it).index_add_(0, ci, u[cu] * norm.unsqueeze(1))
        u, it = u_msg, i_msg
        u_layers.append(u)
        i_layers.append(it)
    return torch.stack(u_layers, dim=0).mean(dim=0), torch.stack(i_layers, dim=0).mean(dim=0)


def standard_scale(X: np.ndarray) -> np.ndarray:
    """distill_recsys.py:172."""
    from sklearn.preprocessing import StandardScaler
    return StandardScaler(with_mean=True, with_std=True).fit_transform(X)


def er_estimator(adj: torch.Tensor, src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """utils_clustgdd.py:151-162."""
    degree = adj @ torch.ones(adj.shape[0])
    values = adj.coalesce().values()
    return values / degree[src] + values / degree[dst]


def attaw_er_estimator(adj: torch.Tensor, ebd: torch.Tensor, src: torch.Tensor, dst: torch.Tensor):
    """utils_clustgdd.py:165-184 (cosine re-weighting, scipy COO rebuild, ER on the re-weighted graph)."""
    import torch.nn.functional as F
    values = adj.coalesce()._values() * F.cosine_similarity(ebd[src], ebd[dst], dim=-1)
    rew = to_tensor_sparse(sp.coo_matrix((values.numpy(), (src.numpy(), dst.numpy())), shape=tuple(adj.shape)))
    degree = rew @ torch.ones(adj.shape[0])
    values = rew.coalesce().values()
    return values / degree[src] + values / degree[dst], rew


def graph_sparse_attaw(adj: torch.Tensor, ratio: float, ebd: torch.Tensor, max_classes: int = None):
    """clustgdd_agent_transduct.py:155-183 ('attaw'): per class  src_prob * dst_prob * ER_low, torch.topk,
    scipy COO rebuild -> torch sparse.  ``max_classes`` bounds the loop for the timed CPU sample."""
    import torch.nn.functional as F
    co = adj.coalesce()
    src, dst = co._indices()[0], co._indices()[1]
    nedges = co._values().shape[0]
    er_low, rew = attaw_er_estimator(adj, ebd, src, dst)
    prob = F.softmax(ebd, dim=-1)
    k = int(nedges * ratio)
    out = []
    for i in range(prob.shape[-1] if max_classes is None else min(max_classes, prob.shape[-1])):
        values = rew.coalesce()._values()
        cp = prob[:, i]
        w = cp[src] * cp[dst] * er_low
        _, idx = torch.topk(w, k)
        g = sp.coo_matrix((values[idx].numpy(), (src[idx].numpy(), dst[idx].numpy())), shape=tuple(adj.shape))
        out.append(to_tensor_sparse(g))
    return out


def host_info() -> dict:
    import os
    info = dict(cpu_count=os.cpu_count(), affinity=len(os.sched_getaffinity(0)), torch_threads=torch.get_num_threads())
    try:
        from threadpoolctl import threadpool_info
        info["threadpools"] = [dict(api=p.get("user_api"), threads=p.get("num_threads")) for p in threadpool_info()]
    except Exception:
        pass
    return info
