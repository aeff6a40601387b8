"""Synthetic inputs of the BASELINE.json shapes.  Host numpy generators with a
fixed seed so that the GPU path and the CPU oracle / reference arm get identical arrays."""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # id: (nodes, undirected/directed input pairs, features, hops, clusters)
    "A": dict(name="cora", n=2708, pairs=5278, f=1433, hops=2, k=140, d_logit=7),
    "B": dict(name="ogbn-arxiv", n=169343, pairs=1166243, f=128, hops=2, k=1000, d_logit=40),
    "E": dict(name="ogbn-products", n=2449029, pairs=61859140, f=100, hops=3, k=10000, d_logit=47),
}
BIPARTITE = {
    "C": dict(name="yelp2018", users=31668, items=38048, inter=1561406, d=64, layers=2),
    "D": dict(name="amazon-book", users=52643, items=91599, inter=2984108, d=64, layers=2),
}


def uniform_graph(n: int, pairs: int, seed: int):
    """`pairs` (u, v) draws with both endpoints uniform in [0, n); self pairs dropped.  The
    builder symmetrises + dedupes; node 0 has no self-loop, so the reference adds I."""
    rs = np.random.RandomState(seed)
    u = rs.randint(0, n, pairs).astype(np.int64)
    v = rs.randint(0, n, pairs).astype(np.int64)
    keep = u != v
    return u[keep], v[keep]


def skewed_graph(n: int, pairs: int, seed: int):
    """Hub-biased variant: half of the destinations are floor(n * r^3)."""
    rs = np.random.RandomState(seed)
    u = rs.randint(0, n, pairs).astype(np.int64)
    v = rs.randint(0, n, pairs).astype(np.int64)
    hub = rs.rand(pairs) < 0.5
    v[hub] = np.floor(n * rs.rand(int(hub.sum())) ** 3).astype(np.int64)
    keep = u != v
    return u[keep], v[keep]


def bipartite_interactions(users: int, items: int, inter: int, seed: int, dup_rate: float = 0.04):
    """user uniform, item Zipf(1.1) over a random item permutation, ~dup_rate duplicate lines."""
    rs = np.random.RandomState(seed)
    n_unique = int(inter * (1 - dup_rate))
    u = rs.randint(0, users, n_unique).astype(np.int64)
    z = rs.zipf(1.1, n_unique)
    perm = rs.permutation(items)
    i = perm[(z - 1) % items].astype(np.int64)
    d = rs.randint(0, n_unique, inter - n_unique)
    u = np.concatenate([u, u[d]])
    i = np.concatenate([i, i[d]])
    p = rs.permutation(inter)
    return u[p], i[p]


def features(n: int, f: int, seed: int, kind: str = "zscore") -> np.ndarray:
    rs = np.random.RandomState(seed)
    X = rs.standard_normal((n, f)).astype(np.float32)
    if kind == "l1":  # Cora-like bag of words rows
        X = np.abs(X)
        X /= X.sum(axis=1, keepdims=True)
    elif kind == "zscore":
        X = (X - X.mean(axis=0)) / X.std(axis=0)
    return np.ascontiguousarray(X, dtype=np.float32)


def clustered_features(n: int, d: int, k: int, seed: int, spread: float = 1.0) -> np.ndarray:
    """Mixture of k Gaussians (what probe logits / SVD embeddings look like to k-means)."""
    rs = np.random.RandomState(seed)
    cen = rs.standard_normal((k, d)).astype(np.float32) * 3
    X = cen[rs.randint(0, k, n)] + spread * rs.standard_normal((n, d)).astype(np.float32)
    return np.ascontiguousarray(X, dtype=np.float32)


def kmeans_init(X: np.ndarray, k: int, seed: int) -> np.ndarray:
    """C0 = X[RandomState(seed).permutation(N)[:K]]."""
    return X[np.random.RandomState(seed).permutation(X.shape[0])[:k]].copy()
