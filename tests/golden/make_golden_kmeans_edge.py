#!/usr/bin/env python
"""Golden vector for the degenerate corner of sklearn's Lloyd step that the reference reaches whenever
n_clusters exceeds the number of distinct rows (clustgdd_agent_transduct.py:105 on duplicated logits):
relocation is skipped (all distances are zero, _k_means_common.pyx:193-196) and _average_centers
(:274-295) fills the empty clusters IN PLACE — an empty cluster below the largest one receives that
cluster's raw SUM, one above it the MEAN.  Run in the build container:
    python tests/golden/make_golden_kmeans_edge.py"""
import os
import warnings
import numpy as np
from sklearn.cluster import KMeans

OUT = os.path.dirname(os.path.abspath(__file__))
pts = np.array([[1.0, 2.0, -1.0], [4.0, -3.0, 0.5], [-2.5, 0.25, 3.0], [0.5, 5.0, 2.0]], dtype=np.float32)
mult = [3, 5, 2, 4]
X = np.concatenate([np.repeat(pts[i:i + 1], m, axis=0) for i, m in enumerate(mult)]).astype(np.float32)
X = X[np.random.RandomState(0).permutation(X.shape[0])]
far = np.array([[40.0, 40.0, 40.0]], dtype=np.float32)
C0 = np.concatenate([far, pts, pts[1:2]]).astype(np.float32)          # cluster 0 and cluster 5 stay empty
out = dict(x=X, c0=C0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for it in (1, 2, 5):
        km = KMeans(n_clusters=6, init=C0, n_init=1, max_iter=it, tol=0, algorithm="lloyd").fit(X)
        out[f"it{it}_centers"], out[f"it{it}_labels"] = km.cluster_centers_, km.labels_
        out[f"it{it}_n_iter"], out[f"it{it}_inertia"] = km.n_iter_, km.inertia_
np.savez_compressed(os.path.join(OUT, "kmeans_edge.npz"), **out)
for it in (1, 2, 5):
    print(it, out[f"it{it}_n_iter"], out[f"it{it}_inertia"], "\n", out[f"it{it}_centers"], out[f"it{it}_labels"])
