#!/usr/bin/env python
"""Golden vector for k-means++ seeding: sklearn's own kmeans_plusplus (the default init behind the
reference's KMeans(n_clusters=n) call, clustgdd_agent_transduct.py:105) with a fixed random_state.
Run in the build container:  python tests/golden/make_golden_kpp.py"""
import os
import numpy as np
from sklearn.cluster import kmeans_plusplus

OUT = os.path.dirname(os.path.abspath(__file__))
rs = np.random.RandomState(5)
cen = rs.randn(30, 6) * 4
X = (cen[rs.randint(0, 30, 4000)] + rs.randn(4000, 6)).astype(np.float32)
out = {}
for k, seed in [(25, 0), (60, 1), (200, 2)]:
    c, idx = kmeans_plusplus(X, k, random_state=np.random.RandomState(seed))
    out[f"k{k}_centers"], out[f"k{k}_indices"], out[f"k{k}_seed"] = c, idx, seed
np.savez_compressed(os.path.join(OUT, "kmeans_plusplus.npz"), x=X, **out)
print("written", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})
