#!/usr/bin/env python
"""Golden vector for SURVEY §8(f) item 3: compute_svd_embeddings (distill_recsys.py:124-155) run by the REFERENCE
on the subset of its shipped Ali-Display training file that the build-stage fixtures already hold, dim = 24.  Stored: the singular values (recovered from the column
norms of the embeddings: |user_emb[:, j]|^2 = sigma_j) and the first four embedding columns (well separated values:
determined up to sign).

    python tests/golden/make_golden_svd.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, _stub_missing_wheels  # noqa: E402


def main():
    sys.path.insert(0, REF)
    _stub_missing_wheels()
    import distill_recsys as dr
    sub = np.load(os.path.join(OUT, "recsys_ali_subset.npz"))          # the Ali-Display subset already in the fixtures
    nu, ni = int(sub["nu"]), int(sub["ni"])
    R = dr.build_interaction_matrix(nu, ni, sub["u"], sub["i"])       # distill_recsys.py:110-117
    dim = 24
    ue, ie = dr.compute_svd_embeddings(R, dim, seed=42)                # :124-155
    sigma = (ue.astype(np.float64) ** 2).sum(0)
    np.savez_compressed(os.path.join(OUT, "svd_ali.npz"), dim=dim, sigma=sigma,
                        user_emb4=ue[:, :4].astype(np.float32), item_emb4=ie[:, :4].astype(np.float32))
    print("sigma", sigma[:6], "...", sigma[-3:], ue.shape, ie.shape)


if __name__ == "__main__":
    main()
