#!/usr/bin/env python
"""Golden vectors for SURVEY §8(f) item 1 — edge scoring + per-class top-k sparsification —
produced by running the REFERENCE's own functions (utils_clustgdd.ER_estimator /
attaw_ER_estimator, ClustGDD.graph_sparse) in the build container.  Same stubbing of the
missing wheels as make_golden.py; nothing is copied from the reference.

    python tests/golden/make_golden_sparsify.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, _stub_missing_wheels  # noqa: E402


def coo_triplets(t):
    """(row, col, val) of a torch sparse tensor sorted row-major (the reference returns top-k order)."""
    t = t.coalesce()
    i, v = t._indices().numpy(), t._values().numpy()
    return i[0].astype(np.int64), i[1].astype(np.int64), v.astype(np.float32)


def main():
    sys.path.insert(0, REF)
    _stub_missing_wheels()
    import deep_robust_utils as dru
    import utils_clustgdd as uc
    import clustgdd_agent_transduct as agent
    import scipy.sparse as sp

    rng = np.random.RandomState(77)
    n, e, C = 400, 2600, 7
    r, c = rng.randint(0, n, e), rng.randint(0, n, e)
    keep = r != c
    M = sp.coo_matrix((np.ones(keep.sum()), (r[keep], c[keep])), shape=(n, n)).tocsr()
    M = M + M.T
    M.data[:] = 1.0
    adj, _ = dru.to_tensor(sp.csr_matrix(M), rng.randn(n, 3).astype(np.float32))
    adj_norm = dru.normalize_adj_tensor(adj, sparse=True)              # what the agent passes (transduct :49)
    ebd = torch.from_numpy((rng.randn(n, C) * 2).astype(np.float32))   # MLP-probe logits stand-in
    co = adj_norm.coalesce()
    src, dst = co._indices()[0], co._indices()[1]

    er = uc.ER_estimator(adj_norm, src, dst)                            # utils_clustgdd.py:151-162
    er_att, rew = uc.attaw_ER_estimator(adj_norm, ebd, src, dst)        # :165-184
    out = dict(n=n, C=C, src=src.numpy(), dst=dst.numpy(), val=co._values().numpy(), ebd=ebd.numpy(),
               er=er.numpy(), er_att=er_att.numpy(), rew_val=rew.coalesce()._values().numpy())
    ratio = 0.3
    out["ratio"] = ratio
    g = agent.ClustGDD.graph_sparse(None, adj_norm, ratio, sp_type="vanilla")      # transduct :131-153
    out["van_row"], out["van_col"], out["van_val"] = coo_triplets(g[0])
    gl = agent.ClustGDD.graph_sparse(None, adj_norm, ratio, ebd=ebd, sp_type="attaw")   # :155-183
    assert len(gl) == C
    for i, gi in enumerate(gl):
        out[f"att{i}_row"], out[f"att{i}_col"], out[f"att{i}_val"] = coo_triplets(gi)
    gs = agent.ClustGDD.graph_sparse(None, adj_norm, ratio, ebd=ebd, sp_type="single")  # :185-203
    out["sin_row"], out["sin_col"], out["sin_val"] = coo_triplets(gs[0])
    np.savez_compressed(os.path.join(OUT, "sparsify.npz"), **out)
    print("sparsify.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if not k.startswith("att") or k.startswith("att0")})


if __name__ == "__main__":
    main()
