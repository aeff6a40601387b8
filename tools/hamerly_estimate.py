#!/usr/bin/env python
"""CPU estimate (numpy / scipy only) of how many rows Hamerly-style bounds could skip per Lloyd iteration at a BASELINE shape
(VERDICT r1 item 9): upper bound = distance to the own centre + its shift, lower bound = second-best distance - largest
other shift.  python tools/hamerly_estimate.py B"""
import sys, importlib.util, numpy as np, scipy.sparse as sp, time
spec = importlib.util.spec_from_file_location("synth", "/root/repo/graph-distillation-for-recommendation_b200/synth.py"); synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
name = sys.argv[1]; scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cfg = synth.CONFIGS[name]
n = int(cfg["n"] * scale); K = int(cfg["k"] * scale); F = cfg["f"]; pairs = int(cfg["pairs"] * scale)
u, v = synth.uniform_graph(n, pairs, 1238)
A = sp.coo_matrix((np.ones(len(u), np.float32), (u, v)), shape=(n, n)).tocsr(); A = A + A.T; A.data[:] = 1; A = A + sp.eye(n, format="csr", dtype=np.float32)
d = np.asarray(A.sum(1)).ravel(); r = 1 / np.sqrt(d); A = sp.diags(r) @ A @ sp.diags(r)
X = synth.features(n, F, 1338).astype(np.float32)
H = X.copy(); T = 0.2 * X
for _ in range(cfg["hops"]): H = 0.8 * (A @ H); T = T + H * 0.2
X = (T - T.mean(0)).astype(np.float32)
rng = np.random.RandomState(1235); C = X[rng.permutation(n)[:K]].copy()
def assign(X, C):
    cn = (C * C).sum(1)
    best = np.empty(n, np.float32); sec = np.empty(n, np.float32); lab = np.empty(n, np.int64)
    for s in range(0, n, 20000):
        D = cn[None, :] - 2 * X[s:s+20000] @ C.T
        idx = np.argpartition(D, 1, axis=1)[:, :2]
        dd = np.take_along_axis(D, idx, 1); o = np.argsort(dd, 1)
        idx = np.take_along_axis(idx, o, 1); dd = np.take_along_axis(dd, o, 1)
        lab[s:s+20000] = idx[:, 0]; best[s:s+20000] = dd[:, 0]; sec[s:s+20000] = dd[:, 1]
    return lab, best, sec
xn = (X * X).sum(1)
lab_prev = None
for it in range(20):
    t0 = time.time()
    lab, best, sec = assign(X, C)
    ub = np.sqrt(np.maximum(best + xn, 0)); lb = np.sqrt(np.maximum(sec + xn, 0))
    Cn = np.zeros_like(C); cnt = np.bincount(lab, minlength=K)
    np.add.at(Cn, lab, X); nz = cnt > 0; Cn[nz] /= cnt[nz, None]; Cn[~nz] = C[~nz]
    delta = np.sqrt(((Cn - C) ** 2).sum(1))
    o = np.argsort(delta)[::-1]; m1, m2 = delta[o[0]], delta[o[1]]
    dmax_other = np.where(lab == o[0], m2, m1)
    skip = (ub + delta[lab]) < (lb - dmax_other)
    changed = -1 if lab_prev is None else int((lab != lab_prev).sum())
    print(f"it {it}: changed {changed}  hamerly-skippable next iter {skip.mean():.3f}  max delta {m1:.3f} median delta {np.median(delta):.4f} median gap {np.median(lb-ub):.4f}  ({time.time()-t0:.1f}s)", flush=True)
    lab_prev = lab; C = Cn
