"""CPU oracle of the distillation core — TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product package
(``graph-distillation-for-recommendation_b200/``) never does.

Each function restates, in plain numpy (integer / index work) or through the plain-C
loops of ``oracle.c`` (floating-point inner loops), the algorithm of the reference call
site it cites.  Paths are relative to ``/root/reference/ClustGDD``; ``sklearn/`` is
scikit-learn's ``sklearn/cluster`` (third-party; the reference pins 1.3.2 in
README.md:14, this image has 1.9.0 — the Lloyd code is the same).

PINNING: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against outputs of the reference's OWN functions run in the build
container: ``tests/golden/make_golden.py`` imports /root/reference + the installed
scikit-learn/scipy/torch, writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks this module against those files on CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc-compile oracle.c (no FMA contraction, OpenMP over rows only)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", src,
                               "-o", _SO, "-lm"])
    return _SO


def _c():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_kmeans_finalize.restype = C.c_double
        _lib.oracle_inertia.restype = C.c_double
        _lib.oracle_inertia_f64.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


I64 = C.c_int64


# =====================================================================================
# Stage 1
# =====================================================================================
def coo_to_csr(row, col, val, shape, symmetrize=False, binarize=False):
    """COO -> CSR, duplicates summed, columns sorted.

    scipy ``sp.csr_matrix((ones, (r, c)))`` utils.py:66-67; ``coo_matrix(...).tocsr()``
    distill_recsys.py:116-117; ``adj + adj.T; adj[adj > 1] = 1`` utils_graphsaint.py:20-22
    (= symmetrize + binarize).  Returns (rowptr int32, colidx int32, vals float32)."""
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    n_rows, n_cols = int(shape[0]), int(shape[1])
    if row.size and (row.min() < 0 or row.max() >= n_rows or col.min() < 0 or col.max() >= n_cols):
        raise ValueError("row/col index exceeds matrix dimensions")
    val = np.ones(row.shape[0], dtype=np.float32) if val is None else np.asarray(val, dtype=np.float32)
    if symmetrize:
        r2 = np.stack([row, col], axis=1).reshape(-1)
        c2 = np.stack([col, row], axis=1).reshape(-1)
        row, col, val = r2, c2, np.repeat(val, 2)
    if row.size == 0:
        return np.zeros(n_rows + 1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32)
    key = row * n_cols + col
    order = np.argsort(key, kind="stable")
    key, val = key[order], val[order]
    head = np.ones(key.shape[0], dtype=bool)
    head[1:] = key[1:] != key[:-1]
    starts = np.flatnonzero(head)
    ukey = key[starts]
    # sequential fp32 sum of each run in input order
    sums = np.add.reduceat(val, starts).astype(np.float32) if not _has_long_runs(starts, key.shape[0]) \
        else _seq_run_sums(val, starts, key.shape[0])
    if binarize:
        sums = np.ones_like(sums)
    urow = (ukey // n_cols).astype(np.int64)
    ucol = (ukey % n_cols).astype(np.int32)
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(rowptr, urow + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return rowptr, ucol, sums.astype(np.float32)


def _has_long_runs(starts, n):
    lens = np.diff(np.append(starts, n))
    return lens.size and lens.max() > 8  # reduceat is pairwise above 8 elements


def _seq_run_sums(val, starts, n):
    ends = np.append(starts[1:], n)
    out = np.empty(starts.shape[0], dtype=np.float32)
    for p, (b, e) in enumerate(zip(starts, ends)):
        s = np.float32(0)
        for v in val[b:e]:
            s = np.float32(s + v)
        out[p] = s
    return out


def csr_transpose(rowptr, colidx, shape):
    """(t_rowptr, t_colidx, t_perm): transposed structure, rows ascending inside a column."""
    n_rows, n_cols = shape
    rows = np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(rowptr))
    order = np.argsort(colidx.astype(np.int64), kind="stable")
    t_rowptr = np.zeros(n_cols + 1, dtype=np.int64)
    np.add.at(t_rowptr, colidx.astype(np.int64) + 1, 1)
    return np.cumsum(t_rowptr).astype(np.int32), rows[order].astype(np.int32), order.astype(np.int32)


def sym_normalize(rowptr, colidx, vals, n, self_loop_mode=2):
    """D^-1/2 (A [+ I]) D^-1/2.   deep_robust_utils.py:180-207 via :245-256 / :408-417.

    * ``if mx[0, 0] == 0: mx = mx + sp.eye(n)`` (:199-200) — mode 2; the sum is float64.
    * rowsum = mx.sum(1); r_inv = rowsum ** -0.5; inf -> 0 (:201-203)
    * (diag(r_inv) . mx) . diag(r_inv) (:204-206), cast to fp32 by
      sparse_mx_to_torch_sparse_tensor (:391).
    Without the identity the matrix stays float32 in the reference and so does the arithmetic.
    Returns (rowptr, colidx, vals f32, deg f64)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    colidx = np.asarray(colidx, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float32)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    if self_loop_mode == 0:
        add = False
    elif self_loop_mode == 1:
        add = True
    else:
        a00 = vals[(rows == 0) & (colidx == 0)].sum() if vals.size else 0.0
        add = bool(a00 == 0)
    if add:
        has_diag = np.zeros(n, dtype=bool)
        has_diag[rows[rows == colidx]] = True
        v64 = vals.astype(np.float64)
        v64[rows == colidx] += 1.0
        new_r = np.flatnonzero(~has_diag)
        r_all = np.concatenate([rows, new_r])
        c_all = np.concatenate([colidx, new_r])
        v_all = np.concatenate([v64, np.ones(new_r.shape[0])])
        order = np.lexsort((c_all, r_all))
        r_all, c_all, v_all = r_all[order], c_all[order], v_all[order]
        deg = np.bincount(r_all, weights=v_all, minlength=n)
        with np.errstate(divide="ignore"):
            r_inv = np.power(deg, -0.5)
        r_inv[np.isinf(r_inv)] = 0.0
        out = ((r_inv[r_all] * v_all) * r_inv[c_all]).astype(np.float32)
        rp = np.zeros(n + 1, dtype=np.int64)
        np.add.at(rp, r_all + 1, 1)
        return np.cumsum(rp).astype(np.int32), c_all.astype(np.int32), out, deg
    deg32 = np.zeros(n, dtype=np.float32)
    np.add.at(deg32, rows, vals)
    with np.errstate(divide="ignore"):
        r_inv = np.power(deg32, np.float32(-0.5)).astype(np.float32)
    r_inv[np.isinf(r_inv)] = 0.0
    out = ((r_inv[rows] * vals).astype(np.float32) * r_inv[colidx]).astype(np.float32)
    return rowptr.astype(np.int32), colidx.astype(np.int32), out, deg32.astype(np.float64)


def sym_normalize_dense(A):
    """deep_robust_utils.py:257-264, fp32:  mx = A + I ; r = rowsum^-1/2 ; inf -> 0 ;
    diag(r) @ mx @ diag(r)."""
    A = np.asarray(A, dtype=np.float32)
    mx = A + np.eye(A.shape[0], dtype=np.float32)
    rowsum = mx.sum(1, dtype=np.float32)
    with np.errstate(divide="ignore"):
        r = (np.float32(1) / np.sqrt(rowsum)).astype(np.float32)
    r[np.isinf(r)] = 0
    return ((r[:, None] * mx).astype(np.float32) * r[None, :]).astype(np.float32)


def bipartite_normalize(rowptr, colidx, w, n_u, n_i, eps=1e-8):
    """distill_recsys.py:329-335 in fp32, scatter-sums in edge (row-major) order."""
    rows = np.repeat(np.arange(n_u, dtype=np.int64), np.diff(rowptr))
    w = np.asarray(w, dtype=np.float32)
    deg_u = np.zeros(n_u, dtype=np.float32)
    deg_i = np.zeros(n_i, dtype=np.float32)
    np.add.at(deg_u, rows, w)
    np.add.at(deg_i, np.asarray(colidx, dtype=np.int64), w)
    e = np.float32(eps)
    norm = w / (np.sqrt(deg_u[rows] + e) * np.sqrt(deg_i[colidx] + e))
    return norm.astype(np.float32), deg_u, deg_i


# =====================================================================================
# Stage 2
# =====================================================================================
def spmm_prop(rowptr, colidx, vals, alpha, X, T=None, beta=0.0):
    """One hop: returns Y = (alpha*A) @ X and updates T += beta*Y in place (oracle.c)."""
    X = _f32(X)
    rows = rowptr.shape[0] - 1
    F = X.shape[1]
    Y = np.empty((rows, F), dtype=np.float32)
    _c().oracle_spmm_prop(I64(rows), I64(F), _p(_i32(rowptr)), _p(_i32(colidx)),
                          _p(None if vals is None else _f32(vals)), C.c_float(alpha), _p(X), I64(X.shape[1]),
                          _p(Y), I64(F), _p(T), I64(0 if T is None else T.shape[1]), C.c_float(beta))
    return Y


def propagate(rowptr, colidx, vals, X, prop_num, alpha):
    """clustgdd_agent_transduct.py:59-65.  Returns (prop_feat, target_feat)."""
    X = _f32(X)
    one_minus = np.float32(1.0 - alpha)
    target = (one_minus * X).astype(np.float32)
    prop = X
    for _ in range(1, int(prop_num)):
        prop = spmm_prop(rowptr, colidx, vals, np.float32(alpha), prop, T=target, beta=one_minus)
    return prop, target


def propagate_f64(rowptr, colidx, vals, X, prop_num, alpha):
    """Same loop evaluated in float64 (values alpha*A still rounded to fp32 first)."""
    X64 = np.ascontiguousarray(X, dtype=np.float64)
    one_minus = float(np.float32(1.0 - alpha))
    target = one_minus * X64
    prop = X64
    rows, F = X64.shape
    for _ in range(1, int(prop_num)):
        Y = np.empty((rows, F), dtype=np.float64)
        _c().oracle_spmm_prop_f64(I64(rows), I64(F), _p(_i32(rowptr)), _p(_i32(colidx)), _p(_f32(vals)),
                                  C.c_float(alpha), _p(prop), I64(F), _p(Y), I64(F))
        prop = Y
        target = target + one_minus * prop
    return prop, target


def lightgcn_propagate(rowptr, colidx, w, u0, i0, num_layers, eps=1e-8):
    """distill_recsys.py:319-353 forward (edge list = row-major COO of the cu x ci CSR)."""
    n_u, n_i = u0.shape[0], i0.shape[0]
    norm, _, _ = bipartite_normalize(rowptr, colidx, w, n_u, n_i, eps)
    t_rowptr, t_colidx, t_perm = csr_transpose(rowptr, colidx, (n_u, n_i))
    u, it = _f32(u0), _f32(i0)
    u_layers, i_layers = [u], [it]
    for _ in range(num_layers):
        u_new = spmm_prop(rowptr, colidx, norm, np.float32(1.0), it)
        i_new = spmm_prop(t_rowptr, t_colidx, norm[t_perm], np.float32(1.0), u)
        u, it = u_new, i_new
        u_layers.append(u)
        i_layers.append(it)
    return (np.stack(u_layers).mean(axis=0, dtype=np.float32), np.stack(i_layers).mean(axis=0, dtype=np.float32))


# =====================================================================================
# Stage 3
# =====================================================================================
def kmeans_assign(X, Cn):
    """fp32 E-step (sklearn/_k_means_lloyd.pyx:196-213).  Returns (labels, best)."""
    X, Cn = _f32(X), _f32(Cn)
    N, D = X.shape
    labels = np.empty(N, dtype=np.int32)
    best = np.empty(N, dtype=np.float32)
    _c().oracle_kmeans_assign_f32(I64(N), I64(Cn.shape[0]), I64(D), _p(X), I64(D), _p(Cn), I64(D), _p(labels), _p(best))
    return labels, best


def kmeans_assign_f64(X, Cn):
    """Exact nearest / runner-up in double.  Returns (labels, second, relative margin)."""
    X, Cn = _f32(X), _f32(Cn)
    N, D = X.shape
    labels = np.empty(N, dtype=np.int32)
    second = np.empty(N, dtype=np.int32)
    margin = np.empty(N, dtype=np.float64)
    _c().oracle_kmeans_assign_f64(I64(N), I64(Cn.shape[0]), I64(D), _p(X), I64(D), _p(Cn), I64(D), _p(labels),
                                  _p(second), _p(margin))
    return labels, second, margin


def labels_match(labels, X, Cn, band=1e-6):
    """BASELINE.json contract: labels must equal the exact argmin wherever the relative
    distance margin exceeds ``band``; inside the band either of the two nearest is accepted.
    Returns (ok, n_in_band, n_bad)."""
    l64, second, margin = kmeans_assign_f64(X, Cn)
    labels = np.asarray(labels)
    inband = margin <= band
    good = (labels == l64) | (inband & (labels == second))
    return bool(good.all()), int(inband.sum()), int((~good).sum())


def segment_sum(X, labels, K):
    """Sequential fp32 per-cluster sums + counts (sklearn/_k_means_lloyd.pyx:215-218)."""
    X = _f32(X)
    N, D = X.shape
    sums = np.empty((K, D), dtype=np.float32)
    counts = np.empty(K, dtype=np.int32)
    _c().oracle_segment_sum(I64(N), I64(K), I64(D), _p(X), I64(D), _p(_i32(labels)), _p(sums), I64(D), _p(counts))
    return sums, counts


def kmeans_finalize(sums, counts, C_old):
    """_average_centers + _center_shift (sklearn/_k_means_common.pyx:274-311)."""
    sums, C_old = _f32(sums), _f32(C_old)
    K, D = sums.shape
    C_new = np.zeros((K, D), dtype=np.float32)
    shift = _c().oracle_kmeans_finalize(I64(K), I64(D), _p(sums), I64(D), _p(_i32(counts)), _p(C_old), I64(D),
                                        _p(C_new), I64(D))
    return C_new, float(shift)


def relocate_empty(X, C_old, labels, sums, counts):
    """_relocate_empty_clusters_dense (sklearn/_k_means_common.pyx:167-211); the idx-th empty
    cluster receives the idx-th farthest sample (sklearn's argpartition order inside the top
    set is unspecified — compare as sets)."""
    empty = np.flatnonzero(counts == 0)
    if empty.size == 0:
        return sums, counts
    X = _f32(X)
    dist = ((X - C_old[labels]) ** 2).sum(axis=1)
    if dist.max() == 0:
        return sums, counts
    far = np.argsort(-dist, kind="stable")[: empty.size]
    sums, counts = sums.copy(), counts.copy()
    for new_c, idx in zip(empty, far):
        old_c = labels[idx]
        sums[old_c] -= X[idx]
        sums[new_c] = X[idx]
        counts[new_c] = 1
        counts[old_c] -= 1
    return sums, counts


def inertia(X, Cn, labels, exact=False):
    X, Cn = _f32(X), _f32(Cn)
    N, D = X.shape
    fn = _c().oracle_inertia_f64 if exact else _c().oracle_inertia
    return float(fn(I64(N), I64(D), _p(X), I64(D), _p(Cn), I64(D), _p(_i32(labels))))


def lloyd_iteration(Xc, Cn):
    """One full Lloyd iteration from shared centres: returns (labels, C_new, counts, shift)."""
    labels, _ = kmeans_assign(Xc, Cn)
    sums, counts = segment_sum(Xc, labels, Cn.shape[0])
    sums, counts = relocate_empty(Xc, Cn, labels, sums, counts)
    C_new, shift = kmeans_finalize(sums, counts, Cn)
    return labels, C_new, counts, shift


def kmeans_fit(X, C0, max_iter=300, tol=1e-4):
    """KMeans(n_clusters=K, init=C0, n_init=1, algorithm='lloyd').fit(X)
    (sklearn/_kmeans.py:1436-1560 + _kmeans_single_lloyd :630-758).
    Returns dict(labels, centers, inertia, n_iter)."""
    X = _f32(X).copy()
    mean = X.mean(axis=0)                       # :1487
    X -= mean                                   # :1489
    Cn = _f32(C0).copy() - mean                 # :1493
    tol_abs = 0.0 if tol == 0 else float(np.mean(np.var(X + mean, axis=0)) * tol)  # :285-293
    labels_old = np.full(X.shape[0], -1, dtype=np.int32)
    strict = False
    n_iter = 0
    for i in range(max_iter):
        n_iter = i + 1
        labels, C_new, _, shift = lloyd_iteration(X, Cn)
        Cn = C_new
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if shift <= tol_abs:
            break
        labels_old = labels
    if not strict:
        labels, _ = kmeans_assign(X, Cn)
    return dict(labels=labels, centers=(Cn + mean).astype(np.float32), inertia=inertia(X, Cn, labels),
                n_iter=n_iter, centers_centered=Cn, mean=mean)


def kmeans_plusplus(X, n_clusters, random_state, n_local_trials=None, cumsum_dtype=np.float64):
    """sklearn's greedy k-means++ (sklearn/_kmeans.py:180-278) with unit sample weights.
    Consumes ``random_state`` exactly as sklearn does.  ``cumsum_dtype=np.float32`` reproduces
    sklearn's own cumulative sum; float64 is what the CUDA kernel accumulates in.
    Returns (centers, indices)."""
    X = _f32(X)
    n, d = X.shape
    if n_local_trials is None:
        n_local_trials = 2 + int(np.log(n_clusters))
    centers = np.empty((n_clusters, d), dtype=np.float32)
    indices = np.full(n_clusters, -1, dtype=np.int64)
    cid = random_state.choice(n, p=np.full(n, 1.0 / n))
    centers[0], indices[0] = X[cid], cid

    def sqdist(C):  # [len(C), n] exact squared distances, fp32 result
        return ((X[None, :, :].astype(np.float64) - C[:, None, :].astype(np.float64)) ** 2).sum(-1).astype(np.float32)

    closest = sqdist(centers[:1])[0]
    pot = closest.astype(np.float64).sum()
    for c in range(1, n_clusters):
        rand_vals = random_state.uniform(size=n_local_trials) * pot
        cand = np.searchsorted(np.cumsum(closest.astype(cumsum_dtype)), rand_vals)
        np.clip(cand, None, n - 1, out=cand)
        dc = np.minimum(closest[None, :], sqdist(X[cand]))
        pots = dc.astype(np.float64).sum(axis=1)
        best = int(np.argmin(pots))
        pot, closest = pots[best], dc[best]
        centers[c], indices[c] = X[cand[best]], cand[best]
    return centers, indices


def standard_scale(X):
    """StandardScaler().fit_transform (distill_recsys.py:172): fp64 mean / population var,
    scale = sqrt(var) (0 -> 1), result fp32((x - fp32(mean)) / fp32(scale))."""
    X = _f32(X)
    mean = X.astype(np.float64).mean(axis=0)
    var = X.astype(np.float64).var(axis=0)
    scale = np.sqrt(var)
    scale[scale == 0] = 1.0
    return ((X - mean.astype(np.float32)) / scale.astype(np.float32)).astype(np.float32)


def cluster_means(X, labels, n):
    """clustgdd_agent_transduct.py:121-125: per-cluster mean, empty cluster -> NaN row."""
    sums, counts = segment_sum(X, labels, n)
    with np.errstate(invalid="ignore", divide="ignore"):
        out = sums / counts[:, None].astype(np.float32)
    out[counts == 0] = np.nan
    return out.astype(np.float32)


# =====================================================================================
# Stage 4
# =====================================================================================
def coarsen_counts(src, dst, labels_src, labels_dst, n_src, n_dst, w=None, drop_diag=False):
    """Segmented edge counting.  distill_recsys.py:184-201 (counts of train LINES, duplicates
    included) and the count/sum matrix implied by graph_compress (transduct :234-250).
    Returns (rowptr, colidx, counts int32, wsum f32|None)."""
    a = np.asarray(labels_src)[np.asarray(src, dtype=np.int64)].astype(np.int64)
    b = np.asarray(labels_dst)[np.asarray(dst, dtype=np.int64)].astype(np.int64)
    keep = (a != b) if drop_diag else np.ones(a.shape[0], dtype=bool)
    key = (a * n_dst + b)[keep]
    ww = None if w is None else np.asarray(w, dtype=np.float32)[keep]
    order = np.argsort(key, kind="stable")
    key = key[order]
    if key.size == 0:
        return np.zeros(n_src + 1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), (None if w is None else np.zeros(0, np.float32))
    head = np.ones(key.shape[0], dtype=bool)
    head[1:] = key[1:] != key[:-1]
    starts = np.flatnonzero(head)
    counts = np.diff(np.append(starts, key.shape[0])).astype(np.int32)
    ukey = key[starts]
    wsum = None
    if ww is not None:
        # fp64 accumulation: the oracle states the exact sum, the kernel's fp32 tree is compared to 1e-5
        wsum = np.add.reduceat(ww[order].astype(np.float64), starts)
    rp = np.zeros(n_src + 1, dtype=np.int64)
    np.add.at(rp, ukey // n_dst + 1, 1)
    return np.cumsum(rp).astype(np.int32), (ukey % n_dst).astype(np.int32), counts, wsum


def graph_compress_dense(labels, rowptr, colidx, vals, n):
    """clustgdd_agent_transduct.py:234-250 as written (dense one-hot), in float64 for
    reference: S = P^T A P with P = onehot / colsum, diagonal removed.  Small n only."""
    N = rowptr.shape[0] - 1
    rows = np.repeat(np.arange(N, dtype=np.int64), np.diff(rowptr))
    sizes = np.bincount(labels, minlength=n).astype(np.float64)
    S = np.zeros((n, n), dtype=np.float64)
    np.add.at(S, (labels[rows], labels[np.asarray(colidx, dtype=np.int64)]), np.asarray(vals, dtype=np.float64))
    with np.errstate(invalid="ignore", divide="ignore"):
        S = S / sizes[:, None] / sizes[None, :]
    np.fill_diagonal(S, 0.0)
    return S


# ---------------------------------------------------------------------------------------
# edge scoring + top-k sparsification (between stages 3 and 4)
# ---------------------------------------------------------------------------------------
def _rows_of(rowptr):
    return np.repeat(np.arange(rowptr.shape[0] - 1, dtype=np.int64), np.diff(rowptr))


def row_degree_f32(rowptr, vals):
    """``adj @ ones`` (utils_clustgdd.py:153-154): fp32 row sums in stored order."""
    n = rowptr.shape[0] - 1
    deg = np.zeros(n, dtype=np.float32)
    np.add.at(deg, _rows_of(rowptr), np.asarray(vals, dtype=np.float32))
    return deg


def er_lower(rowptr, colidx, vals):
    """ER_estimator (utils_clustgdd.py:151-162): values/deg[src] + values/deg[dst] in fp32."""
    vals = np.asarray(vals, dtype=np.float32)
    deg = row_degree_f32(rowptr, vals)
    rows = _rows_of(rowptr)
    return (vals / deg[rows] + vals / deg[np.asarray(colidx, dtype=np.int64)]).astype(np.float32)


def edge_cosine(rowptr, colidx, ebd, eps=1e-8):
    """F.cosine_similarity(ebd[src], ebd[dst], dim=-1) (utils_clustgdd.py:171): ATen divides each vector
    by max(|x|, eps) and then takes the dot product; evaluated here in float64 and rounded (the contract
    on this value is 1e-6 relative, the summation order inside ATen is not specified)."""
    e = np.asarray(ebd, dtype=np.float64)
    nrm = np.maximum(np.sqrt((e * e).sum(1)), eps)
    en = e / nrm[:, None]
    rows = _rows_of(rowptr)
    return (en[rows] * en[np.asarray(colidx, dtype=np.int64)]).sum(1).astype(np.float32)


def attaw_er_lower(rowptr, colidx, vals, ebd):
    """attaw_ER_estimator (utils_clustgdd.py:165-184): values re-weighted by the cosine similarity of the
    end points' embeddings, then ER_lower on the re-weighted graph.  Returns (ER_lower, reweighted values)."""
    rew = (np.asarray(vals, dtype=np.float32) * edge_cosine(rowptr, colidx, ebd)).astype(np.float32)
    return er_lower(rowptr, colidx, rew), rew


def softmax_rows(x):
    """F.softmax(ebd, dim=-1) (clustgdd_agent_transduct.py:158) in fp32."""
    x = np.asarray(x, dtype=np.float32)
    m = x.max(axis=1, keepdims=True)
    ex = np.exp(x - m, dtype=np.float32)
    return (ex / ex.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)


def class_edge_weight(rowptr, colidx, er, prob_col):
    """src_prob * dst_prob * ER_low (clustgdd_agent_transduct.py:164-167), left to right in fp32."""
    rows = _rows_of(rowptr)
    p = np.asarray(prob_col, dtype=np.float32)
    return ((p[rows] * p[np.asarray(colidx, dtype=np.int64)]).astype(np.float32) * np.asarray(er, dtype=np.float32)).astype(np.float32)


def topk_edges(weight, k):
    """torch.topk(weight, k) as an index SET (clustgdd_agent_transduct.py:142,168): every entry above the
    k-th largest value, plus the first entries (in index order) equal to it.  torch leaves the choice among
    equal values unspecified; the multiset of selected weights is what both must agree on."""
    w = np.asarray(weight, dtype=np.float32)
    k = int(k)
    if k <= 0:
        return np.zeros(0, dtype=np.int64)
    thr = np.partition(w, w.shape[0] - k)[w.shape[0] - k]
    gt = np.flatnonzero(w > thr)
    eq = np.flatnonzero(w == thr)[: k - gt.shape[0]]
    return np.sort(np.concatenate([gt, eq]))


def filter_csr(rowptr, colidx, vals, keep_idx):
    """coo_matrix((values[idx], (src[idx], dst[idx]))) (clustgdd_agent_transduct.py:146-151) as a sorted CSR."""
    n = rowptr.shape[0] - 1
    rows = _rows_of(rowptr)[keep_idx]
    rp = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rp, rows + 1, 1)
    return np.cumsum(rp).astype(np.int32), np.asarray(colidx)[keep_idx].astype(np.int32), np.asarray(vals, dtype=np.float32)[keep_idx]


def graph_sparse(rowptr, colidx, vals, ratio, ebd=None, sp_type="vanilla"):
    """ClustGDD.graph_sparse (clustgdd_agent_transduct.py:131-203) for 'vanilla', 'attaw', 'single', 'no_sp':
    list of (rowptr, colidx, vals) CSR triplets."""
    nnz = int(np.asarray(vals).shape[0])
    k = int(nnz * ratio)
    if sp_type == "no_sp":
        return [(rowptr, colidx, vals)]
    if sp_type == "vanilla":
        return [filter_csr(rowptr, colidx, vals, topk_edges(er_lower(rowptr, colidx, vals), k))]
    er, rew = attaw_er_lower(rowptr, colidx, vals, ebd)
    if sp_type == "single":
        return [filter_csr(rowptr, colidx, rew, topk_edges(er, k))]
    if sp_type == "attaw":
        P = softmax_rows(ebd)
        return [filter_csr(rowptr, colidx, rew, topk_edges(class_edge_weight(rowptr, colidx, er, P[:, i]), k))
                for i in range(P.shape[1])]
    raise ValueError(f"unknown sp_type {sp_type!r}")
