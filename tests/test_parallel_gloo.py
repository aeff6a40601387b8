"""world_size-2 tests of the row-partitioned drivers (gdr.parallel) over gloo on CPU.

The product has no CPU kernels, so the compute backend is replaced by ``OracleOps`` — the
oracle behind the same small interface as ``CudaOps`` — and what is under test is the
multi-rank host logic: partition bounds, padded all-gathers, packed all-reduces, the
replicated finalise / convergence decisions, the key-range exchange merge of stage 4 and the
distributed empty-cluster relocation.  Results are compared with the single-process oracle.
"""
import os
import socket
from dataclasses import dataclass

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


@dataclass
class CpuCSR:
    rowptr: torch.Tensor
    colidx: torch.Tensor
    vals: torch.Tensor
    shape: tuple


class OracleOps:
    """CPU stand-in for gdr.parallel.CudaOps (tests only)."""

    def __init__(self):
        from oracle import oracle as o
        self.o = o

    # -- stage 1 (block build): numpy restatement of the block kernels, tests only --
    def edges_route(self, row, col, n_rows, n_cols, rows_per, world, symmetrize):
        r, c = row.numpy().astype(np.int64), col.numpy().astype(np.int64)
        status = 0
        if r.size and (r.min() < 0 or r.max() >= n_rows or c.min() < 0 or c.max() >= n_cols):
            status |= 1
        if ((r == 0) & (c == 0)).any():
            status |= 2
        cb = self._bits(n_cols)
        if symmetrize:
            rr = np.stack([r, c], axis=1).reshape(-1)          # interleaved: pair, mirror, pair, mirror ...
            cc = np.stack([c, r], axis=1).reshape(-1)
        else:
            rr, cc = r, c
        owner = np.minimum(rr // rows_per, world - 1)
        key = (owner << 56) | (rr << cb) | cc
        order = np.argsort(owner, kind="stable")
        return torch.from_numpy(key[order]), [int((owner == w).sum()) for w in range(world)], status

    def csr_from_keys(self, keys, row_lo, n_local, n_cols, binarize):
        cb = self._bits(n_cols)
        key = keys.numpy() & ((1 << 56) - 1)
        rows, cols = (key >> cb) - row_lo, key & ((1 << cb) - 1)
        rp, ci, va = self.o.coo_to_csr(rows, cols, None, (n_local, n_cols), symmetrize=False, binarize=bool(binarize))
        return CpuCSR(torch.from_numpy(rp), torch.from_numpy(ci), torch.from_numpy(va), (n_local, n_cols))

    def bucket_by_owner(self, src, dst, rows_per, world):
        owner = (src // rows_per).numpy()
        order = np.argsort(owner, kind="stable")
        counts = np.bincount(owner, minlength=world)[:world].tolist()
        return torch.stack([src[order], dst[order]], dim=1).contiguous(), counts

    def build_block_csr(self, rows_local, cols, n_local, n):
        rp, ci, va = self.o.coo_to_csr(rows_local.numpy(), cols.numpy(), None, (n_local, n), symmetrize=False, binarize=True)
        return CpuCSR(torch.from_numpy(rp), torch.from_numpy(ci), torch.from_numpy(va), (n_local, n))

    def block_degrees(self, A, row_offset, add_identity):
        rp, ci, va = A.rowptr.numpy().astype(np.int64), A.colidx.numpy().astype(np.int64), A.vals.numpy()
        nl = A.shape[0]
        rows = np.repeat(np.arange(nl), np.diff(rp))
        has_diag = np.zeros(nl, bool)
        has_diag[rows[ci == rows + row_offset]] = True
        if add_identity:
            deg = np.bincount(rows, weights=va.astype(np.float64), minlength=nl) + 1.0
            out_len = np.diff(rp) + (~has_diag)
        else:
            d32 = np.zeros(nl, np.float32)
            np.add.at(d32, rows, va)
            deg, out_len = d32.astype(np.float64), np.diff(rp)
        return torch.from_numpy(deg), torch.from_numpy(np.concatenate([[0], np.cumsum(out_len)]).astype(np.int32))

    def block_fill(self, A, row_offset, add_identity, deg_global, rowptr_out):
        assert add_identity, "the synthetic graphs of these tests take the +I path"
        rp, ci, va = A.rowptr.numpy().astype(np.int64), A.colidx.numpy().astype(np.int64), A.vals.numpy()
        nl = A.shape[0]
        deg = deg_global.numpy()
        rows = np.repeat(np.arange(nl), np.diff(rp))
        v64 = va.astype(np.float64)
        on_diag = ci == rows + row_offset
        v64[on_diag] += 1.0
        has_diag = np.zeros(nl, bool)
        has_diag[rows[on_diag]] = True
        new_r = np.flatnonzero(~has_diag)
        r_all = np.concatenate([rows, new_r])
        c_all = np.concatenate([ci, new_r + row_offset])
        v_all = np.concatenate([v64, np.ones(new_r.shape[0])])
        order = np.lexsort((c_all, r_all))
        r_all, c_all, v_all = r_all[order], c_all[order], v_all[order]
        with np.errstate(divide="ignore"):
            r_inv = np.power(deg, -0.5)
        r_inv[np.isinf(r_inv)] = 0.0
        out = ((r_inv[r_all + row_offset] * v_all) * r_inv[c_all]).astype(np.float32)
        assert np.array_equal(np.bincount(r_all, minlength=nl), np.diff(rowptr_out.numpy()))
        return CpuCSR(rowptr_out, torch.from_numpy(c_all.astype(np.int32)), torch.from_numpy(out), A.shape)

    # -- bipartite --
    def build_block_csr_weighted(self, rows_local, cols, n_local, n_cols):
        rp, ci, va = self.o.coo_to_csr(rows_local.numpy(), cols.numpy(), None, (n_local, n_cols), symmetrize=False, binarize=False)
        return CpuCSR(torch.from_numpy(rp), torch.from_numpy(ci), torch.from_numpy(va), (n_local, n_cols))

    def row_sums(self, A):
        return torch.from_numpy(self.o.row_degree_f32(A.rowptr.numpy(), A.vals.numpy()))

    def bip_norm_block(self, A, deg_rows, deg_cols, eps):
        rows = np.repeat(np.arange(A.shape[0], dtype=np.int64), np.diff(A.rowptr.numpy()))
        e = np.float32(eps)
        w = A.vals.numpy()
        norm = w / (np.sqrt(deg_rows.numpy()[rows] + e) * np.sqrt(deg_cols.numpy()[A.colidx.numpy()] + e))
        return CpuCSR(A.rowptr, A.colidx, torch.from_numpy(norm.astype(np.float32)), A.shape)

    def column_moments(self, X, mean64):
        d = X.numpy().astype(np.float64) - mean64.numpy()[None, :]
        return torch.from_numpy(np.concatenate([d.sum(0), (d * d).sum(0)]))

    def standardize_apply(self, X, mean32, scale32):
        return torch.from_numpy(((X.numpy() - mean32.numpy()) / scale32.numpy()).astype(np.float32))

    def prep_rows(self, x):
        return x.to(torch.float32).contiguous()

    def empty_rows(self, rows, f, like):
        return torch.zeros((rows, f), dtype=torch.float32)

    def scale(self, x, a):
        return torch.from_numpy((np.float32(a) * x.numpy()).astype(np.float32))

    def spmm(self, A, x_full, alpha, target, beta):
        tn = None if target is None else target.numpy()
        y = self.o.spmm_prop(A.rowptr.numpy(), A.colidx.numpy(), A.vals.numpy(), np.float32(alpha),
                             x_full.numpy(), T=tn, beta=np.float32(beta))
        return torch.from_numpy(y)

    def remap_chunk_major(self, A, rows_per, cr, world):
        g = A.colidx.to(torch.int64)
        r, i = g // rows_per, g % rows_per
        c = i // cr
        return CpuCSR(A.rowptr, (c * (world * cr) + r * cr + (i - c * cr)).to(torch.int32), A.vals, A.shape)

    def slice_row_chunks(self, A, bounds):
        rp = A.rowptr.numpy()
        out = []
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            b, e = int(rp[lo]), int(rp[hi])
            out.append(CpuCSR(torch.from_numpy((rp[lo:hi + 1] - rp[lo]).astype(np.int32)), A.colidx[b:e].clone(),
                              A.vals[b:e].clone(), (hi - lo, A.shape[1])))
        return out

    def spmm_rows(self, A, x_full, alpha, y, target, beta, r0):
        rows = A.shape[0]
        if rows == 0:
            return
        tn = None if target is None else np.ascontiguousarray(target[r0:r0 + rows].numpy())
        out = self.o.spmm_prop(A.rowptr.numpy(), A.colidx.numpy(), A.vals.numpy(), np.float32(alpha),
                               np.ascontiguousarray(x_full.numpy()), T=tn, beta=np.float32(beta))
        y[r0:r0 + rows] = torch.from_numpy(out)
        if target is not None:
            target[r0:r0 + rows] = torch.from_numpy(tn)

    def spmm_slab(self, A, x_slab, alpha, y, target, beta, c0):
        w = x_slab.shape[1]
        tn = None if target is None else np.ascontiguousarray(target[:, c0:c0 + w].numpy())
        out = self.o.spmm_prop(A.rowptr.numpy(), A.colidx.numpy(), A.vals.numpy(), np.float32(alpha),
                               np.ascontiguousarray(x_slab.numpy()), T=tn, beta=np.float32(beta))
        y[:, c0:c0 + w] = torch.from_numpy(out)
        if target is not None:
            target[:, c0:c0 + w] = torch.from_numpy(tn)

    def column_sums(self, X):
        x = X.numpy().astype(np.float64)
        return torch.from_numpy(np.concatenate([x.sum(0), (x * x).sum(0)]))

    def center(self, X, mean):
        return X - mean

    def prepare_kmeans(self, Xc, K):
        pass

    def centers_like(self, C, device):
        return C.clone().contiguous()

    def assign(self, Xc, C, labels, labels_prev, n_changed):
        if Xc.shape[0] == 0:
            return
        lab, _ = self.o.kmeans_assign(Xc.numpy(), C.numpy())
        if labels_prev is not None and n_changed is not None:
            n_changed += int((lab != labels_prev.numpy()).sum())
        labels.copy_(torch.from_numpy(lab))

    def segment_sum(self, Xc, labels, K, sums, counts):
        s, c = self.o.segment_sum(Xc.numpy(), labels.numpy(), K)
        sums.copy_(torch.from_numpy(s))
        counts.copy_(torch.from_numpy(c))

    def finalize(self, sums, counts, C_old, C_new):
        cn, shift = self.o.kmeans_finalize(sums.numpy(), counts.numpy(), C_old.numpy())
        C_new.copy_(torch.from_numpy(cn))
        return shift, int((counts == 0).sum())

    def inertia(self, Xc, C, labels):
        if Xc.shape[0] == 0:
            return torch.zeros(1, dtype=torch.float64)
        return torch.tensor([self.o.inertia(Xc.numpy(), C.numpy(), labels.numpy())], dtype=torch.float64)

    def row_dist(self, Xc, C, labels):
        return ((Xc - C[labels.long()]) ** 2).sum(dim=1)

    def label_counts(self, labels, n):
        return torch.bincount(labels.long(), minlength=n).to(torch.int32)

    @staticmethod
    def _bits(n):
        b = 1
        while (1 << b) < n:
            b += 1
        return b

    def coarsen_route(self, A, labels_src, labels_dst, n, world, n_dst=None, src=None, dst=None):
        n_dst = n if n_dst is None else n_dst
        ls, ld = labels_src.numpy().astype(np.int64), labels_dst.numpy().astype(np.int64)
        if A is not None:
            rows = np.repeat(np.arange(A.shape[0]), np.diff(A.rowptr.numpy()))
            a, b, w, drop = ls[rows], ld[A.colidx.numpy().astype(np.int64)], A.vals.numpy(), True
        else:
            a, b, w, drop = ls[src.numpy()], ld[dst.numpy()], None, False
        cr = (n + world - 1) // world
        owner = np.minimum(a // cr, world - 1)
        if drop:
            owner = np.where(a == b, 127, owner)
        key = (owner << 56) | (a << self._bits(n_dst)) | b
        order = np.argsort(owner, kind="stable")
        counts = [int((owner == r).sum()) for r in range(world)]
        return (torch.from_numpy(key[order]), None if w is None else torch.from_numpy(w[order].astype(np.float32)), counts)

    def cluster_stats(self, A, labels_local, n):
        """per coarse row over this rank's rows: (edges, bit pattern of max |w|) — what gdr_cluster_stats returns"""
        lab = labels_local.numpy().astype(np.int64)
        deg = np.diff(A.rowptr.numpy()).astype(np.int64)
        edges = np.zeros(n, np.int64)
        np.add.at(edges, lab, deg)
        rows = np.repeat(np.arange(A.shape[0]), deg)
        wmax = np.zeros(n, np.float32)
        np.maximum.at(wmax, lab[rows], np.abs(A.vals.numpy()))
        return torch.from_numpy(edges.astype(np.int32)), torch.from_numpy(wmax.view(np.int32).copy())

    def coarse_merge_edges(self, keys, w, a_lo, n_rows, n, n_dst=None, stats=None):
        n_dst = n if n_dst is None else n_dst
        if stats is not None:      # the driver hands over the GLOBAL per-cluster stats: every rank must hold the same
            self.last_stats = (stats[0].numpy().copy(), stats[1].numpy().copy())
        bb = self._bits(n_dst)
        key = keys.numpy() & ((1 << 56) - 1)
        order = np.argsort(key, kind="stable")
        key = key[order]
        head = np.ones(key.shape[0], bool)
        head[1:] = key[1:] != key[:-1]
        starts = np.flatnonzero(head)
        ends = np.append(starts[1:], key.shape[0])
        cnt = (ends - starts).astype(np.int32)
        wsum = None
        if w is not None:
            ww = w.numpy()[order]
            wsum = np.zeros(starts.shape[0], np.float32)
            for p, (b0, e0) in enumerate(zip(starts, ends)):      # fp32, exchange order (what the kernel does)
                acc = np.float32(0)
                for j in range(b0, e0):
                    acc = np.float32(acc + ww[j])
                wsum[p] = acc
        ukey = key[starts]
        a = (ukey >> bb) - a_lo
        rp = np.zeros(n_rows + 1, np.int64)
        np.add.at(rp, a + 1, 1)
        return (torch.from_numpy(np.cumsum(rp).astype(np.int32)), torch.from_numpy((ukey & ((1 << bb) - 1)).astype(np.int32)),
                torch.from_numpy(cnt), None if wsum is None else torch.from_numpy(wsum))

    def coarsen_records(self, A, labels_src, labels_dst, n, world, n_dst=None, src=None, dst=None):
        n_dst = n if n_dst is None else n_dst
        if A is not None:
            rows = np.repeat(np.arange(A.shape[0]), np.diff(A.rowptr.numpy()))
            rp, ci, cnt, wsum = self.o.coarsen_counts(rows, A.colidx.numpy(), labels_src.numpy(), labels_dst.numpy(), n, n_dst,
                                                      w=A.vals.numpy(), drop_diag=True)
        else:
            rp, ci, cnt, _ = self.o.coarsen_counts(src.numpy(), dst.numpy(), labels_src.numpy(), labels_dst.numpy(), n, n_dst)
            wsum = np.zeros(cnt.shape[0], np.float32)
        a = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
        key = (a << self._bits(n_dst)) | ci.astype(np.int64)
        wbits = wsum.astype(np.float32).view(np.uint32).astype(np.int64)
        rec = np.stack([key, (cnt.astype(np.int64) << 32) | wbits], axis=1)
        cr = (n + world - 1) // world
        cuts = [int(rp[min(n, r * cr)]) for r in range(world + 1)]
        return torch.from_numpy(rec), [cuts[r + 1] - cuts[r] for r in range(world)]

    def coarse_merge(self, rec, a_lo, n_rows, n, n_dst=None):
        rec = rec.numpy()
        bb = self._bits(n if n_dst is None else n_dst)
        order = np.argsort(rec[:, 0], kind="stable")
        key, val = rec[order, 0], rec[order, 1]
        head = np.ones(key.shape[0], bool)
        head[1:] = key[1:] != key[:-1]
        starts = np.flatnonzero(head)
        cnt = np.add.reduceat(val >> 32, starts).astype(np.int32) if key.size else np.zeros(0, np.int32)
        w = (val & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
        wsum = np.zeros(starts.shape[0], np.float32)
        ends = np.append(starts[1:], key.shape[0])
        for p, (b, e) in enumerate(zip(starts, ends)):          # fp32, source-rank order (what the kernel does)
            acc = np.float32(0)
            for j in range(b, e):
                acc = np.float32(acc + w[j])
            wsum[p] = acc
        ukey = key[starts]
        a = (ukey >> bb) - a_lo
        rp = np.zeros(n_rows + 1, np.int64)
        np.add.at(rp, a + 1, 1)
        return (torch.from_numpy(np.cumsum(rp).astype(np.int32)), torch.from_numpy((ukey & ((1 << bb) - 1)).astype(np.int32)),
                torch.from_numpy(cnt), torch.from_numpy(wsum))

    def coarse_scale(self, rowptr, colidx, wsum, sizes, a_lo, n_rows):
        s = sizes.numpy().astype(np.float32)
        a = np.repeat(np.arange(n_rows), np.diff(rowptr.numpy())) + a_lo
        ia, ib = np.float32(1.0) / s[a], np.float32(1.0) / s[colidx.numpy()]
        return torch.from_numpy(((wsum.numpy() * ia) * ib).astype(np.float32))

    def csr_to_coo(self, rowptr, colidx, vals, n):
        rows = torch.repeat_interleave(torch.arange(n), (rowptr[1:] - rowptr[:-1]).long())
        return torch.sparse_coo_tensor(torch.stack([rows, colidx.long()]), vals, (n, n))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gdr import synth
    from oracle import oracle as o
    n, f, k = 1501, 12, 23          # odd sizes: ragged last block
    u, v = synth.skewed_graph(n, 9000, seed=3)
    rp, ci, va = o.coo_to_csr(u, v, None, (n, n), symmetrize=True, binarize=True)
    rpo, cio, vo, _ = o.sym_normalize(rp, ci, va, n)
    X = synth.clustered_features(n, f, 9, seed=4)
    return n, f, k, rpo, cio, vo, X


def _worker(rank, world, port, case):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gdr import parallel as par
        from gdr import synth
        from oracle import oracle as o
        n, f, k, rpo, cio, vo, X = _inputs()
        part = par.RowPartition(n, world, rank)
        comm = par.Comm(dist)
        ops = OracleOps()
        lo, hi = part.lo, part.hi
        b, e = int(rpo[lo]), int(rpo[hi])
        A_local = CpuCSR(torch.from_numpy((rpo[lo:hi + 1] - rpo[lo]).astype(np.int32)), torch.from_numpy(cio[b:e].copy()),
                         torch.from_numpy(vo[b:e].copy()), (hi - lo, n))
        x_local = torch.from_numpy(X[lo:hi].copy())

        if case == "build":
            # every rank holds a slice of the undirected pair list; exchange-based build == rows of the global build
            u, v = synth.skewed_graph(n, 9000, seed=3)
            per = (u.shape[0] + world - 1) // world
            sl = slice(rank * per, min(u.shape[0], (rank + 1) * per))
            A_blk = par.dist_build_adjacency(comm, part, torch.from_numpy(u[sl].copy()), torch.from_numpy(v[sl].copy()), n, ops=ops)
            assert np.array_equal(A_blk.rowptr.numpy(), A_local.rowptr.numpy())
            assert np.array_equal(A_blk.colidx.numpy(), A_local.colidx.numpy())
            assert np.array_equal(A_blk.vals.numpy(), A_local.vals.numpy())        # bit-identical values
        elif case in ("propagate", "propagate_slabs", "propagate_rows"):
            # one-pass hop / hop pipelined over 3 column slabs / over 3 row chunks (async all-gathers): same bits
            prop, target = par.dist_propagate(comm, part, A_local, x_local, 4, 0.8, ops=ops,
                                              slabs=3 if case == "propagate_slabs" else 1,
                                              row_chunks=3 if case == "propagate_rows" else 1)
            p_ref, t_ref = o.propagate(rpo, cio, vo, X, 4, 0.8)
            assert np.array_equal(prop.numpy(), p_ref[lo:hi])      # row-wise independent: bit-identical
            assert np.array_equal(target.numpy(), t_ref[lo:hi])
        elif case in ("kmeans", "kmeans_empty"):
            C0 = synth.kmeans_init(X, k, seed=5)
            if case == "kmeans_empty":
                C0[3] = C0[2]
                C0[4] = C0[2]
            km = par.DistKMeans(k, C0, max_iter=30, tol=1e-4, ops=ops, comm=comm).fit(x_local)
            ref = o.kmeans_fit(X, C0, max_iter=30, tol=1e-4)
            if case == "kmeans":
                assert km.n_iter_ == ref["n_iter"]
                assert np.array_equal(km.labels_.numpy(), ref["labels"][lo:hi])
                np.testing.assert_allclose(km.cluster_centers_.numpy(), ref["centers"], rtol=1e-5, atol=1e-5)
            assert abs(km.inertia_ - ref["inertia"]) <= (1e-4 if case == "kmeans" else 5e-2) * ref["inertia"]
            # replicated centres are bit-identical on every rank
            cc = km.cluster_centers_.clone()
            mx = cc.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            assert torch.equal(cc, mx)
        elif case == "bipartite":
            # distill_recsys on a row partition: users and items each split over the ranks
            nu, ni, d, L = 301, 207, 8, 2
            uu, ii = synth.bipartite_interactions(nu, ni, 4000, seed=9)
            rs = np.random.RandomState(10)
            u0 = (0.1 * rs.standard_normal((nu, d))).astype(np.float32)
            i0 = (0.1 * rs.standard_normal((ni, d))).astype(np.float32)
            pu, pi = par.RowPartition(nu, world, rank), par.RowPartition(ni, world, rank)
            per = (uu.shape[0] + world - 1) // world
            sl = slice(rank * per, min(uu.shape[0], (rank + 1) * per))
            u_sl, i_sl = torch.from_numpy(uu[sl].copy()), torch.from_numpy(ii[sl].copy())
            R_l, RT_l = par.dist_build_interaction(comm, pu, pi, u_sl, i_sl, ops=ops)
            rp, ci, w = o.coo_to_csr(uu, ii, None, (nu, ni))
            b, e = int(rp[pu.lo]), int(rp[pu.hi])
            assert np.array_equal(R_l.rowptr.numpy(), rp[pu.lo:pu.hi + 1] - rp[pu.lo])
            assert np.array_equal(R_l.colidx.numpy(), ci[b:e]) and np.array_equal(R_l.vals.numpy(), w[b:e])
            A_l, AT_l, du_l, di_l = par.dist_bipartite_normalize(comm, pu, pi, R_l, RT_l, ops=ops)
            norm, du, di = o.bipartite_normalize(rp, ci, w, nu, ni)
            assert np.array_equal(du_l.numpy(), du[pu.lo:pu.hi]) and np.array_equal(di_l.numpy(), di[pi.lo:pi.hi])
            assert np.array_equal(A_l.vals.numpy(), norm[b:e])
            uo, io = par.dist_lightgcn_propagate(comm, pu, pi, A_l, AT_l, torch.from_numpy(u0[pu.lo:pu.hi].copy()),
                                                 torch.from_numpy(i0[pi.lo:pi.hi].copy()), L, ops=ops)
            ur, ir = o.lightgcn_propagate(rp, ci, w, u0, i0, L)
            np.testing.assert_allclose(uo.numpy(), ur[pu.lo:pu.hi], rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(io.numpy(), ir[pi.lo:pi.hi], rtol=1e-6, atol=1e-7)
            xs = par.dist_standard_scale(comm, torch.from_numpy(u0[pu.lo:pu.hi].copy()), ops=ops)
            np.testing.assert_allclose(xs.numpy(), o.standard_scale(u0)[pu.lo:pu.hi], rtol=1e-6, atol=1e-6)
            ncu, nci = 31, 21
            u2cu = rs.randint(0, ncu, nu).astype(np.int32)
            i2ci = rs.randint(0, nci, ni).astype(np.int32)
            rpc, cic, vc = par.dist_build_condensed_bipartite(comm, pu, pi, u_sl, i_sl, torch.from_numpy(u2cu[pu.lo:pu.hi].copy()),
                                                              torch.from_numpy(i2ci[pi.lo:pi.hi].copy()), ncu, nci, ops=ops)
            rp_r, ci_r, cnt_r, _ = o.coarsen_counts(uu, ii, u2cu, i2ci, ncu, nci)
            assert np.array_equal(rpc.numpy(), rp_r) and np.array_equal(cic.numpy(), ci_r)
            assert np.array_equal(vc.numpy(), cnt_r.astype(np.float32)) and int(vc.sum()) == uu.shape[0]
        elif case == "wire":
            # row-padded views (width 7 inside a leading dimension of 8: what _dev.new_padded returns for f % 4 != 0)
            # are non-contiguous; the collectives run on the padded base and leave the padding column zero
            buf = torch.zeros((5, 8))
            v = buf[:, :7]
            v[:] = float(rank + 1)
            assert not v.is_contiguous()
            comm.all_reduce(v)
            assert torch.equal(v, torch.full((5, 7), 3.0)) and float(buf[:, 7].abs().max()) == 0.0
            loc = torch.zeros((3, 8))[:, :7]
            loc[:] = float(rank + 10)
            out_b = torch.zeros((6, 8))
            out = out_b[:, :7]
            comm.all_gather_rows(loc, out)
            assert torch.equal(out[:3], torch.full((3, 7), 10.0)) and torch.equal(out[3:], torch.full((3, 7), 11.0))
            assert float(out_b[:, 7].abs().max()) == 0.0
            h = comm.all_gather_rows(loc, out, async_op=True)
            h.wait()
            with pytest.raises(ValueError):
                comm.all_reduce(torch.zeros((4, 6)).t())
        elif case == "coarsen":
            labels = np.random.RandomState(6).randint(0, 17, n).astype(np.int32)
            labels[labels == 11] = 10      # an empty cluster
            S = o.graph_compress_dense(labels.astype(np.int64), rpo, cio, vo, int(labels.max()) + 1)
            fin = np.isfinite(S)
            for merge in ("records", "route"):
                adj_syn, counts = par.dist_graph_compress(comm, part, torch.from_numpy(labels[lo:hi].copy()), A_local, ops=ops,
                                                          merge=merge)
                got = adj_syn.to_dense().numpy()
                np.testing.assert_allclose(got[fin], S[fin], rtol=1e-5, atol=1e-8)
            rows = np.repeat(np.arange(n), np.diff(rpo))
            _, _, cnt_ref, _ = o.coarsen_counts(rows, cio, labels, labels, 17, 17, drop_diag=True)
            assert np.array_equal(counts.numpy(), cnt_ref)          # integer cell counts, CSR order, independent of the rank count
            ad = adj_syn.coalesce()
            assert ad._nnz() == cnt_ref.shape[0]
            # the routing form hands the owner the per-cluster (edges, max |w|) of the WHOLE graph, identical on every rank
            edges_g, wmax_g = ops.last_stats
            deg_all = np.diff(rpo).astype(np.int64)
            edges_ref = np.zeros(17, np.int64)
            np.add.at(edges_ref, labels, deg_all)
            wmax_ref = np.zeros(17, np.float32)
            np.maximum.at(wmax_ref, labels[rows], np.abs(vo))
            assert np.array_equal(edges_g, edges_ref.astype(np.int32)) and np.array_equal(wmax_g.view(np.float32), wmax_ref)
            # this rank's own key range, not replicated
            a_lo, rp_p, ci_p, v_p, c_p = par.dist_graph_compress(comm, part, torch.from_numpy(labels[lo:hi].copy()), A_local,
                                                                 ops=ops, replicate=False)
            rp_ref, ci_ref, _, _ = o.coarsen_counts(rows, cio, labels, labels, 17, 17, drop_diag=True)
            nr = rp_p.shape[0] - 1
            assert np.array_equal(rp_p.numpy() + rp_ref[a_lo], rp_ref[a_lo:a_lo + nr + 1])
            assert np.array_equal(ci_p.numpy(), ci_ref[rp_ref[a_lo]:rp_ref[a_lo + nr]])
            assert np.array_equal(c_p.numpy(), cnt_ref[rp_ref[a_lo]:rp_ref[a_lo + nr]])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["build", "propagate", "propagate_slabs", "propagate_rows", "kmeans", "kmeans_empty", "coarsen",
                                  "wire", "bipartite"])
def test_world2_gloo(case, oracle):
    mp.spawn(_worker, args=(2, _free_port(), case), nprocs=2, join=True)


def test_slab_bounds_cover_the_columns():
    from gdr.parallel import slab_bounds, default_slabs
    for f, s in [(100, 4), (128, 4), (12, 3), (7, 4), (1433, 4), (64, 2), (3, 2)]:
        b = slab_bounds(f, s)
        assert b[0][0] == 0 and b[-1][1] == f and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert all(c0 % 4 == 0 for c0, _ in b) and len(b) <= s
    assert default_slabs(1, 128) == 1 and default_slabs(8, 100) >= 1


def test_row_partition_bounds():
    from gdr.parallel import RowPartition
    for n, w in [(10, 3), (9, 3), (1, 4), (1501, 2), (169343, 8)]:
        spans = [RowPartition(n, w, r).bounds() for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) == RowPartition(n, w, 0).rows_per


@pytest.mark.parametrize("slabs,row_chunks", [(1, 1), (3, 1), (1, 3), (1, 5)])
def test_world1_drivers_without_process_group(oracle, slabs, row_chunks):
    """The drivers with a single rank and no process group (Comm() is a no-op): the pipelined hop variants, the
    exchange-based build and the coarsening merge all reduce to the single-device result."""
    from gdr import parallel as par
    from gdr import synth
    from oracle import oracle as o
    n, f, k, rpo, cio, vo, X = _inputs()
    part, comm, ops = par.RowPartition(n, 1, 0), par.Comm(), OracleOps()
    A = CpuCSR(torch.from_numpy(rpo.astype(np.int32)), torch.from_numpy(cio.copy()), torch.from_numpy(vo.copy()), (n, n))
    prop, target = par.dist_propagate(comm, part, A, torch.from_numpy(X.copy()), 4, 0.8, ops=ops, slabs=slabs,
                                      row_chunks=row_chunks)
    p_ref, t_ref = o.propagate(rpo, cio, vo, X, 4, 0.8)
    assert np.array_equal(prop.numpy(), p_ref) and np.array_equal(target.numpy(), t_ref)
    if (slabs, row_chunks) == (1, 1):
        u, v = synth.skewed_graph(n, 9000, seed=3)
        B = par.dist_build_adjacency(comm, part, torch.from_numpy(u), torch.from_numpy(v), n, ops=ops)
        assert np.array_equal(B.rowptr.numpy(), rpo) and np.array_equal(B.colidx.numpy(), cio)
        assert np.array_equal(B.vals.numpy(), vo)
        send = torch.arange(12).reshape(6, 2)
        out, counts = comm.all_to_all_rows(send, [6])
        assert torch.equal(out, send) and counts == [6]
        labels = np.random.RandomState(6).randint(0, 17, n).astype(np.int32)
        adj_syn, cnt = par.dist_graph_compress(comm, part, torch.from_numpy(labels), A, ops=ops)
        S = o.graph_compress_dense(labels.astype(np.int64), rpo, cio, vo, 17)
        fin = np.isfinite(S)
        np.testing.assert_allclose(adj_syn.to_dense().numpy()[fin], S[fin], rtol=1e-5, atol=1e-8)
        rows = np.repeat(np.arange(n), np.diff(rpo))
        assert np.array_equal(cnt.numpy(), o.coarsen_counts(rows, cio, labels, labels, 17, 17, drop_diag=True)[2])


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("gather", [False, True])
def test_symmetric_buffer_exchange_plan(world, gather):
    """exchange_plan (the host arithmetic of Comm.symm_exchange / all_gather_symm) replayed for every rank of a world on
    byte buffers: every row lands once, in source-rank order, regions do not overlap, nothing is written out of bounds."""
    import gdr
    from gdr.parallel import exchange_plan
    rng = np.random.RandomState(world + 10 * gather)
    if gather:
        block = rng.randint(0, 50, world)
        block[rng.randint(world)] = 0                       # an empty block
        cnt = [[int(block[s])] * world for s in range(world)]
    else:
        cnt = rng.randint(0, 40, (world, world))
        cnt[rng.randint(world), :] = 0                      # a rank that sends nothing
        cnt[:, rng.randint(world)] = 0                      # a rank that receives nothing
        cnt = cnt.tolist()
    row_bytes = [8, 4, 12]
    plans = [exchange_plan(cnt, me, row_bytes, gather) for me in range(world)]
    size = plans[0]["bytes"]
    assert all(p["bytes"] == size and p["regions"] == plans[0]["regions"] for p in plans)      # same layout on every rank
    bufs = [np.full(size, 0xEE, np.uint8) for _ in range(world)]
    written = [np.zeros(size, np.int32) for _ in range(world)]
    sends = []
    for me in range(world):
        n_send = max(cnt[me]) if gather else sum(cnt[me])
        sends.append([rng.randint(0, 255, n_send * eb).astype(np.uint8) for eb in row_bytes])
    for me, p in enumerate(plans):
        for a, eb in enumerate(row_bytes):
            for dst in range(world):
                n = p["send_cnt"][dst] * eb
                src = sends[me][a][p["send_off"][dst] * eb: p["send_off"][dst] * eb + n]
                off = p["dst_off_bytes"][a][dst]
                assert off % 4 == 0 and off + n <= size
                bufs[dst][off: off + n] = src
                written[dst][off: off + n] += 1
    for d in range(world):
        assert written[d].max() <= 1                                         # no byte written twice
        for a, eb in enumerate(row_bytes):
            reg = plans[d]["regions"][a]
            rows = plans[d]["rows_here"]
            want = np.concatenate([sends[s][a][(0 if gather else sum(cnt[s][:d])) * eb: ((0 if gather else sum(cnt[s][:d])) + cnt[s][d]) * eb]
                                   for s in range(world)] + [np.zeros(0, np.uint8)])
            assert np.array_equal(bufs[d][reg: reg + rows * eb], want)
            assert written[d][reg: reg + rows * eb].min(initial=1) == 1
