#!/usr/bin/env python
"""bench.py — one "step" = one pass of the distillation core over a synthetic graph of a
BASELINE.json shape:  stage 1 (CSR build + D^-1/2 A D^-1/2)  ->  stage 2 (K hops)  ->
stage 3 (20 Lloyd iterations from a fixed init, tol = 0)  ->  stage 4 (P^T A P).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload E|B|A] [--precision fp32|tc]
    python bench.py --impl reference ...        # the reference's CPU path on the host cores

The default workload is config E (BASELINE.json configs[4], ogbn-products-shaped: the shape the target is
quoted on; it fits one GPU) for EVERY N, so the N = 1, 2, 4, 8 lines are one strong-scaling experiment;
--workload B is configs[1] (ogbn-arxiv-shaped).

Prints ONE JSON line (rank 0).  `value` = k-means iterations / s (whole job), the quantity
BASELINE.json's speed-up target is quoted on; `prop` carries the A^K.X GB/s figure of the
same metric string with its own HBM roofline.  See DESIGN.md §Measurement.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LLOYD_ITERS = 20
ALPHA = 0.8


# ----------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]), bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_ev = index, [], threading.Event()
        self.nvml = None
        try:  # in-process NVML: no fork, negligible GIL time (a forked nvidia-smi stalls the launch thread)
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if index < len(ids) and ids[index].strip().isdigit():
                return int(ids[index])
        return index

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flag = lambda bit: "Active" if (r & bit) else "Not Active"
        return [str(sm), str(mx), str(pw), flag(0x8), flag(0x40), flag(0x20), flag(0x4)]

    def sample_now(self):
        """One sample from the CALLING thread (used between timed steps: an NVML query running
        concurrently with the launch thread contends on the driver lock and was measured to
        double the step time)."""
        try:
            if self.nvml is not None:
                self.rows.append(self._sample_nvml())
        except Exception:
            pass

    def run(self):
        if self.nvml is not None:
            return   # NVML available: samples are taken by sample_now() from the main thread
        while not self._stop_ev.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_ev.wait(0.05 if self.nvml is not None else 0.5)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    samples=len(sm), reasons=reasons)


def make_workload(name):
    from gdr import synth
    cfg = dict(synth.CONFIGS[name])
    idx = {"A": 0, "B": 1, "E": 4}[name]
    seed = 1234 + idx
    u, v = synth.uniform_graph(cfg["n"], cfg["pairs"], seed)
    X = synth.features(cfg["n"], cfg["f"], seed + 100, kind="l1" if name == "A" else "zscore")
    cfg.update(u=u, v=v, X=X, seed=seed)
    return cfg


def spmm_bytes(nnz, rows, xrows, F, fused=True, model="min"):
    """compulsory (B_min) or gather (B_gather) bytes of one hop."""
    if model == "min":
        b = nnz * 8 + (rows + 1) * 4 + xrows * F * 4 + rows * F * 4
    else:
        b = nnz * (8 + 4 * F) + (rows + 1) * 4 + rows * F * 4
    return b + (2 * rows * F * 4 if fused else 0)


_THREAD_LIMIT = None


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the reference's CPU path has to run on ALL the host cores
    (the other ranks of the reference arm exit at once), so the BLAS / OpenMP / torch pools are raised at run time."""
    global _THREAD_LIMIT
    n = len(os.sched_getaffinity(0))
    try:
        import torch
        torch.set_num_threads(n)
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_limits
        _THREAD_LIMIT = threadpool_limits(limits=n)   # kept alive for the rest of the process
    except Exception:
        pass
    return n


# ----------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's CPU path (oracle/ref_port.py: scipy + torch CPU sparse + scikit-learn)."""
    if rank != 0:
        return
    import torch
    from oracle import ref_port as rp
    from gdr import synth
    use_all_host_threads()
    w = make_workload(args.workload)
    n, F, K, hops = w["n"], w["f"], w["k"], w["hops"]
    info = rp.host_info()
    if n * K > 1e9:
        return run_reference_bounded(args, w, info)
    t_s1, t_s2, t_s3, t_s4, per_step = [], [], [], [], []
    it_lo, it_hi = 3, 3 + args.ref_kmeans_iters
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        A = rp.build_adjacency(w["u"], w["v"], n)
        adj = rp.to_tensor_sparse(A)
        adj_norm = rp.normalize_adj_tensor_sparse(adj)
        t1 = time.perf_counter()
        _, target = rp.propagate(adj_norm, torch.from_numpy(w["X"]), hops + 1, ALPHA)
        t2 = time.perf_counter()
        tn = target.numpy()
        C0 = synth.kmeans_init(tn, K, w["seed"])
        rp.kmeans_fit(tn, C0, 1)  # first call pays thread-pool start-up; keep it out of the difference
        t3 = time.perf_counter()
        rp.kmeans_fit(tn, C0, it_lo)
        t3a = time.perf_counter()
        km = rp.kmeans_fit(tn, C0, it_hi)
        t3b = time.perf_counter()
        S = rp.graph_compress_sparse(km.labels_.astype(np.int64), adj_norm)
        t4 = time.perf_counter()
        if s >= args.warmup:
            t_s1.append(t1 - t0)
            t_s2.append(t2 - t1)
            # differenced: (fit with it_hi) - (fit with it_lo) cancels validation / centring / final E-step
            d_it = ((t3b - t3a) - (t3a - t3)) / (it_hi - it_lo)
            if d_it <= 0:   # tiny problems converge before it_lo: fall back to the undifferenced fit
                d_it = (t3b - t3a) / max(1, int(km.n_iter_))
            t_s3.append(d_it)
            t_s4.append(t4 - t3b)
            per_step.append(t4 - t0)
    nnz = adj_norm._nnz()
    s_iter = float(np.mean(t_s3))
    iters_per_s = 1.0 / s_iter
    prop_gbs = hops * spmm_bytes(nnz, n, n, F) / np.mean(t_s2) / 1e9
    line = {
        "impl": "reference", "metric": "kmeans_iters_per_s", "value": iters_per_s, "unit": "iters/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(per_step)) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"config {args.workload}: {w['name']}-shaped uniform graph, N={n}, nnz(A_hat)={nnz}, "
                               f"F={F}, hops={hops}, K={K}, k-means D={F}",
                   "lloyd_iters_timed": f"sklearn fit(max_iter={it_hi}) - fit(max_iter={it_lo})"},
        "prop": {"metric": "A^K.X", "value": prop_gbs, "unit": "GB/s", "bytes_model": "B_min",
                 "s_per_hop": float(np.mean(t_s2)) / hops},
        "stages_ms": {"s1_build_normalize": np.mean(t_s1) * 1e3, "s2_propagate": np.mean(t_s2) * 1e3,
                      "s3_kmeans_per_iter": s_iter * 1e3, "s4_coarsen": np.mean(t_s4) * 1e3},
        "cpu_baseline": {"value": iters_per_s, "unit": "iters/s", "cores": info["affinity"], "kind": "port",
                         "sample": f"full config {args.workload}; sklearn/scipy/torch-CPU with all host threads", "host": info},
        "e2e": {"value": iters_per_s, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_reference_bounded(args, w, info):
    """Large workload (config E): the reference's scipy build (`tolil` of 126 M entries) and 4.5 s k-means
    iterations do not fit a few-minute run, so every step is a BOUNDED sample of the same workload: sklearn Lloyd
    on the first CPU_SAMPLE_ROWS rows of the z-scored feature matrix against all K centres, differenced between
    max_iter 1 and 3, the rate scaled by rows/N (an iteration costs the same flops whatever the rows hold, so the
    un-propagated features stand in for the propagated ones; stages 1, 2 and 4 are not timed here)."""
    n, F, K, hops = w["n"], w["f"], w["k"], w["hops"]
    rates, what, per_step = [], "", []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rate, what = cpu_kmeans_rate(w, 1, 3, CPU_SAMPLE_ROWS)
        if s >= args.warmup:
            rates.append(rate)
            per_step.append(time.perf_counter() - t0)
    iters_per_s = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "kmeans_iters_per_s", "value": iters_per_s, "unit": "iters/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(per_step)) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"config {args.workload}: {w['name']}-shaped uniform graph, N={n}, F={F}, hops={hops}, K={K}, "
                               f"k-means D={F}", "bounded": what},
        "stages_ms": {"s1_build_normalize": None, "s2_propagate": None, "s3_kmeans_per_iter": 1e3 / iters_per_s,
                      "s4_coarsen": None},
        "cpu_baseline": {"value": iters_per_s, "unit": "iters/s", "cores": info["affinity"], "kind": "port",
                         "sample": what, "host": info},
        "e2e": {"value": iters_per_s, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------
def run_ours(args, rank, world):
    import torch
    import gdr
    from gdr import synth
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if args.tc_screen:
        from gdr import _lib as _l
        _l.call("gdr_debug_set", b"tc_screen", int(args.tc_screen))
    if world > 1:
        return run_ours_multi(args, rank, world, dev)

    w = make_workload(args.workload)
    n, F, K, hops, seed = w["n"], w["f"], w["k"], w["hops"], w["seed"]
    pk = peaks()

    # inputs resident in HBM before the timed region
    u_d = torch.from_numpy(w["u"]).to(dev)
    v_d = torch.from_numpy(w["v"]).to(dev)
    X_d = torch.from_numpy(w["X"]).to(dev)
    X_pin = torch.from_numpy(w["X"]).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # k-means init: C0 = X_target[perm[:K]] needs the propagated features -> computed once, untimed
    A0 = gdr.sym_normalize(gdr.coo_to_csr(u_d, v_d, None, (n, n), symmetrize=True, binarize=True), 2)
    _, tgt0 = gdr.propagate(A0, X_d, hops + 1, ALPHA)
    perm = torch.from_numpy(np.random.RandomState(seed).permutation(n)[:K].astype(np.int64)).to(dev)
    C0 = tgt0[perm].clone()
    nnz = A0.nnz
    del A0, tgt0

    def step(record):
        e = [ev() for _ in range(5)]
        e[0].record()
        A = gdr.coo_to_csr(u_d, v_d, None, (n, n), symmetrize=True, binarize=True)
        An = gdr.sym_normalize(A, 2)
        e[1].record()
        prop, target = gdr.propagate(An, X_d, hops + 1, ALPHA)
        e[2].record()
        km = gdr.KMeans(n_clusters=K, init=C0, n_init=1, max_iter=LLOYD_ITERS, tol=0, precision=args.precision)
        km.fit(target)
        e[3].record()
        _, adj_syn = gdr.graph_compress(km.labels_, An, [])
        e[4].record()
        # keep only scalars: holding km / adj_syn of every step alive fragments the caching allocator
        # and forces cudaMalloc (a device-wide sync) inside later timed steps
        import types
        nnz_syn = int(adj_syn._nnz())
        return (e, types.SimpleNamespace(n_iter_=km.n_iter_, inertia_=km.inertia_),
                types.SimpleNamespace(_nnz=lambda v=nnz_syn: v))

    # W warm-up steps, continued until the GPU has been busy for >= 2 s: a fresh process starts
    # with the GPU in its idle power state and the first ~0.5 s of work runs at about half speed
    t_warm = time.perf_counter()
    n_warm = 0
    while n_warm < args.warmup or (not args.profile and time.perf_counter() - t_warm < 2.0):
        step(False)
        flush.fill_(1)
        n_warm += 1
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    import ctypes
    from gdr import _lib
    launches0 = gdr.launch_count()
    recs = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)          # L2 flush between timed iterations (untimed)
        torch.cuda.synchronize()
        recs.append(step(True))
        sampler.sample_now()    # clocks / throttle reasons while the GPU is still under load
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    launches = gdr.launch_count() - launches0
    clocks = sampler.stop()
    # roofline leg: the E-step main kernel alone, timed by CUDA-event pairs recorded inside the
    # library on its launch stream (separate pass: event pairs cannot live inside the replayed graph)
    tot_ms, n_l = ctypes.c_double(0), ctypes.c_int64(0)
    _lib.call("gdr_profile_enable", 1)
    for _ in range(3):
        flush.fill_(1)
        step(True)
    torch.cuda.synchronize()
    _lib.call("gdr_profile_collect", ctypes.addressof(tot_ms), ctypes.addressof(n_l))
    _lib.call("gdr_profile_enable", 0)
    assign_ms = np.array([tot_ms.value / max(1, n_l.value)])
    assign_total_ms = assign_ms[0] * (LLOYD_ITERS + 1) * args.steps   # launches per step: 20 iterations + final E-step

    st = np.array([[r[0][i].elapsed_time(r[0][i + 1]) for i in range(4)] for r in recs])  # ms per stage
    step_ms = st.sum(axis=1)
    n_iter = [r[1].n_iter_ for r in recs]
    km_ms = st[:, 2]

    # SpMM kernel time alone (same event mechanism, separate short loop: hops only)
    A_t = gdr.sym_normalize(gdr.coo_to_csr(u_d, v_d, None, (n, n), symmetrize=True, binarize=True), 2)
    _lib.call("gdr_profile_enable", 2)
    for _ in range(5):
        flush.fill_(1)
        gdr.propagate(A_t, X_d, hops + 1, ALPHA)
    _lib.call("gdr_profile_collect", ctypes.addressof(tot_ms), ctypes.addressof(n_l))
    _lib.call("gdr_profile_enable", 0)
    spmm_kernel_ms = tot_ms.value / max(1, n_l.value)
    del A_t
    iters_per_s = float(np.sum(n_iter) / (km_ms.sum() / 1e3))
    prop_ms = st[:, 1]
    # SURVEY §8(d) headline rule: B_gather when X does not fit L2 (N*F*4 > 96 MB), else B_min
    prop_model = "gather" if n * F * 4 > 96e6 else "min"
    b_hop = spmm_bytes(nnz, n, n, F, model=prop_model)
    b_prop = hops * b_hop + 2 * n * F * 4  # + the t = 0 scale pass
    prop_gbs = float(b_prop / (prop_ms.mean() / 1e3) / 1e9)

    # ---- e2e through the public API with HOST buffers (H2D of X, D2H of labels + centres) ----
    tn_host = recs[-1][1]  # keep last km for result checks
    C0_host = C0.cpu().numpy()
    target_host = torch.empty((n, F), dtype=torch.float32).pin_memory()
    target_host.copy_(gdr.propagate(gdr.sym_normalize(gdr.coo_to_csr(u_d, v_d, None, (n, n), symmetrize=True, binarize=True), 2), X_d, hops + 1, ALPHA)[1])
    e2e_t = []
    for i in range(2 if args.profile else max(3, args.steps)):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x_dev = target_host.to(dev, non_blocking=True)
        km = gdr.KMeans(n_clusters=K, init=C0_host, n_init=1, max_iter=LLOYD_ITERS, tol=0, precision=args.precision).fit(x_dev)
        lab_h = km.labels_.cpu()
        cen_h = km.cluster_centers_.cpu()
        torch.cuda.synchronize()
        e2e_t.append((time.perf_counter() - t0, km.n_iter_))
    e2e_t = e2e_t[1:]
    e2e_val = float(sum(k for _, k in e2e_t) / sum(t for t, _ in e2e_t))
    h2d = n * F * 4 + K * F * 4
    d2h = n * 4 + K * F * 4

    # ---- roofline of the dominant kernel (k-means E-step) ----
    # DRAM traffic per launch comes from the committed ncu --set full capture of this same command
    # (profiles/ncu_traffic.json); null when there is no capture for this workload.
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload, {})
    except Exception:
        pass
    flops = 2.0 * n * K * F
    a_tf = float(flops / (assign_ms.mean() / 1e3) / 1e12)
    tc = args.precision in ("tc", "auto") and F <= 128
    two_level = tc and (args.tc_screen >= 2 or (args.tc_screen == 0 and -(-n // 128) >= 4 * 148))
    roofline = {"kernel": ("k_assign_tc two-level screen (tcgen05 1xTF32 all rows -> select -> compact -> 3xTF32 undecided rows)" if two_level
                           else "k_assign_tc (tcgen05 3xTF32)") if tc else "k_assign_simt (exact fp32 FFMA)",
                "bound": "tensor", "achieved": a_tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": a_tf / pk["bf16_sustained"], "traffic": traffic.get("k_assign_tc") if tc else None,
                "peak_source": pk["src"] + " bf16 sustained",
                "note": "useful flops 2NKD per E-step over the time of the whole tensor-core screen; fp32 inputs: TF32 rate = 1/2 bf16, "
                        "3xTF32 emulation ceiling = peak/6, two-level screen ceiling -> peak/2",
                "frac_of_3xtf32_ceiling": a_tf / (pk["bf16_sustained"] / 6.0),
                "launch_ms": float(assign_ms.mean()), "share_of_step": float(assign_total_ms / step_ms.sum())}
    spmm_ms = spmm_kernel_ms
    # SURVEY §8(d) headline rule: B_gather when X does not fit L2 (N*F*4 > 96 MB), else B_min
    model = "gather" if n * F * 4 > 96e6 else "min"
    b_head = spmm_bytes(nnz, n, n, F, model=model)
    roofline_spmm = {"kernel": "k_spmm", "bound": "hbm", "achieved": float(b_head / (spmm_ms / 1e3) / 1e9), "peak": pk["hbm"],
                     "unit": "GB/s", "frac": float(b_head / (spmm_ms / 1e3) / 1e9 / pk["hbm"]),
                     "traffic": traffic.get("k_spmm"), "launch_ms": float(spmm_ms),
                     "bytes_model": "B_gather (X exceeds L2)" if model == "gather" else "B_min (X fits L2; the gathers run out of L2)",
                     "b_min_gbs": float(spmm_bytes(nnz, n, n, F) / (spmm_ms / 1e3) / 1e9),
                     "b_gather_gbs": float(spmm_bytes(nnz, n, n, F, model='gather') / (spmm_ms / 1e3) / 1e9),
                     "peak_source": pk["src"]}

    cpu = cpu_baseline(w, args) if not (args.no_cpu_baseline or args.profile) else None
    line = {
        "metric": "kmeans_iters_per_s", "value": iters_per_s, "unit": "iters/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(step_ms.mean()), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"config {args.workload}: {w['name']}-shaped uniform graph, N={n}, nnz(A_hat)={nnz}, F={F}, "
                               f"hops={hops}, K={K}, k-means D={F}, {LLOYD_ITERS} Lloyd iterations tol=0",
                   "precision": args.precision, "tc_screen": args.tc_screen, "l2": "flushed between timed steps (256 MB write)"},
        "prop": {"metric": "A^K.X", "value": prop_gbs, "unit": "GB/s", "frac_hbm_measured": prop_gbs / pk["hbm"],
                 "frac_hbm_8TBs": prop_gbs / 8000.0, "bytes_model": "B_gather" if prop_model == "gather" else "B_min",
                 "ms": float(prop_ms.mean())},
        "stages_ms": {"s1_build_normalize": float(st[:, 0].mean()), "s2_propagate": float(prop_ms.mean()),
                      "s3_kmeans": float(km_ms.mean()), "s3_kmeans_per_iter": float(km_ms.sum() / np.sum(n_iter)),
                      "s4_coarsen": float(st[:, 3].mean())},
        "roofline": roofline, "roofline_spmm": roofline_spmm, "cpu_baseline": cpu,
        "e2e": {"value": e2e_val, "unit": "iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches), "clocks": clocks, "wall_s": t_wall,
        "result": {"inertia": recs[-1][1].inertia_, "n_iter": int(n_iter[-1]), "syn_nnz": int(recs[-1][2]._nnz())},
    }
    print(json.dumps(line))


def one_gpu_reference(workload):
    """The committed single-GPU measurement of the same workload (for the strong-scaling ratio)."""
    name = {"E": "r1_bench_default_configE_1gpu.json", "B": "r1_bench_configB_v19.json"}.get(workload)
    try:
        d = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        return {"value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"], "source": "profiles/" + name}
    except Exception:
        return None


def run_ours_multi(args, rank, world, dev):
    """N > 1: the same step on the same graph, nodes row-partitioned over the ranks (strong
    scaling).  NCCL all-gather of the propagated rows per hop, all-reduce of the centroid
    partial sums / counts per Lloyd iteration, dense all-reduce merge of the coarsened graph."""
    import torch
    import torch.distributed as dist
    import gdr
    from gdr import parallel as par
    dist.init_process_group("nccl", device_id=dev)
    w = make_workload(args.workload)
    n, F, K, hops, seed = w["n"], w["f"], w["k"], w["hops"], w["seed"]
    pk = peaks()
    part = par.RowPartition(n, world, rank)
    comm = par.Comm(dist)
    ops = par.CudaOps(precision=args.precision)
    u_d = torch.from_numpy(w["u"]).to(dev)
    v_d = torch.from_numpy(w["v"]).to(dev)
    x_local = torch.from_numpy(w["X"][part.lo:part.hi].copy()).to(dev)
    per = (u_d.shape[0] + world - 1) // world
    u_sl = u_d[rank * per: (rank + 1) * per].contiguous()      # this rank's slice of the undirected pair list
    v_sl = v_d[rank * per: (rank + 1) * per].contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # fixed init: rows perm[:K] of the propagated features (computed once on every rank, untimed)
    A0 = gdr.sym_normalize(gdr.coo_to_csr(u_d, v_d, None, (n, n), symmetrize=True, binarize=True), 2)
    _, tgt0 = gdr.propagate(A0, torch.from_numpy(w["X"]).to(dev), hops + 1, ALPHA)
    perm = torch.from_numpy(np.random.RandomState(seed).permutation(n)[:K].astype(np.int64)).to(dev)
    C0 = tgt0[perm].clone()
    nnz = A0.nnz
    del A0, tgt0

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step():
        e = [ev() for _ in range(5)]
        e[0].record()
        A_local = par.dist_build_adjacency(comm, part, u_sl, v_sl, n, ops=ops)   # exchange-based: 1/world of the pairs per rank
        e[1].record()
        prop, target = par.dist_propagate(comm, part, A_local, x_local, hops + 1, ALPHA, ops=ops, slabs=args.slabs,
                                          row_chunks=args.row_chunks)
        e[2].record()
        km = par.DistKMeans(K, C0, max_iter=LLOYD_ITERS, tol=0, ops=ops, comm=comm).fit(target)
        e[3].record()
        adj_syn, _ = par.dist_graph_compress(comm, part, km.labels_, A_local, ops=ops)
        e[4].record()
        # keep only scalars: holding km / adj_syn of every step alive fragments the caching allocator
        # and forces cudaMalloc (a device-wide sync) inside later timed steps
        import types
        nnz_syn = int(adj_syn._nnz())
        return (e, types.SimpleNamespace(n_iter_=km.n_iter_, inertia_=km.inertia_),
                types.SimpleNamespace(_nnz=lambda v=nnz_syn: v))

    t_warm = time.perf_counter()
    n_warm = 0
    while True:   # W warm-up steps, and at least 2 s of work on every rank (same count everywhere)
        go = torch.tensor([1 if (n_warm < args.warmup or time.perf_counter() - t_warm < 2.0) else 0], device=dev)
        dist.all_reduce(go, op=dist.ReduceOp.MAX)
        if int(go.item()) == 0:
            break
        step()
        flush.fill_(1)
        n_warm += 1
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    launches0 = gdr.launch_count()
    recs = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        dist.barrier()
        recs.append(step())
        sampler.sample_now()
    torch.cuda.synchronize()
    dist.barrier()
    t_wall = time.perf_counter() - t0
    launches = gdr.launch_count() - launches0
    clocks = sampler.stop()
    st = torch.tensor([[r[0][i].elapsed_time(r[0][i + 1]) for i in range(4)] for r in recs], dtype=torch.float64, device=dev)
    dist.all_reduce(st, op=dist.ReduceOp.MAX)     # device time, max over ranks, per step and stage
    st = st.cpu().numpy()
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    dist.all_reduce(lt)
    n_iter = [r[1].n_iter_ for r in recs]

    # ---- roofline leg (per GPU): the E-step tensor-core screen of this rank's row block, timed by the
    #      library's own CUDA-event pairs on its launch stream ----
    import ctypes
    from gdr import _lib
    A_local = par.dist_build_adjacency(comm, part, u_sl, v_sl, n, ops=ops)
    _, target_l = par.dist_propagate(comm, part, A_local, x_local, hops + 1, ALPHA, ops=ops)
    tot_ms, n_l = ctypes.c_double(0), ctypes.c_int64(0)
    _lib.call("gdr_profile_enable", 1)
    flush.fill_(1)
    par.DistKMeans(K, C0, max_iter=LLOYD_ITERS, tol=0, ops=ops, comm=comm).fit(target_l)
    torch.cuda.synchronize()
    _lib.call("gdr_profile_collect", ctypes.addressof(tot_ms), ctypes.addressof(n_l))
    _lib.call("gdr_profile_enable", 0)
    assign_ms = tot_ms.value / max(1, n_l.value)
    am = torch.tensor([assign_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(am, op=dist.ReduceOp.MAX)
    assign_ms = float(am.item())

    # ---- e2e: DistKMeans.fit through the public API from pinned HOST rows (H2D of this rank's block of the
    #      propagated features, D2H of its labels + the centres), max over ranks ----
    target_host = torch.empty(tuple(target_l.shape), dtype=torch.float32).pin_memory()
    target_host.copy_(target_l)
    del target_l, A_local
    C0_host = C0.cpu()
    e2e_t = []
    for i in range(max(3, min(args.steps, 5))):
        flush.fill_(1)
        torch.cuda.synchronize()
        dist.barrier()
        t0e = time.perf_counter()
        x_dev = target_host.to(dev, non_blocking=True)
        km = par.DistKMeans(K, C0_host, max_iter=LLOYD_ITERS, tol=0, ops=ops, comm=comm).fit(x_dev)
        lab_h = km.labels_.cpu()
        cen_h = km.cluster_centers_.cpu()
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_t.append((float(tt.item()), km.n_iter_))
    e2e_t = e2e_t[1:]
    e2e_val = float(sum(k for _, k in e2e_t) / sum(t for t, _ in e2e_t))

    if rank == 0:
        step_ms = st.sum(axis=1)
        km_ms, prop_ms = st[:, 2], st[:, 1]
        flops_rank = 2.0 * part.rows_per * K * F
        a_tf = flops_rank / (assign_ms / 1e3) / 1e12
        roofline = {"kernel": "k_assign_tc two-level screen (per GPU, this rank's row block)", "bound": "tensor",
                    "achieved": a_tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": a_tf / pk["bf16_sustained"],
                    "traffic": None, "peak_source": pk["src"] + " bf16 sustained", "launch_ms": assign_ms,
                    "note": "useful flops 2*(N/world)*K*D per E-step over the slowest rank's screen time; fp32 inputs run as TF32"}
        e2e = {"value": e2e_val, "unit": "iters/s", "h2d_bytes_per_step": n * F * 4 + world * K * F * 4,
               "d2h_bytes_per_step": n * 4 + world * K * F * 4}
        prop_model = "gather" if n * F * 4 > 96e6 else "min"   # SURVEY §8(d) headline rule
        b_prop = hops * spmm_bytes(nnz, n, n, F, model=prop_model) + 2 * n * F * 4
        prop_gbs = float(b_prop / (prop_ms.mean() / 1e3) / 1e9)
        iters_per_s = float(np.sum(n_iter) / (km_ms.sum() / 1e3))
        line = {
            "metric": "kmeans_iters_per_s", "value": iters_per_s, "unit": "iters/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(step_ms.mean()), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config {args.workload}: {w['name']}-shaped uniform graph, N={n}, nnz(A_hat)={nnz}, F={F}, "
                                   f"hops={hops}, K={K}, k-means D={F}, {LLOYD_ITERS} Lloyd iterations tol=0",
                       "parallelism": f"row-partition x{world} (stage 1: pair slices, all-to-all by owner, all-gather of degrees; "
                                      f"stage 2: all-gather per hop pipelined over {par.default_row_chunks(world) if args.row_chunks is None else args.row_chunks} row chunks; "
                                      "stage 3: all-reduce of centroid sums/counts per Lloyd iteration; stage 4: dense n x n all-reduce)",
                       "precision": args.precision, "l2": "flushed between timed steps (256 MB write)"},
            "prop": {"metric": "A^K.X", "value": prop_gbs, "unit": "GB/s", "frac_hbm_measured": prop_gbs / (pk["hbm"] * world),
                     "bytes_model": ("B_gather" if prop_model == "gather" else "B_min") + " (whole job, incl. the all-gather time)", "ms": float(prop_ms.mean())},
            "stages_ms": {"s1_build_normalize": float(st[:, 0].mean()), "s2_propagate": float(prop_ms.mean()),
                          "s3_kmeans": float(km_ms.mean()), "s3_kmeans_per_iter": float(km_ms.sum() / np.sum(n_iter)),
                          "s4_coarsen": float(st[:, 3].mean())},
            "roofline": roofline, "cpu_baseline": None, "same_workload_1gpu": one_gpu_reference(args.workload),
            "e2e": e2e, "gpu_launches": int(lt.item()), "clocks": clocks, "wall_s": t_wall,
            "result": {"inertia": recs[-1][1].inertia_, "n_iter": int(n_iter[-1]), "syn_nnz": int(recs[-1][2]._nnz())},
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()


CPU_SAMPLE_ROWS = 200_000   # bounded k-means sample of the large workload (all K centres, first rows of X)


def cpu_kmeans_rate(w, it_lo, it_hi, rows_cap=None):
    """k-means iterations/s of the reference's CPU path (sklearn Lloyd, all host threads) on workload w,
    differenced between two max_iter values so that validation / centring / the final E-step cancel.
    With rows_cap the fit runs on the first rows_cap rows against all K centres and the rate is scaled by
    rows_cap / N (the E-step is linear in the rows and dominates: 2*N*K*D flops per iteration)."""
    from oracle import ref_port as rp
    from gdr import synth
    n, K = w["n"], w["k"]
    X = w["X"]
    C0 = synth.kmeans_init(X, K, w["seed"])
    m = n if rows_cap is None else min(n, int(rows_cap))
    Xs = np.ascontiguousarray(X[:m])
    if m < n:   # the init rows must lie inside the sample
        C0 = synth.kmeans_init(Xs, K, w["seed"])
    rp.kmeans_fit(Xs[:20000], C0[: min(K, 100)], 2)  # warm the thread pools
    t0 = time.perf_counter()
    rp.kmeans_fit(Xs, C0, it_lo)
    t1 = time.perf_counter()
    rp.kmeans_fit(Xs, C0, it_hi)
    t2 = time.perf_counter()
    s_iter = ((t2 - t1) - (t1 - t0)) / float(it_hi - it_lo)
    rate = (1.0 / s_iter) * (m / n)
    what = (f"sklearn KMeans(init=C0,n_init=1,tol=0) on "
            + (f"the full {n}x{X.shape[1]} matrix" if m == n else f"the first {m} of {n} rows (x{X.shape[1]}), rate scaled by {m}/{n}")
            + f", K={K}: fit({it_hi} it) - fit({it_lo} it)")
    return rate, what


def cpu_baseline(w, args):
    """Bounded CPU sample on this box's host cores (oracle-side port of the reference)."""
    from oracle import ref_port as rp
    big = w["n"] * w["k"] > 1e9
    rate, what = cpu_kmeans_rate(w, 1, 3, CPU_SAMPLE_ROWS) if big else cpu_kmeans_rate(w, 2, 12)
    info = rp.host_info()
    return {"value": rate, "unit": "iters/s", "cores": info["affinity"], "kind": "port", "sample": what, "host": info}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["A", "B", "E"])
    ap.add_argument("--precision", default="tc", choices=["fp32", "tc", "auto"])
    ap.add_argument("--slabs", type=int, default=None, help="N > 1: column slabs of the pipelined hop (default: automatic)")
    ap.add_argument("--row-chunks", type=int, default=None, help="N > 1: row chunks of the pipelined hop (default: 4)")
    ap.add_argument("--tc-screen", type=int, default=0, help="debug: 0 auto, 1 direct 3xTF32, 2/3 two-level screen (BN 128/256)")
    ap.add_argument("--ref-kmeans-iters", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: no CPU baseline, minimal e2e leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.workload is None:
        # ONE workload for every N, so that the per-N lines are one strong-scaling experiment: config E
        # (configs[4], ogbn-products-shaped), the shape BASELINE.json's target is quoted on; it fits one GPU
        # (~6 GB resident).  `--workload B` is configs[1] (ogbn-arxiv-shaped), the shape of most parity tests.
        args.workload = "E"
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
