#!/usr/bin/env python
"""Headline benchmark: batched HNSW search QPS at recall@10 on the 1M x 128 f32 SIFT-shaped
synthetic config of BASELINE.json (configs[1]), built and searched on the device.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA)
  python bench.py --impl reference [...]                         the crate's CPU algorithm
                                                                 (oracle port, OpenMP, all cores)

One "step" = one pass of the traversal kernel over one batch of `--nq` queries.
`value`  : queries/s with queries and outputs resident in HBM (CUDA events, max over ranks).
`e2e`    : the same through the host C-ABI call phnsw_search_batch with pinned HOST buffers --
           H2D of the queries and D2H of ids/distances inside the timed region.
`roofline`: algorithmic bytes of the traversal kernel (n_dist * row_bytes +
           sum_layers n_exp * M * 4 + query + results; n_dist / n_exp are counted by the kernel
           and cross-checked here against the CPU oracle on a sample) / kernel time, against the
           measured HBM copy peak of MEASURED_PEAKS.json.
N > 1     : every rank holds a replica of the index and its own query batch (weak scaling in
           queries, no data-path collective); a second, separately reported region runs the
           sharded mode (each rank owns a different 1M-vector sub-index, queries broadcast,
           per-shard top-k all-gathered over NCCL and merged on the device).
Nothing here reads /root/reference.  The oracle is used only as the checker / CPU baseline.
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

METRIC_NAME = "search QPS at recall@10 (1M x 128 f32, L2, ef=300)"


def sift_like(n, dim, seed, latent=16, n_clusters=1024, cs=1.5, noise=0.3):
    """SIFT-shaped synthetic rows: a mixture of 1024 Gaussian clusters on a
    low-dimensional manifold, non-negative, rounded to integers in [0, 218], stored as f32.
    Generated with torch's CPU generator so that both arms see identical data."""
    import torch
    g = torch.Generator().manual_seed(555)
    A = torch.randn(latent, dim, generator=g) / latent ** 0.5
    C = torch.randn(n_clusters, latent, generator=g) * cs
    g = torch.Generator().manual_seed(seed)
    out = torch.empty((n, dim), dtype=torch.float32)
    step = 1 << 18
    for s in range(0, n, step):
        m = min(step, n - s)
        z = torch.randn(m, latent, generator=g) + C[torch.randint(0, n_clusters, (m,), generator=g)]
        x = z @ A + noise * torch.randn(m, dim, generator=g)
        out[s:s + m] = torch.clamp(torch.round(30.0 * x + 80.0), 0.0, 218.0)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_cores():
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1: ask the OS)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def recall_at_k(ids, gt, k):
    hit = 0
    for a, b in zip(ids, gt):
        hit += len(set(int(x) for x in a[:k]) & set(int(x) for x in b[:k]))
    return hit / (len(gt) * k)


def algorithmic_bytes(ndist, nexp, layer_M, dim, nq, k):
    """per query n_dist * row_bytes + sum_l n_exp(l) * M_l * 4 + query + k * 12."""
    return (float(ndist.sum()) * dim * 4 + float((nexp.sum(0) * np.asarray(layer_M)).sum()) * 4
            + nq * (dim * 4 + k * 12))


def run_reference(args):
    """--impl reference: the crate's CPU search path (oracle port; the Rust crate cannot be
    compiled here) on all host cores, on the same data and -- when a device is present -- the
    same 1M graph (built on the device as a fixture and handed over in serialize.rs form);
    each step = a bounded sample of the query batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    cores = host_cores()
    k = 10
    sample_q = min(args.nq, args.ref_queries)
    rows = sift_like(args.n, args.dim, 1234).numpy()
    queries = sift_like(args.nq, args.dim, 4321).numpy()[:sample_q]
    graph = "device-built fixture"
    layers = None
    try:
        import parallel_hnsw_b200 as ph
        if ph.device_count() > 0:
            comp = ph.BigComparator(rows, ph.L2_SQRT)
            gh = ph.Hnsw.generate(comp, seed=1)
            layers = gh.layers()
            gh.close()
            comp.close()
    except Exception as e:  # no device / no library: fall back to a CPU-built sample index
        layers = None
        graph = "unavailable (%s)" % type(e).__name__
    if layers is None:
        n_small = min(args.n, 100000)
        rows = rows[:n_small]
        oh = orc.Hnsw.generate(orc.L2_SQRT, rows, seed=1, improve=False)
        graph = "oracle-built %d-vector index without improve_index" % n_small
    else:
        oh = orc.Hnsw.from_layers(orc.L2_SQRT, rows, layers)
    sp = orc.search_params(args.ef, args.ef, 2)
    for _ in range(args.warmup):
        oh.search(queries=queries[:max(64, sample_q // 8)], sp=sp, max_out=k, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oh.search(queries=queries, sp=sp, max_out=k, nthreads=cores)
    dt = time.perf_counter() - t0
    qps = sample_q * args.steps / dt
    sample = "%d of %d queries per step, %s, %d OpenMP threads" % (sample_q, args.nq, graph, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC_NAME, "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1M x 128 f32 SIFT-shaped synthetic, L2, search ef=%d" % args.ef,
                   "n_vectors": int(rows.shape[0]), "dim": args.dim, "queries_per_step": sample_q},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--vectors", dest="n", type=int, default=1000000)
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--ef", type=int, default=300)
    ap.add_argument("--ref-queries", type=int, default=2000)
    ap.add_argument("--cpu-queries", type=int, default=2000)
    ap.add_argument("--no-improve", action="store_true")
    ap.add_argument("--profile-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed region (ncu --profile-from-start off)")
    ap.add_argument("--sum-order", default="tree", choices=["tree", "sequential"],
                    help="summation order of the traversal kernel's distances (include/phnsw.h)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import parallel_hnsw_b200 as ph

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or ph.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")  # any level >= VERSION prints a banner on stdout; keep
                                          # stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    k = 10
    stream = torch.cuda.current_stream().cuda_stream

    # ---- data + index (replica: same seed on every rank) --------------------------------
    t0 = time.perf_counter()
    rows_h = sift_like(args.n, args.dim, 1234)
    queries_h = sift_like(args.nq, args.dim, 4321 + (rank if world > 1 else 0))
    t_gen = time.perf_counter() - t0
    comp = ph.BigComparator(rows_h.numpy(), ph.L2_SQRT, device=local)
    # one tiny build first: CUDA module load and allocator warm-up are not build throughput
    warm = ph.BigComparator(rows_h.numpy()[:4096], ph.L2_SQRT, device=local)
    ph.Hnsw.generate(warm, seed=1, improve=not args.no_improve).close()
    warm.close()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gh = ph.Hnsw.generate(comp, seed=1, improve=not args.no_improve)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    L = gh.layer_count()
    layer_M = [gh.get_layer_from_top(i)[2] for i in range(L)] if rank == 0 else None

    sp = ph.SearchParameters(args.ef, args.ef, 2)
    tree = args.sum_order == "tree"
    dq = queries_h.to(dev)
    oi = torch.empty((args.nq, k), dtype=torch.int64, device=dev)
    od = torch.empty((args.nq, k), dtype=torch.float32, device=dev)
    oc = torch.empty((args.nq,), dtype=torch.int32, device=dev)
    nd = torch.zeros((args.nq, L), dtype=torch.int32, device=dev)
    ne = torch.zeros((args.nq, L), dtype=torch.int32, device=dev)

    # ---- correctness at full size: exact ground truth + oracle cross-check on a sample ----
    comp.bruteforce_knn(dq[:256], k)  # warm-up (allocator, kernel attributes)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gt, _ = comp.bruteforce_knn(dq, k)
    torch.cuda.synchronize()
    t_gt = time.perf_counter() - t0
    gt_stats = comp.bruteforce_last_stats()
    # secondary number: the sequential summation order (bit-identical to the crate's loops)
    gh.set_sum_order(ph.SUM_SEQUENTIAL)
    for _ in range(args.warmup):
        gh.search_device(dq, sp, oi, od, oc, stream=stream)
    gh.sync(stream)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    seq_steps = max(3, args.steps // 4)
    s0.record()
    for _ in range(seq_steps):
        gh.search_device(dq, sp, oi, od, oc, stream=stream)
    s1.record()
    gh.sync(stream)
    ms_seq = s0.elapsed_time(s1) / seq_steps
    seq_ids, seq_ds = oi.cpu().numpy().astype(np.uint64), od.cpu().numpy().copy()
    gh.set_sum_order(ph.SUM_TREE if tree else ph.SUM_SEQUENTIAL)
    gh.search_device(dq, sp, oi, od, oc, stream=stream, out_ndist=nd, out_nexp=ne)
    gh.sync(stream)
    recall = recall_at_k(oi.cpu().numpy(), gt.cpu().numpy(), k)
    ndist, nexp = nd.cpu().numpy().astype(np.int64), ne.cpu().numpy().astype(np.int64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region 1: device-resident ------------------------------------------------
    for _ in range(args.warmup):
        gh.search_device(dq, sp, oi, od, oc, stream=stream)
    gh.sync(stream)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.nvtx.range_push("timed")
    if args.profile_range:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        gh.search_device(dq, sp, oi, od, oc, stream=stream)
    e1.record()
    barrier()
    if args.profile_range:
        torch.cuda.profiler.stop()
    torch.cuda.nvtx.range_pop()
    gh.sync(stream)
    ms_dev = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- secondary: the same steps issued round robin on two streams, so that the ragged end
    # of one batch (4 query-lengths per launch: ~12 % of the SM-time is tail) overlaps the start
    # of the next -- what a server with back-to-back batches sees
    ss = [torch.cuda.Stream(device=dev) for _ in range(2)]
    outs2 = [(torch.empty_like(oi), torch.empty_like(od), torch.empty_like(oc)) for _ in ss]
    for i, s_ in enumerate(ss):
        gh.search_device(dq, sp, *outs2[i], stream=s_.cuda_stream)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for s_ in ss:
        s_.wait_event(p0)
    for r in range(args.steps):
        gh.search_device(dq, sp, *outs2[r % 2], stream=ss[r % 2].cuda_stream)
    for s_ in ss:
        torch.cuda.current_stream().wait_stream(s_)
    p1.record()
    torch.cuda.synchronize()
    ms_pipe = p0.elapsed_time(p1) / args.steps
    assert torch.equal(outs2[0][0], oi), "two-stream run disagrees with the single-stream run"

    # ---- secondary: operating points (the metric is QPS at recall@10 >= 0.95; the timed
    # configuration above is the crate's default 300 / 300 / 2) --------------------------------
    sweep = []
    if rank == 0:
        for ef_s in (300, 200, 150, 100, 64, 32):
            sp_s = ph.SearchParameters(ef_s, ef_s, 2)
            gh.search_device(dq, sp_s, oi, od, oc, stream=stream)
            gh.sync(stream)
            rec_s = recall_at_k(oi.cpu().numpy(), gt.cpu().numpy(), k)
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record()
            for _ in range(5):
                gh.search_device(dq, sp_s, oi, od, oc, stream=stream)
            w1.record()
            gh.sync(stream)
            sweep.append({"ef": ef_s, "recall_at_10": rec_s,
                          "qps": args.nq * 5 / (w0.elapsed_time(w1) * 1e-3)})
        gh.search_device(dq, sp, oi, od, oc, stream=stream)  # restore the timed run's outputs
        gh.sync(stream)

    # ---- timed region 2: end to end through the host C-ABI call with pinned host buffers --
    q_pin = queries_h.pin_memory().numpy()
    hi = torch.empty((args.nq, k), dtype=torch.int64).pin_memory()
    hd = torch.empty((args.nq, k), dtype=torch.float32).pin_memory()
    hc = torch.empty((args.nq,), dtype=torch.int32).pin_memory()
    import ctypes as C
    from parallel_hnsw_b200 import _native as N

    def e2e_step():
        N.check(N.lib().phnsw_search_batch(
            gh._h, C.c_void_p(q_pin.ctypes.data), None, args.nq, C.byref(sp), 0, None, k,
            C.c_void_p(hi.data_ptr()), C.c_void_p(hd.data_ptr()), C.c_void_p(hc.data_ptr()),
            None, None))

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    assert np.array_equal(hi.numpy(), oi.cpu().numpy()), "host path and device path disagree"

    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])

    # ---- sharded mode (N > 1): own sub-index per rank, broadcast queries, NCCL all-gather ----
    sharded = None
    if world > 1:
        sharded = run_sharded(args, ph, dist, dev, rank, world, k, stream)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (oracle port on the same graph, bounded sample) + parity at full size ----
    from oracle import oracle as orc
    cq = min(args.nq, args.cpu_queries)
    oh = orc.Hnsw.from_layers(orc.L2_SQRT, rows_h.numpy(), gh.layers())
    osp = orc.search_params(args.ef, args.ef, 2)
    cores = host_cores()
    oh.search(queries=queries_h.numpy()[:64], sp=osp, max_out=k, nthreads=cores)
    t0 = time.perf_counter()
    q_ids, q_ds, _, q_nd, q_ne = oh.search(queries=queries_h.numpy()[:cq], sp=osp, max_out=k,
                                           stats=True, nthreads=cores)
    cpu_dt = time.perf_counter() - t0  # the crate's algorithm (sequential sums) is the baseline
    g_ids = oi.cpu().numpy().astype(np.uint64)[:cq]
    # the timed order against the crate's order: BASELINE.json's bar (>= 99.9 % / 1e-5)
    ids_equal_seq = float((g_ids == q_ids).all(1).mean())
    m = (g_ids == q_ids)
    rel_seq = np.abs(od.cpu().numpy()[:cq].astype(np.float64) - q_ds) / np.maximum(np.abs(q_ds), 1e-30)
    max_rel_seq = float(rel_seq[m].max()) if m.any() else None
    seq_dev_equal = float((seq_ids[:cq] == q_ids).all(1).mean())
    if tree:  # and bit for bit against the oracle restating the same tree order
        oh.set_sum_order(1)
        o_ids, o_ds, o_cnt, o_nd, o_ne = oh.search(queries=queries_h.numpy()[:cq], sp=osp, max_out=k,
                                                   stats=True, nthreads=cores)
        oh.set_sum_order(0)
    else:
        o_ids, o_ds, o_nd, o_ne = q_ids, q_ds, q_nd, q_ne
    ids_equal = float((g_ids == o_ids).all(1).mean())
    counters_equal = float(((ndist[:cq] == o_nd.astype(np.int64)).all(1)
                            & (nexp[:cq] == o_ne.astype(np.int64)).all(1)).mean())
    rel = np.abs(od.cpu().numpy()[:cq].astype(np.float64) - o_ds) / np.maximum(np.abs(o_ds), 1e-30)
    max_rel = float(rel[(g_ids == o_ids)].max()) if (g_ids == o_ids).any() else None

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"
    abytes = algorithmic_bytes(ndist, nexp, layer_M, args.dim, args.nq, k)
    kernel_ms = ms_dev / args.steps
    achieved = abytes / (kernel_ms * 1e-3) / 1e9
    qps = world * args.nq * args.steps / (ms_dev * 1e-3)
    e2e_qps = world * args.nq * args.steps / (ms_e2e * 1e-3)
    out = {
        "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": kernel_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "1M x 128 f32 SIFT-shaped synthetic, L2, build + search ef=%d" % args.ef,
                   "n_vectors": args.n, "dim": args.dim, "queries_per_step_per_gpu": args.nq, "k": k,
                   "search": {"number_of_candidates": args.ef, "upper_layer_candidate_count": args.ef,
                              "probe_depth": 2},
                   "sum_order": args.sum_order,
                   "layers_top_first": gh.layer_sizes(),
                   "cache": "inputs larger than L2 (rows %.0f MB + graph %.0f MB vs 126 MB L2)" % (
                       args.n * args.dim * 4 / 1e6, args.n * 48 * 4 / 1e6),
                   "parallelism": "replicas x%d (queries split)" % world if world > 1 else "single GPU"},
        "recall_at_10": recall,
        "build": {"vectors_per_s": args.n / t_build, "seconds": t_build,
                  "improve_index": not args.no_improve, "data_gen_seconds": t_gen},
        "parity": {"sample_queries": cq, "sum_order": args.sum_order,
                   "oracle_same_order": {"ids_equal_frac": ids_equal,
                                         "work_counters_equal_frac": counters_equal,
                                         "max_rel_dist_err": max_rel},
                   "oracle_crate_order": {"ids_equal_frac": ids_equal_seq,
                                          "max_rel_dist_err": max_rel_seq},
                   "sequential_kernel_vs_oracle_ids_equal_frac": seq_dev_equal,
                   "ids_equal_frac": ids_equal_seq, "max_rel_dist_err": max_rel_seq},
        "ground_truth": {
            "what": "exact brute-force kNN of the query batch (recall denominator)",
            "seconds": t_gt, "path": gt_stats["path"],
            "filter_kernel": "tc_filter_kernel (tcgen05 bf16 hi/lo split GEMM, M128 N128 K16)",
            "filter_ms": gt_stats["filter_ms"],
            "filter_tflops": (gt_stats["filter_flops"] / gt_stats["filter_ms"] / 1e9
                              if gt_stats["filter_ms"] > 0 else None),
            "tensor_peak_tflops": peaks.get("bf16_tflops"),
            "frac_of_tensor_peak": (gt_stats["filter_flops"] / gt_stats["filter_ms"] / 1e9
                                    / peaks["bf16_tflops"]
                                    if gt_stats["filter_ms"] > 0 and peaks.get("bf16_tflops") else None),
            "max_candidates_per_query": gt_stats["max_candidates"]},
        "operating_points": {
            "note": "one GPU, same index, ef = number_of_candidates = upper_layer_candidate_count",
            "sweep": sweep,
            "best_qps_at_recall_ge_0.95": max([p_["qps"] for p_ in sweep if p_["recall_at_10"] >= 0.95],
                                              default=None)},
        "two_streams": {"value": world * args.nq / (ms_pipe * 1e-3), "unit": "queries/s",
                        "ms_per_step": ms_pipe,
                        "note": "steps issued alternately on two streams (tails overlap); "
                                "`value` above is the plain single-stream number"},
        "sequential_order": {"value": world * args.nq / (ms_seq * 1e-3), "unit": "queries/s",
                             "ms_per_step": ms_seq,
                             "note": "same kernel with PHNSW_SUM_SEQUENTIAL (the crate's loop bit for bit)"},
        "e2e": {"value": e2e_qps, "unit": "queries/s",
                "h2d_bytes_per_step": int(args.nq * args.dim * 4),
                "d2h_bytes_per_step": int(args.nq * (k * 12 + 4))},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "search_kernel<L2_SQRT, %s>" % ("tree" if tree else "sequential"),
                     "algorithmic_bytes_per_launch": abytes,
                     "n_dist_per_query": float(ndist.sum() / args.nq),
                     "n_exp_per_query": float(nexp.sum() / args.nq)},
        "cpu_baseline": {"value": cq / cpu_dt, "unit": "queries/s", "cores": cores,
                         "kind": "port",
                         "sample": "%d of %d queries on the same device-built graph" % (cq, args.nq)},
        "clocks": clocks,
    }
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                out["roofline"]["traffic"] = json.load(f).get("search_kernel_dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
    if sharded is not None:
        out["sharded"] = sharded
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_sharded(args, ph, dist, dev, rank, world, k, stream):
    """Each rank builds and searches its own sub-index (rows seeded by rank, global id =
    rank * n + local id); queries are broadcast from rank 0; per-shard top-k are all-gathered
    over NCCL and merged by (distance, id) on the device."""
    import torch
    rows_h = sift_like(args.n, args.dim, 1234 + 7919 * (rank + 1))
    comp = ph.BigComparator(rows_h.numpy(), ph.L2_SQRT, device=dev.index)
    gh = ph.Hnsw.generate(comp, seed=1 + rank, improve=not args.no_improve)
    gh.set_sum_order(ph.SUM_TREE if args.sum_order == "tree" else ph.SUM_SEQUENTIAL)
    sp = ph.SearchParameters(args.ef, args.ef, 2)
    dq = sift_like(args.nq, args.dim, 4321).to(dev) if rank == 0 else torch.empty(
        (args.nq, args.dim), dtype=torch.float32, device=dev)
    oi = torch.empty((args.nq, k), dtype=torch.int64, device=dev)
    od = torch.empty((args.nq, k), dtype=torch.float32, device=dev)
    oc = torch.empty((args.nq,), dtype=torch.int32, device=dev)
    gi = torch.empty((world, args.nq, k), dtype=torch.int64, device=dev)
    gd = torch.empty((world, args.nq, k), dtype=torch.float32, device=dev)
    mi = torch.empty((args.nq, k), dtype=torch.int64, device=dev)
    md = torch.empty((args.nq, k), dtype=torch.float32, device=dev)

    from parallel_hnsw_b200.sharded import ShardedHnsw
    sh = ShardedHnsw(gh, rank * args.n, rank, world)

    def step():
        r = sh.search(dq, sp, k, src=0, stream=stream)
        mi.copy_(r[0])
        md.copy_(r[1])

    for _ in range(args.warmup):
        step()
    gh.sync(stream)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    gh.sync(stream)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # recall of the merged result against the exact ground truth over all shards
    gt_i, gt_d = comp.bruteforce_knn(dq, k)
    ggi = torch.empty((world, args.nq, k), dtype=torch.int64, device=dev)
    ggd = torch.empty((world, args.nq, k), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(ggi, gt_i + rank * args.n)
    dist.all_gather_into_tensor(ggd, gt_d)
    ei = torch.empty((args.nq, k), dtype=torch.int64, device=dev)
    ed = torch.empty((args.nq, k), dtype=torch.float32, device=dev)
    ph.merge_topk_device(ggi, ggd, world, args.nq, k, ei, ed, stream)
    torch.cuda.synchronize()
    rec = recall_at_k(mi.cpu().numpy(), ei.cpu().numpy(), k)
    ms = float(t[0])
    gh.close()
    comp.close()
    return {"mode": "sharded sub-indexes + NCCL all-gather top-k merge", "shards": world,
            "vectors_total": world * args.n, "queries_per_step": args.nq,
            "value": args.nq * args.steps / (ms * 1e-3), "unit": "queries/s",
            "ms_per_step": ms / args.steps, "recall_at_10": rec}


if __name__ == "__main__":
    main()
