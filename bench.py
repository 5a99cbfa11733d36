#!/usr/bin/env python
"""Benchmark of the B200 HNSW engine on BASELINE.json's configs.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA)
  python bench.py --impl reference [...]                         the crate's CPU algorithm
                                                                 (oracle port, OpenMP, all cores)

Headline (every N): BASELINE configs[1] -- 1M x 128 f32 SIFT-shaped synthetic, L2, build + batched
search, one "step" = one pass of the traversal kernel over one batch of `--nq` queries.
`value`   : queries/s with queries and outputs resident in HBM (CUDA events, max over ranks).
`e2e`     : the same through the host C-ABI call phnsw_search_batch with pinned HOST buffers --
            H2D of the queries and D2H of ids/distances inside the timed region.
`roofline`: algorithmic bytes of the traversal kernel (n_dist * row_bytes +
            sum_layers n_exp * M * 4 + query + results; n_dist / n_exp are counted by the kernel
            and cross-checked here against the CPU oracle on a sample) / kernel time, against the
            measured HBM copy peak of MEASURED_PEAKS.json.
Secondary blocks in the same JSON line, each with its own build time, recall, parity sample
against the CPU oracle and roofline:
  N = 1      `config3`: configs[2] -- 1M x 1536 cosine, PQ8 (96 codes, K = 256): k-means training
             + encoding on the tensor cores, ADC search + exact re-rank as ONE library call.
  N > 1      `sharded`: configs[3] -- 10M x 96 split N ways, one sub-index per GPU, queries
             broadcast, per-shard top-k exchanged with one ncclAllGather inside the library
             (phnsw_search_batch_sharded), merged on the device.
  N = 8      `config5`: configs[4] -- 8 x 12.5M x 128 generated on the device, PQ8-coded (16 codes
  (or --c5)  per vector), ADC walk + exact re-rank per shard + the same exchange, 10 000-query
             batches.
  N > 1 also: top-level `value` is the replica mode (index replicated, queries split, no
  data-path collective; weak scaling in queries).
Nothing here reads /root/reference.  The oracle is used only as the checker / CPU baseline.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC_NAME = "search QPS at recall@10 (1M x 128 f32, L2, ef=300)"


# ------------------------------------------------------------------------------ synthetic data
def sift_like(n, dim, seed, latent=16, n_clusters=1024, cs=1.5, noise=0.3):
    """SIFT-shaped synthetic rows: a mixture of 1024 Gaussian clusters on a
    low-dimensional manifold, non-negative, rounded to integers in [0, 218], stored as f32.
    Generated with torch's CPU generator so that both arms see identical data."""
    import torch
    g = torch.Generator().manual_seed(555)
    A = torch.randn(latent, dim, generator=g) / latent ** 0.5
    C_ = torch.randn(n_clusters, latent, generator=g) * cs
    g = torch.Generator().manual_seed(seed)
    out = torch.empty((n, dim), dtype=torch.float32)
    step = 1 << 18
    for s in range(0, n, step):
        m = min(step, n - s)
        z = torch.randn(m, latent, generator=g) + C_[torch.randint(0, n_clusters, (m,), generator=g)]
        x = z @ A + noise * torch.randn(m, dim, generator=g)
        out[s:s + m] = torch.clamp(torch.round(30.0 * x + 80.0), 0.0, 218.0)
    return out


class DeviceMixture:
    """Synthetic rows generated in HBM (configs 3-5 are too large to ship from the host): a
    Gaussian mixture of `n_clusters` centres on a `latent`-dimensional manifold embedded in
    `dim` dimensions plus isotropic noise.  The mixture is keyed by `mix_seed` (shared by rows
    and queries and by every rank), the draws by `seed`.
      kind "embedding": unit norm (cosine)                 -- configs[2]
      kind "deep"     : unit norm, used with L2            -- configs[3]
      kind "sift"     : non-negative integers in [0, 218]  -- configs[4] (the PQ-coded 100M)"""

    def __init__(self, kind, dim, dev, mix_seed, latent, n_clusters, spread, noise):
        import torch
        self.kind, self.dim, self.dev, self.noise, self.latent = kind, dim, dev, noise, latent
        g = torch.Generator(device=dev).manual_seed(mix_seed)
        self.basis = torch.randn(latent, dim, generator=g, device=dev) / latent ** 0.5
        self.centers = torch.randn(n_clusters, latent, generator=g, device=dev) * spread

    def rows(self, n, seed):
        import torch
        g = torch.Generator(device=self.dev).manual_seed(seed)
        out = torch.empty((n, self.dim), dtype=torch.float32, device=self.dev)
        step = 1 << 18
        for s in range(0, n, step):
            m = min(step, n - s)
            cl = torch.randint(0, self.centers.shape[0], (m,), generator=g, device=self.dev)
            z = self.centers[cl] + torch.randn(m, self.latent, generator=g, device=self.dev)
            x = z @ self.basis + self.noise * torch.randn(m, self.dim, generator=g, device=self.dev)
            if self.kind == "sift":
                out[s:s + m] = torch.clamp(torch.round(30.0 * x + 80.0), 0.0, 218.0)
            else:
                out[s:s + m] = x / x.norm(dim=1, keepdim=True)
        return out


def mixture_for(kind, dev):
    if kind == "embedding":   # 2048 topic clusters on a 24-d manifold, small isotropic noise
        return DeviceMixture("embedding", 1536, dev, 2024, 24, 2048, 2.0, 0.01)
    if kind == "deep":
        return DeviceMixture("deep", 96, dev, 96, 16, 1024, 1.5, 0.05)
    return DeviceMixture("sift", 128, dev, 555, 16, 1024, 1.5, 0.3)


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
    ids = np.asarray(ids)[:, :k].astype(np.int64)
    gt = np.asarray(gt)[:, :k].astype(np.int64)
    hit = 0
    for a, b in zip(ids, gt):
        hit += len(set(a.tolist()) & set(b.tolist()))
    return hit / (len(gt) * k)


def algorithmic_bytes(ndist, nexp, layer_M, row_bytes, query_bytes, nq, k):
    """per query n_dist * row_bytes + sum_l n_exp(l) * M_l * 4 + query + k * 12."""
    return (float(ndist.sum()) * row_bytes + float((nexp.sum(0) * np.asarray(layer_M)).sum()) * 4
            + nq * (query_bytes + k * 12))


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def hbm_peak(peaks):
    if "hbm_gbs" in peaks:
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6650 (B200_PROFILING.md)"


def main_config(args):
    """`config` of the headline line: identical in both arms (it is derived from the arguments
    only), so that the driver's same-config check compares like with like."""
    return {"workload": "1M x 128 f32 SIFT-shaped synthetic, L2, build + search ef=%d" % args.ef,
            "n_vectors": args.n, "dim": args.dim, "queries_per_step_per_gpu": args.nq, "k": 10,
            "search": {"number_of_candidates": args.ef, "upper_layer_candidate_count": args.ef,
                       "probe_depth": 2},
            "cache": "inputs larger than L2 (rows %.0f MB + graph %.0f MB vs 126 MB L2)" % (
                args.n * args.dim * 4 / 1e6, args.n * 48 * 4 / 1e6)}


FIXTURE_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
import bench
import parallel_hnsw_b200 as ph
if ph.device_count() == 0:
    raise SystemExit(3)
rows = bench.sift_like(%(n)d, %(dim)d, 1234).numpy()
comp = ph.BigComparator(rows, ph.L2_SQRT)
gh = ph.Hnsw.generate(comp, seed=1)
gh.serialize(%(out)r)
"""


def run_reference(args):
    """--impl reference: the crate's CPU search path (oracle port; the Rust crate cannot be
    compiled here) on all host cores, on the same data, the same query batch per step and --
    when a device is present -- the same 1M graph.  That graph is a fixture: a SUBPROCESS builds it
    on the device and writes it in the crate's serialize.rs layout; this process only reads the
    directory with the oracle's loader, so the timing process never maps the product library."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    cores = host_cores()
    k = 10
    queries = sift_like(args.nq, args.dim, 4321).numpy()
    graph = "device-built fixture read from a serialize.rs directory"
    oh = None
    with tempfile.TemporaryDirectory(prefix="phnsw_ref_") as tmp:
        out = os.path.join(tmp, "index")
        code = FIXTURE_SCRIPT % {"root": ROOT, "n": args.n, "dim": args.dim, "out": out}
        env = dict(os.environ)
        for v in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(v, None)
        try:
            r = subprocess.run([sys.executable, "-c", code], env=env, timeout=600,
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            if r.returncode == 0:
                oh = orc.Hnsw.deserialize(out)
            else:
                graph = "fixture subprocess rc %d" % r.returncode
        except Exception as e:  # no device / no library
            graph = "fixture unavailable (%s)" % type(e).__name__
    n_vec = args.n
    if oh is None:
        n_vec = min(args.n, 100000)
        rows = sift_like(n_vec, args.dim, 1234).numpy()
        oh = orc.Hnsw.generate(orc.L2_SQRT, rows, seed=1, improve=False)
        graph = "oracle-built %d-vector index without improve_index (%s)" % (n_vec, graph)
    sp = orc.search_params(args.ef, args.ef, 2)
    for _ in range(args.warmup):
        oh.search(queries=queries, sp=sp, max_out=k, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oh.search(queries=queries, sp=sp, max_out=k, nthreads=cores)
    dt = time.perf_counter() - t0
    qps = args.nq * args.steps / dt
    sample = "%d queries per step (the full batch), %s, %d OpenMP threads" % (args.nq, graph, cores)
    cfg = main_config(args)
    cfg["n_vectors"] = n_vec
    print(json.dumps({
        "impl": "reference", "metric": METRIC_NAME, "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------ helpers (GPU)
class Timer:
    """CUDA-event timing of `steps` calls of fn on torch's current stream."""

    def __init__(self, torch):
        self.torch = torch

    def run(self, fn, steps, warmup, sync=None):
        torch = self.torch
        for _ in range(warmup):
            fn()
        if sync:
            sync()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if sync:
            sync()
        return e0.elapsed_time(e1) / steps


def guarded(name, fn):
    """A secondary block must never take the headline line down with it."""
    try:
        return fn()
    except Exception as e:  # noqa: BLE001
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:300]),
                "trace": traceback.format_exc()[-600:], "block": name}


def cpu_build_baseline(sizes, dim, cores):
    """Reference-style CPU build (oracle restatement of Hnsw::generate incl. improve_index,
    src/lib.rs:825-893, 1546-1603) on the host cores: vectors/s at each size."""
    from oracle import oracle as orc
    out = []
    for n in sizes:
        rows = sift_like(n, dim, 1234).numpy()
        t0 = time.perf_counter()
        orc.Hnsw.generate(orc.L2_SQRT, rows, seed=1, improve=True, nthreads=cores)
        dt = time.perf_counter() - t0
        out.append({"n_vectors": n, "seconds": dt, "vectors_per_s": n / dt})
    return {"kind": "port", "cores": cores, "what": "oracle generate + improve_index, same data "
            "generator and parameters as the device build", "runs": out}


# ------------------------------------------------------------------------------ headline block
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
    ap.add_argument("--cpu-queries", type=int, default=2000)
    ap.add_argument("--cpu-build", default="10000,100000",
                    help="sizes of the CPU build baseline (comma separated, '' = skip)")
    ap.add_argument("--no-improve", action="store_true")
    ap.add_argument("--builds", type=int, default=4,
                    help="full-size builds: the first is reported apart, the rest give the median")
    ap.add_argument("--no-overlap", action="store_true",
                    help="time the plain launches (phnsw_index_set_batch_overlap off)")
    ap.add_argument("--profile-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed region (ncu --profile-from-start off)")
    ap.add_argument("--sum-order", default="tree", choices=["tree", "sequential"],
                    help="summation order of the traversal kernel's distances (include/phnsw.h)")
    ap.add_argument("--blocks", default="auto",
                    help="secondary blocks: auto | none | comma list of config3,sharded,config5")
    ap.add_argument("--block-timeout", type=int, default=700,
                    help="seconds the secondary blocks may take before the headline line is printed without them")
    ap.add_argument("--adc-table", default="q8", choices=["q8", "f32"],
                    help="per-query ADC table form of the timed ADC blocks (include/phnsw.h "
                         "phnsw_pq8_store_set_adc_table); the other form is reported beside it")
    ap.add_argument("--c5-rerank", type=int, default=300,
                    help="config 5: ADC hits re-scored exactly per shard (300 = all candidates, "
                         "what the crate's QuantizedHnsw::search does)")
    ap.add_argument("--c3-n", type=int, default=1000000)
    ap.add_argument("--c4-n", type=int, default=10000000, help="config 4: total vectors over all ranks")
    ap.add_argument("--c5-n", type=int, default=12500000, help="config 5: vectors per rank")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import parallel_hnsw_b200 as ph
    from parallel_hnsw_b200 import _native as N

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
    if args.blocks == "auto":
        blocks = {"config3"} if world == 1 else {"sharded"}
        if world == 8:
            blocks.add("config5")
    elif args.blocks == "none":
        blocks = set()
    else:
        blocks = set(args.blocks.split(","))
    k = 10
    stream = torch.cuda.current_stream().cuda_stream
    peaks = load_peaks()
    ctx = {"torch": torch, "dist": dist, "ph": ph, "N": N, "dev": dev, "rank": rank, "world": world,
           "local": local, "k": k, "stream": stream, "peaks": peaks, "args": args}

    # ---- data + index.  N = 1: the BASELINE configs[1] index.  N > 1: the same workload per GPU,
    # sharded -- rank r owns vectors [r * n, (r + 1) * n) of an N * n vector set (its own seed), builds
    # and searches its own sub-index, every rank searches the SAME query batch (broadcast from rank
    # 0 inside the step) and the per-shard top-k are merged through one NCCL all-gather
    # (phnsw_search_batch_sharded).  `value` counts the (query, shard) searches all ranks did.
    sharded = world > 1
    t0 = time.perf_counter()
    rows_h = sift_like(args.n, args.dim, 1234 + rank)
    queries_h = sift_like(args.nq, args.dim, 4321)
    t_gen = time.perf_counter() - t0
    comp = ph.BigComparator(rows_h.numpy(), ph.L2_SQRT, device=local)
    # one tiny build first: CUDA module load and allocator warm-up are not build throughput
    warm = ph.BigComparator(rows_h.numpy()[:4096], ph.L2_SQRT, device=local)
    ph.Hnsw.generate(warm, seed=1, improve=not args.no_improve).close()
    warm.close()
    torch.cuda.synchronize()
    # build throughput: the first full-size build of the process pays the driver for several GB
    # of construction temporaries (0.0-1.4 s, box dependent); the library keeps them in its own
    # memory pool, so every later build is the steady state.  Both are reported: `seconds` is
    # the median of the builds after the first, `first_build_seconds` the first.
    build_times = []
    gh = None
    for _ in range(max(2, args.builds)):
        if gh is not None:
            gh.close()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gh = ph.Hnsw.generate(comp, seed=1, improve=not args.no_improve)
        torch.cuda.synchronize()
        build_times.append(time.perf_counter() - t0)
    t_build = float(np.median(build_times[1:]))
    # one more, untimed, instrumented build: the work counters of every traversal launch of the
    # build (seed, link and recall searches: 93 % of its kernel time) for build.roofline
    build_work = None
    if rank == 0:
        os.environ["PHNSW_WORK_STATS"] = "1"
        try:
            gw = ph.Hnsw.generate(comp, seed=1, improve=not args.no_improve)
            torch.cuda.synchronize()
            build_work = gw.work_stats()
            gw.close()
        finally:
            os.environ.pop("PHNSW_WORK_STATS", None)
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
    gt, _ = comp.bruteforce_knn(dq, k)  # this rank's shard (ids are shard-local)
    torch.cuda.synchronize()
    t_gt = time.perf_counter() - t0
    gt_stats = comp.bruteforce_last_stats()
    sh = None
    if sharded:
        from parallel_hnsw_b200.sharded import ShardedHnsw
        sh = ShardedHnsw(gh, rank * args.n, rank, world)
        gt_global = sharded_ground_truth(ctx, comp, dq, rank * args.n)
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
    seq_ids = oi.cpu().numpy().astype(np.uint64)
    # ... and the same order with back-to-back batches overlapped (alternate output buffers)
    ms_seq_ov = None
    if not args.no_overlap:
        seq_alt = (torch.empty_like(oi), torch.empty_like(od), torch.empty_like(oc))
        gh.set_batch_overlap(True)
        for i_ in range(args.warmup + seq_steps * 2):
            if i_ == args.warmup:
                gh.sync(stream)
                s0.record()
            gh.search_device(dq, sp, *((oi, od, oc) if i_ & 1 else seq_alt), stream=stream)
        s1.record()
        gh.sync(stream)
        gh.set_batch_overlap(False)
        ms_seq_ov = s0.elapsed_time(s1) / (seq_steps * 2)
        assert np.array_equal(seq_alt[0].cpu().numpy().astype(np.uint64), seq_ids), \
            "sequential order: overlapped launches changed the results"
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
    mi = torch.empty((args.nq, k), dtype=torch.int64, device=dev)   # merged (sharded) results
    md = torch.empty((args.nq, k), dtype=torch.float32, device=dev)

    alt = (torch.empty_like(oi), torch.empty_like(od), torch.empty_like(oc))
    step_no = [0]

    # N > 1: the pipelined sharded step (phnsw_search_batch_sharded_queued): every rank holds the
    # batch, the shard search of a step runs on `stream` chained behind the previous step's (batch
    # overlap), its all-gather + merge on the library's side stream on rotating exchange
    # buffers; four rotating result buffers, one phnsw_comm_flush before the closing event --
    # the flush (all exchanges and merges complete) is inside the timed region
    ring = [(mi, md)] + [(torch.empty_like(mi), torch.empty_like(md)) for _ in range(3)]
    queued = sharded and not args.no_overlap

    def step():
        if queued:
            step_no[0] += 1
            sh.search_queued(dq, sp, k, stream=stream, out=ring[step_no[0] & 3])
        elif sharded:
            sh.search(dq, sp, k, src=0, stream=stream, out=(mi, md))
        else:
            # consecutive steps write different output buffers (two launches may be in flight
            # at once under batch overlap); the last step of a loop of even length lands in `alt`
            step_no[0] += 1
            o = (oi, od, oc) if step_no[0] & 1 else alt
            gh.search_device(dq, sp, *o, stream=stream)

    # back-to-back batches on one stream: the library may start a launch on the SMs the previous
    # one has already left (phnsw_index_set_batch_overlap; same results, checked below).  The
    # plain launches are timed right after as `no_batch_overlap`.
    use_overlap = not args.no_overlap
    gh.set_batch_overlap(use_overlap)
    for _ in range(args.warmup):
        step()
    if queued:
        sh.flush(stream)
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
        step()
    if queued:
        sh.flush(stream)
    e1.record()
    barrier()
    if args.profile_range:
        torch.cuda.profiler.stop()
    torch.cuda.nvtx.range_pop()
    gh.sync(stream)
    ms_dev = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    gh.set_batch_overlap(False)
    ms_plain = None
    ms_blocking = None
    if queued:
        # beside it: the blocking call per step (broadcast of the batch from rank 0 inside the
        # step, everything on one stream, no overlap between steps), same barriers
        q_last = ring[step_no[0] & 3]
        q_ids, q_ds = q_last[0].clone(), q_last[1].clone()
        for _ in range(args.warmup):
            sh.search(dq, sp, k, src=0, stream=stream, out=(mi, md))
        gh.sync(stream)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        b0.record()
        for _ in range(args.steps):
            sh.search(dq, sp, k, src=0, stream=stream, out=(mi, md))
        b1.record()
        barrier()
        gh.sync(stream)
        tb = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        ms_blocking = float(tb[0]) / args.steps
        assert torch.equal(q_ids, mi) and torch.equal(q_ds, md), "pipelined sharded step changed the results"
    if use_overlap and not sharded:
        ov_ids, ov_ds = oi.clone(), od.clone()
        for _ in range(args.warmup):
            step()
        gh.sync(stream)
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record()
        for _ in range(args.steps):
            step()
        n1.record()
        gh.sync(stream)
        ms_plain = n0.elapsed_time(n1) / args.steps
        assert torch.equal(ov_ids, oi) and torch.equal(ov_ds, od), "overlapped launches changed the results"
        assert torch.equal(alt[0], oi) and torch.equal(alt[1], od), "alternate output buffer differs"
    sharded_info = None
    if sharded:
        # beside it: the same shards searched WITHOUT broadcast / exchange / merge (what N
        # independent replicas of this per-GPU workload deliver), max over ranks
        for _ in range(args.warmup):
            gh.search_device(dq, sp, oi, od, oc, stream=stream)
        gh.sync(stream)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        r0.record()
        for _ in range(args.steps):
            gh.search_device(dq, sp, oi, od, oc, stream=stream)
        r1.record()
        barrier()
        gh.sync(stream)
        tl = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        ms_local = float(tl[0])
        # ... and with batch overlap between the steps (what the pipelined step competes with)
        gh.set_batch_overlap(use_overlap)
        for i_ in range(args.warmup + args.steps + 1):
            if i_ == args.warmup:
                gh.sync(stream)
                barrier()
                r0.record()
            o_ = (oi, od, oc) if i_ & 1 else alt
            gh.search_device(dq, sp, *o_, stream=stream)
        r1.record()
        barrier()
        gh.sync(stream)
        gh.set_batch_overlap(False)
        tl = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        ms_local_ov = float(tl[0]) / (args.steps + 1)
        merged_np = mi.cpu().numpy()
        sharded_info = {
            "recall_merged": recall_at_k(merged_np, gt_global, k),
            "ms_local": ms_local, "ms_blocking": ms_blocking, "queued": queued,
            "ms_local_overlap": ms_local_ov,
            # every merged entry that names a vector of this rank's shard must be this rank's own
            # result for that query, and every merged list ascending by (distance, id)
            "merged": merged_np, "merged_d": md.cpu().numpy()}

    # ---- secondary: the same steps issued round robin on two streams, so that the ragged end
    # of one batch overlaps the start of the next -- what a server with back-to-back batches sees
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
            sweep.append({"ef": ef_s, "recall_at_10": round(rec_s, 4),
                          "qps": round(args.nq * 5 / (w0.elapsed_time(w1) * 1e-3))})
        gh.search_device(dq, sp, oi, od, oc, stream=stream)  # restore the timed run's outputs
        gh.sync(stream)

    # ---- timed region 2: end to end through the host C-ABI call with pinned host buffers --
    q_pin = queries_h.pin_memory().numpy()
    hi = torch.empty((args.nq, k), dtype=torch.int64).pin_memory()
    hd = torch.empty((args.nq, k), dtype=torch.float32).pin_memory()
    hc = torch.empty((args.nq,), dtype=torch.int32).pin_memory()

    def e2e_step():
        N.check(N.lib().phnsw_search_batch(
            gh._h, C.c_void_p(q_pin.ctypes.data), None, args.nq, C.byref(sp), 0, None, k,
            C.c_void_p(hi.data_ptr()), C.c_void_p(hd.data_ptr()), C.c_void_p(hc.data_ptr()),
            None, None))

    ms_e2e_sync = None
    if sharded:
        # rank 0: pinned host queries -> H2D -> sharded step (broadcast, search, all-gather,
        # merge) -> D2H of the merged top-k; the other ranks take part in the step
        ms_e2e = timed_e2e_sharded(ctx, sh, gh, queries_h.pin_memory(), dq, sp, 0, args.steps,
                                   args.warmup) * args.steps
    else:
        # (a) one synchronous phnsw_search_batch per step (Hnsw::search's contract)
        for _ in range(args.warmup):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        ms_e2e_sync = (time.perf_counter() - t0) * 1e3
        assert np.array_equal(hi.numpy(), oi.cpu().numpy()), "host path and device path disagree"
        # (b) the same steps as a server drains a queue of batches: phnsw_search_batch_host_async
        # per step (pinned host queries read in place, results written straight to pinned host
        # memory, alternate output buffers), batch overlap on, ONE sync at the end -- every step's
        # host -> device read and device -> host write is inside the timed region
        hi2, hd2, hc2 = (torch.empty_like(hi).pin_memory(), torch.empty_like(hd).pin_memory(),
                         torch.empty_like(hc).pin_memory())
        q_pin_t = queries_h.pin_memory()

        def e2e_async(i):
            o = (hi, hd, hc) if i & 1 else (hi2, hd2, hc2)
            gh.search_host_async(q_pin_t, sp, *o, stream=stream)
        gh.set_batch_overlap(use_overlap)
        for i in range(args.warmup):
            e2e_async(i)
        gh.sync(stream)
        hi.zero_()
        hi2.zero_()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_async(i)
        gh.sync(stream)
        ms_e2e = (time.perf_counter() - t0) * 1e3
        gh.set_batch_overlap(False)
        ref_ids = oi.cpu().numpy()
        assert np.array_equal(hi.numpy(), ref_ids) and np.array_equal(hi2.numpy(), ref_ids), \
            "asynchronous host path and device path disagree"

    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])

    # ---- CPU baseline + parity at full size (rank 0, while the replica index is alive) ------
    main_cpu = None
    if rank == 0:
        main_cpu = guarded("cpu_baseline", lambda: headline_parity(
            ctx, gh, rows_h, queries_h, oi, od, ndist, nexp, seq_ids, tree))
    gh_layers_top_first = gh.layer_sizes()
    gh.close()
    comp.close()
    del rows_h
    torch.cuda.empty_cache()

    if rank == 0:
        out = headline_line(ctx, locals())
    else:
        out = None
    finish(ctx, out, blocks, sweep)


def headline_line(ctx, v):
    """The headline JSON object (rank 0), built before the secondary blocks run."""
    args, world, k, peaks = ctx["args"], ctx["world"], ctx["k"], ctx["peaks"]
    ms_seq_ov = v.get("ms_seq_ov")
    (ndist, nexp, layer_M, ms_dev, ms_e2e, ms_seq, ms_pipe, recall, main_cpu, clocks, tree, t_build,
     t_gen, t_gt, gt_stats, gh_layers_top_first, sharded_info, oi, od) = (v[x] for x in (
         "ndist", "nexp", "layer_M", "ms_dev", "ms_e2e", "ms_seq", "ms_pipe", "recall", "main_cpu",
         "clocks", "tree", "t_build", "t_gen", "t_gt", "gt_stats", "gh_layers_top_first",
         "sharded_info", "oi", "od"))
    build_times, ms_plain, use_overlap = v["build_times"], v["ms_plain"], v["use_overlap"]
    ms_e2e_sync = v["ms_e2e_sync"]
    build_work = v["build_work"]
    cpu_build = None
    if args.cpu_build:
        sizes = [int(x) for x in args.cpu_build.split(",") if x]
        cpu_build = guarded("cpu_build", lambda: cpu_build_baseline(sizes, args.dim, host_cores()))

    peak, peak_src = hbm_peak(peaks)
    abytes = algorithmic_bytes(ndist, nexp, layer_M, args.dim * 4, args.dim * 4, args.nq, k)
    kernel_ms = ms_dev / args.steps
    achieved = abytes / (kernel_ms * 1e-3) / 1e9
    qps = world * args.nq * args.steps / (ms_dev * 1e-3)
    e2e_qps = world * args.nq * args.steps / (ms_e2e * 1e-3)
    cfg = main_config(args)
    out = {
        "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": kernel_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "gpu_launches": args.steps, "config": cfg,
        "recall_at_10": recall,
        "parity": (main_cpu or {}).get("parity"),
        "e2e": {"value": e2e_qps, "unit": "queries/s",
                "h2d_bytes_per_step": int(args.nq * args.dim * 4),
                "d2h_bytes_per_step": int(args.nq * (k * 12 + 4)),
                "what": ("phnsw_search_batch_host_async per step from pinned host buffers (read and written "
                         "in place by the kernel), steps queued on one stream with batch overlap, one "
                         "sync after the last" if ms_e2e_sync else
                         "rank 0: pinned host queries -> H2D -> sharded step -> D2H of the merged top-k"),
                "synchronous_call_per_step": ({"value": world * args.nq * args.steps / (ms_e2e_sync * 1e-3),
                                               "what": "phnsw_search_batch, one blocking call per step"}
                                              if ms_e2e_sync else None)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "search_kernel<L2_SQRT, %s>" % ("tree" if tree else "sequential"),
                     "algorithmic_bytes_per_launch": abytes,
                     "n_dist_per_query": float(ndist.sum() / args.nq),
                     "n_exp_per_query": float(nexp.sum() / args.nq),
                     "traffic_note": "dram bytes per launch are in profiles/ (ncu --set full of "
                                     "this command); not measured in-run"},
        "cpu_baseline": (main_cpu or {}).get("cpu_baseline"),
        "clocks": clocks,
        "sum_order": args.sum_order,
        "parallelism": ("sharded x%d (one sub-index per GPU, one NCCL all-gather + merge per step, "
                        "steps pipelined)" % world) if world > 1 else "single GPU",
        "layers_top_first": gh_layers_top_first,
        "batch_overlap": {
            "on": bool(use_overlap),
            "what": "phnsw_index_set_batch_overlap: back-to-back launches of one stream chained as "
                    "programmatic dependent launches, so the ragged end of a launch is filled by the "
                    "next; identical results (asserted in-run)",
            "no_batch_overlap": ({"value": world * args.nq / (ms_plain * 1e-3), "unit": "queries/s",
                                  "ms_per_step": ms_plain,
                                  "roofline_frac": abytes / (ms_plain * 1e-3) / 1e9 / peak}
                                 if ms_plain else None)},
        "build": {"vectors_per_s": args.n / t_build, "seconds": t_build,
                  "first_build_seconds": build_times[0], "all_build_seconds": build_times,
                  "roofline": ({
                      "bound": "hbm",
                      "what": "algorithmic bytes of the build's traversal launches (SURVEY 8d formula: distance "
                              "evaluations x 4 dim + expansions x M x 4 + one query row per search), counted by "
                              "an instrumented, untimed build of the same index; the scoring / fold kernels "
                              "(7 % of the build's kernel time) are not counted",
                      "distance_evals": build_work["distance_evals"],
                      "searches": build_work["queries"], "search_launches": build_work["launches"],
                      "algorithmic_bytes": build_work["distance_evals"] * args.dim * 4
                                           + build_work["neighbor_list_bytes"] + build_work["queries"] * args.dim * 4,
                      "achieved": (build_work["distance_evals"] * args.dim * 4 + build_work["neighbor_list_bytes"]
                                   + build_work["queries"] * args.dim * 4) / t_build / 1e9,
                      "peak": peak, "unit": "GB/s",
                      "frac": (build_work["distance_evals"] * args.dim * 4 + build_work["neighbor_list_bytes"]
                               + build_work["queries"] * args.dim * 4) / t_build / 1e9 / peak,
                      "note": "over the WHOLE build wall time (host control flow, layer uploads and the "
                              "non-traversal kernels included), sequential summation order at 16 warps per SM"}
                      if build_work else None),
                  "seconds_is": "median of the builds after the first (steady state: construction "
                                "temporaries cached in the library's memory pool)",
                  "improve_index": not args.no_improve, "data_gen_seconds": t_gen,
                  "cpu_baseline_build": cpu_build},
        "sequential_order": {"value": world * args.nq / (ms_seq * 1e-3), "unit": "queries/s",
                             "ms_per_step": ms_seq,
                             "with_batch_overlap": ({"value": world * args.nq / (ms_seq_ov * 1e-3),
                                                     "ms_per_step": ms_seq_ov} if ms_seq_ov else None),
                             "note": "same kernel with PHNSW_SUM_SEQUENTIAL (the crate's loop bit for bit)"},
        "two_streams": {"value": world * args.nq / (ms_pipe * 1e-3), "unit": "queries/s",
                        "ms_per_step": ms_pipe,
                        "note": "steps issued alternately on two streams (tails overlap); "
                                "`value` above is the plain single-stream number"},
        "ground_truth": {
            "what": "exact brute-force kNN of the query batch (recall denominator)",
            "seconds": t_gt, "path": gt_stats["path"],
            "filter_kernel": "tc_filter_kernel (tcgen05 bf16 hi/lo split GEMM, M128 N128 K16)",
            "filter_ms": gt_stats["filter_ms"],
            "filter_tflops_issued": (gt_stats["filter_flops"] / gt_stats["filter_ms"] / 1e9
                                     if gt_stats["filter_ms"] > 0 else None),
            "algorithmic_tflops": (2.0 * args.nq * args.n * args.dim / gt_stats["filter_ms"] / 1e9
                                   if gt_stats["filter_ms"] > 0 else None),
            "tensor_peak_tflops": peaks.get("bf16_tflops"),
            "frac_of_tensor_peak_issued": (gt_stats["filter_flops"] / gt_stats["filter_ms"] / 1e9
                                           / peaks["bf16_tflops"]
                                           if gt_stats["filter_ms"] > 0 and peaks.get("bf16_tflops") else None),
            "frac_of_tensor_peak_algorithmic": (
                2.0 * args.nq * args.n * args.dim / gt_stats["filter_ms"] / 1e9 / peaks["bf16_tflops"]
                if gt_stats["filter_ms"] > 0 and peaks.get("bf16_tflops") else None),
            "split_products_issued (of 3)": (
                gt_stats["filter_flops"] / (2.0 * 128 * ((args.nq + 127) // 128) * 128
                                            * ((args.n + 127) // 128) * 64 * ((args.dim + 63) // 64))
                if gt_stats["filter_flops"] else None),
            "bound": "L2 -> SM: every 128-query block streams all rows (nq/128 x N x dim x 4 B per "
                     "launch, 96 % L2 hits), not the tensor pipe -- issuing one product instead of "
                     "three (operands exact in bf16: zero low parts are skipped) moves the time by 7 %",
            "l2_stream_tbs": (((args.nq + 127) // 128) * args.n * args.dim * 4.0
                              / gt_stats["filter_ms"] / 1e9 if gt_stats["filter_ms"] > 0 else None),
            "max_candidates_per_query": gt_stats["max_candidates"]},
    }
    if main_cpu and "error" in main_cpu:
        out["cpu_baseline_error"] = main_cpu
    if sharded_info:
        # rank 0's shard holds global ids [0, n): its own top-k must reappear in the merged lists
        # wherever the merge kept an id of that range, and merged lists ascend by (distance, id)
        mg, mgd = sharded_info["merged"], sharded_info["merged_d"]
        loc = oi.cpu().numpy()
        own = (mg >= 0) & (mg < args.n)
        ok_rows = [set(mg[i][own[i]].tolist()) <= set(loc[i].tolist()) for i in range(mg.shape[0])]
        asc = [all((mgd[i][j], mg[i][j]) <= (mgd[i][j + 1], mg[i][j + 1]) for j in range(k - 1)
                   if mg[i][j + 1] >= 0) for i in range(mg.shape[0])]
        ms_local = sharded_info["ms_local"] / args.steps
        out["recall_at_10"] = sharded_info["recall_merged"]
        out["gpu_launches"] = 2 * args.steps
        out["sharded_step"] = {
            "what": "N > 1: one sub-index of %d vectors per GPU (%d vectors in all), every rank searches "
                    "the same %d-query batch.  Headline = the PIPELINED step "
                    "(phnsw_search_batch_sharded_queued, one call per step + one phnsw_comm_flush inside "
                    "the timed region): K1 of step i + 1 (epilogue writes global-id records into a "
                    "rotating exchange buffer) overlaps the end of step i's, one ncclAllGather + merge "
                    "per step on the library's side stream.  `value` = (query, shard) searches per "
                    "second over all ranks = N x merged queries/s" % (args.n, args.n * world, args.nq),
            "pipelined": bool(sharded_info["queued"]),
            "merged_queries_per_s": args.nq / (kernel_ms * 1e-3),
            "vectors_total": args.n * world,
            "recall_at_10_merged_vs_exact_over_all_shards": sharded_info["recall_merged"],
            "recall_at_10_rank0_shard_alone": recall,
            "blocking_call_per_step": ({
                "value": world * args.nq / (sharded_info["ms_blocking"] * 1e-3),
                "ms_per_step": sharded_info["ms_blocking"],
                "what": "phnsw_search_batch_sharded: ncclBroadcast(queries) -> K1 -> ncclAllGather -> "
                        "merge on one stream, nothing of step i + 1 starts before step i is merged "
                        "(same results, asserted in-run)"} if sharded_info["ms_blocking"] else None),
            "without_exchange": {"value": world * args.nq / (sharded_info["ms_local_overlap"] * 1e-3),
                                 "ms_per_step": sharded_info["ms_local_overlap"],
                                 "plain_launches_ms_per_step": ms_local,
                                 "what": "the same shards searched with no all-gather / merge, back-to-back "
                                         "with batch overlap (N independent replicas of the per-GPU "
                                         "workload), max over ranks"},
            "exchange_overhead_frac": kernel_ms / sharded_info["ms_local_overlap"] - 1.0,
            "exchange_bytes_per_rank_per_step": int(args.nq * k * 12),
            "merge_parity": {"merged_entries_of_rank0_shard_match_frac": float(np.mean(ok_rows)),
                             "merged_ascending_frac": float(np.mean(asc))}}
    return out


def finish(ctx, out, blocks, sweep):
    """Secondary blocks under a watchdog, then the one JSON line.  A block that hangs (a rank
    that failed inside a collective) must not cost the headline line: after --block-timeout
    seconds every rank leaves and rank 0 prints what it has."""
    torch, dist, args, world, rank = ctx["torch"], ctx["dist"], ctx["args"], ctx["world"], ctx["rank"]
    done = threading.Event()
    lock = threading.Lock()

    def emit(extra_note=None):
        with lock:
            if done.is_set():
                return
            done.set()
            if rank == 0:
                if extra_note:
                    out["blocks_note"] = extra_note
                out["operating_points"] = {
                    "note": "one GPU, same index, ef = number_of_candidates = upper_layer_candidate_count",
                    "sweep": sweep,
                    "best_qps_at_recall_ge_0.95": max(
                        [p_["qps"] for p_ in sweep if p_["recall_at_10"] >= 0.95], default=None)}
                print(json.dumps(out), flush=True)

    def bail():
        emit("secondary blocks did not finish within %d s; skipped" % args.block_timeout)
        os._exit(0)

    wd = threading.Timer(args.block_timeout, bail)
    wd.daemon = True
    wd.start()
    if "config3" in blocks and world == 1:
        r = guarded("config3", lambda: run_config3(ctx))
        if rank == 0:
            out["config3"] = r
        torch.cuda.empty_cache()
    if "sharded" in blocks:
        r = guarded("config4", lambda: run_config4(ctx))
        if rank == 0:
            out["config4"] = r
        torch.cuda.empty_cache()
    if "config5" in blocks:
        r = guarded("config5", lambda: run_config5(ctx))
        if rank == 0:
            out["config5"] = r
        torch.cuda.empty_cache()
    wd.cancel()
    emit()
    if world > 1:
        try:
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass


def headline_parity(ctx, gh, rows_h, queries_h, oi, od, ndist, nexp, seq_ids, tree):
    """CPU baseline (oracle port on the same device-built graph, bounded sample) and parity of
    the timed run's outputs at full size."""
    from oracle import oracle as orc
    args, k = ctx["args"], ctx["k"]
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
    g_ds = od.cpu().numpy()[:cq].astype(np.float64)
    # the timed order against the crate's order: BASELINE.json's bar (>= 99.9 % / 1e-5)
    m = (g_ids == q_ids)
    ids_equal_seq = float(m.all(1).mean())
    rel_seq = np.abs(g_ds - q_ds) / np.maximum(np.abs(q_ds), 1e-30)
    max_rel_seq = float(rel_seq[m].max()) if m.any() else None
    seq_dev_equal = float((seq_ids[:cq] == q_ids).all(1).mean())
    if tree:  # and bit for bit against the oracle restating the same tree order
        oh.set_sum_order(1)
        o_ids, o_ds, _, o_nd, o_ne = oh.search(queries=queries_h.numpy()[:cq], sp=osp, max_out=k,
                                               stats=True, nthreads=cores)
        oh.set_sum_order(0)
    else:
        o_ids, o_ds, o_nd, o_ne = q_ids, q_ds, q_nd, q_ne
    ids_equal = float((g_ids == o_ids).all(1).mean())
    counters_equal = float(((ndist[:cq] == o_nd.astype(np.int64)).all(1)
                            & (nexp[:cq] == o_ne.astype(np.int64)).all(1)).mean())
    rel = np.abs(g_ds - o_ds) / np.maximum(np.abs(o_ds), 1e-30)
    max_rel = float(rel[(g_ids == o_ids)].max()) if (g_ids == o_ids).any() else None
    return {
        "parity": {"sample_queries": cq, "sum_order": args.sum_order,
                   "ids_equal_frac": ids_equal_seq, "max_rel_dist_err": max_rel_seq,
                   "oracle_same_order": {"ids_equal_frac": ids_equal,
                                         "work_counters_equal_frac": counters_equal,
                                         "max_rel_dist_err": max_rel},
                   "oracle_crate_order": {"ids_equal_frac": ids_equal_seq,
                                          "max_rel_dist_err": max_rel_seq},
                   "sequential_kernel_vs_oracle_ids_equal_frac": seq_dev_equal},
        "cpu_baseline": {"value": cq / cpu_dt, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": "%d of %d queries on the same device-built graph" % (cq, args.nq)}}


# ------------------------------------------------------------------------------ config 3
def run_config3(ctx):
    """BASELINE configs[2]: n x 1536 f32 embedding-shaped, cosine, PQ8 (cs 16 -> 96 codes, K 256):
    k-means codebook training + encoding (tcgen05 assignment), ADC search + exact re-rank as one
    library call (phnsw_pq8_search_batch_device), against the full-precision search of the same
    graph."""
    torch, ph, N, dev, k = ctx["torch"], ctx["ph"], ctx["N"], ctx["dev"], ctx["k"]
    args, stream, peaks = ctx["args"], ctx["stream"], ctx["peaks"]
    n, dim, cs, K, nq = args.c3_n, 1536, 16, 256, args.nq
    rerank_k = 100
    mix = mixture_for("embedding", dev)
    t0 = time.perf_counter()
    rows = mix.rows(n, 2024)
    q = mix.rows(nq, 2025)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    comp = ph.BigComparator(rows, ph.COS_HALF, device=dev.index)
    del rows
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    full = ph.Hnsw.generate(comp, seed=1)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    cb = ph.pq8_train(comp, K, cs, kmeans_iters=5, seed=3)
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t0
    train_stats = ph.assign_last_stats()
    t0 = time.perf_counter()
    pq = ph.Pq8Comparator(comp, cb, cs)
    torch.cuda.synchronize()
    t_encode = time.perf_counter() - t0
    enc_stats = ph.assign_last_stats()
    table = ph.ADC_TABLE_Q8 if args.adc_table == "q8" else ph.ADC_TABLE_F32
    pq.set_adc_table(table)
    gh = full.rebind(pq)
    full.set_sum_order(ph.SUM_TREE)
    L = gh.layer_count()
    layer_M = [gh.get_layer_from_top(i)[2] for i in range(L)]
    t0 = time.perf_counter()
    gt, _ = comp.bruteforce_knn(q, k)
    torch.cuda.synchronize()
    t_gt = time.perf_counter() - t0
    gt = gt.cpu().numpy()
    sp = ph.SearchParameters(args.ef, args.ef, 2)
    oi = torch.empty((nq, k), dtype=torch.int64, device=dev)
    od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    oc = torch.empty((nq,), dtype=torch.int32, device=dev)
    tm = Timer(torch)
    steps = max(3, ctx["args"].steps // 2)

    # full-precision search of the same graph (the comparison point)
    ms_full = tm.run(lambda: full.search_device(q, sp, oi, od, oc, stream=stream), steps, 3,
                     lambda: full.sync(stream))
    rec_full = recall_at_k(oi.cpu().numpy(), gt, k)
    # ADC walk alone (ADC distances out), with work counters for the roofline
    ai = torch.empty((nq, rerank_k), dtype=torch.int64, device=dev)
    ad = torch.empty((nq, rerank_k), dtype=torch.float32, device=dev)
    nd = torch.zeros((nq, L), dtype=torch.int32, device=dev)
    ne = torch.zeros((nq, L), dtype=torch.int32, device=dev)
    gh.search_device(q, sp, ai, ad, oc, stream=stream, out_ndist=nd, out_nexp=ne)
    gh.sync(stream)
    rec_adc_raw = recall_at_k(ai.cpu().numpy(), gt, k)
    adc_ids_first = ai.cpu().numpy().astype(np.uint64)
    adc_ds_first = ad.cpu().numpy().copy()
    ndist, nexp = nd.cpu().numpy().astype(np.int64), ne.cpu().numpy().astype(np.int64)
    ms_walk = tm.run(lambda: gh.search_device(q, sp, ai, ad, oc, stream=stream), steps, 3,
                     lambda: gh.sync(stream))
    # the product call: ADC walk + exact re-rank of `rerank_k` hits, one call, one stream
    ms_adc = tm.run(lambda: gh.adc_search_device(q, sp, oi, od, oc, rerank=comp, rerank_k=rerank_k,
                                                 stream=stream), steps, 3, lambda: gh.sync(stream))
    rr_ids = oi.cpu().numpy()
    rr_ds = od.cpu().numpy().copy()
    rec_adc = recall_at_k(rr_ids, gt, k)
    # end to end from pinned host buffers through the host ABI call
    q_pin = q.cpu().pin_memory()
    hi = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    hd = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    hc = torch.empty((nq,), dtype=torch.int32).pin_memory()

    def e2e():
        N.check(N.lib().phnsw_pq8_search_batch(
            gh._h, comp._h, C.c_void_p(q_pin.data_ptr()), nq, C.byref(sp), rerank_k, k,
            C.c_void_p(hi.data_ptr()), C.c_void_p(hd.data_ptr()), C.c_void_p(hc.data_ptr())))
    for _ in range(3):
        e2e()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / steps
    assert np.array_equal(hi.numpy(), rr_ids), "config3: host path and device path disagree"
    # the other table form, same call (secondary)
    other = ph.ADC_TABLE_F32 if table == ph.ADC_TABLE_Q8 else ph.ADC_TABLE_Q8
    pq.set_adc_table(other)
    oi2, od2 = torch.empty_like(oi), torch.empty_like(od)
    ms_other = tm.run(lambda: gh.adc_search_device(q, sp, oi2, od2, oc, rerank=comp, rerank_k=rerank_k,
                                                   stream=stream), steps, 3, lambda: gh.sync(stream))
    rec_other = recall_at_k(oi2.cpu().numpy(), gt, k)
    pq.set_adc_table(table)

    # parity sample against the CPU oracle: the ADC walk (oracle's ADC definition on the same
    # graph, codes and codebook) bit for bit, and the re-ranked distances against the crate's
    # sequential f32 comparator
    parity = guarded("config3.parity", lambda: config3_parity(
        ctx, gh, pq, cb, cs, dim, n, q, sp, adc_ids_first, adc_ds_first, ndist, nexp, rr_ids, rr_ds,
        comp, rerank_k))
    peak, peak_src = hbm_peak(peaks)
    Q = dim // cs
    abytes = algorithmic_bytes(ndist, nexp, layer_M, Q, dim * 4, nq, rerank_k)
    abytes_full = algorithmic_bytes(ndist, nexp, layer_M, dim * 4, dim * 4, nq, k)
    res = {
        "workload": "%d x 1536 f32 embedding-shaped synthetic (2048 clusters on a 24-d manifold, "
                    "unit norm), cosine, PQ8: centroid_size 16 -> 96 u8 codes per vector, K = 256; "
                    "search ef=%d, exact re-rank of %d ADC hits" % (n, args.ef, rerank_k),
        "queries_per_step": nq,
        "value": nq / (ms_adc * 1e-3), "unit": "queries/s", "ms_per_step": ms_adc,
        "what": "phnsw_pq8_search_batch_device: ADC walk + exact re-rank inside the timed region "
                "(one kernel: the warp that finished a query's walk re-ranks it)",
        "adc_table": args.adc_table + (" (per-query tables quantised to u8 by a pre-pass kernel, "
                                       "24 KB per query in flight, integer sums)" if table == ph.ADC_TABLE_Q8
                                       else " (exact f32 entries)"),
        "recall_at_10": rec_adc,
        "other_table_form": {"adc_table": "f32" if table == ph.ADC_TABLE_Q8 else "q8",
                             "value": nq / (ms_other * 1e-3), "ms_per_step": ms_other,
                             "recall_at_10": rec_other},
        "adc_walk_only": {"value": nq / (ms_walk * 1e-3), "ms_per_step": ms_walk,
                          "recall_at_10_before_rerank": rec_adc_raw},
        "full_precision": {"value": nq / (ms_full * 1e-3), "ms_per_step": ms_full,
                           "recall_at_10": rec_full, "sum_order": "tree",
                           "hbm_frac": abytes_full / (ms_full * 1e-3) / 1e9 / peak},
        "adc_over_full_precision": ms_full / ms_adc,
        "e2e": {"value": nq / (ms_e2e * 1e-3), "unit": "queries/s",
                "h2d_bytes_per_step": int(nq * dim * 4), "d2h_bytes_per_step": int(nq * (k * 12 + 4))},
        "roofline": {"bound": "hbm", "kernel": "search_kernel<COS_HALF, ADC> (walk only)",
                     "achieved": abytes / (ms_walk * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": abytes / (ms_walk * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": abytes,
                     "n_dist_per_query": float(ndist.sum() / nq),
                     "n_exp_per_query": float(nexp.sum() / nq),
                     "lut_flops_per_query": 2.0 * dim * K,
                     "walk_includes": "the table pre-pass kernel (adc_lut_q8_kernel) when adc_table = q8",
                     "note": "96 B of codes per distance: the walk is bound by dependent latencies "
                             "and table arithmetic, not by HBM"},
        "build": {"graph_seconds": t_build, "graph_vectors_per_s": n / t_build,
                  "kmeans_train_seconds": t_train, "kmeans_iters": 5,
                  "encode_seconds": t_encode, "assignment_path": enc_stats["path"],
                  "assign_kernel_ms": enc_stats["kernel_ms"],
                  "assign_rows": enc_stats["rows"], "assign_rechecked": enc_stats["rechecked"],
                  "assign_tflops": (enc_stats["flops"] / enc_stats["kernel_ms"] / 1e9
                                    if enc_stats["kernel_ms"] > 0 else None),
                  "assign_gbs": (enc_stats["rows"] * cs * 4 / enc_stats["kernel_ms"] / 1e6
                                 if enc_stats["kernel_ms"] > 0 else None),
                  "train_assign_path": train_stats["path"],
                  "data_gen_seconds": t_gen, "ground_truth_seconds": t_gt},
        "layers_top_first": gh.layer_sizes(),
        "parity": parity,
    }
    gh.close()
    full.close()
    pq.close()
    comp.close()
    return res


def config3_parity(ctx, gh, pq, cb, cs, dim, n, q, sp, adc_ids, adc_ds, ndist, nexp, rr_ids, rr_ds,
                   comp, rerank_k):
    from oracle import oracle as orc
    args = ctx["args"]
    sq = min(200, q.shape[0])
    codes = pq.codes()
    oh = orc.Hnsw.from_layers_codes(orc.COS_HALF, dim, n, gh.layers(), codes, cb, cs)
    orc.attach_pq8(oh, codes, cb, cs, table=pq.adc_table())
    qh = q[:sq].cpu().numpy()
    o = oh.search(queries=qh, sp=orc.search_params(args.ef, args.ef, 2), max_out=rerank_k,
                  stats=True, nthreads=host_cores())
    ids_eq = float((adc_ids[:sq] == o[0]).all(1).mean())
    bits_eq = float((adc_ds[:sq].view(np.uint32) == o[1].view(np.uint32)).all(1).mean())
    ctr_eq = float(((ndist[:sq] == o[3].astype(np.int64)).all(1)
                    & (nexp[:sq] == o[4].astype(np.int64)).all(1)).mean())
    # the re-rank: exact comparator (sequential f32, (1 - sum a*b) / 2) over the oracle's own hits
    uniq = np.unique(o[0][o[0] != np.uint64(0xFFFFFFFFFFFFFFFF)])
    rows_hit = comp.lookup(uniq)
    pos = {int(v): i for i, v in enumerate(uniq)}
    ok_ids = ok_bits = 0
    k = rr_ids.shape[1]
    for i in range(sq):
        hits = [int(v) for v in o[0][i] if v != np.uint64(0xFFFFFFFFFFFFFFFF)]
        ds = np.array([orc.distance(orc.COS_HALF, rows_hit[pos[v]], qh[i]) for v in hits], np.float32)
        order = sorted(range(len(hits)), key=lambda j: (ds[j], hits[j]))[:k]
        want_ids = np.array([hits[j] for j in order], np.int64)
        want_ds = np.array([ds[j] for j in order], np.float32)
        ok_ids += int(np.array_equal(rr_ids[i][:len(order)], want_ids))
        ok_bits += int(np.array_equal(rr_ds[i][:len(order)].view(np.uint32), want_ds.view(np.uint32)))
    return {"sample_queries": sq,
            "adc_walk_vs_oracle": {"ids_equal_frac": ids_eq, "distance_bits_equal_frac": bits_eq,
                                   "work_counters_equal_frac": ctr_eq},
            "rerank_vs_oracle_comparator": {"ids_equal_frac": ok_ids / sq,
                                            "distance_bits_equal_frac": ok_bits / sq},
            "note": "ADC / k-means have no crate analogue (its k-means is dead code, its search "
                    "symmetric): the oracle's definitions are ours -- parity unpinned -- the "
                    "re-rank is the crate's comparator (pq.rs:354-363)"}


# ------------------------------------------------------------------------------ config 4 / 5
def sharded_ground_truth(ctx, comp, dq, id_offset):
    """Exact top-k over ALL shards: per-shard brute force (tensor-core filter + exact re-rank),
    all-gather with torch.distributed, merged by the library's K5 -- deliberately not the
    exchange path under test."""
    torch, dist, ph, k, world = ctx["torch"], ctx["dist"], ctx["ph"], ctx["k"], ctx["world"]
    nq, dev = dq.shape[0], ctx["dev"]
    gt_i, gt_d = comp.bruteforce_knn(dq, k)
    if world == 1:
        return (gt_i + id_offset).cpu().numpy()
    ggi = torch.empty((world, nq, k), dtype=torch.int64, device=dev)
    ggd = torch.empty((world, nq, k), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(ggi, (gt_i + id_offset).contiguous())
    dist.all_gather_into_tensor(ggd, gt_d.contiguous())
    ei = torch.empty((nq, k), dtype=torch.int64, device=dev)
    ed = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ph.merge_topk_device(ggi, ggd, world, nq, k, ei, ed, ctx["stream"])
    torch.cuda.synchronize()
    return ei.cpu().numpy()


def timed_sharded(ctx, sh, gh, dq, sp, rerank_k, steps, warmup):
    """Device-timed sharded steps, barrier + synchronize on both sides, max over ranks."""
    torch, dist, world, dev, stream, k = (ctx["torch"], ctx["dist"], ctx["world"], ctx["dev"],
                                          ctx["stream"], ctx["k"])
    nq = dq.shape[0]
    out = (torch.empty((nq, k), dtype=torch.int64, device=dev),
           torch.empty((nq, k), dtype=torch.float32, device=dev))
    for _ in range(warmup):
        sh.search(dq, sp, k, src=0, stream=stream, rerank_k=rerank_k, out=out)
    gh.sync(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sh.search(dq, sp, k, src=0, stream=stream, rerank_k=rerank_k, out=out)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    gh.sync(stream)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) / steps, out


def timed_sharded_queued(ctx, sh, gh, dq, sp, steps, warmup):
    """Device-timed PIPELINED sharded steps (phnsw_search_batch_sharded_queued, batch overlap on,
    one phnsw_comm_flush before the closing event), barrier + synchronize on both sides, max over
    ranks.  Every rank must already hold `dq`."""
    torch, dist, world, dev, stream, k = (ctx["torch"], ctx["dist"], ctx["world"], ctx["dev"],
                                          ctx["stream"], ctx["k"])
    nq = dq.shape[0]
    ring = [(torch.empty((nq, k), dtype=torch.int64, device=dev),
             torch.empty((nq, k), dtype=torch.float32, device=dev)) for _ in range(4)]
    gh.set_batch_overlap(True)
    try:
        for i in range(warmup):
            sh.search_queued(dq, sp, k, stream=stream, out=ring[i & 3])
        sh.flush(stream)
        gh.sync(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            sh.search_queued(dq, sp, k, stream=stream, out=ring[i & 3])
        sh.flush(stream)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        gh.sync(stream)
    finally:
        gh.set_batch_overlap(False)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) / steps, ring[(steps - 1) & 3]


def timed_e2e_sharded(ctx, sh, gh, q_host_pinned, dq, sp, rerank_k, steps, warmup):
    """The sharded step from HOST buffers: rank 0 uploads the batch, the library broadcasts,
    searches, exchanges and merges, rank 0 reads the merged result back."""
    torch, dist, world, dev, stream, k, rank = (ctx["torch"], ctx["dist"], ctx["world"], ctx["dev"],
                                                ctx["stream"], ctx["k"], ctx["rank"])
    nq = dq.shape[0]
    out = (torch.empty((nq, k), dtype=torch.int64, device=dev),
           torch.empty((nq, k), dtype=torch.float32, device=dev))
    hi = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    hd = torch.empty((nq, k), dtype=torch.float32).pin_memory()

    def step():
        if rank == 0:
            dq.copy_(q_host_pinned, non_blocking=True)
        sh.search(dq, sp, k, src=0, stream=stream, rerank_k=rerank_k, out=out)
        if rank == 0:
            hi.copy_(out[0], non_blocking=True)
            hd.copy_(out[1], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) / steps * 1e3


def run_config4(ctx):
    """BASELINE configs[3]: 10M x 96 f32 Deep-shaped, L2, split over the N ranks: every rank
    builds and searches its own sub-index, the library call phnsw_search_batch_sharded does
    broadcast -> K1 (epilogue writes global-id records into the exchange buffer) -> one
    ncclAllGather -> merge, all on one stream."""
    torch, dist, ph, dev, k = ctx["torch"], ctx["dist"], ctx["ph"], ctx["dev"], ctx["k"]
    args, stream, rank, world, peaks = ctx["args"], ctx["stream"], ctx["rank"], ctx["world"], ctx["peaks"]
    from parallel_hnsw_b200.sharded import ShardedHnsw
    n_shard, dim, nq = args.c4_n // world, 96, args.nq
    mix = mixture_for("deep", dev)
    t0 = time.perf_counter()
    rows = mix.rows(n_shard, 9600 + rank)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    comp = ph.BigComparator(rows, ph.L2_SQRT, device=dev.index)
    rows_h = rows[:0]
    if rank == 0:
        rows_h = rows.cpu()  # the parity sample runs the oracle on rank 0's shard
    del rows
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    gh = ph.Hnsw.generate(comp, seed=1 + rank)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    gh.set_sum_order(ph.SUM_TREE if args.sum_order == "tree" else ph.SUM_SEQUENTIAL)
    sp = ph.SearchParameters(args.ef, args.ef, 2)
    q_host = mix.rows(nq, 9599).cpu().pin_memory()   # the same batch on every rank's generator
    dq = q_host.to(dev) if rank == 0 else torch.zeros((nq, dim), dtype=torch.float32, device=dev)
    sh = ShardedHnsw(gh, rank * n_shard, rank, world)
    steps = args.steps
    ms_block, out = timed_sharded(ctx, sh, gh, dq, sp, 0, steps, args.warmup)
    merged_ids = out[0].cpu().numpy()
    merged_ds = out[1].cpu().numpy()
    # the pipelined step (every rank holds the batch now: the blocking call broadcast it)
    ms, out_q = timed_sharded_queued(ctx, sh, gh, dq, sp, steps, args.warmup)
    assert torch.equal(out_q[0], out[0]) and torch.equal(out_q[1], out[1]), "pipelined step changed the results"
    # the same shard searched alone (no broadcast, no exchange): what the exchange costs
    L = gh.layer_count()
    oi = torch.empty((nq, k), dtype=torch.int64, device=dev)
    od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    oc = torch.empty((nq,), dtype=torch.int32, device=dev)
    nd = torch.zeros((nq, L), dtype=torch.int32, device=dev)
    ne = torch.zeros((nq, L), dtype=torch.int32, device=dev)
    tm = Timer(torch)
    ms_local = tm.run(lambda: gh.search_device(dq, sp, oi, od, oc, stream=stream), steps, 3,
                      lambda: gh.sync(stream))
    gh.search_device(dq, sp, oi, od, oc, stream=stream, out_ndist=nd, out_nexp=ne)
    gh.sync(stream)
    t = torch.tensor([ms_local, t_build], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_local_max, t_build_max = float(t[0]), float(t[1])
    ms_e2e = timed_e2e_sharded(ctx, sh, gh, q_host, dq, sp, 0, max(3, steps // 2), 3)
    gt = sharded_ground_truth(ctx, comp, dq, rank * n_shard)
    rec = recall_at_k(merged_ids, gt, k)
    res = None
    if rank == 0:
        ndist, nexp = nd.cpu().numpy().astype(np.int64), ne.cpu().numpy().astype(np.int64)
        layer_M = [gh.get_layer_from_top(i)[2] for i in range(L)]
        peak, peak_src = hbm_peak(peaks)
        abytes = algorithmic_bytes(ndist, nexp, layer_M, dim * 4, dim * 4, nq, k)
        parity = guarded("sharded.parity", lambda: shard_parity_f32(
            ctx, gh, rows_h.numpy(), q_host.numpy(), oi.cpu().numpy(), od.cpu().numpy(), ndist, nexp,
            merged_ids, merged_ds, rank * n_shard))
        res = {
            "workload": "%d x 96 f32 Deep-shaped synthetic (1024 clusters on a 16-d manifold, unit "
                        "norm), L2, split over %d GPUs (%d vectors per sub-index), search ef=%d" % (
                            n_shard * world, world, n_shard, args.ef),
            "mode": "sharded sub-indexes; pipelined step (phnsw_search_batch_sharded_queued): K1 of "
                    "step i + 1 (global-id records written by the kernel epilogue) overlaps the end of "
                    "step i's, one ncclAllGather + merge per step on the library's side stream, one "
                    "flush inside the timed region",
            "shards": world, "vectors_total": n_shard * world, "queries_per_step": nq,
            "value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms,
            "blocking_call_per_step": {
                "value": nq / (ms_block * 1e-3), "ms_per_step": ms_block,
                "what": "phnsw_search_batch_sharded: ncclBroadcast(queries) -> K1 -> ncclAllGather -> "
                        "merge on one stream (same results, asserted in-run)"},
            "recall_at_10": rec,
            "single_shard": {"value": nq / (ms_local_max * 1e-3), "ms_per_step": ms_local_max,
                             "what": "the slowest rank's own shard searched without broadcast / "
                                     "exchange / merge"},
            "exchange_overhead_frac": ms_block / ms_local_max - 1.0,
            "exchange_overhead_what": "blocking step against the slowest shard's plain launches (the "
                                      "pipelined step also hides the ragged end of each launch)",
            "weak_scaling_efficiency": ms_local_max / ms,
            "e2e": {"value": nq / (ms_e2e * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(nq * dim * 4), "d2h_bytes_per_step": int(nq * k * 12),
                    "what": "rank 0: pinned host queries -> H2D -> sharded step -> D2H of the merged top-k"},
            "exchange_bytes_per_rank_per_step": int(N_slice_bytes(ctx, nq, k)),
            "roofline": {"bound": "hbm", "kernel": "search_kernel<L2_SQRT, tree> on rank 0's shard",
                         "achieved": abytes / (ms_local * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": abytes / (ms_local * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": abytes,
                         "n_dist_per_query": float(ndist.sum() / nq),
                         "n_exp_per_query": float(nexp.sum() / nq)},
            "build": {"seconds_slowest_rank": t_build_max, "vectors_per_s_per_gpu": n_shard / t_build_max,
                      "vectors_per_s_total": n_shard * world / t_build_max, "data_gen_seconds": t_gen},
            "layers_top_first": gh.layer_sizes(),
            "parity": parity,
        }
    sh.comm.close()
    gh.close()
    comp.close()
    return res


def N_slice_bytes(ctx, nq, k):
    return ctx["N"].lib().phnsw_comm_slice_bytes(nq, k)


def shard_parity_f32(ctx, gh, rows_h, q_h, loc_ids, loc_ds, ndist, nexp, merged_ids, merged_ds,
                     id_offset):
    """Rank 0's shard against the oracle on that shard's graph (per-shard parity) and
    the merged list against what rank 0's shard contributes."""
    from oracle import oracle as orc
    args, k = ctx["args"], ctx["k"]
    sq = min(500, q_h.shape[0])
    oh = orc.Hnsw.from_layers(orc.L2_SQRT, rows_h, gh.layers())
    osp = orc.search_params(args.ef, args.ef, 2)
    cores = host_cores()
    tree = args.sum_order == "tree"
    oh.set_sum_order(1 if tree else 0)
    o = oh.search(queries=q_h[:sq], sp=osp, max_out=k, stats=True, nthreads=cores)
    ids_eq = float((loc_ids[:sq].astype(np.uint64) == o[0]).all(1).mean())
    bits_eq = float((loc_ds[:sq].view(np.uint32) == o[1].view(np.uint32)).all(1).mean())
    ctr_eq = float(((ndist[:sq] == o[3].astype(np.int64)).all(1)
                    & (nexp[:sq] == o[4].astype(np.int64)).all(1)).mean())
    crate = {}
    if tree:
        oh.set_sum_order(0)
        c = oh.search(queries=q_h[:sq], sp=osp, max_out=k, nthreads=cores)
        m = loc_ids[:sq].astype(np.uint64) == c[0]
        rel = np.abs(loc_ds[:sq].astype(np.float64) - c[1]) / np.maximum(np.abs(c[1]), 1e-30)
        crate = {"ids_equal_frac": float(m.all(1).mean()),
                 "max_rel_dist_err": float(rel[m].max()) if m.any() else None}
    # every merged entry that carries one of rank 0's ids must be rank 0's (id, distance) pair,
    # and the merged list must be ascending by (distance, id)
    lo, hi = id_offset, id_offset + rows_h.shape[0]
    ok = asc = 0
    for i in range(sq):
        mine = {int(a) + id_offset: b for a, b in zip(o[0][i], o[1][i]) if a != np.uint64(0xFFFFFFFFFFFFFFFF)}
        got = [(float(d), int(v)) for v, d in zip(merged_ids[i], merged_ds[i]) if v >= 0]
        asc += int(got == sorted(got))
        ok += int(all((not (lo <= v < hi)) or (v in mine and np.float32(mine[v]) == np.float32(d))
                      for d, v in got))
    return {"sample_queries": sq, "shard": 0,
            "oracle_same_order": {"ids_equal_frac": ids_eq, "distance_bits_equal_frac": bits_eq,
                                  "work_counters_equal_frac": ctr_eq},
            "oracle_crate_order": crate,
            "merged_entries_of_shard0_match_frac": ok / sq, "merged_ascending_frac": asc / sq}


def run_config5(ctx):
    """BASELINE configs[4]: world x 12.5M x 128 (100M over 8 GPUs), generated shard by shard in
    HBM, PQ8-coded (centroid_size 8 -> 16 u8 codes per vector, K = 256).  Per rank: f32 rows are
    kept (6.4 GB of 180 GB) for the graph build and for the exact re-rank; the search walks the
    16 B/vector codes with per-query tables in shared memory (ADC), re-ranks its hits against
    the f32 rows, and the sharded exchange merges the per-shard top-10.  One batch of 10 000
    queries per step."""
    torch, dist, ph, dev, k = ctx["torch"], ctx["dist"], ctx["ph"], ctx["dev"], ctx["k"]
    args, stream, rank, world, peaks = ctx["args"], ctx["stream"], ctx["rank"], ctx["world"], ctx["peaks"]
    from parallel_hnsw_b200.sharded import ShardedHnsw
    n_shard, dim, cs, K, nq, rerank_k = args.c5_n, 128, 8, 256, args.nq, args.c5_rerank
    mix = mixture_for("sift", dev)
    t0 = time.perf_counter()
    rows = mix.rows(n_shard, 100 + rank)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    comp = ph.BigComparator(rows, ph.L2_SQRT, device=dev.index)
    del rows
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    full = ph.Hnsw.generate(comp, seed=1 + rank)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    cb = ph.pq8_train(comp, K, cs, kmeans_iters=5, seed=3)
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t0
    t0 = time.perf_counter()
    pq = ph.Pq8Comparator(comp, cb, cs)
    torch.cuda.synchronize()
    t_encode = time.perf_counter() - t0
    enc_stats = ph.assign_last_stats()
    pq.set_adc_table(ph.ADC_TABLE_Q8 if args.adc_table == "q8" else ph.ADC_TABLE_F32)
    gh = full.rebind(pq)
    full.close()
    sp = ph.SearchParameters(args.ef, args.ef, 2)
    q_host = mix.rows(nq, 99).cpu().pin_memory()
    dq = q_host.to(dev) if rank == 0 else torch.zeros((nq, dim), dtype=torch.float32, device=dev)
    sh = ShardedHnsw(gh, rank * n_shard, rank, world, rerank=comp)
    steps = args.steps
    ms, out = timed_sharded(ctx, sh, gh, dq, sp, rerank_k, steps, args.warmup)
    merged_ids = out[0].cpu().numpy()
    L = gh.layer_count()
    oi = torch.empty((nq, k), dtype=torch.int64, device=dev)
    od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    oc = torch.empty((nq,), dtype=torch.int32, device=dev)
    ai = torch.empty((nq, rerank_k), dtype=torch.int64, device=dev)
    ad = torch.empty((nq, rerank_k), dtype=torch.float32, device=dev)
    nd = torch.zeros((nq, L), dtype=torch.int32, device=dev)
    ne = torch.zeros((nq, L), dtype=torch.int32, device=dev)
    tm = Timer(torch)
    ms_local = tm.run(lambda: gh.adc_search_device(dq, sp, oi, od, oc, rerank=comp, rerank_k=rerank_k,
                                                   stream=stream), steps, 3, lambda: gh.sync(stream))
    ms_walk = tm.run(lambda: gh.search_device(dq, sp, ai, ad, oc, stream=stream), max(3, steps // 2), 2,
                     lambda: gh.sync(stream))
    gh.search_device(dq, sp, ai, ad, oc, stream=stream, out_ndist=nd, out_nexp=ne)
    gh.sync(stream)
    t = torch.tensor([ms_local, t_build, t_train + t_encode], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_local_max, t_build_max, t_pq_max = float(t[0]), float(t[1]), float(t[2])
    ms_e2e = timed_e2e_sharded(ctx, sh, gh, q_host, dq, sp, rerank_k, max(3, steps // 2), 3)
    gt = sharded_ground_truth(ctx, comp, dq, rank * n_shard)
    rec = recall_at_k(merged_ids, gt, k)
    res = None
    if rank == 0:
        ndist, nexp = nd.cpu().numpy().astype(np.int64), ne.cpu().numpy().astype(np.int64)
        layer_M = [gh.get_layer_from_top(i)[2] for i in range(L)]
        peak, peak_src = hbm_peak(peaks)
        Q = dim // cs
        abytes = algorithmic_bytes(ndist, nexp, layer_M, Q, dim * 4, nq, rerank_k)
        parity = guarded("config5.parity", lambda: shard_parity_adc(
            ctx, gh, pq, cb, cs, dim, n_shard, q_host.numpy(), ai.cpu().numpy(), ad.cpu().numpy(),
            ndist, nexp))
        res = {
            "workload": "%d x 128 SIFT-shaped synthetic generated in HBM (%d vectors per GPU x %d "
                        "GPUs), L2, PQ8-coded: centroid_size 8 -> 16 u8 codes per vector, K = 256; "
                        "ADC walk ef=%d + exact re-rank of %d hits against the resident f32 rows, "
                        "one 10 000-query batch per step" % (n_shard * world, n_shard, world, args.ef, rerank_k),
            "mode": "sharded sub-indexes; one library call per step: ncclBroadcast(queries) -> ADC "
                    "walk -> exact re-rank (global-id records) -> one ncclAllGather -> merge",
            "rows_kept": "f32 rows stay resident per shard (%.1f GB) for the build and the re-rank; "
                         "codes %.2f GB, graph (u32) %.1f GB" % (
                             n_shard * dim * 4 / 1e9, n_shard * 16 / 1e9, n_shard * 48 * 4 / 1e9),
            "shards": world, "vectors_total": n_shard * world, "queries_per_step": nq,
            "adc_table": args.adc_table,
            "value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms,
            "recall_at_10": rec,
            "single_shard": {"value": nq / (ms_local_max * 1e-3), "ms_per_step": ms_local_max,
                             "what": "the slowest rank's ADC walk + re-rank without the exchange"},
            "exchange_overhead_frac": ms / ms_local_max - 1.0,
            "weak_scaling_efficiency": ms_local_max / ms,
            "e2e": {"value": nq / (ms_e2e * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(nq * dim * 4), "d2h_bytes_per_step": int(nq * k * 12)},
            "roofline": {"bound": "hbm", "kernel": "search_kernel<L2_SQRT, ADC> (walk only) on rank 0's shard",
                         "achieved": abytes / (ms_walk * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": abytes / (ms_walk * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": abytes, "walk_ms": ms_walk,
                         "n_dist_per_query": float(ndist.sum() / nq),
                         "n_exp_per_query": float(nexp.sum() / nq)},
            "build": {"graph_seconds_slowest_rank": t_build_max,
                      "graph_vectors_per_s_per_gpu": n_shard / t_build_max,
                      "graph_vectors_per_s_total": n_shard * world / t_build_max,
                      "kmeans_train_plus_encode_seconds": t_pq_max,
                      "assignment_path": enc_stats["path"], "assign_kernel_ms": enc_stats["kernel_ms"],
                      "data_gen_seconds": t_gen},
            "layers_top_first": gh.layer_sizes(),
            "parity": parity,
        }
    sh.comm.close()
    gh.close()
    pq.close()
    comp.close()
    return res


def shard_parity_adc(ctx, gh, pq, cb, cs, dim, n, q_h, adc_ids, adc_ds, ndist, nexp):
    from oracle import oracle as orc
    args = ctx["args"]
    sq = min(200, q_h.shape[0])
    codes = pq.codes()
    oh = orc.Hnsw.from_layers_codes(orc.L2_SQRT, dim, n, gh.layers(), codes, cb, cs)
    orc.attach_pq8(oh, codes, cb, cs, table=pq.adc_table())
    o = oh.search(queries=q_h[:sq], sp=orc.search_params(args.ef, args.ef, 2), max_out=adc_ids.shape[1],
                  stats=True, nthreads=host_cores())
    return {"sample_queries": sq, "shard": 0,
            "adc_walk_vs_oracle": {
                "ids_equal_frac": float((adc_ids[:sq].astype(np.uint64) == o[0]).all(1).mean()),
                "distance_bits_equal_frac": float((adc_ds[:sq].view(np.uint32) == o[1].view(np.uint32)).all(1).mean()),
                "work_counters_equal_frac": float(((ndist[:sq] == o[3].astype(np.int64)).all(1)
                                                   & (nexp[:sq] == o[4].astype(np.int64)).all(1)).mean())},
            "note": "ADC / k-means definitions are the oracle's own (no crate analogue): parity unpinned"}


if __name__ == "__main__":
    main()
