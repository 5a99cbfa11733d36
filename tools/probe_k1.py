"""Developer probe (not part of the product): build the bench index on the device once and
time the traversal kernel.  PHNSW_LIB selects a library variant.
usage: python tools/probe_k1.py [--n N] [--nq NQ ...] [--tag T]"""
import argparse
import hashlib
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph  # noqa: E402
from bench import sift_like  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--nq", type=int, nargs="+", default=[10000])
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--ef", type=int, default=300)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--tag", default="")
ap.add_argument("--no-improve", action="store_true")
ap.add_argument("--order", type=int, nargs="+", default=[0], help="0 sequential, 1 tree")
ap.add_argument("--streams", type=int, default=1)
ap.add_argument("--overlap", action="store_true", help="phnsw_index_set_batch_overlap")
ap.add_argument("--record-events", action="store_true", help="record an event after every launch")
ap.add_argument("--profile", action="store_true", help="cudaProfilerStart/Stop around one launch "
                "(ncu --profile-from-start off)")
args = ap.parse_args()
rows = sift_like(args.n, args.dim, 1234)
comp = ph.BigComparator(rows.numpy(), ph.L2_SQRT)
t = time.time()
gh = ph.Hnsw.generate(comp, seed=1, improve=not args.no_improve)
torch.cuda.synchronize()
tb = time.time() - t
dev = torch.device("cuda:0")
if args.overlap:
    gh.set_batch_overlap(True)
k = 10
sp = ph.SearchParameters(args.ef, args.ef, 2)
st = torch.cuda.current_stream().cuda_stream
for order, nq in [(o, n) for o in args.order for n in args.nq]:
    gh.set_sum_order(order)
    dq = sift_like(nq, args.dim, 4321).to(dev)
    oi = torch.empty((nq, k), dtype=torch.int64, device=dev)
    od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    oc = torch.empty((nq,), dtype=torch.int32, device=dev)
    for _ in range(3):
        gh.search_device(dq, sp, oi, od, oc, stream=st)
    gh.sync(st)
    if args.profile:
        torch.cuda.profiler.start()
        gh.search_device(dq, sp, oi, od, oc, stream=st)
        gh.sync(st)
        torch.cuda.profiler.stop()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.streams == 1:
        evs = [torch.cuda.Event() for _ in range(args.reps)]
        e0.record()
        for r_ in range(args.reps):
            gh.search_device(dq, sp, oi, od, oc, stream=st)
            if args.record_events:
                evs[r_].record()
        e1.record()
        gh.sync(st)
    else:
        ss = [torch.cuda.Stream() for _ in range(args.streams)]
        outs = [(torch.empty_like(oi), torch.empty_like(od), torch.empty_like(oc)) for _ in ss]
        for i, s_ in enumerate(ss):  # warm the per-stream workspaces
            gh.search_device(dq, sp, *outs[i], stream=s_.cuda_stream)
        torch.cuda.synchronize()
        e0.record()
        for s_ in ss:
            s_.wait_event(e0)
        for r in range(args.reps):
            i = r % len(ss)
            gh.search_device(dq, sp, *outs[i], stream=ss[i].cuda_stream)
        for s_ in ss:
            torch.cuda.current_stream().wait_stream(s_)
        e1.record()
        torch.cuda.synchronize()
        oi, od = outs[0][0], outs[0][1]
    ms = e0.elapsed_time(e1) / args.reps
    h = hashlib.sha1(oi.cpu().numpy().tobytes() + od.cpu().numpy().tobytes()).hexdigest()[:12]
    print("PROBE %s order=%d nq=%d: %.3f ms, %.0f QPS, build %.2fs, out sha %s" % (
        args.tag, order, nq, ms, nq / ms * 1e3, tb, h), flush=True)
