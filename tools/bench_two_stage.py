"""Config-5 style latency probe (NOT the bench.py metric): FDE candidate generation + MaxSim rerank of the top-1000 vs the
exhaustive MaxSim scan, single 32-token queries, on one GPU's resident shard.  Prints one JSON line.

  python tools/bench_two_stage.py [--pages 65536] [--cands 1000] [--queries 50]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from morphik_core_b200.fde import TwoStageIndex  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pages", type=int, default=65536)
ap.add_argument("--cands", type=int, default=1000)
ap.add_argument("--queries", type=int, default=50)
ap.add_argument("--k", type=int, default=10)
args = ap.parse_args()

dev = torch.device("cuda", 0)
nq = args.queries
q_host = bench.make_queries(nq)
packed, planted = bench.build_shard(args.pages, dev, 1234, q_host, planted_per_query=10)
two = TwoStageIndex(dtype="bf16")
t0 = time.perf_counter()
two.adopt_packed(packed, [bench.P_PATCH] * args.pages)
torch.cuda.synchronize()
t_build = time.perf_counter() - t0
queries = [q_host[i * 32:(i + 1) * 32].numpy() for i in range(nq)]

lat2, lat1, stage = [], [], []
recall_hits = recall_tot = agree = 0
for i, q in enumerate(queries):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s2, i2, c2 = two.search([q], args.k, n_candidates=args.cands)
    lat2.append((time.perf_counter() - t0) * 1e3)
    stage.append(dict(two.last_timing_ms))
    t0 = time.perf_counter()
    s1, i1, c1 = two.index.search_host([q], args.k)
    lat1.append((time.perf_counter() - t0) * 1e3)
    agree += int(i1[0, 0] == i2[0, 0])
    recall_hits += len(set(i1[0].tolist()) & set(i2[0].tolist()))
    recall_tot += args.k
warm = 5
lat2, lat1, stage = lat2[warm:], lat1[warm:], stage[warm:]
line = {
    "workload": f"{args.pages} pages x 1024 patches bf16 ({args.pages * 262144 / 1e9:.1f} GB) + FDE matrix {args.pages * 20480 / 1e9:.2f} GB, "
                f"single 32-token queries, top-{args.k}",
    "two_stage_p50_ms": float(np.median(lat2)), "two_stage_p95_ms": float(np.percentile(lat2, 95)),
    "fde_candidates_p50_ms": float(np.median([s["fde_candidates_ms"] for s in stage])),
    "maxsim_rerank_p50_ms": float(np.median([s["maxsim_rerank_ms"] for s in stage])),
    "exhaustive_p50_ms": float(np.median(lat1)), "n_candidates": args.cands,
    "top1_agreement": agree / nq, "recall_at_k_vs_exhaustive": recall_hits / recall_tot,
    "fde_build_seconds": t_build, "fde_build_pages_per_s": args.pages / t_build,
}
print(json.dumps(line), flush=True)
