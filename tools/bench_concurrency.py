"""Serving probe: N concurrent query_similar coroutines (different doc_ids filters) against one B200MultiVectorStore,
coalesced (default) vs one GPU pass per request.  Not a bench.py leg; prints one JSON line."""
import argparse
import asyncio
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from morphik_core_b200.catalog import PageRecord  # noqa: E402
from morphik_core_b200.store import B200MultiVectorStore  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pages", type=int, default=65536)
ap.add_argument("--clients", type=int, default=16)
ap.add_argument("--rounds", type=int, default=20)
args = ap.parse_args()

dev = torch.device("cuda", 0)
q_host = bench.make_queries(args.clients)
packed, _ = bench.build_shard(args.pages, dev, 1234, q_host[: 32 * min(args.clients, 32)])
out = {"workload": f"{args.pages} pages x 1024 x 128 bf16, {args.clients} concurrent query_similar calls (32 tokens, k=10, "
                   "each with its own doc_ids filter), rounds of asyncio.gather"}
for label, coalesce in (("coalesced", True), ("one_pass_per_request", False)):
    store = B200MultiVectorStore(mode="bf16", coalesce_queries=coalesce)
    store._index.adopt_packed(packed, [bench.P_PATCH] * args.pages)
    for p in range(args.pages):  # 64 pages per document
        store.catalog.add(PageRecord(f"doc{p // 64}", p % 64, "", {}, None, bench.P_PATCH))
    n_docs = args.pages // 64
    reqs = [dict(query_embedding=q_host[32 * i:32 * (i + 1)].numpy(), k=10,
                 doc_ids=[f"doc{d}" for d in range(0, n_docs, 1 + i % 3)]) for i in range(args.clients)]

    async def round_():
        return await asyncio.gather(*[store.query_similar(**r) for r in reqs])

    asyncio.run(round_())
    t0 = time.perf_counter()
    for _ in range(args.rounds):
        res = asyncio.run(round_())
    dt = time.perf_counter() - t0
    out[label] = {"queries_per_s": args.clients * args.rounds / dt, "round_ms": 1e3 * dt / args.rounds,
                  "last_batch": store.last_coalesced_batch, "top1": res[0][0].document_id if res[0] else None}
    store._index = None
out["speedup"] = out["coalesced"]["queries_per_s"] / out["one_pass_per_request"]["queries_per_s"]
print(json.dumps(out))
