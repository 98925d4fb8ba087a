"""Serving probe: N concurrent query_similar coroutines (different doc_ids filters) against the plugin, coalesced (default)
vs one GPU pass per request.  Not a bench.py leg; prints one JSON line.

  python tools/bench_concurrency.py [--pages 65536] [--clients 16]                                   # one GPU, B200MultiVectorStore
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_concurrency.py --sharded [--clients 32]
                                                                                                     # ShardedB200MultiVectorStore

In --sharded mode every rank adopts a synthetic shard directly (the ingest path is measured by tools/bench_ingest.py), rank 0
mirrors the catalogues and drives the API; ranks > 0 run worker_loop().  Reports queries/s and p50 / p95 of a round of
`--clients` simultaneous callers with distinct doc_ids filters.
"""
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

import faulthandler  # noqa: E402

faulthandler.dump_traceback_later(int(os.environ.get("B200MS_WATCHDOG_S", "400")), exit=True)  # never hang a GPU lease

ap = argparse.ArgumentParser()
ap.add_argument("--pages", type=int, default=65536, help="pages per GPU")
ap.add_argument("--clients", type=int, default=16)
ap.add_argument("--rounds", type=int, default=20)
ap.add_argument("--sharded", action="store_true")
ap.add_argument("--fde-candidates", type=int, default=0)
args = ap.parse_args()
DOC_PAGES = 64


def requests_for(q_host, n_docs_total, doc_name):
    return [dict(query_embedding=q_host[32 * i:32 * (i + 1)].numpy(), k=10,
                 doc_ids=[doc_name(d) for d in range(0, n_docs_total, 1 + i % 3)]) for i in range(args.clients)]


def measure(store, reqs):
    async def round_():
        t0 = time.perf_counter()
        res = await asyncio.gather(*[store.query_similar(**r) for r in reqs])
        return res, time.perf_counter() - t0

    asyncio.run(round_())
    lat = []
    t0 = time.perf_counter()
    for _ in range(args.rounds):
        res, dt = asyncio.run(round_())
        lat.append(dt * 1e3)
    total = time.perf_counter() - t0
    lat = np.sort(np.asarray(lat))
    return {"queries_per_s": args.clients * args.rounds / total, "round_p50_ms": float(lat[len(lat) // 2]),
            "round_p95_ms": float(lat[int(len(lat) * 0.95)]), "last_batch": store.last_coalesced_batch,
            "top1": res[0][0].document_id if res[0] else None, "timing": dict(getattr(store, "last_query_timing", {}))}


if not args.sharded:
    from morphik_core_b200.store import B200MultiVectorStore

    dev = torch.device("cuda", 0)
    q_host = bench.make_queries(args.clients)
    packed, _ = bench.build_shard(args.pages, dev, 1234, q_host[: 32 * min(args.clients, 32)])
    out = {"workload": f"{args.pages} pages x 1024 x 128 bf16, {args.clients} concurrent query_similar calls (32 tokens, k=10, "
                       "each with its own doc_ids filter), rounds of asyncio.gather"}
    for label, coalesce in (("coalesced", True), ("one_pass_per_request", False)):
        store = B200MultiVectorStore(mode="bf16", coalesce_queries=coalesce)
        store._index.adopt_packed(packed, [bench.P_PATCH] * args.pages)
        for p in range(args.pages):
            store.catalog.add(PageRecord(f"doc{p // DOC_PAGES}", p % DOC_PAGES, "", {}, None, bench.P_PATCH))
        out[label] = measure(store, requests_for(q_host, args.pages // DOC_PAGES, lambda d: f"doc{d}"))
        store._index = None
    out["speedup"] = out["coalesced"]["queries_per_s"] / out["one_pass_per_request"]["queries_per_s"]
    print(json.dumps(out))
else:
    import torch.distributed as dist

    from morphik_core_b200.sharded_store import ShardedB200MultiVectorStore

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG")
    dist.init_process_group("nccl", device_id=dev)
    q_host = bench.make_queries(args.clients)
    packed, _ = bench.build_shard(args.pages, dev, 1234 + rank, q_host[: 32 * min(args.clients, 32)])
    out = {"workload": f"{world} GPUs x {args.pages} pages x 1024 x 128 bf16 through ShardedB200MultiVectorStore, {args.clients} concurrent "
                       "query_similar calls (32 tokens, k=10, distinct doc_ids filters), rounds of asyncio.gather"
                       + (f", two-stage with {args.fde_candidates} FDE candidates per rank" if args.fde_candidates else "")}
    n_docs = args.pages // DOC_PAGES
    for label, coalesce in (("coalesced", True), ("one_pass_per_request", False)):
        if args.fde_candidates and label == "coalesced":
            continue  # the two-stage mode takes one filter per pass
        store = ShardedB200MultiVectorStore(mode="bf16", device=local, coalesce_queries=coalesce,
                                            fde_candidates=args.fde_candidates or None)
        store.index.adopt_packed(packed, [bench.P_PATCH] * args.pages)
        if store._two_stage is not None:
            store._two_stage.rebuild_from_index()
        for r in (range(world) if rank == 0 else [rank]):
            for p in range(args.pages):
                store.catalogs[r].add(PageRecord(f"r{r}d{p // DOC_PAGES}", p % DOC_PAGES, "", {}, None, bench.P_PATCH))
        for r in range(world):
            for d in range(n_docs):
                store.doc_rank[f"r{r}d{d}"] = r
            store.rank_rows[r] = args.pages * bench.P_PATCH
        if rank != 0:
            store.worker_loop()
        else:
            out[label] = measure(store, requests_for(q_host, n_docs * world, lambda d: f"r{d % world}d{d // world}"))
            store.close()
        store.index = None
        dist.barrier()
    if rank == 0:
        if "coalesced" in out:
            out["speedup"] = out["coalesced"]["queries_per_s"] / out["one_pass_per_request"]["queries_per_s"]
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
