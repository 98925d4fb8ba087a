"""Probe for the "next" rows of SURVEY 8(f) that sit either side of the scorer (not a bench.py leg): ingest through
store_embeddings (host float32 pages, as the reference hands them over, vs CUDA bf16 tensors = the f-3 fast path), shard
file save/load (f-2) and the embedding-based reranker on 75 candidates (f-4: the reference's candidate count).
Prints one JSON line."""
import argparse
import asyncio
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from morphik_core_b200.models import DocumentChunk  # noqa: E402
from morphik_core_b200.reranker import B200MaxSimReranker  # noqa: E402
from morphik_core_b200.store import B200MultiVectorStore  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pages", type=int, default=4096)
ap.add_argument("--batch", type=int, default=256, help="pages per store_embeddings call")
args = ap.parse_args()
P, D = 1024, 128
g = torch.Generator(device="cuda").manual_seed(3)
dev_pages = [torch.nn.functional.normalize(torch.randn((P, D), generator=g, device="cuda"), dim=1).bfloat16() for _ in range(args.pages)]
host_pages = [p.float().cpu().numpy() for p in dev_pages[: min(args.pages, 1024)]]
out = {"workload": f"{args.pages} pages x {P} x {D}, store_embeddings in calls of {args.batch} pages"}


def ingest(store, pages):
    t0 = time.perf_counter()
    for b0 in range(0, len(pages), args.batch):
        chunks = [DocumentChunk(document_id=f"d{(b0 + i) // 8}", content="", embedding=p, chunk_number=(b0 + i) % 8)
                  for i, p in enumerate(pages[b0:b0 + args.batch])]
        asyncio.run(store.store_embeddings(chunks))
    torch.cuda.synchronize()
    return time.perf_counter() - t0


for mode in ("bf16", "binary"):
    s = B200MultiVectorStore(mode=mode)
    ingest(s, dev_pages[:args.batch])  # warm-up (allocations)
    s.close()
    s = B200MultiVectorStore(mode=mode)
    dt = ingest(s, dev_pages)
    out[f"ingest_cuda_bf16_tensors_{mode}"] = {"pages_per_s": len(dev_pages) / dt, "source_GB_per_s": len(dev_pages) * P * D * 2 / dt / 1e9}
    s.close()
    s = B200MultiVectorStore(mode=mode)
    dt = ingest(s, host_pages)
    out[f"ingest_host_float32_{mode}"] = {"pages_per_s": len(host_pages) / dt, "source_GB_per_s": len(host_pages) * P * D * 4 / dt / 1e9}
    if mode == "bf16":
        with tempfile.TemporaryDirectory() as d:
            t0 = time.perf_counter()
            s.save(d)
            t_save = time.perf_counter() - t0
            size = os.path.getsize(os.path.join(d, "corpus.b2ms"))
            t0 = time.perf_counter()
            s2 = B200MultiVectorStore.load(d)
            torch.cuda.synchronize()
            t_load = time.perf_counter() - t0
            out["shard_file"] = {"bytes": size, "save_GB_per_s": size / t_save / 1e9, "load_GB_per_s": size / t_load / 1e9,
                                 "pages": len(s2.catalog), "note": "tmpfs/overlay file system of the box"}
            s2.close()
    s.close()

rr = B200MaxSimReranker(mode="bf16")
q = host_pages[0][:32]
cands = [DocumentChunk(document_id=f"c{i}", content="", embedding=host_pages[i], chunk_number=0) for i in range(75)]
cands_dev = [DocumentChunk(document_id=f"c{i}", content="", embedding=dev_pages[i], chunk_number=0) for i in range(75)]
for label, cs in (("host_float32_candidates", cands), ("cuda_bf16_candidates", cands_dev)):
    asyncio.run(rr.rerank(q, cs))
    lat = []
    for _ in range(20):
        t0 = time.perf_counter()
        res = asyncio.run(rr.rerank(q, cs))
        lat.append(time.perf_counter() - t0)
    out[f"rerank_75_{label}"] = {"p50_ms": 1e3 * sorted(lat)[len(lat) // 2], "top1": res[0].document_id}
print(json.dumps(out))
