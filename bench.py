#!/usr/bin/env python
"""bench.py -- patch-vectors/sec of the ColPali MaxSim hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (one JSON line on rank 0)
  python bench.py --impl reference [...]                          # the reference's own CPU formulation, same metric

Workload (config.workload): a RESIDENT SHARD of BASELINE configs[1] -- "1M pages x 1024 patches x 128-d bf16, batch-32
queries" needs 262 GB, more than one B200's HBM -- i.e. `--pages` pages (default 262144 = 68.7 GB of bf16
patch vectors per GPU, ~545x the 126 MB L2, so no flush is needed between iterations) scored against a batch of 32
queries x 32 tokens, top-10 per query.  A "step" = one pass of the hot path over the whole shard for the whole batch:
pack queries -> MaxSim scan (tcgen05) -> top-k.  Synthetic, seeded, unit-norm rows with planted relevant pages.

value   whole-job patch-vectors/s with the query batch already resident in HBM (search_device), CUDA events, max over ranks
e2e     same metric through the C-ABI call a plugin user makes (b200ms_search_host): pinned HOST query buffer -> H2D ->
        scan -> top-k -> D2H of the results, host clock around K synchronous calls, max over ranks (N>1 adds the
        NCCL all-gather + merge)
roofline  dominant kernel = maxsim_umma; configs[1] has 1024 resident query tokens => tensor-bound;
          achieved = 2*rows*128*1024 flop per step / CUDA-event time of the scoring launches inside the timed region
hbm_regime  the same shard with ONE 32-token query (the HBM-bound regime the north star's 70 % target is about):
            achieved GB/s = rows*256 B / event time, against the measured copy bandwidth
cpu_baseline  colpali_engine's score_multi_vector formulation (oracle/maxsim_oracle.py: torch einsum on all host cores)
              on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

P_PATCH, DIM, T_TOK = 1024, 128, 32
METRIC = "patch_vectors_per_sec_maxsim"
UNIT = "patch-vectors/s"


def load_traffic():
    """dram__bytes_read+write per patch vector of the dominant kernel, from the committed ncu --set full capture
    (profiles/<round>/traffic.json, written by tools/ncu_summary.py); None if no capture is committed."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    if os.path.isdir(pdir):
        for rnd in sorted(os.listdir(pdir)):
            path = os.path.join(pdir, rnd, "traffic.json")
            if os.path.exists(path):
                with open(path) as f:
                    best = json.load(f)
    return best


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "tflops_burst": float(p["bf16_tflops"]),
                "tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ synthetic workload
def make_queries(n_q: int, seed: int = 4321):
    import torch

    g = torch.Generator().manual_seed(seed)
    q = torch.randn((n_q * T_TOK, DIM), generator=g, dtype=torch.float32)
    return torch.nn.functional.normalize(q, dim=1).contiguous()  # host, [n_q*32, 128]


def build_shard(n_pages: int, device, seed: int, q_host, planted_per_query: int = 10):
    """[n_pages*1024, 128] bf16 unit-norm rows in a 1024-aligned buffer; a few pages per query get 64 noisy copies of the
    query's tokens so that the top-k is meaningful (SURVEY 8d 'Synthetic inputs')."""
    import torch

    rows_total = n_pages * P_PATCH
    buf = torch.empty(rows_total * DIM * 2 + 1024, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % 1024
    packed = buf[off:off + rows_total * DIM * 2]
    rows = packed.view(torch.bfloat16).view(rows_total, DIM)
    g = torch.Generator(device=device).manual_seed(seed)
    chunk = 1 << 21
    for r0 in range(0, rows_total, chunk):
        n = min(chunk, rows_total - r0)
        x = torch.randn((n, DIM), generator=g, device=device, dtype=torch.float32)
        rows[r0:r0 + n] = torch.nn.functional.normalize(x, dim=1).to(torch.bfloat16)
        del x
    n_q = q_host.shape[0] // T_TOK
    qd = q_host.to(device)
    planted = {}
    pg = torch.Generator().manual_seed(seed + 99)
    for qi in range(n_q):
        ids = torch.randint(0, n_pages, (planted_per_query,), generator=pg).tolist()
        planted[qi] = ids
        for j, p in enumerate(ids):
            noise = torch.randn((2 * T_TOK, DIM), generator=g, device=device) * (0.3 / DIM ** 0.5) * (1 + j)
            toks = qd[qi * T_TOK:(qi + 1) * T_TOK].repeat(2, 1) + noise
            rows[p * P_PATCH:p * P_PATCH + 2 * T_TOK] = torch.nn.functional.normalize(toks, dim=1).to(torch.bfloat16)
    return packed, planted


# ------------------------------------------------------------------------------------------------ reference arm / CPU baseline
def set_cpu_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use the host's cores (physical cores by default)."""
    import torch

    want = int(os.environ.get("B200MS_CPU_THREADS", "0")) or max(1, (os.cpu_count() or 2) // 2)
    if torch.get_num_threads() != want:
        torch.set_num_threads(want)


def cpu_reference_rate(n_q: int, budget_s: float, seed: int = 1234):
    """patch-vectors/s of the reference's float scorer (score_multi_vector: pad + einsum + max + sum, batch 128) on the host
    CPU with all torch threads, on a bounded sample of the same workload (same generator family / shapes)."""
    import torch

    from oracle import maxsim_oracle as orc

    set_cpu_threads()
    q = make_queries(n_q).view(n_q, T_TOK, DIM).numpy()
    g = torch.Generator().manual_seed(seed)

    def sample(n_pages):
        x = torch.randn((n_pages, P_PATCH, DIM), generator=g, dtype=torch.float32)
        return torch.nn.functional.normalize(x, dim=2).bfloat16().float().numpy()  # bf16-valued like the GPU shard

    probe = sample(32)
    t0 = time.perf_counter()
    orc.score_multi_vector_port_dense(q, probe)
    t_probe = time.perf_counter() - t0
    t0 = time.perf_counter()
    orc.score_multi_vector_port_dense(q, probe)
    t_probe = min(t_probe, time.perf_counter() - t0)
    n_pages = int(max(32, min(4096, 32 * budget_s / max(t_probe, 1e-4))))
    n_pages = max(128, n_pages // 128 * 128) if n_pages >= 128 else n_pages
    pages = sample(n_pages)
    t0 = time.perf_counter()
    scores = orc.score_multi_vector_port_dense(q, pages)
    dt = time.perf_counter() - t0
    return {"value": n_pages * P_PATCH / dt, "seconds": dt, "pages": n_pages, "threads": torch.get_num_threads(),
            "checksum": float(scores.sum())}


def config0_latency(packed, q_pin, out_pin, k, skip_cpu, n_pages=100, iters=300):
    """BASELINE configs[0] (N=100 pages x 1024 x 128, B_q=1, T=32): per-call latency of b200ms_search_host on the first
    100 pages of the shard (host query in, host top-k out), and the reference's CPU formulation on the same shapes."""
    import numpy as np
    import torch

    from morphik_core_b200.index import MaxSimIndex

    idx0 = MaxSimIndex(device=packed.device.index or 0, dtype="bf16")
    idx0.adopt_packed(packed[: n_pages * P_PATCH * DIM * 2], [P_PATCH] * n_pages)
    q1 = q_pin[:T_TOK]
    o = (out_pin[0][:1], out_pin[1][:1], out_pin[2][:1])
    for _ in range(20):
        idx0.search_host_flat(q1, [T_TOK], k, *o)
    lat = []
    for _ in range(iters):
        t0 = time.perf_counter()
        idx0.search_host_flat(q1, [T_TOK], k, *o)
        lat.append(time.perf_counter() - t0)
    lat = np.sort(np.asarray(lat)) * 1e6
    kern = idx0.score_times_ms(64)
    out = {"workload": f"configs[0]: {n_pages} pages x {P_PATCH} patches x {DIM}-d bf16, 1 query x {T_TOK} tokens, top-{k}",
           "api": "b200ms_search_host (host query in, host top-k out)", "e2e_p50_us": float(lat[len(lat) // 2]),
           "e2e_p95_us": float(lat[int(len(lat) * 0.95)]), "scoring_kernel_us": 1e3 * sum(kern) / len(kern), "calls": iters}
    if not skip_cpu:
        from oracle import maxsim_oracle as orc

        set_cpu_threads()
        qn = q1.view(1, T_TOK, DIM).numpy()
        pages = packed[: n_pages * P_PATCH * DIM * 2].view(torch.bfloat16).view(n_pages, P_PATCH, DIM).float().cpu().numpy()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            orc.score_multi_vector_port_dense(qn, pages)
            ts.append(time.perf_counter() - t0)
        out["cpu_reference_formulation_ms"] = 1e3 * sorted(ts)[len(ts) // 2]
        out["cpu_threads"] = torch.get_num_threads()
    return out


def topk_match_vs_oracle(packed, q_pin, n_q, k, n_pages=512, n_queries=4):
    """Top-k lists of b200ms_search_host on the first n_pages of the shard vs the CPU oracle (fp32 MaxSim on the same
    bf16-valued inputs): identical id lists and the largest relative score difference.  Checker only (oracle/)."""
    import numpy as np
    import torch

    from morphik_core_b200.index import MaxSimIndex
    from oracle import maxsim_oracle as orc

    nq = min(n_q, n_queries)
    sub = MaxSimIndex(device=packed.device.index or 0, dtype="bf16")
    sub.adopt_packed(packed[: n_pages * P_PATCH * DIM * 2], [P_PATCH] * n_pages)
    queries = [q_pin[i * T_TOK:(i + 1) * T_TOK].numpy() for i in range(nq)]
    ts, ti, tc = sub.search_host(queries, k)
    rows = packed[: n_pages * P_PATCH * DIM * 2].view(torch.bfloat16).view(-1, DIM).float().cpu().numpy()
    off = orc.page_offsets([P_PATCH] * n_pages)
    same, err = 0, 0.0
    for i, q in enumerate(queries):
        want = orc.float_maxsim_c(orc.bf16_round_np(q), rows, off)
        ws, wi = orc.topk_np(want, k)
        same += int(ti[i].tolist() == wi.tolist())
        err = max(err, float(np.max(np.abs(ts[i] - ws) / np.maximum(np.abs(ws), 1e-6))))
    return {"sample": f"first {n_pages} pages of the shard, {nq} queries, top-{k}", "id_lists_identical": f"{same}/{nq}",
            "max_rel_score_err": err, "tolerance": 1e-3}


def torch_gpu_reference_rate(rows_bf16, q_dev, n_q, sample_pages=4096, iters=5):
    """The kernel-to-beat on the same GPU: the reference's OWN GPU formulation -- colpali_engine
    score_multi_vector with device="cuda" (fast_multivector_store.py:339,553-555) = einsum("bnd,csd->bcns").max(3).sum(2)
    over 128-page batches -- on a resident sample of the shard (no H2D), in bf16 and in fp32 (TF32 allowed, as the reference
    enables it, colpali_embedding_model.py:32-35).  A reported baseline only; nothing here is on the product path."""
    import torch

    out = {}
    pages = rows_bf16[: sample_pages * P_PATCH].view(sample_pages, P_PATCH, DIM)
    q = q_dev.view(n_q, T_TOK, DIM)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for name, dt in (("bf16", torch.bfloat16), ("fp32_tf32", torch.float32)):
            qq = q.to(dt)

            def once():
                acc = []
                for j in range(0, sample_pages, 128):
                    pb = pages[j:j + 128].to(dt)
                    acc.append(torch.einsum("bnd,csd->bcns", qq, pb).max(dim=3)[0].sum(dim=2))
                return torch.cat(acc, dim=1)

            once()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                scores = once()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            out[name] = {"value": sample_pages * P_PATCH / (ms * 1e-3), "unit": UNIT, "ms": ms,
                         "checksum": float(scores.float().sum())}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    out["sample"] = f"{sample_pages} resident pages x {P_PATCH} patches, {n_q} queries x {T_TOK} tokens, torch.einsum+max+sum, batch 128"
    return out


def run_reference(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_q = args.bq
    vals = []
    for _ in range(args.warmup):
        cpu_reference_rate(n_q, budget_s=min(args.ref_budget, 2.0))
    for _ in range(args.steps):
        vals.append(cpu_reference_rate(n_q, budget_s=args.ref_budget))
    total_pv = sum(v["pages"] * P_PATCH for v in vals)
    total_s = sum(v["seconds"] for v in vals)
    value = total_pv / total_s
    sample = f"{vals[-1]['pages']} pages x {P_PATCH} patches x {DIM}-d (bf16-valued fp32) per step, {n_q} queries x {T_TOK} tokens"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, per_gpu_pages=vals[-1]["pages"], note="CPU arm scores a bounded sample per step"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": vals[-1]["threads"], "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, per_gpu_pages, note=None):
    cfg = {
        "workload": f"configs[1] resident shard: {per_gpu_pages} pages/GPU x {P_PATCH} patches x {DIM}-d bf16, "
                    f"batch-{args.bq} queries x {T_TOK} tokens, top-{args.k} (full config = 1M pages = 262 GB > one B200's HBM)",
        "pages_per_gpu": per_gpu_pages, "patches_per_page": P_PATCH, "dim": DIM, "query_batch": args.bq,
        "query_tokens": T_TOK, "k": args.k, "parallelism": f"document-sharded x{args.gpus}",
        "l2": "inputs larger than L2 (shard >> 126 MB), no flush between iterations",
    }
    if note:
        cfg["note"] = note
    return cfg


# ------------------------------------------------------------------------------------------------ the GPU arm
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from morphik_core_b200.index import MaxSimIndex
    from morphik_core_b200.sharded import ShardedMaxSim

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torch.distributed.run)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the ONE JSON line: with NCCL_DEBUG=VERSION or WARN (set by some images) NCCL prints
        # "NCCL version ..." on stdout at communicator creation; unset it unless the user asked for INFO/TRACE
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        dist.init_process_group("nccl", device_id=device)
    peaks = load_peaks()
    n_pages, n_q, k = args.pages, args.bq, args.k
    rows = n_pages * P_PATCH

    q_host = make_queries(n_q)  # identical on every rank
    packed, planted = build_shard(n_pages, device, seed=1234 + rank, q_host=q_host)
    idx = MaxSimIndex(device=local_rank, dtype="bf16")
    idx.adopt_packed(packed, [P_PATCH] * n_pages)
    id_base = rank * n_pages
    sharded = ShardedMaxSim.from_index(idx, id_base=id_base) if world > 1 else None
    q_lens = [T_TOK] * n_q
    q_dev = q_host.to(device)
    q_pin = q_host.pin_memory()
    out_dev = (torch.empty((n_q, k), dtype=torch.float32, device=device), torch.empty((n_q, k), dtype=torch.int64, device=device),
               torch.empty((n_q,), dtype=torch.int32, device=device))
    out_pin = (torch.empty((n_q, k), dtype=torch.float32).pin_memory(), torch.empty((n_q, k), dtype=torch.int64).pin_memory(),
               torch.empty((n_q,), dtype=torch.int32).pin_memory())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def step_device():
        if world > 1:
            return sharded.search(q_dev, q_lens, k)
        return idx.search_device(q_dev, q_lens, k, out=out_dev)

    def step_e2e():
        if world > 1:
            qd = q_pin.to(device, non_blocking=True)
            ts, ti, tc = sharded.search(qd, q_lens, k)
            return ts.cpu(), ti.cpu(), tc.cpu()
        idx.search_host_flat(q_pin, q_lens, k, *out_pin)
        return out_pin

    # ---- correctness guard before timing: planted pages must come out on top, and the oracle agrees on a sample
    ts, ti, tc = step_device()
    torch.cuda.synchronize(device)
    top1 = ti[:, 0].cpu().tolist()
    if world == 1:
        hits = sum(1 for qi in range(n_q) if top1[qi] - id_base in planted[qi])
        assert hits >= n_q - 1, f"planted pages not retrieved ({hits}/{n_q}): the timed path is not computing MaxSim"

    # ---- value: device-resident inputs, CUDA events, max over ranks
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = idx.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    t_host1 = time.perf_counter()
    clocks = sampler.stop(t_host0, t_host1)
    ms_total = ev0.elapsed_time(ev1)
    launches = idx.launch_count() - launches0
    score_ms = idx.score_times_ms(min(args.steps, 256))  # scoring kernels only, events recorded inside the timed region
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * rows / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the public C-ABI call, host<->device copies inside the timed region
    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * rows * args.steps / float(t.item())
    h2d = n_q * T_TOK * DIM * 4
    d2h = n_q * k * 12 + n_q * 4

    # ---- roofline of the dominant kernel (tensor-bound at 1024 resident query tokens)
    score_ms_avg = sum(score_ms) / len(score_ms)
    flops_per_step = 2.0 * rows * DIM * n_q * T_TOK
    achieved_tf = flops_per_step / (score_ms_avg * 1e-3) / 1e12
    n_mtiles = (n_q * T_TOK + 127) // 128
    # default: CTA-pair kernel, 8 query tiles per pass; B200MS_PAIR_CTA=0 selects the one-CTA W4 kernel (4 tiles per pass)
    pair = os.environ.get("B200MS_PAIR_CTA", "1") != "0" and n_mtiles >= 3
    passes = (n_mtiles + 7) // 8 if pair else (n_mtiles + 3) // 4
    kname, tkey = (("maxsim_umma_pair_kernel<bf16,NM=4>", "maxsim_umma_pair<bf16,NM=4>") if pair else
                   ("maxsim_umma_w4_kernel<bf16,NM=4>", "maxsim_umma<bf16,NM=4>"))
    tr = load_traffic()
    traffic_bytes = traffic_src = None
    if tr and tkey in tr:
        traffic_bytes = tr[tkey]["dram_bytes_per_patch_vector"] * rows  # per launch (one pass)
        traffic_src = tr[tkey]["source"]
    roofline = {
        "kernel": kname, "bound": "tensor", "achieved": achieved_tf,
        "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved_tf / peaks["tflops_sustained"],
        "peak_kind": f"{peaks['source']} cuBLAS bf16 sustained (burst {peaks['tflops_burst']})",
        "frac_of_burst": achieved_tf / peaks["tflops_burst"], "traffic": traffic_bytes, "traffic_source": traffic_src,
        "launches_per_step": passes, "avg_launch_ms": score_ms_avg / passes, "score_ms_per_step": score_ms_avg,
        "algorithmic_flops_per_launch": flops_per_step / passes,
        "hbm_gbs_in_this_regime": passes * rows * DIM * 2 / (score_ms_avg * 1e-3) / 1e9,
    }

    # ---- the HBM-bound regime: one 32-token query over the same shard
    hbm = None
    if rank == 0 or world > 1:
        q1 = q_dev[:T_TOK].contiguous()
        o1 = (out_dev[0][:1], out_dev[1][:1], out_dev[2][:1])
        for _ in range(args.warmup):
            idx.search_device(q1, [T_TOK], k, out=o1)
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            idx.search_device(q1, [T_TOK], k, out=o1)
        e1.record()
        torch.cuda.synchronize(device)
        sm1 = idx.score_times_ms(min(args.steps, 256))
        sm1_avg = sum(sm1) / len(sm1)
        gbs = rows * DIM * 2 / (sm1_avg * 1e-3) / 1e9
        hbm = {"workload": f"same shard, ONE query x {T_TOK} tokens (B_q*T = 32: HBM-bound regime)", "bound": "hbm",
               "kernel": "maxsim_umma_kernel<bf16,NM=1>", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
               "frac": gbs / peaks["hbm_gbs"], "peak_kind": f"{peaks['source']} copy bandwidth",
               "patch_vectors_per_sec": rows / (sm1_avg * 1e-3), "score_ms": sm1_avg,
               "step_ms": e0.elapsed_time(e1) / args.steps,
               "algorithmic_bytes_per_launch": rows * DIM * 2,
               "traffic": (tr["maxsim_umma<bf16,NM=1>"]["dram_bytes_per_patch_vector"] * rows
                           if tr and "maxsim_umma<bf16,NM=1>" in tr else None)}

    # ---- configs[0], the reference's own CPU-runnable case (100 pages, one 32-token query): single-call latency through the
    #      C-ABI with host buffers, next to the reference formulation on the host CPU for the same shapes
    cfg0 = None
    if rank == 0 and world == 1:
        try:
            cfg0 = config0_latency(packed, q_pin, out_pin, k, args.no_cpu_baseline)
        except Exception as e:  # noqa: BLE001  (an extra leg must never break the bench line)
            cfg0 = {"error": repr(e)[:200]}

    # ---- the reference's own GPU formulation on the same GPU (rank 0, N=1 only)
    torch_gpu = None
    if rank == 0 and world == 1:
        try:
            torch_gpu = torch_gpu_reference_rate(packed.view(torch.bfloat16).view(-1, DIM), q_dev, n_q)
        except Exception as e:  # noqa: BLE001  (a baseline must never break the bench line)
            torch_gpu = {"error": repr(e)[:200]}

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_rate(n_q, budget_s=args.cpu_budget)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "host_cpus": os.cpu_count(),
               "sample": f"{r['pages']} pages x {P_PATCH} patches x {DIM}-d, {n_q} queries x {T_TOK} tokens, "
                         f"{r['seconds']:.1f} s of score_multi_vector (torch einsum, fp32) on the host"}

        try:  # "top-k match vs reference" (BASELINE metric): the C-ABI path against the oracle on a sample of the shard
            cpu["topk_match"] = topk_match_vs_oracle(packed, q_pin, n_q, k)
        except Exception as e:  # noqa: BLE001
            cpu["topk_match"] = {"error": repr(e)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(args, n_pages),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "b200ms_search_host (C-ABI, pinned host buffers)" if world == 1 else
                           "pinned H2D + ShardedMaxSim.search (NCCL all-gather + merge) + D2H"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "hbm_regime": hbm, "cpu_baseline": cpu,
            "torch_gpu_reference_formulation": torch_gpu, "config0_latency": cfg0,
            "page_scores_per_sec": value / P_PATCH * n_q, "top1_sample": top1[:4],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--pages", type=int, default=262144, help="pages per GPU in the resident shard")
    ap.add_argument("--bq", type=int, default=32, help="queries per batch")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--ref-budget", type=float, default=6.0, help="seconds of CPU work per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
