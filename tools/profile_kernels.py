"""Small driver for ncu captures (NOT a benchmark: numbers printed under a profiler are never bench values).

Builds a synthetic shard (default 32768 pages x 1024 x 128, 8.6 GB bf16) and launches each scoring kernel a few times
in a fixed order so that `ncu -k regex:<name> -s <skip> -c <count>` picks a warm launch:
  bf16:   4x bq=32 (2 launches of maxsim_umma<0,4> each), then 4x bq=1 (maxsim_umma<0,1>)
  int8:   4x bq=32 (maxsim_umma<1,8>), 4x bq=1 (maxsim_umma<1,1>)          [--int8]
  binary: 4x bq=8 (maxsim_b1<8>), 4x bq=1 (maxsim_b1<1>)                   [--binary]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from morphik_core_b200.index import MaxSimIndex  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pages", type=int, default=32768)
ap.add_argument("--int8", action="store_true")
ap.add_argument("--binary", action="store_true")
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--bqs", type=str, default="32,1", help="query batch sizes to drive (bf16 / int8)")
args = ap.parse_args()

dev = torch.device("cuda", 0)
q_host = bench.make_queries(32)
packed, _ = bench.build_shard(args.pages, dev, 1234, q_host)
q_dev = q_host.to(dev)
lens = [bench.P_PATCH] * args.pages


def drive(idx, bqs):
    for bq in bqs:
        q = q_dev[: bq * 32].contiguous()
        for _ in range(args.reps):
            idx.search_device(q, [32] * bq, 10)
        torch.cuda.synchronize()
        ms = idx.score_times_ms(args.reps)
        print(f"{idx.dtype_name} bq={bq}: score ms {['%.3f' % m for m in ms]}", flush=True)


idx = MaxSimIndex(dtype="bf16")
idx.adopt_packed(packed, lens)
drive(idx, [int(b) for b in args.bqs.split(",")])
rows = packed.view(torch.bfloat16).view(-1, 128)
if args.int8:
    i8 = MaxSimIndex(dtype="int8")
    buf = torch.empty(rows.shape[0] * 128 + 1024, dtype=torch.uint8, device=dev)
    off = (-buf.data_ptr()) % 1024
    p8 = buf[off:off + rows.shape[0] * 128]
    step = 1 << 22
    for r0 in range(0, rows.shape[0], step):
        p8.view(torch.int8).view(-1, 128)[r0:r0 + step] = torch.clamp(torch.round(rows[r0:r0 + step].float() * 127.0), -127, 127).to(torch.int8)
    i8.adopt_packed(p8, lens)
    drive(i8, [int(b) for b in args.bqs.split(",")])
if args.binary:
    b1 = MaxSimIndex(dtype="binary")
    bits = torch.empty((rows.shape[0], 16), dtype=torch.uint8, device=dev)
    step = 1 << 22
    for r0 in range(0, rows.shape[0], step):
        bits[r0:r0 + step] = b1.sign_pack(rows[r0:r0 + step])
    braw = torch.empty(bits.numel() + 1024, dtype=torch.uint8, device=dev)
    off = (-braw.data_ptr()) % 1024
    pb = braw[off:off + bits.numel()]
    pb.copy_(bits.reshape(-1))
    b1.adopt_packed(pb, lens)
    drive(b1, [8, 1])
