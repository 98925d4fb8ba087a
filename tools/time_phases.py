"""Phase timing of one single-query search_device step (CUDA events around each C-ABI call) -- diagnostics only."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from morphik_core_b200 import _native as nat
from morphik_core_b200.index import MaxSimIndex, _vp, _aligned_bytes

pages = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = torch.device("cuda", 0)
q_host = bench.make_queries(32)
packed, _ = bench.build_shard(pages, dev, 1234, q_host)
idx = MaxSimIndex(dtype="bf16"); idx.adopt_packed(packed, [1024] * pages); idx._attach()
for bq in (1, 32):
    q = q_host[: bq * 32].to(dev).contiguous(); lens = [32] * bq; lens_c = nat.i32_array(lens)
    groups = bq; gp = max((groups + 3) // 4 * 4, 4)
    q_packed = _aligned_bytes(gp * 32 * 256, dev)
    ld = (pages + 31) // 32 * 32
    scores = torch.zeros((gp, ld), dtype=torch.float32, device=dev)
    goff = (ctypes.c_int32 * (bq + 1))(); ng = ctypes.c_int(0)
    ts = torch.empty((bq, 10), dtype=torch.float32, device=dev); ti = torch.empty((bq, 10), dtype=torch.int64, device=dev); tc = torch.empty((bq,), dtype=torch.int32, device=dev)
    st = idx._stream()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    acc = np.zeros(4)
    for it in range(12):
        ev[0].record()
        idx.h.check(nat.lib.b200ms_pack_queries(idx.h.ptr, _vp(q), nat.F32, lens_c, bq, _vp(q_packed), nat.BF16, ctypes.c_float(127.0), goff, ctypes.byref(ng), st))
        ev[1].record()
        idx.h.check(nat.lib.b200ms_score(idx.h.ptr, _vp(q_packed), ng.value, lens_c, goff, bq, _vp(scores), ld, st))
        ev[2].record()
        idx.h.check(nat.lib.b200ms_topk(idx.h.ptr, _vp(scores), nat.F32, pages, ld, goff, bq, None, 10, ctypes.c_float(1.0), 0, _vp(ts), _vp(ti), _vp(tc), st))
        ev[3].record()
        torch.cuda.synchronize()
        if it >= 2:
            acc += [ev[i].elapsed_time(ev[i + 1]) for i in range(3)] + [ev[0].elapsed_time(ev[3])]
    print(f"bq={bq} pages={pages}: pack {acc[0]/10:.3f} ms, score {acc[1]/10:.3f} ms, topk {acc[2]/10:.3f} ms, total {acc[3]/10:.3f} ms", flush=True)
