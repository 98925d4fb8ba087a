"""Diagnostics for the CTA-pair kernel on a tiny corpus: prints where it differs from the one-CTA kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from morphik_core_b200.index import MaxSimIndex

rng = np.random.default_rng(1)
def rows(n):
    x = rng.standard_normal((n, 128)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)
for lens, ur in (([128, 128], 128), ([128, 128, 128, 128], 128), ([300, 40, 1000, 7, 128], 256)):
    pages = [rows(n) for n in lens]
    queries = [rows(32) for _ in range(16)]
    a = MaxSimIndex(dtype="bf16"); a.set_option("pair_cta", 1); a.set_option("unit_rows", ur); a.add_pages(pages)
    b = MaxSimIndex(dtype="bf16"); b.set_option("pair_cta", 0); b.set_option("unit_rows", ur); b.add_pages(pages)
    ga, gb = a.score_matrix(queries), b.score_matrix(queries)
    print("lens", lens, "equal", np.array_equal(ga, gb), "max diff", np.abs(ga - gb).max(), flush=True)
    if not np.array_equal(ga, gb):
        np.set_printoptions(precision=3, linewidth=200)
        print("pair:\n", ga[:, :8]); print("ref:\n", gb[:, :8])
        if ga.shape[1] >= 2:
            print("swapped-pages match:", np.allclose(ga[:, [1, 0]], gb[:, :2]))
