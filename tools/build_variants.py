"""Developer A/B aid: build variants of libb200ms.so that differ only in -D flags, next to the product library
(morphik-core_b200/lib/ab/libb200ms_<name>.so -- git-ignored, travels with a gpurun snapshot).  A process picks one with
B200MS_LIB=<path> (see _native.py); the product path never does.

  python tools/build_variants.py base= exact=-DB200MS_INT_MAX_FLOAT=0 one=-DB200MS_EPI_ONE_PHASE=1
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morphik_core_b200 import build_native as bn  # noqa: E402

REBUILD = ["maxsim_umma.cu", "maxsim_umma_pair.cu"]  # the translation units the A/B macros touch


def main():
    bn.build()
    out_dir = os.path.join(bn.LIB_DIR, "ab")
    os.makedirs(out_dir, exist_ok=True)
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition("=")
        defs = [f for f in flags.split(",") if f]
        objs, procs = [], []
        for src in bn.SOURCES:
            if src in REBUILD:
                obj = os.path.join(out_dir, f"{name}_{src.replace('.cu', '.o')}")
                procs.append(subprocess.Popen([bn._nvcc()] + bn.NVCC_FLAGS + defs + ["-c", os.path.join(bn.CSRC, src), "-o", obj],
                                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
            else:
                obj = os.path.join(bn.LIB_DIR, src.replace(".cu", ".o"))
            objs.append(obj)
        for p in procs:
            out, _ = p.communicate()
            spills = [ln for ln in out.splitlines() if "spill" in ln and "0 bytes spill stores, 0 bytes spill loads" not in ln]
            if p.returncode != 0:
                print(out)
                raise SystemExit(f"variant {name}: nvcc failed")
            if spills:
                print(f"variant {name}: SPILLS\n" + "\n".join(spills[:8]))
        lib = os.path.join(out_dir, f"libb200ms_{name}.so")
        subprocess.run([bn._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-ldl"], check=True)
        print(lib)


if __name__ == "__main__":
    main()
