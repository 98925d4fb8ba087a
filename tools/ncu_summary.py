"""Turn ncu captures (gpurun_out/*.ncu-rep) into the small committed summaries under profiles/<round>/.

  python tools/ncu_summary.py r01 gpurun_out/prof_umma_nm4.ncu-rep:maxsim_umma<bf16,NM=4>:PAGES ...

Writes profiles/<round>/<name>_summary.csv (selected raw metrics per captured launch) and updates
profiles/<round>/traffic.json with dram bytes per patch vector (read+write, per launch / rows per launch).
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ("Kernel Name", "gpu__time_duration", "dram__bytes", "dram__throughput", "gpu__dram_throughput", "pipe_tensor", "pipe_alu",
        "warps_active", "registers_per_thread", "sm__throughput", "cycles_elapsed.avg", "issue_active", "smsp__inst_executed.sum",
        "tc_wavefronts_mem_shared", "mem_tensor_cycles", "launch__grid_size", "launch__block_size", "lts__t_bytes.sum")


def main():
    rnd = sys.argv[1]
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", rnd)
    os.makedirs(out_dir, exist_ok=True)
    tpath = os.path.join(out_dir, "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for spec in sys.argv[2:]:
        rep, label, pages = spec.rsplit(":", 2)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        keep = [i for i, h in enumerate(hdr) if any(k in h for k in KEYS)]
        name = os.path.splitext(os.path.basename(rep))[0]
        with open(os.path.join(out_dir, f"{name}_summary.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow([hdr[i] for i in keep])
            w.writerow([units[i] for i in keep])
            for r in rows[2:]:
                w.writerow([r[i] for i in keep])
        rd, wr, du = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")

        def to_bytes(v, unit):
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]

        last = rows[-1]
        total = to_bytes(last[rd], units[rd]) + to_bytes(last[wr], units[wr])
        n_rows = int(pages) * 1024
        traffic[label] = {"dram_bytes_per_launch": total, "rows_per_launch": n_rows, "dram_bytes_per_patch_vector": total / n_rows,
                          "duration": f"{last[du]} {units[du]}", "source": f"profiles/{rnd}/{name}_summary.csv (ncu --set full, {pages} pages)"}
        print(label, traffic[label])
    json.dump(traffic, open(tpath, "w"), indent=1)


if __name__ == "__main__":
    main()
