"""Extract per-kernel counters from ncu reports into a small committed JSON (development tool).

    python tools/ncu_counters.py profiles/r2_ncu_counters.json gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]

bench.py reads `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum of one nn1_sweep_kernel launch at the headline
workload) from this file by kernel name instead of carrying a literal."""
import csv
import json
import os
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct"]
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3}


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    kernels = {}
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            if name in kernels:
                continue                                  # first launch of each kernel name
            c = {"report": os.path.basename(rep)}
            for m in METRICS:
                if m in hdr and r[hdr.index(m)]:
                    v = float(r[hdr.index(m)].replace(",", ""))
                    u = units[hdr.index(m)]
                    if m.startswith("dram__bytes"):
                        v *= SCALE.get(u, 1.0)
                    elif m == "gpu__time_duration.sum":
                        v *= SCALE.get(u, 1.0)            # microseconds
                    c[m] = v
            st = sorted(((float(r[i]), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i]), reverse=True)
            c["top_stalls_per_issue"] = {h.split("issue_stalled_")[1].split("_per_issue")[0]: round(v, 3) for v, h in st[:6]}
            kernels[name] = c
    json.dump({"source": "ncu --set full --clock-control none, one launch per kernel name; bytes in B, durations in us (cold cache, serialised)",
               "kernels": kernels}, open(out_path, "w"), indent=1)
    print("wrote", out_path, len(kernels), "kernels")


if __name__ == "__main__":
    main()
