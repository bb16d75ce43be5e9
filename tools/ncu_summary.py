"""Summarise an .ncu-rep (development tool): headline metrics, stall reasons, executed-instruction regions.
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [--regions]"""
import csv
import subprocess
import sys
from collections import Counter


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_allocated",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:70s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        st = [(float(r[i]), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i]]
        print("  stalls per issue:", ", ".join(f"{h.split('issue_stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, h in sorted(st, reverse=True)[:8]))
    if "--regions" not in sys.argv:
        return
    rows = page(rep, "source")
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    iS, iI, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    op, tot = Counter(), 0
    for r in data:
        s = r[iS].split()
        o = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
        op[o] += int(r[iI]); tot += int(r[iI])
    print("  opcodes:", ", ".join(f"{o} {100 * n / tot:.1f}%" for o, n in op.most_common(14)))
    prev, start, acc, sm, out = None, 0, 0, 0, []
    for k, r in enumerate(data):
        n = int(r[iI])
        if prev is None or abs(n - prev) > 0.02 * max(prev, 1):
            if prev is not None:
                out.append((start, k - 1, prev, acc, sm))
            start, acc, sm = k, 0, 0
        prev = n; acc += n; sm += int(r[iSm])
    out.append((start, len(data) - 1, prev, acc, sm))
    ts = sum(o[4] for o in out)
    for s, e, n, a, m in out:
        if a > 0.01 * tot or m > 0.01 * ts:
            print(f"  lines {s:5d}-{e:5d} ({e - s + 1:4d} instr) x{n / 1e3:9.1f}K  = {a / 1e6:7.1f}M ({100 * a / tot:4.1f}% instr, {100 * m / ts:4.1f}% samples)  {data[s][iS].strip()[:44]}")


if __name__ == "__main__":
    main()
