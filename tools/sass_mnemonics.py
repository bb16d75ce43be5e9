"""Per-kernel SASS mnemonic counts of libpcdist.so (development tool; runs without a GPU).

    python tools/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt

Evidence that the library is sm_100a code using the Blackwell instructions the design relies on: packed fp32x2 math (FFMA2 /
FADD2 / FMUL2), three-input min (FMNMX3), warp reductions (CREDUX), 1-D TMA bulk copies (UBLKCP) behind mbarriers (SYNCS),
programmatic dependent launch (ACQBULK / griddepcontrol lowers to it) -- and no tensor-core instruction (HMMA / UTC*MMA), as
BASELINE.json's north star asks for this K = 3 fp32 path."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "3dpointcloudattack_b200", "libpcdist.so")
WATCH = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FMNMX3", "FMNMX", "CREDUX", "REDUX", "VOTE", "UBLKCP", "SYNCS", "ACQBULK", "LDS", "STS", "LDG", "STG",
         "ATOMG", "REDG", "RED", "ATOM", "SHFL", "DFMA", "DMUL", "DADD", "MUFU", "BAR", "HMMA", "UTCHMMA", "UTCQMMA", "WARPSYNC"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    print("library:", os.path.relpath(LIB, ROOT), " architectures:", arch)
    fn, counts, totals = None, {}, collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            counts[fn][m.group(1)] += 1
            totals[m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"{len(counts)} kernels; mnemonic totals:", ", ".join(f"{k} {totals[k]}" for k in WATCH if totals[k]))
    print("tensor-core instructions (HMMA / UTC*MMA):", sum(v for k, v in totals.items() if "MMA" in k))
    print()
    for f, name in sorted(zip(counts, demangle), key=lambda t: t[1]):
        c = counts[f]
        short = re.sub(r"\(.*", "", name)
        print(f"{short}: {sum(c.values())} instr | " + ", ".join(f"{k} {c[k]}" for k in WATCH if c[k]))


if __name__ == "__main__":
    main()
