"""Per-kernel SASS opcode histogram of the built library (profiles/r02/sass_histogram.txt).

    python tools/sass_histogram.py [path/to/lib.so] > profiles/r02/sass_histogram.txt

What the mnemonics prove (B200_PROFILING.md): UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk (the TMA engine, 1-D bulk form), UTMALDG = cp.async.bulk.tensor, DMMA = mma.sync f64 (the only fp64
tensor instruction there is), HMMA = legacy mma.sync (the optional BOPY_B200_F32_ENGINE=mma_sync engine and its peak probe).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bopy_b200", "lib", "libbopy_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "DMMA", "HMMA", "DFMA", "FFMA",
         "F2F", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "BAR"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.split("\n")
        return dict(zip(names, out))
    except (OSError, subprocess.CalledProcessError):
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    names = demangle(list(kernels))
    total = collections.Counter()
    print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass), sm_100a")
    print("# kernel | instructions | " + " ".join(WATCH))
    for k, c in kernels.items():
        total.update(c)
        short = re.sub(r"\(.*", "", names[k]).replace("void ", "").replace("bopy::", "")
        row = " ".join(f"{op}={c[op]}" for op in WATCH if c[op])
        print(f"{short:110s} | {sum(c.values()):6d} | {row}")
    print("# whole library: " + " ".join(f"{op}={total[op]}" for op in WATCH))


if __name__ == "__main__":
    main()
