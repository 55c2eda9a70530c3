#!/usr/bin/env python
"""profiles/sass_summary.txt: which Blackwell instructions each kernel of libsgic.so contains.

    python scripts/sass_summary.py > profiles/sass_summary.txt

`cuobjdump -sass` of the shipped library, one line per kernel with the counts of the mnemonics that show the
sm_100a paths are really there: UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor load), UBLKCP (TMA bulk copy), LDTM
(tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier), FHFMA (mixed-precision FMA), REDUX, plus register count
from `cuobjdump -res-usage`.  Runs on the CPU-only build box."""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "searchable-generative-image-compression_b200" / "libsgic.so"
MNEMONICS = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "FHFMA", "HFMA2", "FFMA", "REDUX", "ATOMG", "LDS.128",
             "LDG", "STG", "ELECT", "R2UR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+)", ln)
        if m and cur:
            regs[cur] = int(m.group(1))
    funcs = sass.split("Function : ")[1:]
    names = [f.split("\n", 1)[0].strip() for f in funcs]
    pretty = demangle(names)
    print(f"# cuobjdump -sass {LIB.relative_to(ROOT)}  ({LIB.stat().st_size} bytes)")
    arch = re.search(r"arch = (sm_\w+)", sass)
    print(f"# arch: {arch.group(1) if arch else '?'}; kernels: {len(funcs)}")
    print("# kernel | registers | instructions | " + " ".join(MNEMONICS))
    rows = []
    for name, body in zip(names, funcs):
        ins = [l for l in body.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
        c = Counter()
        for l in ins:
            for m in MNEMONICS:
                if re.search(r"\b" + re.escape(m), l):
                    c[m] += 1
        short = re.sub(r"\(.*", "", pretty.get(name, name)).replace("void sgic::", "")
        rows.append((short, regs.get(name, -1), len(ins), c))
    for short, r, n, c in sorted(rows):
        print(f"{short} | {r} | {n} | " + " ".join(f"{m}={c[m]}" for m in MNEMONICS if c[m]))
    tot = Counter()
    for _, _, _, c in rows:
        tot.update(c)
    print("# total: " + " ".join(f"{m}={tot[m]}" for m in MNEMONICS))


if __name__ == "__main__":
    sys.exit(main())
