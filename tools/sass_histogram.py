"""Per-kernel SASS opcode evidence for libstad.so (no GPU needed): for every kernel the counts of the Blackwell-native
instructions — UTCHMMA (tcgen05.mma, incl. the .2CTA pair form), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA
loads / stores), UTCBAR (tcgen05.commit), SYNCS (mbarrier), MUFU.EX2 / MUFU.TANH — next to the legacy tensor path
(HMMA: must be 0) and the instruction total.
    python tools/sass_histogram.py [path/to/libstad.so] > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU.EX2", "MUFU.TANH",
        "FFMA2", "HMMA", "total"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "simple-tad_b200", "libstad.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for w in WANT:
                if w == "UTCHMMA.2CTA":
                    if op.startswith("UTCHMMA") and ".2CTA" in op:
                        cur[w] += 1
                elif w != "total" and (op == w or op.startswith(w + ".")):
                    cur[w] += 1
    names = demangle(list(per))
    print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass; static counts, not executed counts)")
    print("# " + " | ".join(f"{w:>12s}" for w in WANT) + " | kernel")
    tot = collections.Counter()
    for k, c in per.items():
        short = re.sub(r"\(anonymous namespace\)::|stad::", "", names.get(k, k))
        short = re.sub(r"\(.*", "", short)
        print("  " + " | ".join(f"{c[w]:12d}" for w in WANT) + " | " + short)
        tot.update(c)
    print("  " + " | ".join(f"{tot[w]:12d}" for w in WANT) + " | ALL KERNELS")


if __name__ == "__main__":
    main()
