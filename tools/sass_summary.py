"""Instruction mix of every kernel in libabm.so (cuobjdump -sass): loads/stores by width, shared-memory atomics, barriers,
warp collectives.  `python tools/sass_summary.py > profiles/<round>_sass_mnemonics.txt`"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "animal-movement-abm_b200", "libabm.so")
GROUPS = [("LDG.128", r"^LDG\.E(\.\w+)*\.128"), ("LDG.64", r"^LDG\.E(\.\w+)*\.64"), ("LDG.32", r"^LDG\.E(?!.*\.(64|128|U8|S8|U16|S16))"), ("LDG.8/16", r"^LDG\.E.*\.(U8|S8|U16|S16)"),
          ("STG.128", r"^STG\.E(\.\w+)*\.128"), ("STG.64", r"^STG\.E(\.\w+)*\.64"), ("STG.32", r"^STG\.E(?!.*\.(64|128|U8|S8|U16|S16))"), ("STG.8/16", r"^STG\.E.*\.(U8|S8|U16|S16)"),
          ("LDS", r"^LDS"), ("STS", r"^STS"), ("ATOMS", r"^ATOMS"), ("ATOMG/RED", r"^(ATOMG|RED)"), ("BAR", r"^BAR"), ("SHFL/VOTE/MATCH/REDUX", r"^(SHFL|VOTE|MATCH|REDUX)"),
          ("IMAD/IADD/LOP", r"^(IMAD|IADD|LOP|SHF|LEA|ISETP|SEL|PRMT)"), ("F64", r"^D(ADD|MUL|FMA|SETP)"), ("BRA", r"^BRA")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["total"] += 1
            for name, pat in GROUPS:
                if re.match(pat, op):
                    kernels[cur][name] += 1
                    break
    print(f"# {os.path.relpath(LIB, ROOT)}: static SASS instruction counts per kernel (sm_100a)")
    names = [g[0] for g in GROUPS]
    for k, c in kernels.items():
        if c["total"] < 200:
            continue
        short = re.sub(r"\(.*", "", k)[:110]
        print(f"\n{short}\n  total {c['total']}: " + ", ".join(f"{n} {c[n]}" for n in names if c[n]))


if __name__ == "__main__":
    main()
