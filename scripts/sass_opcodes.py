#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS opcodes in the shipped library -> profiles/sass_opcodes.txt.

    python scripts/sass_opcodes.py            # after smart-nar_fast_tts_b200/build.py

UTCHMMA / UTCQMMA = tcgen05.mma; LDTM / STTM = tcgen05.ld / st (TMEM); UTMALDG / UTMASTG = TMA tensor load / store;
UTCBAR = tcgen05.commit -> mbarrier; SYNCS = mbarrier ops; ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents;
LDGSTS = cp.async; UBLKCP = bulk copy; HMMA / FFMA for contrast."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "smart-nar_fast_tts_b200", "libfs2_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "ACQBULK", "PREEXIT", "LDGSTS", "UBLKCP",
       "HMMA", "FFMA2", "FFMA", "MUFU"]


def kernel_name(demangled: str) -> str:
    """`void ns::k<(int)2>(args...)` -> `k<(int)2>`: cut at the '(' that opens the parameter list (depth 0 of <...>)."""
    s = demangled[5:] if demangled.startswith("void ") else demangled
    depth = 0
    for i, ch in enumerate(s):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return s[:i]
    return s


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    per[cur][o] += 1
    names = subprocess.run(["cu++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    out = [f"# {os.path.relpath(LIB, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass; scripts/sass_opcodes.py)",
           "# " + " ".join(f"{o:>8}" for o in OPS) + "  kernel"]
    tot = collections.Counter()
    for (mangled, c), name in zip(per.items(), names):
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
        name = kernel_name(name)
        out.append("  " + " ".join(f"{c[o]:>8}" for o in OPS) + "  " + name)
        tot.update(c)
    out.append("  " + " ".join(f"{tot[o]:>8}" for o in OPS) + "  TOTAL")
    dst = os.path.join(ROOT, "profiles", "sass_opcodes.txt")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[-1:]), "->", dst, f"({len(per)} kernels)")


if __name__ == "__main__":
    sys.exit(main())
