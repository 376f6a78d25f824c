#!/usr/bin/env python
"""Static check of the built objects: in every kernel, no global load may be scheduled before the
griddepcontrol.wait (SASS: ACQBULK) unless it is a non-coherent load of weights that was written on purpose
(LDG.E.CONSTANT before ACQBULK is reported; the source must then be checked by hand).

    python scripts/check_pdl_sass.py        # exit code 1 when a suspicious load is found
"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bad = 0
for obj in sorted(glob.glob(os.path.join(ROOT, "smart-nar_fast_tts_b200", "build", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for fn in sass.split("Function : ")[1:]:
        name = fn.split("\n", 1)[0].strip()
        ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", fn)
        try:
            w = next(i for i, x in enumerate(ins) if "ACQBULK" in x)
        except StopIteration:
            print(f"NO-WAIT   {os.path.basename(obj)}: {name[:90]}")
            bad += 1
            continue
        early = [x.strip() for x in ins[:w] if re.search(r"\b(LDG|LD\.E|UTMALDG|LDGSTS|ATOM|RED|STG)\b", x.split("(")[0])]
        # the tcgen05 GEMMs stage LayerNorm gamma / beta / the final dot vector (weights no kernel of the forward writes)
        # before the wait on purpose: up to three predicated non-coherent loads
        if "tc_conv_gemm" in name and len(early) <= 3 and all("LDG.E.CONSTANT" in x for x in early):
            print(f"weights   {os.path.basename(obj)}: {name[:70]}: {len(early)} LDG.E.CONSTANT before the wait (expected)")
            continue
        if early:
            print(f"EARLY     {os.path.basename(obj)}: {name[:70]}: {early[:4]}")
            bad += 1
print("kernels with findings:", bad)
sys.exit(1 if bad else 0)
