"""Opcode histogram of a built object's SASS, per kernel: python tools/sass_histogram.py soundsym_b200/csrc/_obj/dtw_h2.o [> profiles/...]
(the evidence that the scans really are tcgen05 / TMEM / bulk-TMA kernels: UTCHMMA, LDTM, UBLKCP, UTCBAR, VHMNMX, FMNMX3 ...)."""
import collections
import re
import subprocess
import sys

for obj in sys.argv[1:]:
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    print("==== %s" % obj)
    for f in re.split(r"\n\s+Function : ", txt)[1:]:
        name = f.split("\n")[0]
        ops = collections.Counter()
        for l in f.split("\n"):
            m = re.match(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_x]+)*)", l)
            if m:
                ops[m.group(1).split(".")[0] + ("." + m.group(1).split(".")[1] if m.group(1).startswith(("LDTM", "UTC", "UBLKCP")) and "." in m.group(1) else "")] += 1
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        top = ops.most_common(48)
        proof = [(k, v) for k, v in ops.items() if k.startswith(("UTC", "LDTM", "STTM", "UBLKCP", "UTMA", "USETMAXREG", "VHMNMX", "FMNMX3", "HMMA")) and (k, v) not in top]
        print("%s\n   %d instructions: %s%s" % (demangled[:140], sum(ops.values()), ", ".join("%s %d" % kv for kv in top),
                                               ("; [tensor / TMEM / TMA] " + ", ".join("%s %d" % kv for kv in sorted(proof))) if proof else ""))
