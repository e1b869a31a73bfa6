#!/usr/bin/env python
"""Per-kernel SASS opcode census of the in-tree library (cuobjdump -sass): the evidence that bulk TMA (UBLKCP), mbarrier
(SYNCS), cluster barriers (UCGABAR / barrier.cluster), PDL (ACQBULK / griddepcontrol), REDUX and the FP32 scoring loop are
what the source says.  usage: sass_listing.py <out_dir>   (writes sass_<file>.txt per .cu)"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "rdpn6d_b200", "librdpn6d_b200.so")
out_dir = sys.argv[1]
os.makedirs(out_dir, exist_ok=True)
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, check=True, capture_output=True)
    for cubin in sorted(os.listdir(td)):
        name = cubin.split(".")[0]
        if name.startswith("librdpn"):
            continue
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(td, cubin)], capture_output=True, text=True).stdout
        fn, census = None, collections.OrderedDict()
        for line in txt.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
                census[fn] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
            if m and fn:
                census[fn][m.group(1)] += 1
        with open(os.path.join(out_dir, "sass_%s.txt" % name), "w") as f:
            f.write("# cuobjdump -sass of rdpn6d_b200/librdpn6d_b200.so (%s): instruction count per opcode and kernel\n" % cubin)
            for fn, c in census.items():
                tot = sum(c.values())
                f.write("\n== %s   (%d instructions)\n" % (fn[:160], tot))
                keys = sorted(c, key=lambda k: (-c[k], k))
                f.write("  " + "  ".join("%s:%d" % (k, c[k]) for k in keys) + "\n")
print("written to", out_dir)
