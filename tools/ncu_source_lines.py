#!/usr/bin/env python3
"""Per-source-line warp-stall histogram of one kernel from an `ncu --set full --import-source on` report:
    python tools/ncu_source_lines.py gpurun_out/r2o_fused.ncu-rep _Z13k_gmres_fusedILi2EEv9FusedArgs wb_fused > profiles/...
ncu's CSV export of the source page carries the sampling columns only per SASS instruction, so the instruction
offsets are mapped back to source lines with `nvdisasm -g` on the cubin of the SAME build (the library is compiled
with -lineinfo).  Samples are per warp and include the producer warp's spin and every wait at a barrier."""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, mangled, unit = sys.argv[1], sys.argv[2], sys.argv[3]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "waiwera_b200", "libwaiwera_b200.so")], cwd=tmp,
                   capture_output=True)
    cubin = [f for f in glob.glob(os.path.join(tmp, "*.cubin")) if os.path.basename(f).startswith(unit + ".")][0]
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(dis) if l.startswith(".text.%s:" % mangled)][0]
    cur, omap = None, {}
    for l in dis[start + 1:]:
        if l.startswith(".text."):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
        if m and cur:
            omap[int(m.group(1), 16)] = cur
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    ia, isamp = hdr.index("Address"), hdr.index("# Samples")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(rows[2][ia], 16)
    by, why, tot = collections.Counter(), collections.defaultdict(collections.Counter), 0
    for r in rows[2:]:
        key = omap.get(int(r[ia], 16) - base, ("?", 0))
        s = int(r[isamp] or 0)
        tot += s
        by[key] += s
        for i in stall:
            if int(r[i] or 0):
                why[key][hdr[i][6:]] += int(r[i])
    src = {}
    print("# %s, kernel %s: %d warp samples; share, file:line, source, top stall reasons" % (os.path.basename(rep), mangled, tot))
    for (f, ln), s in by.most_common(40):
        if f not in src:
            p = os.path.join(ROOT, "waiwera_b200", "csrc", f)
            src[f] = open(p).read().split("\n") if os.path.exists(p) else []
        text = src[f][ln - 1].strip()[:88] if 0 < ln <= len(src[f]) else ""
        print("%5.1f%%  %s:%d  %-90s | %s" % (100.0 * s / tot, f, ln, text,
                                           ", ".join("%s %d" % kv for kv in why[(f, ln)].most_common(3))))


if __name__ == "__main__":
    main()
