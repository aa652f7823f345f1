#!/usr/bin/env python
"""Per-source-line executed-instruction shares of the kernels in an `ncu --set full --import-source on` report.

usage: python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top-n]

ncu's CSV source page lists SASS with per-instruction counters but no line numbers; the line table comes from
`nvdisasm --print-line-info` on the cubins of front_end_b200/libfe_b200.so (built with -lineinfo).  The two are joined
by instruction offset inside the function (functions are matched by name substring and instruction count).  Runs here
(no GPU needed)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def disassemble():
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "front_end_b200", "libfe_b200.so")], cwd=tmp,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    funcs = {}
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        name, line, fname = None, None, None
        for l in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
            if m:
                name = m.group(1)
                funcs[name] = []
                continue
            if name is None:
                continue
            if l.strip().startswith(".section"):
                name = None
                continue
            m = re.search(r'//## File "(.*?)", line (\d+)', l)
            if m:
                fname, line = m.group(1), int(m.group(2))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                funcs[name].append((int(m.group(1), 16), fname, line))
    return funcs


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    funcs = disassemble()
    rows = list(csv.reader(io.StringIO(out)))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            kname = rows[i][1]
            hdr = rows[i + 1]
            ci, ai = hdr.index("Instructions Executed"), hdr.index("Address")
            j = i + 2
            sass = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                r = rows[j]
                if len(r) > ci and r[ai].startswith("0x"):
                    sass.append((int(r[ai], 16), int(r[ci])))
                j += 1
            i = j
            if want not in kname:
                continue
            base = re.sub(r"^void ", "", kname).split("(")[0].split("<")[0].split("::")[-1]
            cands = [n for n, ins in funcs.items() if base in n and len(ins) == len(sass)]
            print("== %s  (%d SASS instructions, %.3g warp instructions executed)" % (kname[:110], len(sass), sum(n for _, n in sass)))
            if not cands:
                print("   no line table matched")
                continue
            table = funcs[cands[0]]
            a0 = sass[0][0]
            off = {o: (f, l) for o, f, l in table}
            per = defaultdict(int)
            for a, n in sass:
                per[off.get(a - a0, (None, None))] += n
            tot = float(sum(per.values())) or 1.0
            srcs = {}
            for (f, l), n in sorted(per.items(), key=lambda kv: -kv[1])[:topn]:
                text = "?"
                if f and os.path.exists(f):
                    srcs.setdefault(f, open(f).read().splitlines())
                    text = srcs[f][l - 1].strip()[:120] if l and l <= len(srcs[f]) else "?"
                print("   %5.1f%%  %s:%s  %s" % (100 * n / tot, os.path.basename(f) if f else "?", l, text))
        else:
            i += 1


if __name__ == "__main__":
    main()
