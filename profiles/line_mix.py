"""Executed instructions per pair by SOURCE LINE (and opcode) from an .ncu-rep captured with
--import-source on (kernels built with -lineinfo). Usage:
    python profiles/line_mix.py rep.ncu-rep pairs_per_launch [top_n]
"""
import collections
import csv
import io
import subprocess
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
import _ncu_pages  # noqa: E402
import sys

rep, pairs = sys.argv[1], float(sys.argv[2])
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = _ncu_pages.page(rep, "srcsass")
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
lines = collections.defaultdict(lambda: [0, 0, collections.Counter(), ""])
cur_file, cur_line, cur_text = "", "", ""
iE = None
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        iE = r.index("Instructions Executed")
        continue
    if r[0] == "Function Name" or iE is None or len(r) <= iE:
        continue
    if r[0]:  # a CUDA source line: its SASS rows follow
        cur_line, cur_text = r[0], r[1].strip()
        continue
    try:
        n = int(r[iE])
    except ValueError:
        continue
    toks = r[3].strip().split()
    if not toks:
        continue
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    e = lines[(cur_file, cur_line)]
    e[3] = cur_text
    if op in FP64:
        e[0] += n
    else:
        e[1] += n
    e[2][op] += n
tot64 = sum(e[0] for e in lines.values()) * 32 / pairs
toto = sum(e[1] for e in lines.values()) * 32 / pairs
print(f"per pair: FP64 {tot64:.1f}, other {toto:.1f}")
rank = sorted(lines.items(), key=lambda kv: -(2 * kv[1][0] + kv[1][1]))
for (f, ln), e in rank[:top_n]:
    ops = ", ".join(f"{op} {n * 32 / pairs:.1f}" for op, n in e[2].most_common(6))
    print(f"{e[0] * 32 / pairs:6.1f} {e[1] * 32 / pairs:6.1f}  {f}:{ln:>4}  {e[3][:70]:70s} | {ops}")
