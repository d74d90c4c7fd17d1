"""Dynamic opcode mix per pair from an .ncu-rep source page. Usage:
    python profiles/opcode_mix.py rep.ncu-rep pairs_per_launch
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
raw = _ncu_pages.page(rep, "source")
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
ops, tot = collections.Counter(), 0
for r in rows[2:]:
    if len(r) <= iE:
        continue
    try:
        n = int(r[iE])
    except ValueError:
        continue
    toks = r[iS].strip().split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    ops[op] += n
    tot += n
fp64 = sum(n for op, n in ops.items() if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"kernel: {rows[0][1]}")
print(f"thread instructions per pair: total {tot * 32 / pairs:.1f}, FP64 pipe {fp64 * 32 / pairs:.1f}, "
      f"other {(tot - fp64) * 32 / pairs:.1f}; issue model 2*FP64 + other = "
      f"{(2 * fp64 + tot - fp64) * 32 / pairs:.1f}")
for op, n in ops.most_common(24):
    print(f"  {op:8s} {n * 32 / pairs:8.1f}")
