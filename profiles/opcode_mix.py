"""Dynamic opcode mix per pair from the source page of an ncu report (every captured kernel, and
their sum when there are several). Usage:
    python profiles/opcode_mix.py rep.ncu-rep|exported-base pairs_per_launch
Counts are WARP instructions x 32 / pairs ("thread-instruction slots per pair", the unit of the
issue model in DESIGN.md section 4); `lanes` is the average number of active threads.
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _ncu_pages  # noqa: E402

rep, pairs = sys.argv[1], float(sys.argv[2])
sections = _ncu_pages.opcode_sections(rep)
combined = collections.Counter()
for name, ops, lanes in sections:
    combined.update(ops)
if len(sections) > 1:
    sections = sections + [("ALL KERNELS ABOVE", combined, 0.0)]
for name, ops, lanes in sections:
    tot = sum(ops.values())
    fp64 = sum(n for op, n in ops.items() if op in _ncu_pages.FP64_OPS)
    print(f"kernel: {name}")
    print(f"thread instructions per pair: total {tot * 32 / pairs:.1f}, FP64 pipe {fp64 * 32 / pairs:.1f}, "
          f"other {(tot - fp64) * 32 / pairs:.1f}; issue model 2*FP64 + other = "
          f"{(2 * fp64 + tot - fp64) * 32 / pairs:.1f}" + (f"; active lanes {lanes:.1f} of 32" if lanes else ""))
    for op, n in ops.most_common(24):
        print(f"  {op:8s} {n * 32 / pairs:8.1f}")
    print()
