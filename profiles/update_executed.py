"""Record the ncu-measured executed instructions per pair of a kernel build in
profiles/executed_per_pair.json, keyed by a hash of the kernel sources it depends on (bench.py
prints the numbers only while that hash matches the sources of the running build). Usage:

    python profiles/update_executed.py WORKLOAD rep.ncu-rep PAIRS_PER_LAUNCH SOURCE_TXT [DRAM_BYTES]

WORKLOAD: layer_gz | c1_gz | tensor | mag_b | eqs | tess_gz; SOURCE_TXT: the committed summary
under profiles/ this entry is taken from; DRAM_BYTES: dram__bytes_read.sum + write.sum of one
FULL-SIZE launch (from a separate --metrics capture), optional.
"""
import collections
import csv
import io
import json
import os
import subprocess
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
import _ncu_pages  # noqa: E402
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

PRISM_FILES = ["hb200_fast.cuh", "hb200_kernels.cuh", "hb200_math.cuh", "hb200_tables.h", "hb200_xmath.cuh"]
FILES = {"layer_gz": PRISM_FILES, "c1_gz": PRISM_FILES, "tensor": PRISM_FILES, "mag_b": PRISM_FILES,
         "eqs": ["hb200_kernels.cuh", "hb200_math.cuh", "hb200_tables.h", "hb200_xmath.cuh"],
         "tess_gz": ["hb200_tess.cuh", "hb200_trig.cuh", "hb200_math.cuh", "hb200_tables.h", "hb200_xmath.cuh"]}

workload, rep, pairs, source = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
dram = int(sys.argv[5]) if len(sys.argv) > 5 else None
sections = _ncu_pages.opcode_sections(rep)  # several kernels (tesseroids: root pass + walks): their sum
ops = collections.Counter()
for _name, sec_ops, _lanes in sections:
    ops.update(sec_ops)
tot = sum(ops.values())
fp64 = sum(n for op, n in ops.items() if op in _ncu_pages.FP64_OPS)
kernel_name = " + ".join(name for name, _o, _l in sections)
path = bench.EXECUTED_FILE
table = json.load(open(path)) if os.path.exists(path) else {}
entry = {"kernel": kernel_name, "fp64": round(fp64 * 32 / pairs, 2), "other": round((tot - fp64) * 32 / pairs, 2),
         "pairs_per_launch": pairs, "files": FILES[workload], "sha16": bench.csrc_sha16(FILES[workload]),
         "source": source}
if dram is not None:
    entry["dram_bytes_per_launch"] = dram
elif workload in table and "dram_bytes_per_launch" in table[workload]:
    entry["dram_bytes_per_launch"] = table[workload]["dram_bytes_per_launch"]
table[workload] = entry
if workload == "layer_gz":
    table["c1_gz"] = dict(entry, source=source + " (same kernel as layer_gz)")
    table["c1_gz"].pop("dram_bytes_per_launch", None)
json.dump(table, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(entry))
