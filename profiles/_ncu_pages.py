"""Pages of an ncu report as CSV text. `rep` is either an .ncu-rep (ncu is run on it) or the base
name of pages exported on the GPU box by scripts/gpu_step.sh::export_pages (BASE.raw.csv,
BASE.source.csv.gz, BASE.srcsass.csv.gz) — the reports themselves (20-40 MB each) do not fit the
64 MiB that one gpurun call may bring back."""
import gzip
import os
import subprocess

ARGS = {"raw": ["--page", "raw", "--csv"], "source": ["--page", "source", "--csv"],
        "srcsass": ["--page", "source", "--csv", "--print-source", "cuda,sass"]}


def page(rep, name):
    if rep.endswith(".ncu-rep") and os.path.exists(rep):
        return subprocess.run(["ncu", "-i", rep, *ARGS[name]], capture_output=True, text=True).stdout
    base = rep[:-len(".ncu-rep")] if rep.endswith(".ncu-rep") else rep
    for path, opener in ((f"{base}.{name}.csv", open), (f"{base}.{name}.csv.gz", gzip.open)):
        if os.path.exists(path):
            with opener(path, "rt") as fh:
                return fh.read()
    raise FileNotFoundError(f"no ncu report or exported '{name}' page for {rep}")


def opcode_sections(rep):
    """[(kernel name, Counter opcode -> warp instructions executed, avg active threads)] of the
    source page, one entry per captured launch (ncu prints every launch of a multi-kernel report
    twice: identical sections are dropped)."""
    import collections
    import csv
    import hashlib
    import io

    rows = list(csv.reader(io.StringIO(page(rep, "source"))))
    sections, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            sections.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    out, seen = [], set()
    for sec in sections:
        digest = hashlib.sha256(repr(sec["rows"]).encode()).hexdigest()
        if digest in seen or not sec["rows"]:
            continue
        seen.add(digest)
        hdr = sec["rows"][0]
        iS, iE, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        ops, threads = collections.Counter(), 0
        for r in sec["rows"][1:]:
            if len(r) <= iT:
                continue
            try:
                n = int(r[iE])
                threads += int(r[iT])
            except ValueError:
                continue
            toks = r[iS].strip().split()
            if not toks:
                continue
            op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
            ops[op] += n
        total = sum(ops.values())
        out.append((sec["name"], ops, threads / total if total else 0.0))
    return out


FP64_OPS = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
