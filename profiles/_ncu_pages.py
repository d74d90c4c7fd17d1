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
