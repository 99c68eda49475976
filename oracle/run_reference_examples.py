#!/usr/bin/env python
"""Builds and runs ALL of the reference's own example programs (reference examples/*.c, unmodified)
twice -- linked with the unmodified reference (oracle/_ref/libalis_ref.a) and linked with this
repo's host layer over the CPU oracle engine (tests/bin/libalis_oracle.so, test infrastructure) --
and compares their text outputs.  Container only (needs /root/reference); the GPU twin of this
check is tests/test_gpu_examples.py.  Writes a JSON summary (default profiles/).

  python oracle/run_reference_examples.py [--work DIR] [--reuse] [--timeout S] [--out FILE]

Comparison: every .txt table, per column, |a-b| / max|column of the same file| (a value that is
rounding noise next to the physical columns of its file -- e.g. the S-polarised intensity of a
P-polarised emitter, 1e-16 against 0.4 -- is compared against the largest column)."""
import argparse
import glob
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ALIS_REFERENCE", "/root/reference")
# command lines of the examples that take arguments; examples that consume another example's output
# run in that example's directory (Waveguide_Propagation reads the .h5 mode profiles that
# Waveguide_Dispersion 3.0 wrote)
ARGS = {"Cross_Polarization_T": ["300", "120", "80", "200"], "Waveguide_Dispersion": ["3.0"]}
RUN_IN = {"Waveguide_Propagation": "Waveguide_Dispersion"}


def build(src, exe, mine):
    if mine:
        cmd = ["gcc", "-std=c99", "-O2", "-fopenmp", "-w", "-I" + os.path.join(ROOT, "include"), src,
               "-L" + os.path.join(ROOT, "tests", "bin"), "-lalis_oracle", "-Wl,-rpath," + os.path.join(ROOT, "tests", "bin"), "-lm", "-o", exe]
    else:
        cmd = ["gcc", "-std=c99", "-O3", "-fopenmp", "-w", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(REF, "src"), src,
               os.path.join(ROOT, "oracle", "_ref", "libalis_ref.a"), "-lm", "-o", exe]
    return subprocess.run(cmd, capture_output=True, text=True).returncode == 0


def tables(root):
    out = {}
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith(".txt") and not f.startswith("log"):
                try:
                    out[os.path.relpath(os.path.join(d, f), root)] = np.atleast_2d(np.loadtxt(os.path.join(d, f)))
                except Exception:
                    pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--work", default="/tmp/refex")
    ap.add_argument("--reuse", action="store_true", help="compare the outputs already under --work (run_<name>_{ref,mine}/)")
    ap.add_argument("--timeout", type=int, default=400)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_reference_examples_cpu.json"))
    ap.add_argument("--only", default="", help="comma-separated example names: rerun these and merge into --out")
    a = ap.parse_args()
    os.makedirs(a.work, exist_ok=True)
    report = {"host_threads": os.cpu_count(), "engine_of_our_twin": "CPU oracle (tests/bin/libalis_oracle.so)", "examples": {}}
    only = [x for x in a.only.split(",") if x]
    if only and os.path.exists(a.out):
        report = json.load(open(a.out))
    for src in sorted(glob.glob(os.path.join(REF, "examples", "*.c"))):
        name = os.path.basename(src)[:-2]
        if only and name not in only:
            continue
        rec = {"args": ARGS.get(name, [])}
        for tag in ("ref", "mine"):
            d = os.path.join(a.work, f"run_{RUN_IN.get(name, name)}_{tag}")
            if not a.reuse:
                if name not in RUN_IN:
                    subprocess.run(["rm", "-rf", d])
                os.makedirs(d, exist_ok=True)
                rec[tag + "_builds"] = build(src, os.path.join(d, name), tag == "mine")
                env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "tests", "bin") + os.pathsep + os.path.join(ROOT, "oracle"))
                t0 = time.time()
                try:
                    r = subprocess.run(["./" + name] + ARGS.get(name, []), cwd=d, env=env, stdout=open(os.path.join(d, f"log_{name}.txt"), "w"), stderr=subprocess.STDOUT, timeout=a.timeout)
                    rec[tag + "_rc"] = r.returncode
                except subprocess.TimeoutExpired:
                    rec[tag + "_rc"] = "timeout"
                rec[tag + "_seconds"] = round(time.time() - t0, 1)
        ta, tb = tables(os.path.join(a.work, f"run_{RUN_IN.get(name, name)}_ref")), tables(os.path.join(a.work, f"run_{RUN_IN.get(name, name)}_mine"))
        h5 = [os.path.relpath(os.path.join(d_, f), os.path.join(a.work, f"run_{RUN_IN.get(name, name)}_ref"))
              for d_, _, fs in os.walk(os.path.join(a.work, f"run_{RUN_IN.get(name, name)}_ref")) for f in fs if f.endswith(".h5")]
        if h5:
            same = sum(open(os.path.join(a.work, f"run_{RUN_IN.get(name, name)}_ref", f), "rb").read() ==
                       (open(os.path.join(a.work, f"run_{RUN_IN.get(name, name)}_mine", f), "rb").read() if os.path.exists(os.path.join(a.work, f"run_{RUN_IN.get(name, name)}_mine", f)) else b"") for f in h5)
            rec["h5_files_byte_identical"] = f"{same}/{len(h5)}"
        cmp_ = {}
        for k in sorted(set(ta) | set(tb)):
            if k not in ta or k not in tb:
                cmp_[k] = "only in " + ("ours" if k in tb else "reference")
            elif ta[k].shape != tb[k].shape:
                cmp_[k] = f"shape {ta[k].shape} vs {tb[k].shape}"
            else:
                cmp_[k] = float(np.abs(ta[k] - tb[k])[:, 1:].max() / (np.abs(ta[k])[:, 1:].max() + 1e-300)) if ta[k].shape[1] > 1 else float(np.abs(ta[k] - tb[k]).max())
        rec["max_error_relative_to_largest_column"] = cmp_
        report["examples"][name] = rec
        print(name, rec, file=sys.stderr)
    with open(a.out, "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
