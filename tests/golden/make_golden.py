#!/usr/bin/env python
"""Regenerates tests/golden/stepdump_hashes.json from the UNMODIFIED reference
(oracle/_ref/bin/stepdump_ref, built by oracle/build_ref.sh from /root/reference).
Runs in the build container only (the reference is not present on the GPU box).

For every configuration in oracle/harness/configs.h the complete field state after the
default number of steps is hashed (sha256 over the raw dump, so parity is bit-exact) and
per-array L2 norms are recorded for diagnostics.
"""
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import helpers  # noqa: E402

out = {}
with tempfile.TemporaryDirectory() as tmp:
    for cfg in helpers.CONFIGS:
        path = os.path.join(tmp, cfg + ".bin")
        log = helpers.run_stepdump("stepdump_ref", cfg, path, cwd=tmp)
        a = helpers.read_dump(path)
        out[cfg] = {
            "sha256": helpers.sha256_file(path),
            "shape": list(a.shape),
            "l2": [[[float(np.sqrt((a[m, s, c].astype(np.float64) ** 2).sum())) for c in range(3)] for s in range(2)] for m in range(a.shape[0])],
            "summary": log.strip().splitlines()[-1],
        }
        print(cfg, out[cfg]["sha256"][:16], out[cfg]["summary"])
with open(os.path.join(helpers.GOLDEN, "stepdump_hashes.json"), "w") as f:
    json.dump(out, f, indent=1)

# monitor values (probes, flux, absorption, phasers, far field ...) printed by oracle/harness/monitors.c
# linked against the reference; one OpenMP thread so that its `omp reduction` sums run in index order
mon = {}
for cfg in helpers.MONITOR_CONFIGS:
    mon[cfg] = helpers.run_monitors("monitors_ref", cfg, threads=1)
    print(cfg, len(mon[cfg]), "monitor values")
with open(os.path.join(helpers.GOLDEN, "monitor_values.json"), "w") as f:
    json.dump(mon, f, indent=1)

# the remaining public API (clone / reboot / removeSources / removeObjects / LT / nkFile / phasor fluxes):
# oracle/harness/api_misc.c linked against the reference
import subprocess  # noqa: E402
with tempfile.TemporaryDirectory() as tmp:
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([os.path.join(helpers.REF_BIN, "api_misc_ref")], cwd=tmp, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-500:]
    vals = helpers.parse_tag_values(r.stdout)
    assert vals and vals[-1][0] == "done"
    print("api_misc", len(vals), "values")
with open(os.path.join(helpers.GOLDEN, "api_misc_values.json"), "w") as f:
    json.dump(vals, f, indent=1)

# slices of a graded grid resampled onto the uniform raster (oracle/harness/graded_slices.c): sha256 per file
with tempfile.TemporaryDirectory() as tmp:
    files = helpers.run_graded_slices("graded_slices_ref", tmp)
    print("graded_slices", len(files), "files")
with open(os.path.join(helpers.GOLDEN, "graded_slices.json"), "w") as f:
    json.dump(files, f, indent=1)
