"""Round-2 measurement prep: wall time (host call incl. the D2H of the result and the synchronise)
of the device-side monitor calls through the engine C ABI, with the algorithmic bytes each one
moves, on one GPU.  Usage (GPU box):

    python profiles/monitors_probe.py [n]            # n^3-cell box, default 512
    ncu --set full --clock-control none -k regex:"eval_box|reduce_absmax|column_sums|surface_sums|geometry_fill" \
        -o gpurun_out/monitors python profiles/monitors_probe.py 256

ALISB_PROBE_ORACLE=1 runs the same script against the CPU oracle (logic check without a GPU)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alis_b200 import capi, synthetic  # noqa: E402

lib = capi.load_oracle() if os.environ.get("ALISB_PROBE_ORACLE") else capi.load_engine()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
w = synthetic.BoxWorld(n, n, n)
e = w.make_engine(lib)
e.run(20)
e.set_spacing(0, 2, np.full(n + 2, 10.0, np.float32))
out = {"backend": lib.alisb_backend().decode(), "cells": n ** 3}


def timed(name, fn, bytes_moved, reps=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    dt = (time.perf_counter() - t0) / reps
    out[name] = {"ms": round(dt * 1e3, 4), "algorithmic_bytes": bytes_moved, "GBps": round(bytes_moved / dt * 1e-9, 1)}
    return r


whole = [1, n - 1, 1, n - 1, 1, n - 1]
cells = (n - 1) ** 3
plane = [1, n - 1, 1, n - 1, n // 2, n // 2]
ee, sz, ez = e.field_fn(capi.FN_EE), e.field_fn(capi.FN_S, 0, 2), e.field_fn(capi.FN_NODE, 0, 2)
timed("worldMax_EE (reduce_absmax, 3 arrays)", lambda: e.reduce_absmax(ee, whole), 12 * cells)
timed("worldMax_Ez (reduce_absmax, 1 array)", lambda: e.reduce_absmax(ez, whole), 4 * cells)
timed("worldSum_Sz (column_sums, 4 arrays)", lambda: e.column_sums(sz, whole), 16 * cells + 4 * (n - 1) ** 2)
timed("sliceSnap_EE_XY (eval_box, plane)", lambda: e.eval_box(ee, plane), (12 * 2 + 4) * (n - 1) ** 2)
timed("sliceSnap_Ez_XY (eval_box, plane)", lambda: e.eval_box(ez, plane), (4 * 2 + 4) * (n - 1) ** 2)

# far-field surface sums: the job list of a 360x180 AzimuthalMap over two z faces (Metasurface_Farfield layout)
m = min(n - 2, 256)
ph = capi.Phaser(e, [1, m, 1, m, n // 4, 3 * n // 4], 0x30, 1)
ph.accumulate(np.array([0.6 + 0.8j], np.complex64), np.array([0.8 - 0.6j], np.complex64))
ndir = 4096
rng = np.random.RandomState(1)
W = (rng.randn(ndir * 2 * m) + 1j * rng.randn(ndir * 2 * m)).astype(np.complex64)
jobs = [(4 + (q & 1), (q >> 1) & 3, (q // 8) * 2 * m, (q // 8) * 2 * m + m) for q in range(ndir * 8)]
timed(f"surface_sums ({ndir} directions x 2 faces x 4 terms, {m}x{m} faces)", lambda: ph.surface_sums(0, jobs, W), 8 * m * m * len(jobs), reps=2)

# geometry: one sphere + a perforated slab, subN = 12, on the whole box
objs = [("BALL", [[0.3 * n * 10, 0.6 * n * 10, 0], [0.3 * n * 10, 0.6 * n * 10, 0], [0.55 * n * 10, 0.85 * n * 10, 0]], []),
        ("DIFFERENCE", [[2, 0, 0], [0, 0, 0], [0, 0, 0]],
         [("BOX", [[-3e38, 3e38, 0], [-3e38, 3e38, 0], [0.4 * n * 10, 0.5 * n * 10, 0]], []),
          ("DLATTICE", [[420, 0, 0], [727.46, 0, 0], [0, 0, 0]], [("RODZ", [[-126, 126, 0], [-126, 126, 0], [-3e38, 3e38, 0]], [])])])]
g = capi.GeometryDesc(lib, objs, [0, 0, 0], [n + 1, n + 1, n + 1], 12, lambda a, p: np.float32(10.0) * np.float32(p))
b = timed("geometry_bounds (whole grid)", lambda: g.bounds([1, n, 1, n, 1, n]), 0, reps=2)
box = [int(v) for v in b[1]]
timed("geometry_fill (perforated slab box, subN=12)", lambda: g.fill(1, box), 12 * (box[1] - box[0] + 1) * (box[3] - box[2] + 1) * (box[5] - box[4] + 1), reps=2)
print(json.dumps(out, indent=1))
