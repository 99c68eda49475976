"""The ordered boundary-copy list resolved into one gather (alis_b200/csrc/copy_resolve.h, kernel copy_resolved).

CPU part: tests/bin/rc_check (tests/native/rc_check.cpp, built by alis_b200/build_host.sh) applies random ordered copy
lists, and the periodic / hexagonal H-pass lists of reference update.c:311-339, once copy by copy as the reference does
(copyBoundary, update.c:182-191) and once through the resolved plan; both must agree bit for bit.  The same program checks
the TFSF strip lists regrouped per cell (alis_b200/csrc/strip_resolve.h, kernel strip_resolved) against the strips added
one after the other (update(), update.c:193-202).
GPU part: worlds whose passes end in such lists step identically with the resolved kernel and with the level-by-level
list kernels (ALISB_RESOLVE=0), and identically to the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RC = os.path.join(ROOT, "tests", "bin", "rc_check")


def test_resolved_plan_equals_ordered_copies_on_cpu():
    assert os.path.exists(RC), "tests/bin/rc_check is missing: run __graft_entry__.build()"
    r = subprocess.run([RC, "400"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    word, cases, resolved, refused, _, bound, _, entries, _, multi, _, cells = r.stdout.split()
    assert word == "ok" and int(cases) == 804 and int(resolved) + int(refused) == int(cases)
    assert int(resolved) >= 700          # the walks must not give up on ordinary lists
    assert 0 < int(bound) < int(entries)  # both kinds of entries (ordered among themselves / free) were exercised
    assert 0 < int(multi) < int(cells)    # strip cells with one and with several ordered terms


class BlochBox:
    """a BoxWorld whose passes end in the ordered copy lists of a world that is Bloch-periodic along x, y and z
    (update.c:284-297 and 311-339 with iAlt = iNUM-1 etc.) or hexagonal along y (update.c:289-292, 322-331)"""

    def __new__(cls, shape, hexagonal=False, complex_field=1):
        from alis_b200 import synthetic
        w = synthetic.BoxWorld(*shape, num=4, complex_field=complex_field)
        nx, ny, nz = shape
        X, Y, Z = 0, 1, 2
        cx, sx, cy, sy, cz, sz = 0.8, 0.6, 0.28, 0.96, -0.6, 0.8
        cyp, syp, cym, sym = 0.96, 0.28, -0.6, 0.8
        ia, ja, ka = nx - 1, ny - 1, nz - 1
        half = nx // 2
        E = [(0, X, cx, -sx, [nx + 1, nx + 1, 1, ny, 1, nz], [-ia, 0, 0]),
             (0, Z, cz, -sz, [1, nx, 1, ny, nz + 1, nz + 1], [0, 0, -ka])]
        if hexagonal:
            E += [(0, Y, cym, -sym, [1, half, ny + 1, ny + 1, 1, nz], [nx - half, 1 - ny, 0]),
                  (0, Y, cyp, -syp, [half + 1, nx + 1, ny + 1, ny + 1, 1, nz], [-half, 1 - ny, 0])]
        else:
            E += [(0, Y, cy, -sy, [1, nx, ny + 1, ny + 1, 1, nz], [0, -ja, 0])]
        H = [(1, Y, cx, -sx, [nx + 1, nx + 1, 1, ny + 1, 1, nz + 1], [-ia, 0, 0]),
             (1, Z, cx, -sx, [nx + 1, nx + 1, 1, ny + 1, 1, nz + 1], [-ia, 0, 0]),
             (1, X, cz, -sz, [1, nx + 1, 1, ny + 1, nz + 1, nz + 1], [0, 0, -ka]),
             (1, Y, cz, -sz, [1, nx + 1, 1, ny + 1, nz + 1, nz + 1], [0, 0, -ka]),
             (1, Y, cx, sx, [1, 1, 1, ny + 1, 1, nz + 1], [ia, 0, 0]),
             (1, Z, cx, sx, [1, 1, 1, ny + 1, 1, nz + 1], [ia, 0, 0]),
             (1, X, cz, sz, [1, nx + 1, 1, ny + 1, 1, 1], [0, 0, ka]),
             (1, Y, cz, sz, [1, nx + 1, 1, ny + 1, 1, 1], [0, 0, ka])]
        if hexagonal:
            for comp in (X, Z):
                H += [(1, comp, cyp, syp, [1, half, 1, 1, 1, nz], [nx - half, ny - 1, 0])]
            for comp in (X, Z):
                H += [(1, comp, cym, -sym, [1, half, ny + 1, ny + 1, 1, nz], [nx - half, 1 - ny, 0])]
            for comp in (X, Z):
                H += [(1, comp, cym, sym, [half + 1, nx + 1, 1, 1, 1, nz], [-half, ny - 1, 0])]
            for comp in (X, Z):
                H += [(1, comp, cyp, -syp, [half + 1, nx + 1, ny + 1, ny + 1, 1, nz], [-half, 1 - ny, 0])]
        else:
            H += [(1, Z, cy, -sy, [1, nx + 1, ny + 1, ny + 1, 1, nz + 1], [0, -ja, 0]),
                  (1, X, cy, -sy, [1, nx + 1, ny + 1, ny + 1, 1, nz + 1], [0, -ja, 0]),
                  (1, Z, cy, sy, [1, nx + 1, 1, 1, 1, nz + 1], [0, ja, 0]),
                  (1, X, cy, sy, [1, nx + 1, 1, 1, 1, nz + 1], [0, ja, 0])]
        w.copies = lambda: (E, H)
        return w


@pytest.mark.gpu
@pytest.mark.parametrize("shape,hexagonal,complex_field", [
    ((40, 36, 32), False, 1),
    ((40, 36, 32), True, 1),
    ((51, 29, 33), False, 0),        # real fields: plain factors instead of rotations
    ((150, 140, 130), False, 1),     # more entries than a thread keeps in registers: the staged variant of the kernel
])
def test_resolved_copies_step_like_ordered_copies(oracle_lib, engine_lib, monkeypatch, shape, hexagonal, complex_field):
    from tests import helpers
    w = BlochBox(shape, hexagonal, complex_field)
    g, o = w.make_engine(engine_lib), w.make_engine(oracle_lib)
    monkeypatch.setenv("ALISB_RESOLVE", "0")
    g0 = w.make_engine(engine_lib)                    # level-by-level list kernels
    monkeypatch.delenv("ALISB_RESOLVE")
    helpers.seed_fields([g, o, g0], tuple(n + 2 for n in shape), complex_field=complex_field)
    for e in (g, o, g0):
        e.run(12)
    helpers.assert_fields_identical(g, o, f"resolved copies vs oracle {shape}")
    helpers.assert_fields_identical(g0, o, f"list kernels vs oracle {shape}")
    assert g.launch_count() <= g0.launch_count()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["nanoblock", "tfsf_normal", "metasurface", "phc", "focused_tfsf"])
@pytest.mark.parametrize("knobs", [
    {"ALISB_RESOLVE": "0"},                                           # the level-by-level list kernels (copy_list, strip_list)
    {"ALISB_MAILBOX": "0", "ALISB_AUX_ASIDE": "0", "ALISB_CONC": "0"},  # copies + stream sync, auxiliary worlds in line, no side streams
], ids=["list-kernels", "in-line"])
def test_named_configs_with_the_launch_bound_knobs_off(cfg, knobs, tmp_path, monkeypatch):
    """the default path (resolved copies and strips, mailbox, auxiliary worlds on a side stream) is what every other GPU test
    runs; here the same reference-built worlds step through the paths those replaced and must give the same golden state"""
    from tests import helpers
    if not helpers.have_ref_bin("stepdump_bridge_cuda"):
        pytest.skip("oracle/_ref harness not built")
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    out = str(tmp_path / "g.bin")
    helpers.run_stepdump("stepdump_bridge_cuda", cfg, out, cwd=str(tmp_path))
    assert helpers.sha256_file(out) == helpers.golden_hashes()[cfg]["sha256"], f"{cfg} with {knobs} differs from the reference golden"
