"""The host C layer (alis_b200/host: our own createWorld / putObjects / planewave / monitors ...)
validated WITHOUT a GPU: linked against the CPU oracle engine (tests/bin/libalis_oracle.so, test
infrastructure) it must reproduce the unmodified reference bit for bit -- the complete field
state of the 10 configurations (sha256 goldens) and every monitor value of monitors.c (far-field
values, whose surface sums are engine calls, to 1e-5: helpers.monitor_mismatches)."""
import os
import subprocess

import pytest

from tests import helpers

needs_bins = pytest.mark.skipif(not helpers.have_test_bin("stepdump_mine_oracle"), reason="tests/bin not built (run __graft_entry__.build())")


@needs_bins
@pytest.mark.parametrize("cfg", helpers.CONFIGS)
def test_own_setup_reproduces_reference_fields(cfg, tmp_path):
    out = str(tmp_path / "m.bin")
    helpers.run_mine_stepdump("stepdump_mine_oracle", cfg, out, cwd=str(tmp_path))
    assert helpers.sha256_file(out) == helpers.golden_hashes()[cfg]["sha256"], f"{cfg}: own host layer + oracle engine differs from the reference"


@needs_bins
@pytest.mark.parametrize("cfg", helpers.MONITOR_CONFIGS)
def test_monitor_values_equal_reference(cfg):
    got = helpers.run_monitors("monitors_mine_oracle", cfg, direct_phasor=True)
    want = helpers.golden_monitors()[cfg]
    assert len(want) > 35
    bad = helpers.monitor_mismatches(got, want)
    assert not bad, f"{cfg}: {len(bad)} monitor values differ from the reference: {bad[:5]}"


@needs_bins
@pytest.mark.parametrize("cfg", ["nanoblock", "dipole_pml", "phc"])
def test_device_monitors_equal_host_view_monitors(cfg):
    """slices / world reductions / absorption evaluated by the engine calls (row f1) against the
    same monitors evaluated through the host views: identical to the printed digit"""
    dev = helpers.run_monitors("monitors_mine_oracle", cfg)
    host = helpers.run_monitors("monitors_mine_oracle", cfg, host_monitors=True)
    assert len(dev) > 250 and dev == host


@needs_bins
@pytest.mark.parametrize("cfg", ["metasurface", "dipole_band"])
def test_phasor_arrays_are_coherent_for_direct_user_access(cfg):
    """user code indexing P->zMax.Ex[i][j][k] itself (reference examples/Cross_Polarization_T.c:3-19):
    the page-protection guard refreshes the host views on first touch -- same values as with an
    explicit alisSyncPhaser and as the reference; with the guard off the views are stale (zero)"""
    want = {w[0]: w[1] for w in helpers.golden_monitors()[cfg]}
    for direct in (True, "omp", False):        # "omp": eight threads touch the protected arrays at the same time
        got = {g[0]: g[1] for g in helpers.run_monitors("monitors_mine_oracle", cfg, direct_phasor=direct)}
        for tag in ("phasor_direct_mid", "phasor_direct_end"):
            assert got[tag] == want[tag] and float(want[tag]) != 0, (tag, direct)
    import os
    import subprocess
    env = dict(os.environ, HARNESS_DIRECT_PHASOR="1", ALIS_PHASER_GUARD="0")
    r = subprocess.run([os.path.join(helpers.TEST_BIN, "monitors_mine_oracle"), cfg, "0"], env=env, capture_output=True, text=True, timeout=900)
    stale = dict(line.split() for line in r.stdout.splitlines() if line.startswith("phasor_direct"))
    assert float(stale["phasor_direct_mid"]) == 0 and float(stale["phasor_direct_end"]) == 0


@needs_bins
def test_remaining_public_api_equals_reference(tmp_path):
    """cloneWorld, rebootWorld, removeSources, removeObjects + new objects, LT, nkFile, phaserPoynting*,
    deleteCloneWorld / deletePhasers / deleteWorld (oracle/harness/api_misc.c): every value as the
    unmodified reference prints it"""
    got, want = helpers.run_api_misc("api_misc_mine_oracle", str(tmp_path)), helpers.golden_api_misc()
    assert len(want) >= 28 and got == want, [(g, w) for g, w in zip(got, want) if g != w][:5]


def test_graded_grid_slices_match_reference_golden(tmp_path):
    """oracle/harness/graded_slices.c with the host layer on the oracle engine: every slice file (1-D, 2-D
    text incl. the refractive-index map, 3-D HDF5) of a world graded along x, y, z has the sha256 of the
    file the unmodified reference wrote (tests/golden/graded_slices.json, made by make_golden.py)"""
    want = helpers.golden_graded_slices()
    assert len(want) == 12
    assert helpers.run_graded_slices("graded_slices_mine_oracle", str(tmp_path)) == want


def test_spectrum_worker_thread_gives_the_inline_spectrum(tmp_path):
    """writeSpectrum's running DFT on the worker thread (alis_b200/host/output.c, default with the CUDA engine) against the
    in-line loop of reference output.c:211-215: same float operations per bin in the same order, so the written spectrum must
    be byte-identical.  Run here on the CPU engine with the worker forced on."""
    import shutil
    exe = os.path.join(helpers.TEST_BIN, "phc_band_diagram_oracle")
    if not os.path.exists(exe):
        pytest.skip("tests/bin/phc_band_diagram_oracle not built")
    outs = []
    for mode in ("0", "1"):
        d = tmp_path / f"mode{mode}"
        d.mkdir()
        shutil.copy(exe, d / "phc_band_diagram")
        env = dict(os.environ)
        env["ALIS_SPECTRUM_THREAD"] = mode
        env["LD_LIBRARY_PATH"] = helpers.TEST_BIN + os.pathsep + env.get("LD_LIBRARY_PATH", "")
        env.setdefault("OMP_NUM_THREADS", "4")
        r = subprocess.run(["./phc_band_diagram", "1.2", "1600"], cwd=d, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        txt = sorted(p for p in d.rglob("*.txt"))
        assert txt, "no spectrum was written"
        outs.append([p.read_bytes() for p in txt])
    assert outs[0] == outs[1], "the worker thread's spectrum differs from the in-line one"
    assert len(outs[0][0].splitlines()) > 6
