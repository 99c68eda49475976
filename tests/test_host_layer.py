"""The host C layer (alis_b200/host: our own createWorld / putObjects / planewave / monitors ...)
validated WITHOUT a GPU: linked against the CPU oracle engine (tests/bin/libalis_oracle.so, test
infrastructure) it must reproduce the unmodified reference bit for bit -- the complete field
state of the 10 configurations (sha256 goldens) and every monitor value of monitors.c (far-field
values, whose surface sums are engine calls, to 1e-5: helpers.monitor_mismatches)."""
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
    got = helpers.run_monitors("monitors_mine_oracle", cfg)
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
