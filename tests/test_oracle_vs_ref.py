"""Pins the CPU oracle (oracle/alis_oracle.c) to the unmodified reference.

(1) golden: the full field state produced by the oracle engine (driven through the bridge from
    a reference-built world) must hash to the value recorded from the reference itself
    (tests/golden/stepdump_hashes.json, made by tests/golden/make_golden.py);
(2) live: when oracle/_ref holds the reference build, rerun the reference and compare dumps.
Bit-exact (sha256 of the raw fp32 dump)."""
import os

import pytest

from tests import helpers

needs_bridge = pytest.mark.skipif(not helpers.have_ref_bin("stepdump_bridge_oracle"),
                                  reason="oracle/_ref harness not built (needs /root/reference at build time)")


@needs_bridge
@pytest.mark.parametrize("cfg", helpers.CONFIGS)
def test_oracle_matches_golden(cfg, tmp_path):
    out = str(tmp_path / "o.bin")
    helpers.run_stepdump("stepdump_bridge_oracle", cfg, out, cwd=str(tmp_path))
    gold = helpers.golden_hashes()[cfg]
    assert helpers.sha256_file(out) == gold["sha256"], f"{cfg}: oracle field state differs from the reference golden"


@needs_bridge
@pytest.mark.parametrize("cfg", ["dipole_pml", "nanoblock", "phc"])
def test_reference_rerun_matches_golden(cfg, tmp_path):
    """the reference itself is deterministic across runs/thread counts"""
    out = str(tmp_path / "r.bin")
    os.environ["OMP_NUM_THREADS"] = "3"
    try:
        helpers.run_stepdump("stepdump_ref", cfg, out, cwd=str(tmp_path))
    finally:
        os.environ.pop("OMP_NUM_THREADS", None)
    assert helpers.sha256_file(out) == helpers.golden_hashes()[cfg]["sha256"]
