"""One world on several engines behind the public C API (alis_b200/host/shard.c): ALIS_DEVICES
names the devices, the host layer cuts the world into z slabs -- grid, tables, CPML slabs, material blocks, ADE
boxes, sources, TFSF strips (auxiliary worlds replicated per slab), boundary copies -- and moves one plane of
tangential components per interface and half step (alisb_halo_exchange).

The complete field state after the harness run must hash to the golden of the UNMODIFIED reference
(tests/golden/stepdump_hashes.json), exactly as for the undivided world:
  CPU suite  host layer / reference set-up through the bridge + CPU oracle engine, 2 and 3 slabs
  GPU suite  the same with the CUDA engine: two and three slabs on ONE GPU (ALIS_DEVICES=0,0: runs on the
             single-GPU test box) and one slab per GPU where the box has several (peer copies over NVLink)."""
import os

import numpy as np
import pytest

from tests import helpers
from tests import test_refex

SHARDABLE = ["dipole_pml", "vacuum_cpml", "pec_box", "nanoblock", "tfsf_normal", "phc", "metasurface", "dipole_band",
             "aniso_ade", "bloch_wide", "guided_y", "guided_xz", "shape_zoo", "focused_beam", "focused_tfsf"]
QUICK = ["nanoblock", "tfsf_normal", "phc", "metasurface", "aniso_ade", "guided_xz", "dipole_pml", "focused_tfsf"]


def _check(binary, cfg, devices, tmp_path, monkeypatch):
    monkeypatch.setenv("ALIS_DEVICES", devices)
    out = str(tmp_path / "dump.bin")
    log = helpers.run_stepdump(binary, cfg, out, cwd=str(tmp_path))
    gold = helpers.golden_hashes()[cfg]
    if helpers.sha256_file(out) != gold["sha256"]:
        monkeypatch.delenv("ALIS_DEVICES")
        ref = str(tmp_path / "single.bin")
        helpers.run_stepdump(binary, cfg, ref, cwd=str(tmp_path))
        a, b = helpers.read_dump(out), helpers.read_dump(ref)
        bad = np.argwhere(a != b)
        raise AssertionError(f"{cfg} on devices {devices}: {len(bad)} values differ from the undivided run, first at [m,set,comp,a,b,c]={bad[:1].tolist()}, "
                             f"rel L2 {helpers.rel_l2(a, b):.3e}; undivided run matches the golden: {helpers.sha256_file(ref) == gold['sha256']}\n{log[-400:]}")


def _bin(name):
    return os.path.exists(os.path.join(helpers.TEST_BIN, name)) or os.path.exists(os.path.join(helpers.REF_BIN, name))


@pytest.mark.skipif(not _bin("stepdump_mine_oracle"), reason="host layer not built")
@pytest.mark.parametrize("devices", ["0,0", "0,0,0"])
@pytest.mark.parametrize("cfg", QUICK)
def test_z_slabs_match_the_reference_golden(cfg, devices, tmp_path, monkeypatch):
    _check(os.path.join(helpers.TEST_BIN, "stepdump_mine_oracle"), cfg, devices, tmp_path, monkeypatch)


@pytest.mark.skipif(not helpers.have_ref_bin("stepdump_bridge_oracle"), reason="oracle/_ref harness not built")
@pytest.mark.parametrize("cfg", ["nanoblock", "metasurface", "bloch_wide", "shape_zoo"])
def test_z_slabs_from_the_reference_setup(cfg, tmp_path, monkeypatch):
    """worlds set up by the UNMODIFIED reference, cut and stepped through the bridge"""
    _check("stepdump_bridge_oracle", cfg, "0,0,0", tmp_path, monkeypatch)


@pytest.mark.skipif(not _bin("stepdump_mine_oracle"), reason="host layer not built")
def test_what_is_not_cut(tmp_path, monkeypatch):
    """a world periodic along z runs undivided (with a notice, still the golden); chi2 mixing between worlds is refused"""
    monkeypatch.setenv("ALIS_DEVICES", "0,0")
    out = str(tmp_path / "dump.bin")
    exe = os.path.join(helpers.TEST_BIN, "stepdump_mine_oracle")
    helpers.run_stepdump(exe, "pbc_slab", out, cwd=str(tmp_path))
    assert helpers.sha256_file(out) == helpers.golden_hashes()["pbc_slab"]["sha256"]
    with pytest.raises(RuntimeError, match="not available on a world cut into z slabs"):
        helpers.run_stepdump(exe, "chi2_shg", out, cwd=str(tmp_path))


@test_refex.needs_bins
def test_reference_driver_on_two_slabs(tmp_path, monkeypatch):
    """the reference's own Dipole_PML driver (225 PNG frames from sliceSnap) on a world cut in two: the monitors
    gather their regions from both slabs through the host views"""
    monkeypatch.setenv("ALIS_DEVICES", "0,0")
    _, n = test_refex._twin("Dipole_PML", "oracle", tmp_path, 900)
    assert n["png"] > 10


@pytest.mark.skipif(not _bin("monitors_mine_oracle"), reason="host layer not built")
@pytest.mark.parametrize("cfg", ["metasurface", "nanoblock", "dipole_band"])
def test_monitors_on_three_slabs_equal_reference(cfg):
    """every monitor of oracle/harness/monitors.c (~400 values: field functions, probes, flux boxes, absorption, world and
    slice reductions, time-averaged and frequency-domain slices, phasor flux, far field, direct phasor access) on a world cut
    into three slabs against the UNMODIFIED reference: regions are gathered from all slabs, phasor faces are cut along z and
    stitched, far-field surface sums are added over the slabs.  The flux integrals are summed on the host in double for cut
    worlds (tolerance 5e-6, as for the device reduction); everything else must agree to the printed digit."""
    got = helpers.run_monitors("monitors_mine_oracle", cfg, direct_phasor=True, extra_env={"ALIS_DEVICES": "0,0,0"})
    want = helpers.golden_monitors()[cfg]
    bad = helpers.monitor_mismatches(got, want, device_flux=True)
    assert not bad, f"{cfg}: {len(bad)} monitor values differ from the reference: {bad[:5]}"


# ---- CUDA engine ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("cfg,devices", [(c, "0,0") for c in SHARDABLE] + [(c, "0,0,0") for c in QUICK])
def test_z_slabs_on_one_gpu(cfg, devices, tmp_path, monkeypatch):
    """several slabs on ONE GPU: the whole decomposition and the peer copies, on the single-GPU test box"""
    _check(os.path.join(helpers.TEST_BIN, "stepdump_mine_cuda"), cfg, devices, tmp_path, monkeypatch)


@pytest.mark.gpu
@pytest.mark.skipif(not helpers.have_ref_bin("stepdump_bridge_cuda"), reason="oracle/_ref harness not built")
@pytest.mark.parametrize("cfg", ["nanoblock", "metasurface", "bloch_wide"])
def test_z_slabs_from_the_reference_setup_on_the_gpu(cfg, tmp_path, monkeypatch):
    _check("stepdump_bridge_cuda", cfg, "0,0", tmp_path, monkeypatch)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["nanoblock", "phc", "metasurface", "tfsf_normal", "aniso_ade", "vacuum_cpml"])
def test_one_slab_per_gpu(cfg, tmp_path, monkeypatch):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    _check(os.path.join(helpers.TEST_BIN, "stepdump_mine_cuda"), cfg, ",".join(str(d) for d in range(min(n, 4))), tmp_path, monkeypatch)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["Nanoblock_Scattering", "Metasurface_Farfield"])
def test_reference_driver_on_two_slabs_on_the_gpu(name, tmp_path, monkeypatch):
    """the reference's own drivers with their full monitor sets (flux boxes, slices; running DFT on cut faces and the
    far-field map) on a world cut in two"""
    if not os.path.exists(test_refex._exe(name, "ref")):
        pytest.skip("reference twin not built")
    monkeypatch.setenv("ALIS_DEVICES", "0,0")
    test_refex._twin(name, "cuda", tmp_path, 1500)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", helpers.MONITOR_CONFIGS)
def test_monitors_on_two_slabs_on_the_gpu(cfg):
    got = helpers.run_monitors("monitors_mine_cuda", cfg, direct_phasor=True, extra_env={"ALIS_DEVICES": "0,0"})
    want = helpers.golden_monitors()[cfg]
    bad = helpers.monitor_mismatches(got, want, device_flux=True)
    assert not bad, f"{cfg}: {len(bad)} monitor values differ from the reference: {bad[:5]}"
