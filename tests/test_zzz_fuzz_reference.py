"""Randomised differential test against the UNMODIFIED reference (oracle/_ref/bin/stepdump_ref):
worlds drawn from a seed -- cell sizes, time step, flat axes, boundary types per axis (PML depths
incl. true CPML, mirror, PEC, periodic, Bloch/complex), graded grids, sub-pixel sampling, up to three
objects of dielectric / dispersive / anisotropic matter, a normal-incidence TFSF plane wave in every
fourth world, one or two dipoles (oracle/harness/configs.h,
cfg_fuzz) -- are stepped by the reference and by this repo; the complete field state must be
identical to the bit.  No goldens: both sides run live, so the reference build must be present
(it travels to the GPU box with the snapshot).

CPU: own host layer over the oracle engine (both rasterisers) and reference set-up + oracle engine
through the bridge.  GPU: own host layer + CUDA engine, reference set-up + CUDA engine."""
import os
import subprocess

import pytest

from tests import helpers

REF = os.path.join(helpers.REF_BIN, "stepdump_ref")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="reference twin not built (oracle/build_ref.sh needs /root/reference)")
CPU_SEEDS = list(range(1, 121))
# (kept late in the alphabet: new GPU coverage runs after the established suite)
# 3-D CPML, anisotropic cells, a flat x axis, Bloch-periodic z (15, 58: no golden configuration has a periodic z), TFSF (with and without objects, Num 24), flat axis + graded grid, Bloch/complex
# with mirror and PEC sides, three objects, graded 3-D grid
GPU_SEEDS = [5, 7, 8, 9, 15, 16, 20, 23, 29, 32, 41, 58, 59, 67, 76]


def dump(exe, seed, out, cwd, **env):
    e = dict(os.environ, FUZZ_SEED=str(seed), **env)
    r = subprocess.run([exe, "fuzz", "0", out], cwd=cwd, env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, f"{os.path.basename(exe)} seed {seed}: rc {r.returncode}\n{r.stdout[-800:]}\n{r.stderr[-800:]}"
    if not os.path.exists(out):
        # a drawn object that owns no grid node: both libraries print the reference's message and exit(0)
        assert "is zero!" in r.stdout, r.stdout[-400:]
        return "volume-of-object-is-zero", "no dump"
    return helpers.sha256_file(out), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""


@needs_ref
@pytest.mark.parametrize("seed", CPU_SEEDS)
def test_random_world_matches_reference_cpu(seed, tmp_path):
    if not helpers.have_test_bin("stepdump_mine_oracle"):
        pytest.skip("tests/bin not built")
    t = str(tmp_path)
    want, what = dump(REF, seed, os.path.join(t, "r.bin"), t, OMP_NUM_THREADS="2")
    got, _ = dump(os.path.join(helpers.TEST_BIN, "stepdump_mine_oracle"), seed, os.path.join(t, "m.bin"), t)
    assert got == want, f"seed {seed}: own host layer (engine rasteriser) differs from the reference [{what}]"
    got, _ = dump(os.path.join(helpers.TEST_BIN, "stepdump_mine_oracle"), seed, os.path.join(t, "h.bin"), t, ALIS_DEVICE_GEOMETRY="0")
    assert got == want, f"seed {seed}: own host layer (host rasteriser) differs from the reference [{what}]"
    got, _ = dump(os.path.join(helpers.REF_BIN, "stepdump_bridge_oracle"), seed, os.path.join(t, "b.bin"), t)
    assert got == want, f"seed {seed}: reference set-up + oracle engine differs from the reference [{what}]"


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("seed", GPU_SEEDS)
def test_random_world_matches_reference_gpu(seed, tmp_path):
    t = str(tmp_path)
    want, what = dump(REF, seed, os.path.join(t, "r.bin"), t)
    got, _ = dump(os.path.join(helpers.TEST_BIN, "stepdump_mine_cuda"), seed, os.path.join(t, "m.bin"), t)
    assert got == want, f"seed {seed}: drop-in library on the GPU differs from the reference [{what}]"
    got, _ = dump(os.path.join(helpers.REF_BIN, "stepdump_bridge_cuda"), seed, os.path.join(t, "b.bin"), t)
    assert got == want, f"seed {seed}: reference set-up + CUDA engine differs from the reference [{what}]"


@needs_ref
@pytest.mark.parametrize("seed", list(range(1, 41)))
def test_random_world_cut_into_z_slabs_cpu(seed, tmp_path):
    """the same worlds cut into three z slabs behind the public API (ALIS_DEVICES, host/shard.c): flat and thin z axes get
    fewer slabs, worlds periodic along z run uncut -- the field state is the reference's either way (160 seeds x {2, 3}
    slabs were identical when this was written)"""
    if not helpers.have_test_bin("stepdump_mine_oracle"):
        pytest.skip("tests/bin not built")
    t = str(tmp_path)
    want, what = dump(REF, seed, os.path.join(t, "r.bin"), t, OMP_NUM_THREADS="2")
    got, _ = dump(os.path.join(helpers.TEST_BIN, "stepdump_mine_oracle"), seed, os.path.join(t, "m.bin"), t, ALIS_DEVICES="0,0,0")
    assert got == want, f"seed {seed}: world cut into three slabs differs from the reference [{what}]"


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("seed", [5, 8, 16, 23, 41, 59, 67, 76])
def test_random_world_cut_into_z_slabs_gpu(seed, tmp_path):
    t = str(tmp_path)
    want, what = dump(REF, seed, os.path.join(t, "r.bin"), t)
    got, _ = dump(os.path.join(helpers.TEST_BIN, "stepdump_mine_cuda"), seed, os.path.join(t, "m.bin"), t, ALIS_DEVICES="0,0")
    assert got == want, f"seed {seed}: world cut into two slabs on the GPU differs from the reference [{what}]"


MONITOR_SEEDS_CPU = list(range(1, 25))
MONITOR_SEEDS_GPU = [2, 9, 14, 23]


@needs_ref
@pytest.mark.parametrize("seed", MONITOR_SEEDS_CPU)
def test_random_world_monitors_match_reference_cpu(seed):
    """the monitor harness (about 400 values: every field function, probes, flux boxes, world and slice
    reductions, slice checksums, time-averaged and frequency-domain slices, a phaser list on the
    world's absorbing faces with phasor flux, far field and direct array access) on random 3-D worlds
    large enough for its fixed probe positions (FUZZ_BIG): identical to the printed digit, far field
    within 1e-5"""
    if not helpers.have_test_bin("monitors_mine_oracle"):
        pytest.skip("tests/bin not built")
    env = {"FUZZ_SEED": str(seed), "FUZZ_BIG": "1"}
    want = helpers.run_monitors("monitors_ref", "fuzz", threads=1, extra_env=env)
    got = helpers.run_monitors("monitors_mine_oracle", "fuzz", extra_env=env, direct_phasor=True)
    assert len(want) > 350
    assert not helpers.monitor_mismatches(got, want), f"seed {seed}"


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("seed", MONITOR_SEEDS_GPU)
def test_random_world_monitors_match_reference_gpu(seed):
    env = {"FUZZ_SEED": str(seed), "FUZZ_BIG": "1"}
    want = helpers.run_monitors("monitors_ref", "fuzz", threads=1, extra_env=env)
    got = helpers.run_monitors("monitors_mine_cuda", "fuzz", extra_env=env)
    assert len(want) > 350
    assert not helpers.monitor_mismatches(got, want, device_flux=True), f"seed {seed}"


@pytest.mark.gpu
@pytest.mark.parametrize("nfreq", [9, 96, 131], ids=["9freq", "96freq", "131freq-two-launches"])
def test_parallel_phaser_kernel_bit_exact(nfreq):
    """kernel phaser_accumulate: every face and frequency of one updatePhaser in one launch, one thread per
    (face cell, frequency), phase factors as kernel arguments; lists longer than 96 frequencies take one
    launch per 96.  Accumulators identical to the oracle's loops (reference phaser.c:83-110)"""
    import numpy as np
    from alis_b200 import capi, synthetic
    eng, ora = capi.load_engine(), capi.load_oracle()
    w = synthetic.BoxWorld(48, 44, 40, num=6)
    g, o = w.make_engine(eng), w.make_engine(ora)
    helpers.seed_fields([g, o], (50, 46, 42), seed=7)
    freqs = np.linspace(1 / 800.0, 1 / 400.0, nfreq).astype(np.float32)
    box = [8, 41, 8, 37, 8, 33]
    pg, po = capi.Phaser(g, box, 0x3F, len(freqs)), capi.Phaser(o, box, 0x3F, len(freqs))
    for _ in range(8):
        for e in (g, o):
            e.step_E()
            e.step_H()
        _, _, tE, tH = g.time()
        wE = np.exp(2j * np.pi * freqs.astype(np.float64) * tE).astype(np.complex64)
        wH = np.exp(2j * np.pi * freqs.astype(np.float64) * tH).astype(np.complex64)
        pg.accumulate(wE, wH)
        po.accumulate(wE, wH)
    for f in sorted({0, 1, nfreq//2, 95 % nfreq, 96 % nfreq, nfreq - 1}):
        for face in range(6):
            for which in range(4):
                a, b = pg.read(f, face, which), po.read(f, face, which)
                assert np.array_equal(a.view(np.float32), b.view(np.float32)), (f, face, which)
                assert np.abs(a).max() > 0


# ---- the GPU paths that have never run on a GPU yet come last ----------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["metasurface", "dipole_band"])
def test_phasor_guard_on_gpu(cfg):
    """direct user access to the phasor arrays with the accumulators on the GPU: the SIGSEGV handler
    pulls them from the device (kept at the end of the suite: it is the one path that calls CUDA from
    a signal handler)"""
    want = helpers.golden_monitors()[cfg]
    for mode in (True, "omp"):         # "omp": eight threads fault on the protected arrays at once
        got = helpers.run_monitors("monitors_mine_cuda", cfg, direct_phasor=mode)
        assert not helpers.monitor_mismatches(got, want, device_flux=True), mode


@pytest.mark.gpu
def test_remaining_public_api_on_gpu(tmp_path):
    """cloneWorld, rebootWorld, removeSources, removeObjects + new objects ... (oracle/harness/api_misc.c)
    with the CUDA engine: every value as the unmodified reference prints it (flux values reduced on
    the device to 5e-6)"""
    got, want = helpers.run_api_misc("api_misc_mine_cuda", str(tmp_path)), helpers.golden_api_misc()
    assert not helpers.monitor_mismatches(got, want, device_flux=True)


@pytest.mark.gpu
def test_graded_grid_slices_on_gpu(tmp_path):
    """slices of a world graded along x, y, z, fields from the CUDA engine, resampled onto the uniform
    raster by the host layer: files identical (sha256) to the unmodified reference's"""
    assert helpers.run_graded_slices("graded_slices_mine_cuda", str(tmp_path)) == helpers.golden_graded_slices()
