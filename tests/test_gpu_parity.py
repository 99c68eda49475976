"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle on the same
seeded inputs, bit for bit; against the reference goldens for the named configurations; and
size-independent properties at larger sizes.  All integer/float results are compared exactly:
the kernels are compiled without FMA contraction and use the reference's expression order."""
import ctypes as C
import os

import numpy as np
import pytest

from alis_b200 import capi, synthetic
from tests import helpers

pytestmark = pytest.mark.gpu


def both(oracle_lib, engine_lib, shape, **kw):
    w = synthetic.BoxWorld(*shape, **kw)
    return w, w.make_engine(engine_lib), w.make_engine(oracle_lib)


def compare_psi(g, o):
    for s in range(2):
        for ax in range(3):
            for sd in range(2):
                f, l = g.grid.pml[s][ax][sd][0], g.grid.pml[s][ax][sd][1]
                if f > l:
                    continue
                box = [1, g.nx, 1, g.ny, 1, g.nz]
                box[2 * ax], box[2 * ax + 1] = f, l
                for w in range(2):
                    a = g.read_box(capi.ARR_PSI, [s, 0, ax, sd, w], box)
                    b = o.read_box(capi.ARR_PSI, [s, 0, ax, sd, w], box)
                    assert np.array_equal(a, b), f"psi set={s} axis={ax} side={sd} which={w}"


@pytest.mark.parametrize("shape,num,steps", [
    ((76, 64, 56), 12, 40),      # the synthetic config at the golden size
    ((37, 41, 45), 5, 25),       # odd sizes, ragged quads
    ((131, 33, 40), 8, 20),      # two x tiles, ragged tail
    ((260, 30, 29), 6, 12),      # three x tiles
    ((64, 64, 64), 24, 30),      # Num=24 -> true CPML (kappa 10, alpha 0.01)
])
def test_box_world_seeded_bit_exact(oracle_lib, engine_lib, shape, num, steps):
    w, g, o = both(oracle_lib, engine_lib, shape, num=num)
    helpers.seed_fields([g, o], tuple(n + 2 for n in shape))
    for e in (g, o):
        e.run(steps)
    assert g.time() == o.time()
    helpers.assert_fields_identical(g, o, f"box {shape}")
    compare_psi(g, o)


def test_golden_vacuum_cpml_on_gpu(engine_lib):
    import hashlib
    w = synthetic.BoxWorld(76, 64, 56)
    e = w.make_engine(engine_lib)
    e.run(60)
    hdr = b"ALISDUMP" + np.array([76, 64, 56, 0], np.int32).tobytes()
    got = hashlib.sha256(hdr + e.read_all_fields().tobytes()).hexdigest()
    assert got == helpers.golden_hashes()["vacuum_cpml"]["sha256"]
    assert e.launch_count() > 0


@pytest.mark.skipif(not helpers.have_ref_bin("stepdump_bridge_cuda"), reason="oracle/_ref harness not built")
@pytest.mark.parametrize("cfg", helpers.CONFIGS)
def test_named_configs_match_reference_golden(cfg, tmp_path):
    """reference-built world (unmodified reference setup code) stepped by the CUDA engine"""
    out = str(tmp_path / "g.bin")
    helpers.run_stepdump("stepdump_bridge_cuda", cfg, out, cwd=str(tmp_path))
    gold = helpers.golden_hashes()[cfg]
    if helpers.sha256_file(out) != gold["sha256"]:
        ref = str(tmp_path / "o.bin")
        helpers.run_stepdump("stepdump_bridge_oracle", cfg, ref, cwd=str(tmp_path))
        a, b = helpers.read_dump(out), helpers.read_dump(ref)
        bad = np.argwhere(a != b)
        raise AssertionError(f"{cfg}: {len(bad)} values differ from the oracle, first {bad[:1].tolist()}, rel L2 {helpers.rel_l2(a, b):.3e}")


def test_read_write_box_roundtrip(engine_lib):
    w = synthetic.BoxWorld(45, 38, 33, num=3)
    e = w.make_engine(engine_lib, with_source=False)
    rng = np.random.RandomState(1)
    full = rng.rand(47, 40, 35).astype(np.float32)
    e.write_field(1, 2, 0, full)
    assert np.array_equal(e.read_field(1, 2, 0), full)
    box = [3, 9, 0, 39, 5, 5]
    sub = rng.rand(7, 40, 1).astype(np.float32)
    e.write_box(capi.ARR_FIELD, [1, 2, 0], box, sub)
    full[3:10, :, 5:6] = sub
    assert np.array_equal(e.read_field(1, 2, 0), full)
    assert np.array_equal(e.read_box(capi.ARR_FIELD, [1, 2, 0], box), sub)


def test_linearity_at_scale(engine_lib):
    """size-independent property at a size the oracle cannot reach in seconds: scaling the
    source by 2 scales every field value by exactly 2 (power-of-two scaling is exact in fp32
    for normal numbers; only values that passed through the denormal range at the leading
    edge of the wave may differ, by ~1e-45 absolute)."""
    shape = (384, 320, 288)
    w1 = synthetic.BoxWorld(*shape, amplitude=1.0)
    w2 = synthetic.BoxWorld(*shape, amplitude=2.0)
    e1, e2 = w1.make_engine(engine_lib), w2.make_engine(engine_lib)
    for e in (e1, e2):
        e.run(90)
    k = w1.ax[2].arr(0)
    box = [1, shape[0], 1, shape[1], k - 2, k + 3]
    for s in range(2):
        for c in range(3):
            a, b = e1.read_box(capi.ARR_FIELD, [s, c, 0], box), e2.read_box(capi.ARR_FIELD, [s, c, 0], box)
            big = np.abs(b) > 1e-25
            assert big.sum() > 1000
            mism = (2 * a[big] != b[big]).mean()
            assert mism < 1e-4, mism
            assert np.abs(2 * a - b).max() <= 1e-6 * np.abs(b).max()
    assert np.abs(e1.read_box(capi.ARR_FIELD, [0, 2, 0], box)).max() > 0


def test_two_slabs_at_scale_equal_single(engine_lib):
    """sharded == single, bit for bit, on a grid with two x tiles per row and several z chunks"""
    _slab_check(engine_lib, (300, 260, 250), 12)


def _cudart():
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    raise RuntimeError("libcudart not found")


def test_two_slabs_on_one_gpu_equal_single(engine_lib):
    """the multi-GPU decomposition with both slabs on cuda:0 and device-to-device halo copies"""
    _slab_check(engine_lib, (70, 50, 62), 45)


def _slab_check(engine_lib, shape, steps, fuse=False):
    rt = _cudart()
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    w = synthetic.BoxWorld(*shape)
    single = w.make_engine(engine_lib)
    helpers.seed_fields([single], tuple(n + 2 for n in shape), seed=11)
    full = single.read_all_fields()
    parts = [w.make_engine(engine_lib, rank=r, world=2) for r in range(2)]
    for r, e in enumerate(parts):          # same seeded state, split along z (ghost planes included)
        c0, c1 = e.shard
        for s in range(2):
            for c in range(3):
                e.write_field(s, c, 0, np.ascontiguousarray(full[0, s, c, :, :, c0 - 1:c1 + 2]))
    single.run(steps)
    ref = single.read_all_fields()
    lo, hi = parts

    def xchg(set_, src, dst):
        for comp in range(2):
            sp, _, nb = src.halo_info(set_, 0, comp)
            _, rp, _ = dst.halo_info(set_, 0, comp)
            src.synchronize()
            dst.synchronize()
            assert rt.cudaMemcpy(rp, sp, nb, 3) == 0
    for _ in range(steps):
        for e in parts:
            e.step_E()
            if fuse:
                e.flush(True)          # issue the deferred E pass as the fused variant (the H pass follows)
        xchg(capi.SET_E, lo, hi)
        for e in parts:
            e.step_H()
        xchg(capi.SET_H, hi, lo)
    got = np.concatenate([e.read_all_fields()[..., 1:-1] for e in parts], axis=-1)
    assert np.array_equal(got, ref[..., 1:-1])
    assert np.abs(ref).max() > 0


def test_phaser_accumulation_bit_exact(oracle_lib, engine_lib):
    w, g, o = both(oracle_lib, engine_lib, (48, 44, 40), num=6)
    helpers.seed_fields([g, o], (50, 46, 42), seed=7)
    box = [8, 41, 8, 37, 8, 33]
    freqs = np.array([1 / 400.0, 1 / 520.0, 1 / 650.0], np.float32)
    pg, po = capi.Phaser(g, box, 0x3F, 3), capi.Phaser(o, box, 0x3F, 3)
    for step in range(12):
        for e in (g, o):
            e.step_E()
            e.step_H()
        _, _, tE, tH = g.time()
        wE = np.exp(2j * np.pi * freqs.astype(np.float64) * tE).astype(np.complex64)
        wH = np.exp(2j * np.pi * freqs.astype(np.float64) * tH).astype(np.complex64)
        pg.accumulate(wE, wH)
        po.accumulate(wE, wH)
    for f in range(3):
        for face in range(6):
            for which in range(4):
                a, b = pg.read(f, face, which), po.read(f, face, which)
                assert np.array_equal(a.view(np.float32), b.view(np.float32)), (f, face, which)
                assert np.abs(a).max() > 0


def test_split_pass_equals_whole_pass_on_gpu(oracle_lib, engine_lib):
    """the overlap split (send plane first, then the rest) is bit-identical to the whole pass, on
    the side of a shard that faces a neighbour (rank 0: E pass, last rank: H pass)"""
    w = synthetic.BoxWorld(140, 44, 50, num=6)
    for rank, split_set in ((0, capi.SET_E), (1, capi.SET_H)):
        g, o = w.make_engine(engine_lib, rank=rank, world=2), w.make_engine(oracle_lib, rank=rank, world=2)
        helpers.seed_fields([g, o], (142, 46, g.nz + 2), seed=9)
        for _ in range(15):
            o.step_E()
            o.step_H()
            for s in (capi.SET_E, capi.SET_H):
                if s == split_set:
                    g.step_part(s, capi.PART_SEND_PLANE)
                    g.step_part(s, capi.PART_REST)
                else:
                    g.step_part(s, capi.PART_ALL)
        assert g.time() == o.time()
        helpers.assert_fields_identical(g, o, f"split pass on GPU, rank {rank}")
        compare_psi(g, o)


@pytest.mark.parametrize("alias", [False, True])
def test_chi2_coupling_seeded_bit_exact(oracle_lib, engine_lib, alias):
    """coupleTWM (update.c:356-391) through the C ABI on seeded dense fields: every sweep, the float and
    the double products, a zero tensor entry (skipped sweep) and the signal-is-its-own-pump case."""
    shape = (45, 38, 41)
    rng = np.random.RandomState(7)
    box = [5, 33, 4, 30, 7, 36]
    dims = (box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1)
    fill = [rng.rand(*dims).astype(np.float32) for _ in range(3)]
    eps = [(0.2 + rng.rand(shape[0] + (c == 0), shape[1] + (c == 1), shape[2] + (c == 2))).astype(np.float32) for c in range(3)]
    d = (rng.rand(3, 6).astype(np.float32) - 0.5)
    d[1, 2] = 0.0
    zero = {k: [[0.0] * 8] * 3 for k in ("JJ", "JK", "JE", "KJ", "KK", "KE")}
    out = []
    for lib in (engine_lib, oracle_lib):
        w = synthetic.BoxWorld(*shape, num=6)
        sig, p1, p2 = (w.make_engine(lib) for _ in range(3))
        for q, e in enumerate((sig, p1, p2)):
            helpers.seed_fields([e], tuple(n + 2 for n in shape), seed=100 + q, scale=0.5)
        sig.set_inv_eps(*eps)
        sig.add_dispersive(box, 0, 0, [0.0, 0.0, 0.0], zero, fill)
        for _ in range(2):
            sig.chi2_couple(0, 0, sig if alias else p1, p2, d, 0.37)
        out.append(sig.read_all_fields())
        for e in (sig, p1, p2):
            e.close()
    assert np.array_equal(out[0], out[1])
    assert np.abs(out[0]).max() > 0


# ---- fused E+H stepping (temporal blocking): must be bit-identical to the two-pass oracle ----------
@pytest.fixture(params=[(0, 0), (0, 1), (1, 0)], ids=["out-of-place", "out-of-place-zslabs", "in-place-wavefront"])
def fusion_env(monkeypatch, request):
    """engines created under this fixture fuse whatever they can, however small the grid; both fused
    kernels: yee_fused (double-buffered fields, halo recomputed; also its opt-in variant that carries the z CPML
    slabs inside the fused box) and yee_wave (in place, tile wavefront)"""
    monkeypatch.setenv("ALISB_WAVE", str(request.param[0]))
    monkeypatch.setenv("ALISB_FUSE_PMZ", str(request.param[1]))
    monkeypatch.setenv("ALISB_FUSE", "1")
    monkeypatch.setenv("ALISB_FUSE_MIN", "1")
    monkeypatch.setenv("ALISB_SMALL", "1")
    monkeypatch.setenv("ALISB_FUSE_FRAC", "0.01")


_add_materials = helpers.add_materials


@pytest.mark.parametrize("shape,num,steps,materials", [
    ((160, 72, 64), 12, 30, False),    # the fused box is three 32-cell groups wide; the dipole's planes are excluded
    ((300, 61, 47), 6, 20, False),     # ragged rows (61-13 not a multiple of the tile), several x segments
    ((96, 40, 90), 8, 25, False),      # long z: several z-chunks per range
    ((160, 72, 64), 12, 30, True),     # material arrays and an ADE object inside the fused box
    ((96, 40, 90), 8, 25, True),
    ((96, 64, 76), 24, 30, False),     # true CPML (Num=24: kappa, alpha): more than half of the fused planes lie in z slabs
])
@pytest.mark.parametrize("rows", [8, 16])
def test_fused_stepping_bit_exact(oracle_lib, engine_lib, fusion_env, monkeypatch, shape, num, steps, materials, rows):
    monkeypatch.setenv("ALISB_FCHUNK", "7")
    monkeypatch.setenv("ALISB_FROWS", str(rows))
    w, g, o = both(oracle_lib, engine_lib, shape, num=num)
    if materials:
        _add_materials([g, o], shape)
    helpers.seed_fields([g, o], tuple(n + 2 for n in shape))
    for e in (g, o):
        e.run(steps)
    assert g.time() == o.time()
    assert g.fused_launch_count() >= steps, "the fused kernel did not run"
    helpers.assert_fields_identical(g, o, f"fused box {shape}")
    compare_psi(g, o)
    # interleaved access: reading between the passes forces the plain variant, results must not change
    for e in (g, o):
        for _ in range(3):
            e.step_E()
            e.read_box(capi.ARR_FIELD, [0, 2, 0], [1, 4, 1, 4, 1, 4])
            e.step_H()
        e.step_E()
        e.step_E()
        e.step_H()
        e.run(2)
    helpers.assert_fields_identical(g, o, f"fused box {shape} after interleaved access")
    compare_psi(g, o)


def test_fused_flush_then_access(oracle_lib, engine_lib, fusion_env):
    """alisb_flush(e, 1) issues the fused kernel (E' and, speculatively, H' and the H set's z-slab psi of the fused box);
    a field write before the H pass must make the H pass recompute all of it from the changed E"""
    shape = (160, 56, 44)
    in_place = os.environ.get("ALISB_WAVE") == "1"
    w, g, o = both(oracle_lib, engine_lib, shape, num=8)
    helpers.seed_fields([g, o], tuple(n + 2 for n in shape))
    rng = np.random.RandomState(5)
    patch = rng.rand(6, 5, 4).astype(np.float32) * 1e-3
    for e in (g, o):
        e.run(3)
        e.step_E()
        e.flush(True)
        if in_place and e is g:
            with pytest.raises(capi.EngineError):      # in-place stepping has already formed H': writes are refused here
                e.write_box(capi.ARR_FIELD, [0, 0, 0], [60, 65, 20, 24, 3, 6], patch)
        elif not in_place:
            e.write_box(capi.ARR_FIELD, [0, 0, 0], [60, 65, 20, 24, 3, 6], patch)      # inside the low z slab
        e.step_H()
        e.run(2)
        e.step_E()
        e.flush(True)
        e.read_box(capi.ARR_FIELD, [1, 2, 0], [1, 4, 1, 4, 1, 4])
        e.step_H()
        e.run(3)
    helpers.assert_fields_identical(g, o, "flush, then access, then H pass")
    compare_psi(g, o)


def test_fused_write_between_passes(oracle_lib, engine_lib, fusion_env):
    """a field write after the (deferred) E pass must reach the H pass of the same step"""
    shape = (160, 48, 40)
    w, g, o = both(oracle_lib, engine_lib, shape, num=6)
    helpers.seed_fields([g, o], tuple(n + 2 for n in shape))
    rng = np.random.RandomState(3)
    patch = rng.rand(6, 5, 4).astype(np.float32)
    for e in (g, o):
        e.run(3)
        e.step_E()
        e.write_box(capi.ARR_FIELD, [0, 1, 0], [60, 65, 20, 24, 18, 21], patch)
        e.step_H()
        e.run(4)
    helpers.assert_fields_identical(g, o, "write between passes")


def test_fused_two_slabs_equal_single(engine_lib, fusion_env):
    _slab_check(engine_lib, (160, 50, 62), 30, fuse=True)



@pytest.mark.skipif(not helpers.have_ref_bin("stepdump_bridge_cuda"), reason="oracle/_ref harness not built")
@pytest.mark.parametrize("cfg", helpers.CONFIGS)
def test_named_configs_with_forced_fusion(cfg, tmp_path, fusion_env):
    """the 14 reference configurations (TFSF, ADE, material arrays, Bloch-complex fields, SYM/PBC/BBC/HBC
    boundaries, chi2 mixing between worlds) with fused stepping forced on wherever a fusable box exists:
    the field state must still hash to the reference golden"""
    out = str(tmp_path / "g.bin")
    log = helpers.run_stepdump("stepdump_bridge_cuda", cfg, out, cwd=str(tmp_path))
    gold = helpers.golden_hashes()[cfg]
    if helpers.sha256_file(out) != gold["sha256"]:
        ref = str(tmp_path / "o.bin")
        helpers.run_stepdump("stepdump_bridge_oracle", cfg, ref, cwd=str(tmp_path))
        a, b = helpers.read_dump(out), helpers.read_dump(ref)
        bad = np.argwhere(a != b)
        raise AssertionError(f"{cfg} (fused): {len(bad)} values differ from the oracle, first {bad[:1].tolist()}, rel L2 {helpers.rel_l2(a, b):.3e}\n{log[-500:]}")
    print(cfg, [l for l in log.splitlines() if l.startswith("fused_launches")])
    if cfg in ("vacuum_cpml", "nanoblock", "bloch_wide"):     # the grids wide enough for a fused box (>= one 32-cell group between the x slabs)
        assert "fused_launches=0" not in log, f"{cfg}: the fused kernel did not run"
