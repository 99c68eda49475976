"""Near-to-far-field surface sums (alisb_phaser_surface_sums).

CPU: the oracle's restatement of the reference's double loops (farfld.c:248-459, weights folded
by the caller) against an independent numpy complex128 evaluation on a seeded phasor state.
GPU: the CUDA kernel against the oracle on the same accumulators -- both form the same fp32
products and accumulate in double, so they agree to double rounding."""
import numpy as np
import pytest

from alis_b200 import capi, synthetic
from tests import helpers


def accumulate_some(engines, box, freqs, steps=6):
    ph = [capi.Phaser(e, box, 0x3F, len(freqs)) for e in engines]
    for _ in range(steps):
        for e in engines:
            e.step_E()
            e.step_H()
        _, _, tE, tH = engines[0].time()
        wE = np.exp(2j * np.pi * freqs.astype(np.float64) * tE).astype(np.complex64)
        wH = np.exp(2j * np.pi * freqs.astype(np.float64) * tH).astype(np.complex64)
        for p in ph:
            p.accumulate(wE, wH)
    return ph


def random_jobs(ph, rng, njobs):
    """jobs over all faces/accumulators with shared and private weight vectors"""
    weights, jobs, offs = [], [], {}
    pos = 0
    for q in range(njobs):
        face, which = int(rng.randint(6)), int(rng.randint(4))
        nlo, nhi = ph.face_extents(face)
        key = (face, q % 3)                       # every third job of a face shares its vectors
        if key not in offs:
            wl = (rng.randn(nlo) + 1j * rng.randn(nlo)).astype(np.complex64)
            wh = (rng.randn(nhi) + 1j * rng.randn(nhi)).astype(np.complex64)
            wl[0] = 0                             # the folded "outer" vector starts with a zero
            offs[key] = (pos, pos + nlo)
            weights += [wl, wh]
            pos += nlo + nhi
        jobs.append((face, which, offs[key][0], offs[key][1]))
    return jobs, np.concatenate(weights)


def numpy_sums(ph, freq, jobs, W):
    out = []
    for face, which, lo, hi in jobs:
        nlo, nhi = ph.face_extents(face)
        A = ph.read(freq, face, which).astype(np.complex128)
        A = A.reshape(nlo, nhi)                   # the face block is (lower axis, higher axis), z fastest
        out.append(np.einsum("u,uv,v->", W[lo:lo + nlo].astype(np.complex128), A, W[hi:hi + nhi].astype(np.complex128)))
    return np.array(out)


def test_oracle_surface_sums_match_numpy(oracle_lib):
    w = synthetic.BoxWorld(40, 36, 32, num=5)
    o = w.make_engine(oracle_lib)
    helpers.seed_fields([o], (42, 38, 34), seed=11)
    freqs = np.array([1 / 400.0, 1 / 610.0], np.float32)
    (po,) = accumulate_some([o], [6, 35, 7, 30, 5, 28], freqs)
    rng = np.random.RandomState(5)
    jobs, W = random_jobs(po, rng, 40)
    for f in range(2):
        got, want = po.surface_sums(f, jobs, W), numpy_sums(po, f, jobs, W)
        scale = np.abs(want).max()
        assert scale > 0
        assert np.abs(got - want).max() <= 2e-6 * scale      # fp32 products, double sums
    # the degenerate cases: no jobs, a job on a face that is not accumulated
    assert po.surface_sums(0, [], W).shape == (0,)


@pytest.mark.gpu
def test_cuda_surface_sums_match_oracle(oracle_lib, engine_lib):
    w = synthetic.BoxWorld(48, 44, 40, num=6)
    g, o = w.make_engine(engine_lib), w.make_engine(oracle_lib)
    helpers.seed_fields([g, o], (50, 46, 42), seed=7)
    freqs = np.array([1 / 400.0, 1 / 520.0, 1 / 650.0], np.float32)
    pg, po = accumulate_some([g, o], [8, 41, 8, 37, 8, 33], freqs)
    rng = np.random.RandomState(9)
    for njobs in (1, 7, 300, 5000):               # a single query up to a sweep-sized batch
        jobs, W = random_jobs(pg, rng, njobs)
        for f in (0, 2):
            a, b = pg.surface_sums(f, jobs, W), po.surface_sums(f, jobs, W)
            scale = np.abs(b).max()
            assert scale > 0
            assert np.abs(a - b).max() <= 1e-12 * scale, (njobs, f, np.abs(a - b).max() / scale)
    with pytest.raises(capi.EngineError):
        pg.surface_sums(0, [(0, 0, W.size, 0)], W)          # weight vector outside the buffer


@pytest.mark.gpu
def test_cuda_surface_sums_single_face(oracle_lib, engine_lib):
    """only the z faces are accumulated (the Metasurface_Farfield layout): jobs on other faces are refused"""
    w = synthetic.BoxWorld(56, 30, 44, num=6)
    g = w.make_engine(engine_lib)
    helpers.seed_fields([g], (58, 32, 46), seed=3)
    p = capi.Phaser(g, [1, 56, 1, 30, 9, 36], 0x30, 1)
    g.run(3)
    p.accumulate(np.array([1 + 0.5j], np.complex64), np.array([0.25 - 1j], np.complex64))
    nlo, nhi = p.face_extents(5)
    assert (nlo, nhi) == (56, 30)
    W = np.ones(nlo + nhi, np.complex64)
    got = p.surface_sums(0, [(5, 0, 0, nlo), (4, 3, 0, nlo)], W)
    want = [p.read(0, 5, 0).astype(np.complex128).sum(), p.read(0, 4, 3).astype(np.complex128).sum()]
    assert np.allclose(got, want, rtol=1e-6, atol=1e-6 * np.abs(want).max())
    with pytest.raises(capi.EngineError):
        p.surface_sums(0, [(0, 0, 0, 0)], W)
