"""Device-side field functions (alisb_eval_box / alisb_reduce_absmax / alisb_column_sums, SURVEY 8
row f1; reference fields.c:130-681, slices.c:382-427, output.c:56-101).

CPU: the oracle (which evaluates the shared restatement csrc/monitor_eval.h) against independent
numpy float32 formulas written from the reference's definitions; the host layer's use of these
calls is pinned against the unmodified reference by tests/test_host_layer.py (goldens).
GPU: the CUDA kernels against the oracle on seeded fields -- every function, with and without
mirror folding -- bit for bit."""
import numpy as np
import pytest

from alis_b200 import capi, synthetic
from tests import helpers

F32 = np.float32
ALL_FNS = ([(capi.FN_RAW, s, c, 0, 0) for s in range(2) for c in range(3)]
           + [(capi.FN_NODE, s, c, 0, 0) for s in range(2) for c in range(3)]
           + [(capi.FN_EE, 0, 0, 0, 0), (capi.FN_HH, 0, 0, 0, 0), (capi.FN_UH, 0, 0, 0, 0)]
           + [(capi.FN_EXH, 0, a, b, 0) for a in range(3) for b in range(3) if a != b]
           + [(capi.FN_S, 0, c, 0, 0) for c in range(3)]
           + [(capi.FN_TDIAG, 0, c, 0, 0) for c in range(3)]
           + [(capi.FN_TOFF, 0, a, b, 0) for (a, b) in ((0, 1), (2, 0), (1, 2))]
           + [(capi.FN_DIVE, 0, 0, 0, 0)])


def make(lib, shape=(40, 36, 32), seed=5, steps=3, complex_field=0):
    w = synthetic.BoxWorld(*shape, num=5, complex_field=complex_field)
    e = w.make_engine(lib)
    helpers.seed_fields([e], tuple(n + 2 for n in shape), complex_field=complex_field, seed=seed)
    e.run(steps)
    return e


def test_oracle_field_functions_match_numpy(oracle_lib):
    e = make(oracle_lib)
    A = e.read_all_fields()[0]                      # [set][comp][a][b][c]
    E, H = A[0], A[1]
    box = [2, 30, 3, 28, 1, 25]                     # index == array coordinate here (org = 0)
    sl = np.s_[2:31, 3:29, 1:26]

    def sh(X, da=0, db=0, dc=0):
        return X[2 + da:31 + da, 3 + db:29 + db, 1 + dc:26 + dc]

    got = e.eval_box(e.field_fn(capi.FN_NODE, 0, 0), box)                       # Ex = (Ex[i] + Ex[i+1]) / 2
    assert np.array_equal(got, F32(0.5) * (sh(E[0]) + sh(E[0], 1)))
    got = e.eval_box(e.field_fn(capi.FN_NODE, 1, 2), box)                       # Hz = 4-point average in the xy plane
    assert np.array_equal(got, F32(0.25) * (((sh(H[2]) + sh(H[2], 1)) + sh(H[2], 0, 1)) + sh(H[2], 1, 1)))
    got = e.eval_box(e.field_fn(capi.FN_RAW, 1, 1), box)
    assert np.array_equal(got, H[1][sl])
    ee = F32(0)
    for c, d in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
        ee = ee + (sh(E[c]) * sh(E[c]) + sh(E[c], *d) * sh(E[c], *d))
    got = e.eval_box(e.field_fn(capi.FN_EE), box)
    assert np.array_equal(got, F32(0.5) * ee)
    # Sz = ExHy - EyHx, each product 0.25 * (E0*(H0+H0') + E1*(H1+H1')) with H paired along the third axis
    exhy = F32(0.25) * (sh(E[0]) * (sh(H[1]) + sh(H[1], 0, 0, 1)) + sh(E[0], 1) * (sh(H[1], 1) + sh(H[1], 1, 0, 1)))
    eyhx = F32(0.25) * (sh(E[1]) * (sh(H[0]) + sh(H[0], 0, 0, 1)) + sh(E[1], 0, 1) * (sh(H[0], 0, 1) + sh(H[0], 0, 1, 1)))
    got = e.eval_box(e.field_fn(capi.FN_S, 0, 2), box)
    assert np.array_equal(got, exhy - eyhx)
    assert np.abs(got).max() > 0
    # reductions agree with the evaluated block
    f = e.field_fn(capi.FN_S, 0, 2)
    assert e.reduce_absmax(f, box) == np.abs(got).max()
    dz = np.linspace(1.0, 2.0, e.nz + 2).astype(F32)
    e.set_spacing(0, 2, dz)
    cols = e.column_sums(f, box)
    want = np.zeros(got.shape[:2], F32)
    for k in range(box[4] + 1, box[5]):
        want = want + dz[k] * got[:, :, k - box[4]]
    for k in (box[4], box[5]):
        want = (want.astype(np.float64) + 0.5 * np.float64(dz[k]) * got[:, :, k - box[4]].astype(np.float64)).astype(F32)
    assert np.array_equal(cols, want)


def test_oracle_mirror_folding(oracle_lib):
    """index -i on a mirror axis reads its image with the component's parity (fields.c:130-291)"""
    e = make(oracle_lib)
    org = (5, 4, 3)                                  # index 0 sits at array coordinate 5/4/3
    cosk = (1.0, -1.0, 1.0)
    pos = [1, 6, 1, 5, 1, 4]
    neg = [-6, -1, -5, -1, -4, -1]
    for (fn, s, c, c2, m) in ALL_FNS:
        f = e.field_fn(fn, s, c, c2, m, sym=(1, 1, 1), cosk=cosk, org=org)
        a, b = e.eval_box(f, pos), e.eval_box(f, neg)[::-1, ::-1, ::-1]
        assert np.array_equal(np.abs(a), np.abs(b)), (fn, s, c, c2)
        assert np.abs(a).max() > 0
    # a node-centred E component is even along its own axis, odd along the other two
    f = e.field_fn(capi.FN_NODE, 0, 0, sym=(1, 0, 0), cosk=(1.0, 1.0, 1.0), org=org)
    assert np.array_equal(e.eval_box(f, [-3, -1, 2, 4, 2, 4])[::-1], e.eval_box(f, [1, 3, 2, 4, 2, 4]))
    f = e.field_fn(capi.FN_NODE, 0, 1, sym=(1, 0, 0), cosk=(1.0, 1.0, 1.0), org=org)
    assert np.array_equal(e.eval_box(f, [-3, -1, 2, 4, 2, 4])[::-1], -e.eval_box(f, [1, 3, 2, 4, 2, 4]))


def test_eval_rejects_boxes_outside_the_arrays(oracle_lib):
    e = make(oracle_lib, steps=0)
    f = e.field_fn(capi.FN_EE)
    with pytest.raises(capi.EngineError):
        e.eval_box(f, [0, e.nx + 1, 1, 4, 1, 4])       # the stencil would read nx+2
    with pytest.raises(capi.EngineError):
        e.eval_box(f, [-1, 4, 1, 4, 1, 4])             # negative index without a mirror plane
    e.eval_box(e.field_fn(capi.FN_RAW, 0, 0), [0, e.nx + 1, 0, e.ny + 1, 0, e.nz + 1])
    with pytest.raises(capi.EngineError):
        e.column_sums(f, [1, 4, 1, 4, 1, 4])           # no spacing table yet


@pytest.mark.gpu
@pytest.mark.parametrize("complex_field", [0, 1])
def test_cuda_field_functions_bit_exact(oracle_lib, engine_lib, complex_field):
    shape = (67, 45, 38)                               # ragged in x
    g, o = make(engine_lib, shape, complex_field=complex_field), make(oracle_lib, shape, complex_field=complex_field)
    helpers.assert_fields_identical(g, o)
    dz = np.linspace(0.5, 3.0, shape[2] + 2).astype(F32)
    for e in (g, o):
        e.set_spacing(0, 2, dz)
    plain = dict(sym=(0, 0, 0), cosk=(1.0, 1.0, 1.0), org=(0, 0, 0))
    mirror = dict(sym=(1, 1, 1), cosk=(1.0, -1.0, -1.0), org=(20, 15, 12))
    fns = list(ALL_FNS)
    if complex_field:
        fns += [(capi.FN_NODE, s, c, 0, 1) for s in range(2) for c in range(3)]
    for (fn, s, c, c2, m) in fns:
        for kw, box in ((plain, [0, shape[0], 1, shape[1], 0, shape[2]]), (plain, [3, 3, 2, 40, 5, 5]),
                        (mirror, [-19, 40, -14, 25, -11, 20]), (mirror, [-7, -2, 0, 0, -9, 9])):
            if fn == capi.FN_RAW and box[0] == 0:
                box = [0, shape[0] + 1, 0, shape[1] + 1, 0, shape[2] + 1]
            a, b = g.eval_box(g.field_fn(fn, s, c, c2, m, **kw), box), o.eval_box(o.field_fn(fn, s, c, c2, m, **kw), box)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (fn, s, c, c2, m, box)
            assert g.reduce_absmax(g.field_fn(fn, s, c, c2, m, **kw), box) == o.reduce_absmax(o.field_fn(fn, s, c, c2, m, **kw), box)
        for kw, box in ((plain, [1, shape[0], 1, shape[1], 1, shape[2]]), (plain, [5, 5, 1, 30, 7, 7]), (mirror, [-19, 40, -14, 25, 0, 20])):
            a, b = g.column_sums(g.field_fn(fn, s, c, c2, m, **kw), box), o.column_sums(o.field_fn(fn, s, c, c2, m, **kw), box)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), ("columns", fn, s, c, c2, m, box)
    # the imaginary field set of a real world is identically zero (fields.c:179)
    if not complex_field:
        assert not g.eval_box(g.field_fn(capi.FN_NODE, 0, 0, 0, 1), [1, 5, 1, 5, 1, 5]).any()
    with pytest.raises(capi.EngineError):
        g.eval_box(g.field_fn(capi.FN_EE), [0, shape[0] + 1, 1, 4, 1, 4])


@pytest.mark.gpu
def test_cuda_absmax_skips_nans_like_the_reference(oracle_lib, engine_lib):
    g, o = make(engine_lib, steps=0), make(oracle_lib, steps=0)
    f = np.zeros((42, 38, 34), F32)
    f[5, 6, 7], f[9, 9, 9], f[20, 20, 20] = np.nan, -3.5, 2.0
    for e in (g, o):
        e.write_field(0, 2, 0, f)
    box = [1, 39, 1, 35, 1, 31]
    a, b = g.reduce_absmax(g.field_fn(capi.FN_RAW, 0, 2), box), o.reduce_absmax(o.field_fn(capi.FN_RAW, 0, 2), box)
    assert a == b == F32(3.5)


def _absorbing_pair(libs, shape, box, steps=4):
    w = synthetic.BoxWorld(*shape, num=5)
    engines = [w.make_engine(lib) for lib in libs]
    helpers.add_materials(engines, shape, box=box)
    helpers.seed_fields(engines, tuple(n + 2 for n in shape), seed=13)
    dzE = np.linspace(1.0, 2.0, shape[2] + 2).astype(F32)
    dzH = np.linspace(2.0, 1.5, shape[2] + 2).astype(F32)
    for e in engines:
        e.set_spacing(0, 2, dzE)
        e.set_spacing(1, 2, dzH)
        e.run(steps)
    return engines


def test_oracle_absorption_columns_match_numpy(oracle_lib):
    """objectAbsorption's z columns (output.c:23-47,375-404) against a numpy float32 restatement"""
    shape, box = (30, 26, 22), [8, 20, 6, 18, 5, 15]
    (o,) = _absorbing_pair([oracle_lib], shape, box)
    got = o.absorption_columns(0, box)
    sl = np.s_[box[0]:box[1] + 1, box[2]:box[3] + 1, box[4]:box[5] + 1]
    dzE = np.linspace(1.0, 2.0, shape[2] + 2).astype(F32)[box[4]:box[5] + 1]
    dzH = np.linspace(2.0, 1.5, shape[2] + 2).astype(F32)[box[4]:box[5] + 1]
    loss = [F32(0.012), F32(0.009), F32(0.015)]
    for c in range(3):
        E = o.read_field(0, c)[sl]
        inv = o.read_box(capi.ARR_INVEPS, [c], box)
        fill = o.read_box(capi.ARR_ADE_FILL, [0, c], box)
        cur = (F32(2) * fill * E) / (F32(1) + F32(1) / loss[c])
        for p in range(2):
            cur = cur + o.read_box(capi.ARR_ADE_J, [0, 0, p, c], box)
        cur = cur / inv
        terms = (dzH if c == 2 else dzE) * E * cur
        want = np.zeros(terms.shape[:2], F32)
        for k in range(terms.shape[2]):
            want = want + terms[:, :, k]
        assert np.array_equal(got[:, :, c], want), c
        assert np.abs(want).max() > 0


@pytest.mark.gpu
def test_cuda_absorption_columns_bit_exact(oracle_lib, engine_lib):
    shape = (45, 33, 28)
    for box in ([10, 30, 8, 25, 6, 20], [1, 45, 1, 33, 1, 28]):        # interior box; the whole grid (end-plane corrections)
        g, o = _absorbing_pair([engine_lib, oracle_lib], shape, box)
        helpers.assert_fields_identical(g, o)
        a, b = g.absorption_columns(0, box), o.absorption_columns(0, box)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), box
        assert np.abs(b).max() > 0
        sub = [box[0] + 2, box[1] - 3, box[2] + 1, box[3] - 1, box[4], box[5]]
        assert np.array_equal(g.absorption_columns(0, sub), b[2:-3, 1:-1])
    with pytest.raises(capi.EngineError):
        g.absorption_columns(1, box)


def _accumulate(e, steps, with_imag):
    fn = e.field_fn(capi.FN_NODE, 1, 2)                       # Hz, the field of PhC_Band_Diagram's sliceFreqDom
    box = [2, 30, 3, 28, 10, 10]
    acc = e.accum(fn, box, with_imag)
    want_re = np.zeros((29, 26, 1), F32)
    want_im = np.zeros((29, 26, 1), F32)
    for n in range(1, steps + 1):
        e.step_E()
        e.step_H()
        wr, wi = F32(np.cos(F32(0.37) * n)), F32(np.sin(F32(0.37) * n))
        acc.add(float(wr), float(wi))
        v = e.eval_box(fn, box)
        want_re += v * wr                                     # float32, one rounding per product and per sum: slices.c:352-353
        want_im += v * wi
    return acc, want_re, want_im


def test_oracle_running_sums_match_numpy(oracle_lib):
    """alisb_accum_*: the per-point sums of sliceFreqDom / sliceTimeAvg (reference slices.c:282-379)"""
    e = make(oracle_lib)
    acc, want_re, want_im = _accumulate(e, 6, True)
    re, im = acc.read()
    assert np.array_equal(re, want_re) and np.array_equal(im, want_im) and np.abs(re).max() > 0
    acc.clear()
    re, im = acc.read()
    assert not re.any() and not im.any()
    acc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("with_imag", [True, False])
def test_cuda_running_sums_bit_exact(oracle_lib, engine_lib, with_imag):
    """kernel eval_accumulate (asynchronous, one launch per call) against the oracle's loops"""
    g, o = make(engine_lib), make(oracle_lib)
    ag, _, _ = _accumulate(g, 8, with_imag)
    ao, want_re, want_im = _accumulate(o, 8, with_imag)
    for (a, b) in zip(ag.read(), ao.read()):
        if a is not None:
            assert np.array_equal(a, b)
    assert np.array_equal(ag.read()[0], want_re)
    ag.close()
    ao.close()
