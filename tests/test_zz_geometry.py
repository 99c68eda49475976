"""Geometry rasterisation through the engine ABI (alisb_geometry_bounds / alisb_geometry_fill, SURVEY 8
row f3; reference shapes.c:17-221, object.c:70-121,229-262).

CPU: the shape interpreter (csrc/geometry_eval.h, evaluated by the oracle) is pinned against the
unmodified reference through the golden field hashes -- the own host layer with
ALIS_DEVICE_GEOMETRY=1 must still reproduce them -- and against numpy restatements of a few shapes.
GPU: the CUDA kernels against the oracle on a zoo of every shape type, bit for bit, and the golden
configurations through the public API with the device rasteriser on."""
import os
import subprocess

import numpy as np
import pytest

from alis_b200 import capi
from tests import helpers

F32 = np.float32
INF = float(np.finfo(np.float32).max)
needs_bins = pytest.mark.skipif(not helpers.have_test_bin("stepdump_mine_oracle"), reason="tests/bin not built (run __graft_entry__.build())")
OBJECT_CONFIGS = ["nanoblock", "phc", "metasurface", "aniso_ade", "pbc_slab", "tfsf_normal", "shape_zoo", "guided_y"]


def uniform(dx):
    return lambda axis, pos: F32(dx[axis]) * F32(pos)


def zoo():
    """one object per shape type (earlier objects win), spread over a 40^3-ish index box"""
    E = [[0, 0, 0]] * 3

    def ext(x0, x1, y0, y1, z0, z1):
        return [[x0, x1, 0], [y0, y1, 0], [z0, z1, 0]]

    prim = ["BALL", "DIAMOND", "BICONEX", "BICONEY", "BICONEZ", "RODX", "RODY", "RODZ", "DRODX", "DRODY", "DRODZ",
            "HRODXY", "HRODYZ", "HRODZX", "HRODXZ", "HRODYX", "HRODZY"]
    objs = []
    for q, name in enumerate(prim):
        cx, cy, cz = 60 + 130 * (q % 3), 55 + 120 * ((q // 3) % 3), 50 + 170 * (q // 9)
        objs.append((name, ext(cx - 47, cx + 41, cy - 33, cy + 38, cz - 29, cz + 44), []))
    objs.append(("TRANSLATION", [[13.5, -7.25, 21], [0, 0, 0], [0, 0, 0]], [("BALL", ext(10, 90, 300, 380, 20, 95), [])]))
    objs.append(("ROTATION", [[20, 35, 50], [200, 200, 200], [0, 0, 0]], [("BOX", ext(-60, 70, -25, 30, -40, 45), [])]))
    objs.append(("COMBINATION", [[2, 0, 0], [0, 0, 0], [0, 0, 0]], [("RODZ", ext(300, 340, 300, 350, 10, 200), []), ("BOX", ext(320, 390, 330, 360, 50, 90), [])]))
    objs.append(("DIFFERENCE", [[2, 0, 0], [0, 0, 0], [0, 0, 0]],
                 [("BOX", ext(-INF, INF, -INF, INF, 300, 360), []),
                  ("DLATTICE", [[90, 0, 0], [155.88, 0, 0], [0, 0, 0]], [("RODZ", ext(-27, 27, -27, 27, -INF, INF), [])])]))
    objs.append(("INTERSECTION", [[2, 0, 0], [0, 0, 0], [0, 0, 0]], [("BALL", ext(5, 120, 5, 130, 370, 395), []), ("DRODZ", ext(20, 100, 0, 150, 365, 399), [])]))
    objs.append(("LATTICE", [[70, 3, 0], [0, 0, 0], [0, 0, 0]], [("BOX", ext(150, 170, 10, 30, 365, 395), [])]))
    objs.append(("CLATTICEZ", [[5, 0, 0], [300, 100, 0], [0, 0, 0]], [("BALL", ext(330, 370, 90, 110, 370, 395), [])]))
    objs.append(("CLATTICEX", [[3, 0, 0], [330, 200, 0], [0, 0, 0]], [("BOX", ext(10, 30, 320, 340, 230, 250), [])]))
    objs.append(("CLATTICEY", [[4, 0, 0], [100, 280, 0], [0, 0, 0]], [("DIAMOND", ext(120, 160, 350, 390, 270, 300), [])]))
    objs.append(("BOX", [[200, 200, 200], [150, 0, 0], [0, 0, 0]], []))      # centre + half-width form (degenerate z pair)
    return objs


def make_desc(lib, subn, n=40, dx=(10.0, 10.0, 10.0)):
    return capi.GeometryDesc(lib, zoo(), [0, 0, 0], [n + 1, n + 1, n + 1], subn, uniform(dx))


def test_oracle_fill_factors_match_numpy(oracle_lib):
    """a ball and a box rasterised with subN = 4: numpy restatement of object.c:78-121"""
    objs = [("BALL", [[55, 185, 0], [40, 200, 0], [70, 170, 0]], []), ("BOX", [[0, 400, 0], [0, 400, 0], [0, 123, 0]], [])]
    subn, n = 4, 24
    g = capi.GeometryDesc(oracle_lib, objs, [0, 0, 0], [n + 1, n + 1, n + 1], subn, uniform((10.0, 10.0, 10.0)))
    box = [1, n, 1, n, 1, n]
    b = g.bounds([0, n, 0, n, 0, n])
    assert b[0].tolist() == [6, 18, 5, 19, 8, 16] and b[1].tolist() == [0, n, 0, n, 0, 12]
    fx, fy, fz = g.fill(0, box)

    def inside_ball(x, y, z):
        u = [(F32(2) * v - F32(lo) - F32(hi)) / (F32(hi) - F32(lo)) for v, lo, hi in ((x, 55, 185), (y, 40, 200), (z, 70, 170))]
        return (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) < 1

    def want(comp):
        idx = [np.arange(1, n + 1, dtype=np.float64) - (0.5 if comp == a else 0.0) for a in range(3)]
        offs = [(0.5 * (1 - subn + 2 * q) / subn) for q in range(subn)]
        acc = np.zeros((n, n, n), F32)
        for ox in offs:
            for oy in offs:
                for oz in offs:
                    x = F32(10) * (idx[0][:, None, None] + ox).astype(F32)
                    y = F32(10) * (idx[1][None, :, None] + oy).astype(F32)
                    z = F32(10) * (idx[2][None, None, :] + oz).astype(F32)
                    acc += inside_ball(x, y, z).astype(F32)
        return acc / F32(subn ** 3)

    for comp, got in enumerate((fx, fy, fz)):
        w = want(comp)
        partial = (w > 0) & (w < 1)
        assert partial.sum() > 100
        # fully sampled cells agree exactly; uniform cells take the reference's 9-point shortcut
        assert np.array_equal(got[partial & (got > 0) & (got < 1)], w[partial & (got > 0) & (got < 1)])
        assert np.abs(got - w).max() <= 8.0 / subn ** 3


def test_geometry_rejects_bad_descriptions(oracle_lib):
    g = make_desc(oracle_lib, 2, n=12)
    with pytest.raises(capi.EngineError):
        g.fill(0, [0, 14, 1, 4, 1, 4])             # beyond the coordinate tables
    with pytest.raises(capi.EngineError):
        g.fill(99, [1, 4, 1, 4, 1, 4])
    g._nodes[0].type = 77
    with pytest.raises(capi.EngineError):
        g.bounds([1, 4, 1, 4, 1, 4])


@needs_bins
@pytest.mark.parametrize("mode", ["1", "0"])
@pytest.mark.parametrize("cfg", OBJECT_CONFIGS)
def test_host_layer_rasterisers_reproduce_reference(cfg, mode, tmp_path):
    """own host layer with the engine rasteriser (ALIS_DEVICE_GEOMETRY=1, the default: shape interpreter
    in the oracle engine here) and with the host rasteriser (=0): the complete field state hashes to
    the golden of the unmodified reference either way"""
    out = str(tmp_path / "m.bin")
    env = dict(os.environ, ALIS_DEVICE_GEOMETRY=mode)
    r = subprocess.run([os.path.join(helpers.TEST_BIN, "stepdump_mine_oracle"), cfg, "0", out], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    assert "device_geometry_objects" in r.stderr and ("device_geometry_objects 0" in r.stderr) == (mode == "0"), r.stderr[-300:]
    assert helpers.sha256_file(out) == helpers.golden_hashes()[cfg]["sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("subn", [1, 3, 12])
def test_cuda_rasteriser_bit_exact(oracle_lib, engine_lib, subn):
    n, dx = (40, (10.0, 10.0, 10.0)) if subn < 12 else (26, (16.0, 15.5, 16.0))     # both cover the 400 nm zoo
    g, o = make_desc(engine_lib, subn, n=n, dx=dx), make_desc(oracle_lib, subn, n=n, dx=dx)
    whole = [0, n, 0, n, 0, n]
    bg, bo = g.bounds(whole), o.bounds(whole)
    assert np.array_equal(bg, bo)
    owned = [q for q in range(len(bo)) if bo[q][0] <= bo[q][1]]
    assert len(owned) >= 20
    partial = 0
    for q in owned:
        box = [max(int(bo[q][0]) - 1, 0), min(int(bo[q][1]) + 2, n + 1), max(int(bo[q][2]) - 1, 0), min(int(bo[q][3]) + 2, n + 1), max(int(bo[q][4]) - 1, 0), min(int(bo[q][5]) + 2, n + 1)]
        box = [box[0] + (box[0] == 0), box[1], box[2] + (box[2] == 0), box[3], box[4] + (box[4] == 0), box[5]]
        fg, fo = g.fill(q, box), o.fill(q, box)
        for c in range(3):
            assert np.array_equal(fg[c].view(np.uint32), fo[c].view(np.uint32)), (q, c)
            partial += int(((fo[c] > 0) & (fo[c] < 1)).sum())
    assert subn == 1 or partial > 1000
    with pytest.raises(capi.EngineError):
        g.fill(0, [0, n + 5, 1, 4, 1, 4])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["1", "0"])
@pytest.mark.parametrize("cfg", ["nanoblock", "phc", "aniso_ade", "shape_zoo"])
def test_dropin_with_device_and_host_rasteriser(cfg, mode, tmp_path):
    out = str(tmp_path / "m.bin")
    env = dict(os.environ, ALIS_DEVICE_GEOMETRY=mode)
    r = subprocess.run([os.path.join(helpers.TEST_BIN, "stepdump_mine_cuda"), cfg, "0", out], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    assert "device_geometry_objects" in r.stderr and ("device_geometry_objects 0" in r.stderr) == (mode == "0"), r.stderr[-300:]
    assert helpers.sha256_file(out) == helpers.golden_hashes()[cfg]["sha256"]
