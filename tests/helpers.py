"""Shared test helpers (test infrastructure; may use the oracle)."""
import hashlib
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
GOLDEN = os.path.join(ROOT, "tests", "golden")

CONFIGS = ["dipole_pml", "vacuum_cpml", "pec_box", "pbc_slab", "nanoblock", "tfsf_normal",
           "phc", "metasurface", "dipole_band", "aniso_ade", "bloch_wide", "chi2_shg", "chi2_twm", "chi2_alias",
           "guided_y", "guided_xz", "shape_zoo", "focused_beam", "focused_tfsf"]


MONITOR_CONFIGS = ["nanoblock", "metasurface", "dipole_band", "phc", "dipole_pml", "aniso_ade"]
TEST_BIN = os.path.join(ROOT, "tests", "bin")


def have_test_bin(name):
    return os.path.exists(os.path.join(TEST_BIN, name))


def run_monitors(binary, config, threads=None, steps=0, host_monitors=False, direct_phasor=False, extra_env=None):
    """runs oracle/harness/monitors.c (reference- or own-library build) -> ordered [(tag, text value)]
    host_monitors: ALIS_HOST_MONITORS=1, i.e. every monitor through the host views instead of the
    device-side evaluation"""
    exe = os.path.join(REF_BIN if binary.endswith("_ref") else TEST_BIN, binary)
    env = dict(os.environ)
    env.pop("ALIS_HOST_MONITORS", None)
    if host_monitors:
        env["ALIS_HOST_MONITORS"] = "1"
    env.update(extra_env or {})
    env.pop("HARNESS_DIRECT_PHASOR", None)
    if direct_phasor:                    # the harness indexes the phasor arrays without refreshing them first
        env["HARNESS_DIRECT_PHASOR"] = "omp" if direct_phasor == "omp" else "1"
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    r = subprocess.run([exe, config, str(steps)], env=env, capture_output=True, text=True, timeout=1800)
    if r.returncode != 0:
        raise RuntimeError(f"{binary} {config} failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    vals = []
    for line in r.stdout.splitlines():
        parts = line.split()
        if len(parts) == 2 and not line.startswith(("Preparing", "Putting")):
            try:
                float(parts[1])
            except ValueError:
                continue
            vals.append([parts[0], parts[1]])
    return vals


def parse_tag_values(stdout):
    """'tag value' lines of a harness program -> ordered [[tag, text value]]"""
    vals = []
    for line in stdout.splitlines():
        parts = line.split()
        if len(parts) == 2 and not line.startswith(("Preparing", "Putting")):
            try:
                float(parts[1])
            except ValueError:
                continue
            vals.append([parts[0], parts[1]])
    return vals


def run_api_misc(binary, cwd):
    r = subprocess.run([os.path.join(TEST_BIN, binary)], cwd=cwd, capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        raise RuntimeError(f"{binary} failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return parse_tag_values(r.stdout)


def run_graded_slices(binary, cwd):
    """oracle/harness/graded_slices.c: {file name: sha256} of the slice files it writes under cwd"""
    exe = os.path.join(REF_BIN if binary.endswith("_ref") else TEST_BIN, binary)
    r = subprocess.run([exe], cwd=cwd, capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        raise RuntimeError(f"{binary} failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    out = os.path.join(cwd, "graded_slices.out")
    return {f: sha256_file(os.path.join(out, f)) for f in sorted(os.listdir(out))}


def golden_graded_slices():
    with open(os.path.join(GOLDEN, "graded_slices.json")) as f:
        return json.load(f)


def golden_api_misc():
    with open(os.path.join(GOLDEN, "api_misc_values.json")) as f:
        return json.load(f)


def monitor_mismatches(got, want, device_flux=False):
    """Value-by-value comparison of two monitors.c outputs.  Everything must be identical to the
    printed digit (%.9g) except
      * far-field values (tags ff*): the surface sums run on the device / in the oracle with fp32
        products and double accumulation, the reference sums in float with an OpenMP reduction --
        tolerance 1e-5 (the north-star bound for spectra) of the run's far-field scale;
      * with device_flux, the Poynting integrals (reduced on the device in double): 5e-6."""
    assert [g[0] for g in got] == [w[0] for w in want]
    val = {w[0]: float(w[1]) for w in want}
    # far-field values are compared against the scale of their kind in the same run: amplitudes (re/im
    # pairs) against the largest |f|, intensities / fluxes against the largest intensity -- a component
    # that vanishes by symmetry is cancellation noise in the reference as well as here
    amp = [abs(v) for k, v in val.items() if k.startswith("ff") and k.endswith(("_re", "_im")) and np.isfinite(v)]
    inten = [abs(v) for k, v in val.items() if k.startswith("ff") and not k.endswith(("_re", "_im")) and np.isfinite(v)]
    amp_scale, inten_scale = max(amp + [0.0]), max(inten + [0.0])
    bad = []
    for g, w in zip(got, want):
        tag, a, b = g[0], g[1], w[1]
        if a == b:
            continue
        x, y = float(a), float(b)
        if x != x and y != y:
            continue                    # NaN in both (the sign a NaN prints with is arbitrary)
        if tag.startswith("ff"):
            scale = amp_scale if tag.endswith(("_re", "_im")) else inten_scale
            if abs(x - y) <= 1e-5 * max(scale, abs(x), abs(y)):
                continue
        if device_flux and tag.startswith(("poynting", "flux_")) and abs(x - y) <= 5e-6 * max(abs(x), abs(y)):
            continue
        bad.append((g, w))
    return bad


def run_mine_stepdump(binary, config, out, steps=0, cwd=None):
    exe = os.path.join(TEST_BIN, binary)
    r = subprocess.run([exe, config, str(steps), out], cwd=cwd, capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        raise RuntimeError(f"{binary} {config} failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return r.stdout


def golden_monitors():
    with open(os.path.join(GOLDEN, "monitor_values.json")) as f:
        return json.load(f)


def have_ref_bin(name):
    return os.path.exists(os.path.join(REF_BIN, name))


def run_stepdump(binary, config, out, steps=0, cwd=None):
    exe = os.path.join(REF_BIN, binary)
    env = dict(os.environ)
    env.setdefault("OMP_NUM_THREADS", str(min(8, os.cpu_count() or 1)))
    r = subprocess.run([exe, config, str(steps), out], cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        raise RuntimeError(f"{binary} {config} failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return r.stdout


def read_dump(path):
    b = open(path, "rb").read()
    assert b[:8] == b"ALISDUMP"
    nx, ny, nz, cf = np.frombuffer(b[8:24], dtype=np.int32)
    a = np.frombuffer(b[24:], dtype=np.float32).reshape(cf + 1, 2, 3, nx + 2, ny + 2, nz + 2)
    return a


def sha256_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def golden_hashes():
    with open(os.path.join(GOLDEN, "stepdump_hashes.json")) as f:
        return json.load(f)


def rel_l2(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    d = np.sqrt(((a - b) ** 2).sum())
    n = np.sqrt((b ** 2).sum())
    return float(d / n) if n > 0 else float(d)


def seed_fields(engines, shape, complex_field=0, seed=12345, scale=1e-3):
    """Same LCG-seeded dense fields (uniform in +-scale) into every engine."""
    rng = np.random.RandomState(seed)
    for m in range(complex_field + 1):
        for s in range(2):
            for c in range(3):
                f = np.zeros(shape, np.float32)
                f[1:-1, 1:-1, 1:-1] = ((rng.rand(*[n - 2 for n in shape]) * 2 - 1) * scale).astype(np.float32)
                for e in engines:
                    e.write_field(s, c, m, f)


def add_materials(engines, shape, seed=21, box=None, zero_some=False):
    """random inverse permittivity everywhere and one dispersive box (1 Drude + 1 critical-point pole,
    lossy) in the middle of the grid -- the same on every engine"""
    rng = np.random.RandomState(seed)
    eps = [(0.3 + 0.7 * rng.rand(shape[0] + (c == 0), shape[1] + (c == 1), shape[2] + (c == 2))).astype(np.float32) for c in range(3)]
    if zero_some:                                   # cells without material response (objectAbsorption skips them)
        for a in eps:
            a[rng.rand(*a.shape) < 0.1] = 0
    if box is None:
        box = [shape[0] // 3, 2 * shape[0] // 3, shape[1] // 3, 2 * shape[1] // 3, shape[2] // 3, 2 * shape[2] // 3]
    dims = (box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1)
    fill = [rng.rand(*dims).astype(np.float32) for _ in range(3)]
    coef = {k: [[0.0] * 8 for _ in range(3)] for k in ("JJ", "JK", "JE", "KJ", "KK", "KE")}
    for c in range(3):
        coef["JJ"][c][0], coef["JE"][c][0] = 0.93 - 0.01 * c, 0.011 + 0.001 * c
        coef["JJ"][c][1], coef["JK"][c][1], coef["JE"][c][1] = 0.9, -0.05 - 0.01 * c, 0.007
        coef["KJ"][c][1], coef["KK"][c][1], coef["KE"][c][1] = 0.04, 0.91, 0.003 + 0.001 * c
    for e in engines:
        e.set_inv_eps(*eps)
        e.add_dispersive(box, 1, 2, [0.012, 0.009, 0.015], coef, fill)
    return box


def assert_fields_identical(e1, e2, what=""):
    a, b = e1.read_all_fields(), e2.read_all_fields()
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} of {a.size} values differ; first at [m,set,comp,a,b,c]={bad[0].tolist()} "
                             f"{a[tuple(bad[0])]!r} vs {b[tuple(bad[0])]!r}; rel L2 = {rel_l2(a, b):.3e}")
    assert np.abs(a).max() > 0, f"{what}: fields are all zero - the test is vacuous"


def read_png(path):
    """decodes an 8-bit RGB, non-interlaced PNG (what both png writers here produce; all five scanline
    filters are handled) -> uint8 array (height, width, 3)"""
    import struct
    import zlib
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n", path
    pos, idat, w = 8, b"", None
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert zlib.crc32(tag + body) == struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0], "bad chunk crc"
        if tag == b"IHDR":
            w, h, depth, colour, _, _, interlace = struct.unpack(">IIBBBBB", body)
            assert depth == 8 and colour == 2 and interlace == 0, "only 8-bit RGB, non-interlaced"
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3*w)
    out = np.zeros((h, 3*w), np.uint8)
    for y in range(h):
        f, line = raw[y, 0], raw[y, 1:].astype(np.int32)
        up = out[y - 1].astype(np.int32) if y else np.zeros(3*w, np.int32)
        if f == 0:
            cur = line
        elif f == 2:
            cur = (line + up) & 255
        else:
            cur = np.zeros(3*w, np.int32)
            for x in range(3*w):
                a = cur[x - 3] if x >= 3 else 0
                c = up[x - 3] if x >= 3 else 0
                b = up[x]
                if f == 1:
                    pred = a
                elif f == 3:
                    pred = (a + b)//2
                else:
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if pa <= pb and pa <= pc else (b if pb <= pc else c)
                cur[x] = (line[x] + pred) & 255
        out[y] = cur
    return out.reshape(h, w, 3)


def write_png(path, img):
    import struct
    import zlib
    h, w, _ = img.shape
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body))
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))
