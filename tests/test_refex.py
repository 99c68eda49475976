"""The reference's OWN example drivers (reference examples/*.c, unmodified, FULL monitor sets) against
this library: tests/bin/refex_<Name>_cuda (include/alis.h + alis_b200/libalis.so) next to
oracle/_ref/bin/refex_<Name>_ref (the unmodified reference, all host threads).  Both binaries are built
by __graft_entry__.build() from the sources where they lie under /root/reference (no copy) and travel to
the GPU box.  Every output file is compared:

  *.txt  value by value, column-relative (writers print %g / %e; the reference's OpenMP float
         reductions depend on the thread count in the last digits, and the far-field sums are formed
         in another order: tolerance 2e-4 of the column maximum, far inside which the observed
         differences lie -- see DESIGN.md section 6)
  *.h5   byte-identical
  *.png  decoded and compared pixel by pixel (the reference writes them through oracle/shim/png.h);
         colour levels may differ by one step where a value sits on a rounding boundary

The CPU suite runs the short ones with the oracle engine; the GPU suite runs all 15 runnable drivers
(Contour_Example needs a user-supplied contour file and cannot run anywhere as it stands)."""
import json
import os
import shutil
import subprocess
import time

import numpy as np
import pytest

from tests import helpers

# name -> (command line arguments, directory of another example whose outputs it consumes)
REFEX = {
    "Dipole_PML": ([], None),
    "Nanoblock_Scattering": ([], None),
    "PhC_Band_Diagram": ([], None),
    "Metasurface_Farfield": ([], None),
    "Dipole_Band": ([], None),
    "Planewave": ([], None),
    "Focused_Beam": ([], None),
    "Far-Field": ([], None),
    "Bowtie_Nanoantenna": ([], None),
    "Nanowire_Scattering": ([], None),
    "Light_Emitting_Diode": ([], None),
    "PhC_Cavity": ([], None),
    "Cross_Polarization_T": (["300", "120", "80", "200"], None),
    "Waveguide_Dispersion": (["3.0"], None),
    "Waveguide_Propagation": ([], "Waveguide_Dispersion"),
}
NAMED = ["Dipole_PML", "Nanoblock_Scattering", "PhC_Band_Diagram", "Metasurface_Farfield", "Dipole_Band"]


def _exe(name, tag):
    return os.path.join(helpers.REF_BIN if tag == "ref" else helpers.TEST_BIN, f"refex_{name}_{tag}")


def _run(name, tag, cwd, timeout):
    exe = _exe(name, tag)
    local = os.path.join(cwd, name)                 # argv[0] names the output directory
    shutil.copy(exe, local)
    env = dict(os.environ)
    env.pop("ALIS_DEVICE_GEOMETRY_MIN", None)        # the library's own defaults: these runs are also the timed ones
    env["LD_LIBRARY_PATH"] = os.pathsep.join([os.path.join(helpers.ROOT, "alis_b200"), helpers.TEST_BIN, os.path.join(helpers.ROOT, "oracle"), env.get("LD_LIBRARY_PATH", "")])
    t0 = time.time()
    r = subprocess.run(["./" + name] + REFEX[name][0], cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    dt = time.time() - t0
    assert r.returncode == 0, f"{name} ({tag}) rc={r.returncode}\n{r.stdout[-1500:]}\n{r.stderr[-1500:]}"
    if os.path.isfile(local):                       # createWorld's `rm -fr <argv[0]>...` normally removed it already
        os.remove(local)
    return dt, r.stdout


def _files(root):
    out = {}
    for d, _, fs in os.walk(root):
        for f in fs:
            out[os.path.relpath(os.path.join(d, f), root)] = os.path.join(d, f)
    return out


def compare_outputs(name, d_ref, d_my):
    """-> dict of what was compared; asserts on any mismatch"""
    a, b = _files(d_ref), _files(d_my)
    assert sorted(a) == sorted(b), (sorted(set(a) - set(b)), sorted(set(b) - set(a)))
    n = {"txt": 0, "h5": 0, "png": 0, "png_pixels_off_by_one": 0, "png_pixels": 0, "txt_max_err": 0.0}
    for k in sorted(a):
        if k.endswith(".txt"):
            x, y = np.atleast_2d(np.loadtxt(a[k])), np.atleast_2d(np.loadtxt(b[k]))
            assert x.shape == y.shape, k
            fin = np.isfinite(x)
            assert np.array_equal(fin, np.isfinite(y)), k
            x, y = np.where(fin, x, 0), np.where(fin, y, 0)
            scale = np.abs(x).max(axis=0, keepdims=True) + 1e-30
            # a column of rounding noise next to the physical columns of its file (a vanishing
            # polarisation, 1e-16 against 0.4) is compared against the largest column
            scale = np.maximum(scale, 1e-6*scale.max())
            err = float((np.abs(x - y)/scale).max()) if x.size else 0.0
            assert err < 2e-4, f"{name}/{k}: max column-relative error {err:.3e}"
            n["txt"] += 1
            n["txt_max_err"] = max(n["txt_max_err"], err)
        elif k.endswith(".h5"):
            assert open(a[k], "rb").read() == open(b[k], "rb").read(), f"{name}/{k} differs"
            n["h5"] += 1
        elif k.endswith(".png"):
            p, q = helpers.read_png(a[k]), helpers.read_png(b[k])
            assert p.shape == q.shape, f"{name}/{k}: {p.shape} vs {q.shape}"
            diff = np.abs(p.astype(np.int16) - q.astype(np.int16))
            n["png"] += 1
            n["png_pixels"] += int(diff.size)
            n["png_pixels_off_by_one"] += int((diff > 0).sum())
            # far-field maps (Log-*.png, the Dipole_Band profile) come from sums formed in another order;
            # everything else is computed from bit-identical fields
            lim = 3 if ("Log-" in k or name in ("Dipole_Band",)) else 1
            assert diff.max() <= lim and (diff > 0).mean() < 0.02, f"{name}/{k}: max colour difference {diff.max()}, {100*(diff > 0).mean():.2f}% of the values differ"
    assert n["txt"] + n["h5"] + n["png"] > 0, "the example wrote nothing"
    return n


def _twin(name, tag, tmp_path, timeout):
    dep = REFEX[name][1]
    recs = {}
    for who in ("ref", tag):
        d = tmp_path / who
        d.mkdir()
        t = 0.0
        if dep:                                         # its inputs: run the producer in the same directory first
            _run(dep, who, str(d), timeout)
        t, out = _run(name, who, str(d), timeout)
        recs[who] = (t, out)
    n = compare_outputs(name, str(tmp_path / "ref"), str(tmp_path / tag))
    return recs, n


needs_bins = pytest.mark.skipif(not (os.path.exists(_exe("Dipole_PML", "ref")) and os.path.exists(_exe("Dipole_PML", "oracle"))),
                                reason="refex binaries not built (needs /root/reference at build time)")


@needs_bins
@pytest.mark.parametrize("name", ["Dipole_PML", "Metasurface_Farfield", "Planewave", "Focused_Beam"])
def test_reference_driver_on_the_oracle_engine(name, tmp_path):
    """host layer + CPU oracle engine: every file the unmodified driver writes, PNG frames included"""
    _, n = _twin(name, "oracle", tmp_path, 900)
    if name in ("Dipole_PML", "Metasurface_Farfield"):
        assert n["png"] > 10


def test_read_png_roundtrip(tmp_path):
    img = (np.arange(7*5*3) % 251).astype(np.uint8).reshape(5, 7, 3)
    helpers.write_png(str(tmp_path / "a.png"), img)
    assert np.array_equal(helpers.read_png(str(tmp_path / "a.png")), img)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMED + [k for k in REFEX if k not in NAMED])
def test_reference_driver_on_the_gpu(name, tmp_path):
    """the unmodified driver, full monitor set, CUDA engine vs the unmodified reference on the host cores"""
    assert os.path.exists(_exe(name, "cuda")), "run __graft_entry__.build() where /root/reference is mounted"
    if not os.path.exists(_exe(name, "ref")):
        pytest.skip("reference twin not built")
    recs, n = _twin(name, "cuda", tmp_path, 1500)
    rec = {"example": name, "args": REFEX[name][0], "seconds_b200": round(recs["cuda"][0], 2), "seconds_reference_cpu": round(recs["ref"][0], 2),
           "host_threads": os.cpu_count(), "compared": n}
    if name == "Dipole_Band":          # the three self-normalised numbers (reference examples/Dipole_Band.c:26-28)
        kat = [float(line.rsplit(":", 1)[1]) for line in recs["cuda"][1].splitlines()
               if line.startswith(("Total radiated power", "Intensity to (0,0)", "Radiated power to hemisphere"))]
        assert len(kat) == 3 and all(abs(v - 1) < 0.01 for v in kat), kat
        rec["known_answers"] = kat
    print("\nREFEX_TIMING " + json.dumps(rec))
    os.makedirs(os.path.join(helpers.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(helpers.ROOT, "gpurun_out", "refex_timing.json"), "a") as f:
        f.write(json.dumps(rec) + "\n")
