"""End-to-end drop-in check on the GPU: the example drivers (examples/*.c, the four named
configurations + the Dipole_Band known-answer test) built against alis_b200/libalis.so, run next
to their twins built against the unmodified reference (oracle/_ref/bin/*_ref, CPU, all host
threads).  Text outputs must agree value by value (the writers print %g, and the reference's
OpenMP reductions are thread-count dependent in the last digits, hence rtol 2e-4 / atol 1e-6 of
the column maximum); wall times are printed and stored in gpurun_out/examples_timing.json."""
import json
import os
import shutil
import subprocess
import time

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

EXAMPLES = {
    "dipole_pml": [],
    "nanoblock_scattering": [],
    "phc_band_diagram": ["1.2", "8000"],
    "metasurface_farfield": [],
    "dipole_band": [],
}


def run(exe, args, cwd):
    shutil.copy(exe, os.path.join(cwd, os.path.basename(exe)))
    t0 = time.time()
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(helpers.ROOT, "alis_b200") + os.pathsep + env.get("LD_LIBRARY_PATH", "")   # the copy's $ORIGIN rpath no longer points there
    r = subprocess.run(["./" + os.path.basename(exe)] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=1500)
    dt = time.time() - t0
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return dt, r.stdout


def collect_txt(root):
    out = {}
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith(".txt"):
                rel = os.path.relpath(os.path.join(d, f), root)
                rel = rel.replace("_ref.out", ".out").replace("_cuda.out", ".out")
                out[rel] = np.atleast_2d(np.loadtxt(os.path.join(d, f)))
    return out


@pytest.mark.parametrize("name", list(EXAMPLES))
def test_example_matches_reference(name, tmp_path):
    ref_exe = os.path.join(helpers.REF_BIN, name + "_ref")
    my_exe = os.path.join(helpers.TEST_BIN, name + "_cuda")
    assert os.path.exists(my_exe), "run __graft_entry__.build()"
    if not os.path.exists(ref_exe):
        pytest.skip("reference twin not built")
    d_ref, d_my = tmp_path / "ref", tmp_path / "mine"
    d_ref.mkdir()
    d_my.mkdir()
    t_my, out_my = run(my_exe, EXAMPLES[name], str(d_my))
    t_ref, out_ref = run(ref_exe, EXAMPLES[name], str(d_ref))
    a, b = collect_txt(str(d_ref)), collect_txt(str(d_my))
    assert a and sorted(a) == sorted(b), (sorted(a), sorted(b))
    for k in a:
        assert a[k].shape == b[k].shape, k
        scale = np.abs(a[k]).max(axis=0, keepdims=True) + 1e-30
        err = np.abs(a[k] - b[k]) / scale
        assert err.max() < 2e-4, f"{name}/{k}: max column-relative error {err.max():.3e}"
    if name == "dipole_band":          # the three self-normalised numbers (reference examples/Dipole_Band.c:26-28)
        kat = [float(line.rsplit(":", 1)[1]) for line in out_my.splitlines()
               if line.startswith(("Total radiated power", "Intensity to (0,0)", "Radiated power to hemisphere"))]
        assert len(kat) == 3 and all(abs(v - 1) < 0.01 for v in kat), kat
    rec = {"example": name, "args": EXAMPLES[name], "seconds_b200": round(t_my, 2), "seconds_reference_cpu": round(t_ref, 2), "host_threads": os.cpu_count()}
    print("\nEXAMPLE_TIMING " + json.dumps(rec))
    os.makedirs(os.path.join(helpers.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(helpers.ROOT, "gpurun_out", "examples_timing.json"), "a") as f:
        f.write(json.dumps(rec) + "\n")
