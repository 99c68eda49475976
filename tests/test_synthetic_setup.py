"""The Python mirror of the reference setup for the synthetic all-PML box (alis_b200/synthetic.py,
used by bench.py) reproduces the reference: its descriptors + the oracle engine give the same
bits as the unmodified reference ran on the same world (golden `vacuum_cpml`)."""
import hashlib

import numpy as np

from alis_b200 import capi, synthetic
from tests import helpers


def dump_bytes(e, n):
    hdr = b"ALISDUMP" + np.array([n[0], n[1], n[2], 0], np.int32).tobytes()
    return hdr + e.read_all_fields().tobytes()


def test_box_world_matches_reference_golden(oracle_lib):
    w = synthetic.BoxWorld(76, 64, 56)            # == cfg_vacuum_cpml in oracle/harness/configs.h
    e = w.make_engine(oracle_lib)
    e.run(60)
    assert e.time() == (60, 302.5, 300.0, 302.5)
    got = hashlib.sha256(dump_bytes(e, w.n)).hexdigest()
    assert got == helpers.golden_hashes()["vacuum_cpml"]["sha256"]


def test_cpml_tables_shape_and_grading():
    w = synthetic.BoxWorld(40, 36, 32, num=12)
    for s in range(2):
        for a in range(3):
            dtd, P, C1, C2 = w.tables[s][a]
            n = w.n[a]
            assert dtd.shape == (n + 2,) and dtd[0] == 0 and dtd[n + 1] == 0
            assert np.all(dtd[1:n + 1] == np.float32(0.5))          # Num<24: kappa = 1
            inside = (P != 0)
            assert inside.sum() == 2 * 12
            assert np.all(C1[inside] <= 0) and np.all(C2[inside] >= 0) and np.all(C1 == -C2)
    w24 = synthetic.BoxWorld(80, 80, 80, num=24)                    # Num>=24: true CPML, kappa up to 10
    assert w24.tables[0][0][0][1:81].min() < 0.06


def test_algorithmic_bytes():
    w = synthetic.BoxWorld(256, 256, 256)
    assert w.algorithmic_bytes_per_step() == 72 * 256 ** 3 + 32 * 6 * 12 * 256 ** 2
