"""Host-side mirror (Python) of the reference setup for the SYNTHETIC benchmark worlds:
an all-PML vacuum box with one point dipole (reference examples/Dipole_PML.c
and the `vac_cpml` probe).  It produces exactly the descriptors the reference's
createWorld + pointDipole would hand to updateE/updateH:

  * index ranges                      reference world.c:222-234
  * dtd? tables                       world.c:148-150   (uniform grid, world.c:48-50)
  * CPML / ADE-PML tables             world.c:98-132    (makeCpmlCoeff)
  * PML slab ranges                   update.c:119-124
  * ordered boundary copies           update.c:284-297, 311-339
  * point-dipole box/amplitude        source.c:87-124, waveform Pulse source.c:26-35

The float/double mix of every C expression is reproduced (np.float32 casts) and checked
bit-for-bit against the reference in tests/test_synthetic_setup.py.
Also provides z-slab sharding of such a world for the multi-GPU bench.
"""
import ctypes as C
import math

import numpy as np

from . import capi

F = np.float32
_libm = C.CDLL("libm.so.6")
for _n in ("sinf", "cosf", "expf", "floorf", "sqrtf"):
    getattr(_libm, _n).restype = C.c_float
    getattr(_libm, _n).argtypes = [C.c_float]
_libm.powf.restype = C.c_float
_libm.powf.argtypes = [C.c_float, C.c_float]
PI = F(3.14159265358979323846)


def powf(a, b):
    return F(_libm.powf(F(a), F(b)))


class Axis:
    """One axis of an all-PML uniform world."""

    def __init__(self, n, d, num):
        interior = n - 2 * (num + 1)
        if interior < 3:
            raise ValueError("grid too small for the PML")
        self.n, self.d, self.num = n, F(d), num
        self.Min = -(interior // 2)
        self.Max = self.Min + interior - 1
        self.MIN = self.Min - (num + 1)
        self.MAX = self.Max + (num + 1)
        assert self.MAX - self.MIN + 1 == n

    def itox(self, i):                       # world.c:48 (float arithmetic)
        return F(self.d * F(i))

    def arr(self, i):                        # index -> array coordinate
        return i - (self.MIN - 1)


def axis_tables(ax: Axis, dt, offset):
    """dtd, P, P1(C1), P2(C2) in array coordinates (n+2 floats), world.c:148,98-132."""
    n, num = ax.n, ax.num
    dt = F(dt)
    dtd = np.zeros(n + 2, F)
    for i in range(ax.MIN, ax.MAX + 1):
        dx = F(ax.itox(i - 0.5 * (offset - 1)) - ax.itox(i - 0.5 * (offset + 1)))
        dtd[ax.arr(i)] = F(dt / dx)
    P, C1, C2 = np.zeros(n + 2, F), np.zeros(n + 2, F), np.zeros(n + 2, F)
    sigma_max, kappa_max = F(1), F(1 if num < 24 else 10)
    alpha_max = F(0 if num < 24 else 0.01)
    sigma_pow, alpha_pow = F(3.5), F(1)
    for k in range(1, num + 1):
        x = (k - 0.5 * offset) / num                                        # double
        sigma = F(0.8 * float(sigma_max) * float(F(sigma_pow + F(1))) * float(powf(x, sigma_pow)))
        kappa = F(F(1) + F(F(kappa_max - F(1)) * powf(x, sigma_pow)))
        alpha = F(alpha_max * powf((num - (k - 0.5 * offset)) / num, alpha_pow))
        for idx in (ax.Min - k - 1 + offset, ax.Max + k + 1):
            a = ax.arr(idx)
            dtd[a] = F(dtd[a] / kappa)
            s = float(F(F(sigma * dtd[a]) + F(alpha * dt)))
            ca = F((1 - 0.5 * s) / (1 + 0.5 * s))
            cb = F(float(F(sigma * dtd[a])) / (1 + 0.5 * s))
            P[a] = ca
            C1[a] = F(-dtd[a] * cb)
            C2[a] = F(dtd[a] * cb)
    return dtd, P, C1, C2


def pulse_params(wavelength, supplement, dt):
    """source.c:28-33"""
    wavelength, supplement, dt = F(wavelength), F(supplement), F(dt)
    frequency = F(F(2) * PI / wavelength)
    deviation = F(F(F(2) * F(wavelength * wavelength)) / F(PI * supplement))
    peaktime = F(dt * F(_libm.floorf(F(F(PI * deviation) / dt))))
    n_active = int(F(F(2) * F(_libm.floorf(F(F(PI * deviation) / dt)))))
    return frequency, deviation, peaktime, n_active


def pulse_waveform(frequency, deviation, peaktime):
    """source.c:34 evaluated with the C library's sinf/expf (same libm as the reference)."""
    def wave(t, phase):
        t, phase = F(t), F(phase)
        u = F(t - peaktime)
        q = F(u / deviation)
        return F(F(_libm.sinf(F(F(frequency * u) - phase))) * F(_libm.expf(F(-F(q * q)))))
    return wave


class BoxWorld:
    """All-PML vacuum box of nx x ny x nz cells (PML included) with an Ez point dipole."""

    def __init__(self, nx, ny, nz, dx=10.0, dt=5.0, num=12, wavelength=400.0, supplement=800.0, amplitude=1.0, complex_field=0):
        self.complex_field = complex_field             # 1: a second (imaginary) field set, as in Bloch-periodic worlds
        self.ax = [Axis(nx, dx, num), Axis(ny, dx, num), Axis(nz, dx, num)]
        self.n = (nx, ny, nz)
        self.dt, self.num = F(dt), num
        self.tables = [[axis_tables(a, dt, off) for a in self.ax] for off in (0, 1)]   # [set][axis]
        self.freq, self.dev, self.peak, self.n_active = pulse_params(wavelength, supplement, dt)
        self.wave = pulse_waveform(self.freq, self.dev, self.peak)
        # source.c:103-118: Ez dipole at (0,0,0): box 1x1x2 along z, amplitude array
        amp = F(F(amplitude) * F(F(F(_libm.sqrtf(F(1.5 / float(PI)))) * F(wavelength)) / F(self.dt * self.dt)))
        i, j, k = 0, 0, 0
        dtdxE, dtdyE = self.tables[0][0][0], self.tables[0][1][0]
        dtdzH = self.tables[1][2][0]
        half = F(amp * 0.5)
        self.src_box = [self.ax[0].arr(i), self.ax[0].arr(i), self.ax[1].arr(j), self.ax[1].arr(j), self.ax[2].arr(k), self.ax[2].arr(k + 1)]
        self.src_amp = np.array([F(F(F(half * dtdxE[self.ax[0].arr(i)]) * dtdyE[self.ax[1].arr(j)]) * dtdzH[self.ax[2].arr(kk)])
                                 for kk in (k, k + 1)], F).reshape(1, 1, 2)

    # -- descriptors ------------------------------------------------------------------
    def pml_ranges(self):
        """[set][axis][side] -> (first,last) in array coords, update.c:119-124"""
        out = []
        for s in range(2):
            alt = 1 if s == 0 else -1
            per_axis = []
            for a in self.ax:
                lo = (a.arr(a.MIN - (alt - 1) // 2), a.arr(a.Min - (alt + 3) // 2))
                hi = (a.arr(a.Max + 2), a.arr(a.MAX))
                per_axis.append((lo, hi))
            out.append(per_axis)
        return out

    def copies(self):
        """update.c:284-297 (E) and :311-339 (H) for PML on all faces (cos=1, sin=0, Alt=1)."""
        X, Y, Z = 0, 1, 2
        nx, ny, nz = self.n
        E = [(0, X, 1.0, -0.0, [nx + 1, nx + 1, 1, ny, 1, nz], [-1, 0, 0]),
             (0, Z, 1.0, -0.0, [1, nx, 1, ny, nz + 1, nz + 1], [0, 0, -1]),
             (0, Y, 1.0, -0.0, [1, nx, ny + 1, ny + 1, 1, nz], [0, -1, 0])]
        H = [(1, Y, 1.0, -0.0, [nx + 1, nx + 1, 1, ny + 1, 1, nz + 1], [-1, 0, 0]),
             (1, Z, 1.0, -0.0, [nx + 1, nx + 1, 1, ny + 1, 1, nz + 1], [-1, 0, 0]),
             (1, X, 1.0, -0.0, [1, nx + 1, 1, ny + 1, nz + 1, nz + 1], [0, 0, -1]),
             (1, Y, 1.0, -0.0, [1, nx + 1, 1, ny + 1, nz + 1, nz + 1], [0, 0, -1]),
             (1, Y, 1.0, 0.0, [1, 1, 1, ny + 1, 1, nz + 1], [1, 0, 0]),
             (1, Z, 1.0, 0.0, [1, 1, 1, ny + 1, 1, nz + 1], [1, 0, 0]),
             (1, X, 1.0, 0.0, [1, nx + 1, 1, ny + 1, 1, 1], [0, 0, 1]),
             (1, Y, 1.0, 0.0, [1, nx + 1, 1, ny + 1, 1, 1], [0, 0, 1]),
             (1, Z, 1.0, -0.0, [1, nx + 1, ny + 1, ny + 1, 1, nz + 1], [0, -1, 0]),
             (1, X, 1.0, -0.0, [1, nx + 1, ny + 1, ny + 1, 1, nz + 1], [0, -1, 0]),
             (1, Z, 1.0, 0.0, [1, nx + 1, 1, 1, 1, nz + 1], [0, 1, 0]),
             (1, X, 1.0, 0.0, [1, nx + 1, 1, 1, 1, nz + 1], [0, 1, 0])]
        return E, H

    # -- engine construction (optionally one z-slab of the world) -----------------------
    def shard_range(self, rank, world):
        """planes [c0,c1] (global array coords) owned by `rank` of `world` z-slabs"""
        nz = self.n[2]
        base, rem = divmod(nz, world)
        c0 = 1 + rank * base + min(rank, rem)
        c1 = c0 + base - 1 + (1 if rank < rem else 0)
        return c0, c1

    def make_engine(self, lib, device=0, rank=0, world=1, with_source=True):
        nx, ny, nz = self.n
        c0, c1 = self.shard_range(rank, world)
        nzl = c1 - c0 + 1
        off = c0 - 1                                   # local c = global c - off
        g = capi.Grid()
        g.nx, g.ny, g.nz, g.complex_field, g.dt = nx, ny, nzl, self.complex_field, float(self.dt)
        g.kshard_lo, g.nz_global, g.device = c0, nz, device
        pr = self.pml_ranges()
        for s in range(2):
            for a in range(3):
                for sd in range(2):
                    f, l = pr[s][a][sd]
                    if a == 2:                          # clip the z slabs to the shard
                        f, l = max(f, c0) - off, min(l, c1) - off
                        if f > l:
                            f, l = 1, 0
                    g.pml[s][a][sd][0], g.pml[s][a][sd][1] = f, l
        e = capi.Engine(lib, g)
        for s in range(2):
            for a in range(3):
                t = self.tables[s][a]
                if a == 2:
                    t = tuple(np.ascontiguousarray(x[off:off + nzl + 2]) for x in t)
                e.set_axis_tables(s, a, *t)
        # boundary copies: x/y faces clipped to the shard, z faces only on the end ranks
        for s, ops in enumerate(self.copies()):
            loc = []
            for (cs, comp, cosk, sink, box, alt) in ops:
                b = list(box)
                if alt[2] != 0:                         # a z-face copy
                    if (b[4] == nz + 1 and rank != world - 1) or (b[4] == 1 and rank != 0):
                        continue
                    b[4] -= off
                    b[5] -= off
                else:
                    lo = c0 if rank > 0 else 1
                    hi = c1 if rank < world - 1 else b[5]
                    b[4], b[5] = max(b[4], lo) - off, min(b[5], hi) - off
                    if b[4] > b[5]:
                        continue
                loc.append((cs, comp, cosk, sink, b, alt))
            e.set_copies(s, loc)
        if with_source:
            sb = list(self.src_box)
            k0, k1 = max(sb[4], c0), min(sb[5], c1)
            if k0 <= k1:
                amp = self.src_amp[:, :, k0 - sb[4]:k1 - sb[4] + 1]
                sb[4], sb[5] = k0 - off, k1 - off
                e.add_source(capi.SET_E, 2, 0, sb, self.wave, self.n_active, Amplitude=amp)
        e.shard = (c0, c1)
        return e

    def cells(self):
        return self.n[0] * self.n[1] * self.n[2]

    def algorithmic_bytes_per_step(self):
        """72 B per cell-step + 32 B per CPML slab cell (per slab it lies in)."""
        nx, ny, nz = self.n
        slab = 2 * self.num * (ny * nz + nx * nz + nx * ny)
        return 72 * nx * ny * nz + 32 * slab
