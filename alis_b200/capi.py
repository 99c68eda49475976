"""ctypes binding of the engine C ABI (include/alis_b200.h).

Two shared libraries export this ABI:
  * alis_b200/libalis_b200_engine.so  -- the product: hand-written sm_100a CUDA kernels
  * oracle/liboracle_alis.so          -- the CPU oracle, TEST INFRASTRUCTURE ONLY

`load_engine()` loads the CUDA library and raises if it is missing -- there is no CPU
fallback.  `load_oracle()` may only be called from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ENGINE_SO = os.path.join(HERE, "libalis_b200_engine.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle_alis.so")

MAXPOLES = 6
SET_E, SET_H = 0, 1
PART_ALL, PART_SEND_PLANE, PART_REST = 0, 1, 2
ARR_FIELD, ARR_INVEPS, ARR_PSI, ARR_ADE_J, ARR_ADE_K, ARR_ADE_FILL = range(6)

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)


class Grid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("complex_field", C.c_int32),
                ("dt", C.c_float), ("pml", C.c_int32 * 2 * 2 * 3 * 2),  # [set][axis][side][first,last]
                ("kshard_lo", C.c_int32), ("nz_global", C.c_int32), ("device", C.c_int32),
                ("reserved", C.c_int32 * 5)]


class Dispersive(C.Structure):
    _fields_ = [("box", C.c_int32 * 6), ("nj", C.c_int32), ("nk", C.c_int32), ("a", C.c_float * 3),
                ("JJ", C.c_float * MAXPOLES * 3), ("JK", C.c_float * MAXPOLES * 3), ("JE", C.c_float * MAXPOLES * 3),
                ("KJ", C.c_float * MAXPOLES * 3), ("KK", C.c_float * MAXPOLES * 3), ("KE", C.c_float * MAXPOLES * 3),
                ("fill", f32p * 3)]


WAVEFORM_FN = C.CFUNCTYPE(C.c_float, C.c_void_p, C.c_float, C.c_float)


class Source(C.Structure):
    _fields_ = [("set", C.c_int32), ("comp", C.c_int32), ("m", C.c_int32), ("box", C.c_int32 * 6),
                ("amplitude", C.c_float), ("phase", C.c_float), ("Amplitude", f32p), ("Phase", f32p),
                ("n_active", C.c_uint64), ("waveform", WAVEFORM_FN), ("user", C.c_void_p)]


class FieldFn(C.Structure):
    """alisb_field_fn: one of the reference's point-wise field functions (fields.c) + mirror info"""
    _fields_ = [("fn", C.c_int32), ("set", C.c_int32), ("comp", C.c_int32), ("comp2", C.c_int32), ("m", C.c_int32),
                ("sym", C.c_int32 * 3), ("cosk", C.c_float * 3), ("org", C.c_int32 * 3)]


FN_RAW, FN_NODE, FN_EE, FN_HH, FN_UH, FN_EXH, FN_S, FN_TDIAG, FN_TOFF, FN_DIVE = range(10)


class Shape(C.Structure):
    """alisb_shape: one node of a flattened object tree (reference shapes.c)"""
    _fields_ = [("type", C.c_int32), ("child", C.c_int32), ("nchild", C.c_int32), ("aux", C.c_int32), ("f", C.c_float * 3 * 3)]


class Geometry(C.Structure):
    _fields_ = [("nobjects", C.c_int32), ("roots", i32p), ("nnodes", C.c_int32), ("nodes", C.POINTER(Shape)),
                ("naux", C.c_int32), ("aux", f32p), ("subn", C.c_int32), ("lo", C.c_int32 * 3), ("hi", C.c_int32 * 3),
                ("coord", f32p * 3)]


SHAPES = ("BOX BALL DIAMOND BICONEX BICONEY BICONEZ RODX RODY RODZ DRODX DRODY DRODZ HRODXY HRODYZ HRODZX HRODXZ HRODYX HRODZY "
          "TRANSLATION ROTATION COMBINATION INTERSECTION DIFFERENCE LATTICE DLATTICE CLATTICEX CLATTICEY CLATTICEZ").split()
SHAPE = {name: q for q, name in enumerate(SHAPES)}


class Copy(C.Structure):
    _fields_ = [("set", C.c_int32), ("comp", C.c_int32), ("cosk", C.c_float), ("sink", C.c_float),
                ("box", C.c_int32 * 6), ("alt", C.c_int32 * 3)]


def _declare(lib):
    vp, i, u64, f = C.c_void_p, C.c_int, C.c_uint64, C.c_float
    sig = {
        "alisb_abi_version": (i, []),
        "alisb_last_error": (C.c_char_p, []),
        "alisb_backend": (C.c_char_p, []),
        "alisb_create": (i, [C.POINTER(Grid), C.POINTER(vp)]),
        "alisb_destroy": (None, [vp]),
        "alisb_set_stream": (i, [vp, vp]),
        "alisb_set_axis_tables": (i, [vp, i, i, f32p, f32p, f32p, f32p]),
        "alisb_set_inv_eps": (i, [vp, f32p, f32p, f32p]),
        "alisb_add_dispersive": (i, [vp, C.POINTER(Dispersive)]),
        "alisb_clear_dispersive": (i, [vp]),
        "alisb_add_source": (i, [vp, C.POINTER(Source)]),
        "alisb_add_tfsf": (i, [vp, vp]),
        "alisb_set_copies": (i, [vp, i, C.POINTER(Copy), i]),
        "alisb_clear_sources": (i, [vp]),
        "alisb_step_E": (i, [vp]),
        "alisb_step_H": (i, [vp]),
        "alisb_step_part": (i, [vp, i, i]),
        "alisb_run": (i, [vp, i]),
        "alisb_synchronize": (i, [vp]),
        "alisb_get_time": (i, [vp, C.POINTER(u64), C.POINTER(f), C.POINTER(f), C.POINTER(f)]),
        "alisb_set_time": (i, [vp, u64, f, f, f]),
        "alisb_reset_fields": (i, [vp]),
        "alisb_read_box": (i, [vp, i, i32p, i32p, f32p]),
        "alisb_write_box": (i, [vp, i, i32p, i32p, f32p]),
        "alisb_set_spacing": (i, [vp, i, i, f32p]),
        "alisb_plane_flux": (i, [vp, C.c_void_p, i, C.POINTER(C.c_double)]),
        "alisb_eval_box": (i, [vp, C.POINTER(FieldFn), i32p, f32p]),
        "alisb_accum_create": (i, [vp, C.POINTER(FieldFn), i32p, i, C.POINTER(vp)]),
        "alisb_accum_add": (i, [vp, f, f]),
        "alisb_accum_read": (i, [vp, f32p, f32p]),
        "alisb_accum_clear": (i, [vp]),
        "alisb_accum_destroy": (None, [vp]),
        "alisb_reduce_absmax": (i, [vp, C.POINTER(FieldFn), i32p, f32p]),
        "alisb_column_sums": (i, [vp, C.POINTER(FieldFn), i32p, f32p]),
        "alisb_absorption_columns": (i, [vp, i, i32p, f32p]),
        "alisb_geometry_bounds": (i, [i, C.POINTER(Geometry), i32p, i32p]),
        "alisb_geometry_fill": (i, [i, C.POINTER(Geometry), i, i32p, f32p, f32p, f32p]),
        "alisb_chi2_couple": (i, [vp, i, i, vp, vp, f32p, f]),
        "alisb_set_fusion": (i, [vp, i]),
        "alisb_flush": (i, [vp, i]),
        "alisb_begin_pass": (i, [vp, i]),
        "alisb_fused_launch_count": (u64, [vp]),
        "alisb_halo_info": (i, [vp, i, i, i, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "alisb_halo_pack": (i, [vp, i]),
        "alisb_halo_unpack": (i, [vp, i]),
        "alisb_halo_exchange": (i, [vp, vp, i]),
        "alisb_phaser_create": (i, [vp, i32p, i, i, C.POINTER(vp)]),
        "alisb_phaser_destroy": (None, [vp]),
        "alisb_phaser_accumulate": (i, [vp, f32p, f32p]),
        "alisb_phaser_read": (i, [vp, i, i, i, f32p]),
        "alisb_phaser_surface_sums": (i, [vp, i, C.c_void_p, i, f32p, C.c_size_t, C.POINTER(C.c_double)]),
        "alisb_launch_count": (u64, [vp]),
        "alisb_set_profiling": (i, [vp, i]),
        "alisb_last_run_kernel_ms": (i, [vp, C.POINTER(f), C.POINTER(i)]),
        "alisb_profile_dominant": (i, [vp, C.POINTER(f), C.POINTER(i), C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError => the library does not export the ABI
        fn.restype, fn.argtypes = res, args
    return sig


ABI_SYMBOLS = None


def _load(path, what):
    global ABI_SYMBOLS
    if not os.path.exists(path):
        raise RuntimeError(f"{what} not built: {path} is missing (run `python -c 'import __graft_entry__ as g; g.build()'`)")
    lib = C.CDLL(path)
    ABI_SYMBOLS = sorted(_declare(lib))
    return lib


def load_engine():
    """The CUDA engine.  Raises if the extension is missing -- no CPU fallback."""
    return _load(ENGINE_SO, "CUDA engine")


def load_oracle():
    """CPU oracle (test infrastructure).  Never call this from product code."""
    return _load(ORACLE_SO, "CPU oracle")


def _fp(a):
    return a.ctypes.data_as(f32p) if a is not None else None


def _ip(seq):
    return (C.c_int32 * len(seq))(*[int(v) for v in seq])


class EngineError(RuntimeError):
    pass


class Engine:
    """Thin object wrapper over one alisb_engine (CUDA or oracle, depending on `lib`)."""

    def __init__(self, lib, grid: Grid):
        self.lib, self.grid = lib, grid
        self.h = C.c_void_p()
        self._keep = []
        self._ck(lib.alisb_create(C.byref(grid), C.byref(self.h)))
        self.nx, self.ny, self.nz = grid.nx, grid.ny, grid.nz

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(f"{self.lib.alisb_backend().decode()}: {self.lib.alisb_last_error().decode()} (rc={rc})")

    def close(self):
        if self.h:
            self.lib.alisb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- coefficients -------------------------------------------------------------------
    def set_axis_tables(self, set_, axis, dtd, P=None, P1=None, P2=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float32) for a in (dtd, P, P1, P2)]
        self._ck(self.lib.alisb_set_axis_tables(self.h, set_, axis, *[_fp(a) for a in arrs]))

    def set_inv_eps(self, cx, cy, cz):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float32) for a in (cx, cy, cz)]
        self._ck(self.lib.alisb_set_inv_eps(self.h, *[_fp(a) for a in arrs]))

    def add_dispersive(self, box, nj, nk, a, coef, fill):
        d = Dispersive()
        d.box = _ip(box)
        d.nj, d.nk = nj, nk
        d.a = (C.c_float * 3)(*a)
        for name in ("JJ", "JK", "JE", "KJ", "KK", "KE"):
            arr = getattr(d, name)
            for c in range(3):
                for p in range(nk):
                    arr[c][p] = float(coef[name][c][p])
        keep = [np.ascontiguousarray(f, dtype=np.float32) for f in fill]
        for c in range(3):
            d.fill[c] = _fp(keep[c])
        self._ck(self.lib.alisb_add_dispersive(self.h, C.byref(d)))

    def add_source(self, set_, comp, m, box, waveform, n_active, amplitude=1.0, phase=0.0, Amplitude=None, Phase=None):
        """waveform: python callable (t, phase) -> float, evaluated on the host every step."""
        s = Source()
        s.set, s.comp, s.m = set_, comp, m
        s.box = _ip(box)
        s.amplitude, s.phase = amplitude, phase
        A = None if Amplitude is None else np.ascontiguousarray(Amplitude, dtype=np.float32)
        P = None if Phase is None else np.ascontiguousarray(Phase, dtype=np.float32)
        s.Amplitude, s.Phase = _fp(A), _fp(P)
        s.n_active = int(n_active)
        cb = WAVEFORM_FN(lambda user, t, ph: float(waveform(t, ph)))
        self._keep.append(cb)
        s.waveform = cb
        self._ck(self.lib.alisb_add_source(self.h, C.byref(s)))

    def set_copies(self, set_, ops):
        arr = (Copy * max(1, len(ops)))()
        for q, (cs, comp, cosk, sink, box, alt) in enumerate(ops):
            arr[q].set, arr[q].comp, arr[q].cosk, arr[q].sink = cs, comp, cosk, sink
            arr[q].box = _ip(box)
            arr[q].alt = _ip(alt)
        self._ck(self.lib.alisb_set_copies(self.h, set_, arr, len(ops)))

    # --- stepping -------------------------------------------------------------------------
    def step_E(self):
        self._ck(self.lib.alisb_step_E(self.h))

    def step_H(self):
        self._ck(self.lib.alisb_step_H(self.h))

    def step_part(self, set_, part):
        """part: 0 whole pass, 1 only the plane the z-neighbour needs, 2 the rest (see alis_b200.h)"""
        self._ck(self.lib.alisb_step_part(self.h, set_, part))

    def run(self, steps):
        self._ck(self.lib.alisb_run(self.h, steps))

    def synchronize(self):
        self._ck(self.lib.alisb_synchronize(self.h))

    def set_stream(self, cuda_stream):
        self._ck(self.lib.alisb_set_stream(self.h, C.c_void_p(cuda_stream)))

    def time(self):
        n, t, tE, tH = C.c_uint64(), C.c_float(), C.c_float(), C.c_float()
        self._ck(self.lib.alisb_get_time(self.h, C.byref(n), C.byref(t), C.byref(tE), C.byref(tH)))
        return n.value, t.value, tE.value, tH.value

    def launch_count(self):
        return int(self.lib.alisb_launch_count(self.h))

    def set_profiling(self, on):
        self._ck(self.lib.alisb_set_profiling(self.h, int(on)))

    def last_run_kernel_ms(self):
        ms, n = C.c_float(), C.c_int()
        self._ck(self.lib.alisb_last_run_kernel_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def profile_dominant(self):
        """(avg ms per launch, launches, algorithmic bytes per launch) of the dominant kernel"""
        ms, n, by = C.c_float(), C.c_int(), C.c_double()
        self._ck(self.lib.alisb_profile_dominant(self.h, C.byref(ms), C.byref(n), C.byref(by)))
        return (ms.value / n.value if n.value else 0.0), n.value, by.value

    # --- data -----------------------------------------------------------------------------
    def read_box(self, arr, ids, box):
        ids = list(ids) + [0] * (6 - len(ids))
        shape = (box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1)
        out = np.empty(shape, dtype=np.float32)
        self._ck(self.lib.alisb_read_box(self.h, arr, _ip(ids), _ip(box), _fp(out)))
        return out

    def write_box(self, arr, ids, box, data):
        ids = list(ids) + [0] * (6 - len(ids))
        data = np.ascontiguousarray(data, dtype=np.float32)
        assert data.shape == (box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1)
        self._ck(self.lib.alisb_write_box(self.h, arr, _ip(ids), _ip(box), _fp(data)))

    def full_box(self):
        return [0, self.nx + 1, 0, self.ny + 1, 0, self.nz + 1]

    def read_field(self, set_, comp, m=0):
        """(nx+2, ny+2, nz+2) array in the reference's [a][b][c] order."""
        return self.read_box(ARR_FIELD, [set_, comp, m], self.full_box())

    def write_field(self, set_, comp, m, data):
        self.write_box(ARR_FIELD, [set_, comp, m], self.full_box(), data)

    def read_all_fields(self):
        ncf = self.grid.complex_field + 1
        return np.stack([np.stack([np.stack([self.read_field(s, c, m) for c in range(3)]) for s in range(2)]) for m in range(ncf)])

    # --- monitor reductions ------------------------------------------------------------------
    def set_spacing(self, set_, axis, d):
        d = np.ascontiguousarray(d, dtype=np.float32)
        self._ck(self.lib.alisb_set_spacing(self.h, set_, axis, _fp(d)))

    def plane_flux(self, planes):
        """planes: [(axis, at, b0, b1, c0, c1)] -> array (n, 2) of the two partial sums per plane"""
        arr = np.ascontiguousarray(np.array(planes, dtype=np.int32).reshape(-1, 6))
        out = np.zeros((arr.shape[0], 2), np.float64)
        self._ck(self.lib.alisb_plane_flux(self.h, arr.ctypes.data_as(C.c_void_p), arr.shape[0], out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def field_fn(self, fn, set_=0, comp=0, comp2=0, m=0, sym=(0, 0, 0), cosk=(1.0, 1.0, 1.0), org=(0, 0, 0)):
        f = FieldFn()
        f.fn, f.set, f.comp, f.comp2, f.m = fn, set_, comp, comp2, m
        f.sym, f.cosk, f.org = (C.c_int32 * 3)(*sym), (C.c_float * 3)(*cosk), (C.c_int32 * 3)(*org)
        return f

    def eval_box(self, f, ibox):
        """field function f (field_fn(...)) on the inclusive INDEX box -> (ni, nj, nk) float32"""
        out = np.empty((ibox[1] - ibox[0] + 1, ibox[3] - ibox[2] + 1, ibox[5] - ibox[4] + 1), np.float32)
        self._ck(self.lib.alisb_eval_box(self.h, C.byref(f), _ip(ibox), _fp(out)))
        return out

    def accum(self, f, ibox, with_imag=True):
        """running sums of field function f over the INDEX box kept by the engine (sliceTimeAvg / sliceFreqDom)"""
        return Accum(self, f, ibox, with_imag)

    def reduce_absmax(self, f, ibox):
        out = np.zeros(1, np.float32)
        self._ck(self.lib.alisb_reduce_absmax(self.h, C.byref(f), _ip(ibox), _fp(out)))
        return out[0]

    def column_sums(self, f, ibox):
        out = np.empty((ibox[1] - ibox[0] + 1, ibox[3] - ibox[2] + 1), np.float32)
        self._ck(self.lib.alisb_column_sums(self.h, C.byref(f), _ip(ibox), _fp(out)))
        return out

    def absorption_columns(self, obj, box):
        """(na, nb, 3): the z sums xs, ys, zs of E.J per column of ADE object obj (output.c:375-404)"""
        out = np.empty((box[1] - box[0] + 1, box[3] - box[2] + 1, 3), np.float32)
        self._ck(self.lib.alisb_absorption_columns(self.h, obj, _ip(box), _fp(out)))
        return out

    def set_fusion(self, on):
        """fused E+H stepping where a step is only the two curls (include/alis_b200.h)"""
        self._ck(self.lib.alisb_set_fusion(self.h, 1 if on else 0))

    def fused_launch_count(self):
        return int(self.lib.alisb_fused_launch_count(self.h))

    def flush(self, h_pass_follows=False):
        self._ck(self.lib.alisb_flush(self.h, 1 if h_pass_follows else 0))

    def begin_pass(self, set_):
        self._ck(self.lib.alisb_begin_pass(self.h, set_))

    def chi2_couple(self, obj, m, pump1, pump2, d, alt):
        """coupleTWM (update.c:356-391): mixes E(pump1) x E(pump2) into this engine's E inside ADE object obj"""
        d = np.ascontiguousarray(d, dtype=np.float32).reshape(18)
        self._ck(self.lib.alisb_chi2_couple(self.h, obj, m, pump1.h, pump2.h, _fp(d), alt))

    # --- halo -----------------------------------------------------------------------------
    def halo_info(self, set_, m, comp):
        sp, rp, nb = C.c_void_p(), C.c_void_p(), C.c_size_t()
        self._ck(self.lib.alisb_halo_info(self.h, set_, m, comp, C.byref(sp), C.byref(rp), C.byref(nb)))
        return sp.value, rp.value, nb.value

    def halo_pack(self, set_):
        self._ck(self.lib.alisb_halo_pack(self.h, set_))

    def halo_unpack(self, set_):
        self._ck(self.lib.alisb_halo_unpack(self.h, set_))

    def halo_exchange_with_upper(self, upper, set_):
        """both slabs in this process: my top E plane -> upper's ghost plane (set E) / upper's bottom H plane -> my ghost plane (set H)"""
        self._ck(self.lib.alisb_halo_exchange(self.h, upper.h, set_))


class Accum:
    def __init__(self, eng, f, ibox, with_imag):
        self.eng, self.shape, self.with_imag = eng, (ibox[1] - ibox[0] + 1, ibox[3] - ibox[2] + 1, ibox[5] - ibox[4] + 1), with_imag
        self.h = C.c_void_p()
        eng._ck(eng.lib.alisb_accum_create(eng.h, C.byref(f), _ip(ibox), int(with_imag), C.byref(self.h)))

    def add(self, wr, wi=0.0):
        self.eng._ck(self.eng.lib.alisb_accum_add(self.h, wr, wi))

    def read(self):
        re = np.empty(self.shape, np.float32)
        im = np.empty(self.shape, np.float32) if self.with_imag else None
        self.eng._ck(self.eng.lib.alisb_accum_read(self.h, _fp(re), _fp(im) if self.with_imag else None))
        return re, im

    def clear(self):
        self.eng._ck(self.eng.lib.alisb_accum_clear(self.h))

    def close(self):
        if self.h:
            self.eng.lib.alisb_accum_destroy(self.h)
            self.h = None


class Phaser:
    def __init__(self, engine: Engine, box, face_mask, nfreq):
        self.e, self.box, self.mask, self.nfreq = engine, list(box), face_mask, nfreq
        self.h = C.c_void_p()
        engine._ck(engine.lib.alisb_phaser_create(engine.h, _ip(box), face_mask, nfreq, C.byref(self.h)))

    def accumulate(self, expE, expH):
        a = np.ascontiguousarray(np.asarray(expE, dtype=np.complex64)).view(np.float32)
        b = np.ascontiguousarray(np.asarray(expH, dtype=np.complex64)).view(np.float32)
        self.e._ck(self.e.lib.alisb_phaser_accumulate(self.h, _fp(a), _fp(b)))

    def read(self, freq, face, which):
        box = list(self.box)
        ax, sd = face // 2, face % 2
        box[2 * ax] = box[2 * ax + 1] = self.box[2 * ax + 1] if sd else self.box[2 * ax]
        shape = (box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1)
        out = np.empty(shape + (2,), dtype=np.float32)
        self.e._ck(self.e.lib.alisb_phaser_read(self.h, freq, face, which, _fp(out)))
        return out.view(np.complex64)[..., 0]

    def face_extents(self, face):
        """(n_lo, n_hi): face extents along its lower / higher in-face axis"""
        n = [self.box[1] - self.box[0] + 1, self.box[3] - self.box[2] + 1, self.box[5] - self.box[4] + 1]
        ax = face // 2
        return n[1 if ax == 0 else 0], n[1 if ax == 2 else 2]

    def surface_sums(self, freq, jobs, weights):
        """jobs: [(face, which, w_lo, w_hi)] with offsets into `weights` (complex64, 1-D);
        returns complex128 per job: sum_u sum_v w_lo[u] A[u][v] w_hi[v]  (farfld.c:248-459)"""
        J = np.ascontiguousarray(np.array(jobs, dtype=np.int64).reshape(-1, 4).astype(np.uint32)).view(np.int32)
        W = np.ascontiguousarray(np.asarray(weights, dtype=np.complex64)).view(np.float32)
        out = np.zeros((J.shape[0], 2), np.float64)
        self.e._ck(self.e.lib.alisb_phaser_surface_sums(self.h, freq, J.ctypes.data_as(C.c_void_p), J.shape[0], _fp(W), W.size // 2,
                                                        out.ctypes.data_as(C.POINTER(C.c_double))))
        return out[:, 0] + 1j * out[:, 1]

    def close(self):
        if self.h:
            self.e.lib.alisb_phaser_destroy(self.h)
            self.h = C.c_void_p()


class GeometryDesc:
    """Flattened object list + coordinate tables for alisb_geometry_bounds / alisb_geometry_fill.

    objects: list of nested tuples (type_name, f(3x3 list), [children]); earlier objects win.
    coords(axis, position) -> coordinate (the world's itox/jtoy/ktoz) as a python callable."""

    def __init__(self, lib, objects, lo, hi, subn, coords, device=0):
        self.lib, self.device, self.subn = lib, device, subn
        nodes, aux, roots = [], [], []

        def add(obj, slot):
            name, f, kids = obj
            n = nodes[slot]
            n.type = SHAPE[name]
            for a in range(3):
                for b in range(3):
                    n.f[a][b] = float(f[a][b])
            if name == "ROTATION":
                n.aux = len(aux)
                for q in range(3):
                    ang = np.float32(np.float32(f[0][q]) * np.pi / 180.0)
                    aux.extend([np.cos(ang, dtype=np.float32), np.sin(ang, dtype=np.float32)])
            if name.startswith("CLATTICE"):
                n.aux = len(aux)
                cnt = int(f[0][0])
                for q in range(cnt):
                    th = np.float32(2 * np.pi * q / cnt)
                    aux.extend([np.cos(th, dtype=np.float32), np.sin(th, dtype=np.float32)])
            if kids:
                n.child, n.nchild = len(nodes), len(kids)
                first = len(nodes)
                nodes.extend(Shape() for _ in kids)
                for q, k in enumerate(kids):
                    add(k, first + q)

        for obj in objects:
            roots.append(len(nodes))
            nodes.append(Shape())
            add(obj, roots[-1])
        self._nodes = (Shape * len(nodes))(*nodes)
        self._roots = np.array(roots, np.int32)
        self._aux = np.array(aux if aux else [0.0], np.float32)
        slots = subn + 3
        edge = np.float32((subn - 1.0) / (subn * 2.0))
        self._coord = []
        for a in range(3):
            t = np.zeros((hi[a] - lo[a] + 1, 2, slots), np.float32)
            for q, idx in enumerate(range(lo[a], hi[a] + 1)):
                for half in range(2):
                    base = np.float32(idx - 0.5 * half)
                    t[q, half, 0] = coords(a, base)
                    t[q, half, 1] = coords(a, np.float32(base - edge))
                    t[q, half, 2] = coords(a, np.float32(base + edge))
                    for s_ in range(subn):
                        d = 1 - subn + 2 * s_
                        t[q, half, 3 + s_] = coords(a, np.float32(np.float64(base) + 0.5 * d / subn))
            self._coord.append(np.ascontiguousarray(t))
        g = Geometry()
        g.nobjects, g.roots = len(roots), self._roots.ctypes.data_as(i32p)
        g.nnodes, g.nodes = len(nodes), C.cast(self._nodes, C.POINTER(Shape))
        g.naux, g.aux = len(aux), _fp(self._aux)
        g.subn = subn
        g.lo, g.hi = (C.c_int32 * 3)(*lo), (C.c_int32 * 3)(*hi)
        for a in range(3):
            g.coord[a] = _fp(self._coord[a])
        self.g = g

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(f"{self.lib.alisb_backend().decode()}: {self.lib.alisb_last_error().decode()} (rc={rc})")

    def bounds(self, ibox):
        out = np.zeros((self.g.nobjects, 6), np.int32)
        self._ck(self.lib.alisb_geometry_bounds(self.device, C.byref(self.g), _ip(ibox), out.ctypes.data_as(i32p)))
        return out

    def fill(self, obj, ibox):
        shape = (ibox[1] - ibox[0] + 1, ibox[3] - ibox[2] + 1, ibox[5] - ibox[4] + 1)
        f = [np.empty(shape, np.float32) for _ in range(3)]
        self._ck(self.lib.alisb_geometry_fill(self.device, C.byref(self.g), obj, _ip(ibox), _fp(f[0]), _fp(f[1]), _fp(f[2])))
        return f
