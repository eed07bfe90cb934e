"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracle and the compiled reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (osiris_b200/) never does.

Two checkers live here:
  * Oracle(prec)  -- oracle/pic_oracle.c, our restatement of the reference algorithm;
  * RefSimd(isa, prec) -- the reference's own C SIMD kernels compiled by oracle/Makefile from
    /root/reference/source/simd into oracle/_ref/ (prebuilt files travel to the GPU box).
All arrays use the reference's Fortran layouts: x(D,np) -> numpy (np, D); grids
f(3,S1,S2,S3) -> numpy (S3, S2, S1, 3).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the oracle (and, when /root/reference exists, oracle/_ref).  Building the checker
    is not using it."""
    libs = [os.path.join(HERE, f) for f in ("libpic_oracle_f32.so", "libpic_oracle_f64.so", "libcpu_driver.so")]
    srcs = [os.path.join(HERE, f) for f in ("pic_oracle.c", "cpu_driver.c")]
    need = force or not all(os.path.exists(f) for f in libs)
    if not need:  # a restatement edited after the last build
        need = max(os.path.getmtime(f) for f in srcs) > min(os.path.getmtime(f) for f in libs)
    if os.path.isdir("/root/reference/source/simd") and not os.path.exists(
        os.path.join(REFDIR, "libref_avx2_f32.so")
    ):
        need = True
    if need:
        subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)


def dtype_of(prec):
    return np.float32 if prec in ("f32", 4, np.float32) else np.float64


class Oracle:
    def __init__(self, prec="f32"):
        build()
        self.dtype = dtype_of(prec)
        name = "libpic_oracle_f32.so" if self.dtype == np.float32 else "libpic_oracle_f64.so"
        self.lib = C.CDLL(os.path.join(HERE, name))
        self.creal = C.c_float if self.dtype == np.float32 else C.c_double
        rp = np.ctypeslib.ndpointer(self.dtype, flags="C_CONTIGUOUS")
        L = self.lib
        L.oracle_advance_deposit.argtypes = [
            C.c_int, C.c_int, _i32p, rp, rp, rp, C.c_int, C.c_int, self.creal, rp, rp,
            _i32p, _i32p, rp, _i32p, _i32p, _f64p, C.c_double, _f64p]
        L.oracle_advance_deposit.restype = None
        L.oracle_accelerate_deposit.argtypes = [
            C.c_int, C.c_int, _i32p, rp, rp, rp, C.c_long, self.creal, self.creal, self.creal, C.c_int,
            rp, _i32p, _i32p, _f64p, C.c_double, _f64p]
        L.oracle_accelerate_deposit.restype = None
        for fn in (L.oracle_dbdt,):
            fn.argtypes = [C.c_int, rp, rp, _i32p, _i32p, _i32p, _f64p, C.c_double]
            fn.restype = None
        L.oracle_dedt.argtypes = [C.c_int, rp, rp, rp, _i32p, _i32p, _i32p, _f64p, C.c_double]
        L.oracle_dedt.restype = None
        L.oracle_yee_advance.argtypes = [C.c_int, rp, rp, rp, _i32p, _i32p, _i32p, _f64p, C.c_double]
        L.oracle_yee_advance.restype = None
        L.oracle_fei_advance.argtypes = [C.c_int, rp, rp, rp, _i32p, _i32p, _i32p, _f64p, C.c_double, C.c_int, rp, rp]
        L.oracle_fei_advance.restype = None
        L.oracle_update_periodic.argtypes = [C.c_int, rp, C.c_int, _i32p, _i32p, _i32p, C.c_int, C.c_int]
        L.oracle_update_periodic.restype = None
        L.oracle_particles_periodic.argtypes = [C.c_int, _i32p, C.c_long, _i32p]
        L.oracle_particles_periodic.restype = None
        L.oracle_sort_idx.argtypes = [C.c_int, _i32p, C.c_long, _i32p, _i64p]
        L.oracle_sort_idx.restype = None
        L.oracle_field_energy.argtypes = [C.c_int, rp, _i32p, _i32p, _i32p, _f64p, _f64p]
        L.oracle_field_energy.restype = None
        L.oracle_deposit_rho.argtypes = [C.c_int, C.c_int, _i32p, rp, rp, C.c_long, _f64p, _i32p, _i32p]
        L.oracle_deposit_rho.restype = None
        L.oracle_move_window_vdf.argtypes = [C.c_int, rp, C.c_int, _i32p, _i32p, _i32p, _i32p]
        L.oracle_move_window_vdf.restype = None
        L.oracle_particles_bc.argtypes = [C.c_int, _i32p, rp, rp, C.c_long, _i32p, _i32p,
                                          np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")]
        L.oracle_particles_bc.restype = None
        L.oracle_smooth_setup.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, _i32p, _i32p, rp]
        L.oracle_smooth_setup.restype = None
        L.oracle_smooth.argtypes = [C.c_int, rp, _i32p, _i32p, _i32p, _i32p, C.c_int, rp]
        L.oracle_smooth.restype = None
        L.oracle_spline.argtypes = [C.c_int, self.creal, rp]
        L.oracle_splineh.argtypes = [C.c_int, self.creal, rp]
        L.oracle_ntrim.argtypes = [self.creal]
        L.oracle_ntrim.restype = C.c_int

    # -- helpers ---------------------------------------------------------------------------
    @staticmethod
    def _i(a):
        return np.ascontiguousarray(a, dtype=np.int32)

    def spline(self, order, x, half=False):
        w = np.zeros(order + 1, self.dtype)
        (self.lib.oracle_splineh if half else self.lib.oracle_spline)(order, x, w)
        return w

    def ntrim(self, x):
        return self.lib.oracle_ntrim(x)

    def advance_deposit(self, dim, order, ix, x, u, q, rqm, e, b, emf_gc, j, j_gc, dx, dt,
                        i0=1, i1=None):
        """In place on ix, x, u, j.  e/b/j: numpy grids (.., S2, S1, 3); *_gc = (gcl[D], gcu[D]).
        Returns the time-centred kinetic energy sum(q * u^2/(gamma+1))."""
        n = q.shape[0] if i1 is None else i1
        esz = self._i(e.shape[::-1][1:1 + dim])
        jsz = self._i(j.shape[::-1][1:1 + dim])
        ene = np.zeros(1, np.float64)
        self.lib.oracle_advance_deposit(
            dim, order, ix, x, u, q, i0, n, rqm, e, b, esz, self._i(emf_gc[0]), j, jsz,
            self._i(j_gc[0]), np.ascontiguousarray(dx, np.float64), float(dt), ene)
        return float(ene[0])

    def accelerate_deposit(self, dim, order, ix, x, u, q, u_frac, q_frac, ufl1, use_particle_uacc, j, j_gc, dx, dt):
        """accelerate_deposit (spec/os-spec-accel.f90): in place on ix, x, j; returns sum(q u^2/(gamma+1))."""
        jsz = self._i(j.shape[::-1][1:1 + dim])
        ene = np.zeros(1, np.float64)
        self.lib.oracle_accelerate_deposit(dim, order, ix, x, u, q, q.shape[0], u_frac, q_frac, ufl1,
                                           int(bool(use_particle_uacc)), j, jsz, self._i(j_gc[0]),
                                           np.ascontiguousarray(dx, np.float64), float(dt), ene)
        return float(ene[0])

    def yee_advance(self, dim, e, b, j, nx, gc, dx, dt):
        self.lib.oracle_yee_advance(dim, e, b, j, self._i(nx), self._i(gc[0]), self._i(gc[1]),
                                    np.ascontiguousarray(dx, np.float64), float(dt))

    def fei_advance(self, dim, e, b, j, nx, gc, dx, dt, coef_e, coef_b):
        """advance_fei (emf/solver/os-emf-solver-fei.f03:2331-2363) with the given stencil coefficients"""
        ce, cb = np.ascontiguousarray(coef_e, self.dtype), np.ascontiguousarray(coef_b, self.dtype)
        self.lib.oracle_fei_advance(dim, e, b, j, self._i(nx), self._i(gc[0]), self._i(gc[1]),
                                    np.ascontiguousarray(dx, np.float64), float(dt), len(ce), ce, cb)

    def dbdt(self, dim, b, e, nx, gc, dx, dt):
        self.lib.oracle_dbdt(dim, b, e, self._i(nx), self._i(gc[0]), self._i(gc[1]),
                             np.ascontiguousarray(dx, np.float64), float(dt))

    def dedt(self, dim, e, b, j, nx, gc, dx, dt):
        self.lib.oracle_dedt(dim, e, b, j, self._i(nx), self._i(gc[0]), self._i(gc[1]),
                             np.ascontiguousarray(dx, np.float64), float(dt))

    def update_periodic(self, dim, f, nx, gc, mode, axes=None):
        """mode 'replace' | 'add'; axes sequentially x, y, z as update_boundary_vdf does."""
        m = 0 if mode == "replace" else 1
        for ax in (range(dim) if axes is None else axes):
            self.lib.oracle_update_periodic(dim, f, f.shape[-1], self._i(nx), self._i(gc[0]),
                                            self._i(gc[1]), ax, m)

    def particles_periodic(self, dim, ix, nx):
        self.lib.oracle_particles_periodic(dim, ix, ix.shape[0], self._i(nx))

    def sort_idx(self, dim, ix, nx):
        idx = np.zeros(ix.shape[0], np.int64)
        self.lib.oracle_sort_idx(dim, ix, ix.shape[0], self._i(nx), idx)
        return idx

    def field_energy(self, dim, f, nx, gc, dx):
        out = np.zeros(3, np.float64)
        self.lib.oracle_field_energy(dim, f, self._i(nx), self._i(gc[0]), self._i(gc[1]),
                                     np.ascontiguousarray(dx, np.float64), out)
        return out

    def deposit_rho(self, dim, order, ix, x, q, grid_shape, gc):
        rho = np.zeros(grid_shape, np.float64)
        size = self._i(grid_shape[::-1][:dim])
        self.lib.oracle_deposit_rho(dim, order, ix, x, q, q.shape[0], rho, size, self._i(gc[0]))
        return rho


    def move_window_vdf(self, dim, f, nx, gc, nmove):
        nm = np.zeros(3, np.int32)
        nm[:dim] = nmove
        self.lib.oracle_move_window_vdf(dim, f, f.shape[-1], self._i(nx), self._i(gc[0]), self._i(gc[1]), nm)

    PART_BC = {"periodic": 1, "absorbing": 2, "specular": 3}

    def particles_bc(self, dim, ix, x, p, nx, bc):
        """bc = ([lower per axis], [upper per axis]) of 'periodic' | 'absorbing' | 'specular'; in place on
        ix, x, p; returns the keep mask."""
        code = np.zeros(6, np.int32)
        for sd in range(2):
            for d in range(dim):
                code[sd * 3 + d] = self.PART_BC[bc[sd][d]]
        keep = np.ones(ix.shape[0], np.uint8)
        self.lib.oracle_particles_bc(dim, ix, x, p, ix.shape[0], self._i(nx), code, keep)
        return keep.astype(bool)

    SMOOTH_TYPES = {"none": 0, "custom": 1, "binomial": 2, "compensated": 3, "5pass": 4, "tri6": 5,
                    "tri8": 6, "digital": 11}

    def smooth_setup(self, dim, types, orders, swfj=None, digital_fc=None, digital_A=None):
        """setup_smooth: returns (order[dim], max_order, scoef[dim, 2*max_order+1])."""
        ty = self._i([self.SMOOTH_TYPES[t] for t in types])
        sw = np.zeros((3, 15, 3), np.float64) if swfj is None else np.ascontiguousarray(swfj, np.float64)
        fc = np.ascontiguousarray(digital_fc if digital_fc is not None else [0.5] * 3, np.float64)
        A = np.ascontiguousarray(digital_A if digital_A is not None else [60.0] * 3, np.float64)
        order = np.zeros(3, np.int32)
        mo = np.zeros(1, np.int32)
        sc = np.zeros((3, 31), self.dtype)
        self.lib.oracle_smooth_setup(dim, ty, self._i(orders), sw, fc, A, order, mo, sc.reshape(-1))
        m = int(mo[0])
        W = 2 * m + 1
        return order[:dim].copy(), m, sc.reshape(-1)[:dim * W].reshape(dim, W).copy()

    def smooth(self, dim, f, nx, gc, order, max_order, scoef):
        self.lib.oracle_smooth(dim, f, self._i(nx), self._i(gc[0]), self._i(gc[1]), self._i(order),
                               int(max_order), np.ascontiguousarray(scoef, self.dtype))


class RefSimd:
    """The reference's own compiled kernels (source/simd), e.g. RefSimd('avx2', 'f32').
    Signature of every entry point: simd/os-spec-push-avx2.c:1141-1145."""

    def __init__(self, isa="avx2", prec="f32"):
        build()
        self.dtype = dtype_of(prec)
        tag = "f32" if self.dtype == np.float32 else "f64"
        path = os.path.join(REFDIR, f"libref_{isa}_{tag}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.isa = isa

    @staticmethod
    def available(isa="avx2", prec="f32"):
        tag = "f32" if dtype_of(prec) == np.float32 else "f64"
        return os.path.exists(os.path.join(REFDIR, f"libref_{isa}_{tag}.so"))

    def advance_deposit(self, dim, order, ix, x, u, q, rqm, e, b, emf_gc, j, j_gc, dx, dt,
                        i0=1, i1=None):
        """Arrays must be 32-byte aligned and padded to a multiple of 16 particles (the callee
        writes dummy particles past i1).  In place.  Returns the energy accumulated."""
        fn = getattr(self.lib, f"vadvance_deposit_{dim}d_s{order}_")
        n = q.shape[0] if i1 is None else i1
        for a in (ix, x, u, q, e, b, j):
            assert a.ctypes.data % 32 == 0 and a.flags.c_contiguous
        esz = np.ascontiguousarray(e.shape[::-1][1:1 + dim], np.int32)
        jsz = np.ascontiguousarray(j.shape[::-1][1:1 + dim], np.int32)
        eoff = np.ascontiguousarray(emf_gc[0], np.int32)
        joff = np.ascontiguousarray(j_gc[0], np.int32)
        dxa = np.ascontiguousarray(dx, np.float64)
        ene = C.c_double(0.0)
        rq = (C.c_float if self.dtype == np.float32 else C.c_double)(rqm)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        fn(p(ix), p(x), p(u), p(q), C.byref(C.c_int(i0)), C.byref(C.c_int(n)), C.byref(rq),
           p(e), p(b), p(esz), p(eoff), p(j), p(jsz), p(joff), p(dxa),
           C.byref(C.c_double(dt)), C.byref(ene))
        return ene.value


def aligned(shape, dtype, pad_rows=16):
    """32-byte aligned zeroed array whose first axis has `pad_rows` spare rows behind it (the
    reference pads particle buffers to p_vecwidth = 16, system/os-sys-multi.f03:59)."""
    shape = tuple(np.atleast_1d(shape))
    n0 = shape[0] + pad_rows
    count = int(np.prod((n0,) + shape[1:]))
    raw = np.zeros(count * np.dtype(dtype).itemsize + 64, np.uint8)
    off = (-raw.ctypes.data) % 32
    full = raw[off:off + count * np.dtype(dtype).itemsize].view(dtype).reshape((n0,) + shape[1:])
    return full[:shape[0]]
