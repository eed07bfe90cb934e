"""ctypes binding of include/osiris_b200.h -- the reference-side stub a Python host would add.

The library is the product: if libosiris_b200.so is missing this module raises, and every compute
entry point returns OSGPU_ERR_NO_DEVICE without a GPU.  There is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("OSGPU_LIB") or os.path.join(HERE, "libosiris_b200.so")  # (OSGPU_LIB: kernel-variant A/B runs)

OSGPU_OK = 0
ERR_NAMES = {-1: "INVALID", -2: "CUDA", -3: "NCCL", -4: "CAPACITY", -5: "STABILITY", -6: "NO_DEVICE"}
NCCL_ID_BYTES = 128


class OsgpuError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"osgpu error {status} ({ERR_NAMES.get(status, '?')}): {text}")
        self.status = status


class Config(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("order", C.c_int32), ("real_bytes", C.c_int32), ("nspecies", C.c_int32),
        ("nx", C.c_int32 * 3), ("gc_emf", (C.c_int32 * 3) * 2), ("gc_jay", (C.c_int32 * 3) * 2),
        ("dx", C.c_double * 3), ("dt", C.c_double),
        ("node_dims", C.c_int32 * 3), ("node_coords", C.c_int32 * 3), ("periodic", C.c_int32 * 3),
        ("rank", C.c_int32), ("nranks", C.c_int32), ("device", C.c_int32), ("smooth_order", C.c_int32),
        ("no_tags", C.c_int32), ("part_bc", (C.c_int32 * 3) * 2),
        ("topology", C.c_int32), ("neighbor", (C.c_int32 * 3) * 2), ("emf_bc", (C.c_int32 * 3) * 2),
    ]


class Phasespace(C.Structure):
    """osgpu_phasespace (include/osiris_b200.h)"""
    _fields_ = [
        ("ndims", C.c_int32), ("x_or_p", C.c_int32 * 3), ("xp_dim", C.c_int32 * 3), ("ps_type", C.c_int32),
        ("n", C.c_int32 * 3), ("use_ene_bin", C.c_int32),
        ("min", C.c_double * 3), ("max", C.c_double * 3), ("ene_bin", C.c_double * 2),
        ("xmin", C.c_double * 3), ("center", C.c_double * 3), ("ixmin", C.c_int32 * 3),
    ]


TOPO_MPI, TOPO_OLD, TOPO_HOST = 0, 1, 2


_lib = None


def load():
    """Load libosiris_b200.so (built in-tree by osiris_b200/build.py); raise if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise ImportError(
            f"{LIBPATH} not found: build it with `python -m osiris_b200.build` "
            "(the CUDA library is the product; there is no fallback)")
    L = C.CDLL(LIBPATH)
    vp, i32p, i64p, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)
    L.osgpu_version.restype = C.c_char_p
    L.osgpu_last_error.restype = C.c_char_p
    L.osgpu_cart_rank.argtypes = [C.c_int, C.c_int, i32p, i32p]
    L.osgpu_cart_coords.argtypes = [C.c_int, C.c_int, i32p, C.c_int, i32p]
    L.osgpu_cart_neighbors.argtypes = [C.c_int, C.c_int, i32p, i32p, C.c_int, i32p]
    L.osgpu_exchange_ranges.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    L.osgpu_exchange_ranges_move.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    L.osgpu_min_gc.argtypes = [C.c_int, i32p]
    L.osgpu_field_wall_op.argtypes = [C.c_int] * 8 + [i32p]
    L.osgpu_current_wall_op.argtypes = [C.c_int] * 4 + [i32p]
    L.osgpu_nccl_unique_id.argtypes = [vp]
    L.osgpu_init.argtypes = [C.POINTER(vp), C.POINTER(Config), vp]
    L.osgpu_finalize.argtypes = [vp]
    L.osgpu_species_create.argtypes = [vp, C.c_int, C.c_double, C.c_int64]
    L.osgpu_species_upload.argtypes = [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp]
    L.osgpu_species_upload_begin.argtypes = [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp]
    L.osgpu_species_upload_end.argtypes = [vp, C.c_int]
    L.osgpu_species_count.argtypes = [vp, C.c_int]
    L.osgpu_species_count.restype = C.c_int64
    L.osgpu_species_download.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    L.osgpu_species_download_async.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    L.osgpu_species_push.argtypes = [vp, C.c_int]
    L.osgpu_species_update_boundary.argtypes = [vp, C.c_int]
    L.osgpu_species_accelerate.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]
    L.osgpu_species_energy.argtypes = [vp, C.c_int, dp]
    L.osgpu_species_deposit_charge.argtypes = [vp, C.c_int, vp, C.c_int]
    L.osgpu_species_phasespace.argtypes = [vp, C.c_int, C.POINTER(Phasespace), vp, C.c_int]
    L.osgpu_current_zero.argtypes = [vp]
    L.osgpu_move_window.argtypes = [vp, vp]
    L.osgpu_species_inject.argtypes = [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp]
    L.osgpu_fields_upload.argtypes = [vp, vp, vp]
    L.osgpu_fields_download.argtypes = [vp, vp, vp]
    L.osgpu_current_upload.argtypes = [vp, vp]
    L.osgpu_current_download.argtypes = [vp, vp]
    L.osgpu_step.argtypes = [vp, C.c_int, dp]
    for name in ("osgpu_emf_update_boundary", "osgpu_particles_update_boundary", "osgpu_sort",
                 "osgpu_current_update_boundary", "osgpu_emf_advance", "osgpu_sync", "osgpu_current_smooth",
                 "osgpu_emf_update_particle_fld"):
        getattr(L, name).argtypes = [vp]
    L.osgpu_current_set_smooth.argtypes = [vp, vp, C.c_int, vp]
    L.osgpu_emf_set_solver.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.osgpu_emf_set_smooth.argtypes = [vp, vp, C.c_int, vp]
    L.osgpu_advance_deposit.argtypes = [vp, dp]
    L.osgpu_field_energy.argtypes = [vp, dp]
    L.osgpu_phase_times_ms.argtypes = [vp, dp]
    L.osgpu_launch_count.argtypes = [vp]
    L.osgpu_launch_count.restype = C.c_int64
    L.osgpu_stream.argtypes = [vp]
    L.osgpu_stream.restype = vp
    L.osgpu_host_alloc.argtypes = [C.c_size_t]
    L.osgpu_host_alloc.restype = vp
    L.osgpu_host_free.argtypes = [vp]
    L.osgpu_vadvance_deposit.argtypes = [
        C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_double, vp, vp, vp, vp, vp, vp, vp,
        vp, C.c_double, dp]
    _lib = L
    return L


def check(status):
    if status != OSGPU_OK:
        raise OsgpuError(status, load().osgpu_last_error().decode())


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pinned_empty(shape, dtype):
    """numpy array over page-locked memory (osgpu_host_alloc); keeps itself alive via .base."""
    L = load()
    shape = tuple(np.atleast_1d(shape))
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = L.osgpu_host_alloc(max(n, 1))
    if not p:
        raise OsgpuError(-2, L.osgpu_last_error().decode())
    buf = (C.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _pinned_keep[arr.ctypes.data] = p
    return arr


_pinned_keep = {}


def pinned_free(arr):
    p = _pinned_keep.pop(arr.ctypes.data, None)
    if p:
        load().osgpu_host_free(p)


EXPORTS = [
    "osgpu_version", "osgpu_last_error", "osgpu_device_count", "osgpu_cart_rank", "osgpu_cart_coords",
    "osgpu_cart_neighbors", "osgpu_exchange_ranges", "osgpu_min_gc", "osgpu_nccl_unique_id", "osgpu_init",
    "osgpu_finalize", "osgpu_species_create", "osgpu_species_upload", "osgpu_species_count",
    "osgpu_species_download", "osgpu_fields_upload", "osgpu_fields_download", "osgpu_current_upload",
    "osgpu_current_download", "osgpu_step", "osgpu_emf_update_boundary", "osgpu_advance_deposit",
    "osgpu_particles_update_boundary", "osgpu_sort", "osgpu_current_update_boundary", "osgpu_emf_advance",
    "osgpu_field_energy", "osgpu_phase_times_ms", "osgpu_launch_count", "osgpu_stream", "osgpu_sync",
    "osgpu_host_alloc", "osgpu_host_free", "osgpu_vadvance_deposit", "osgpu_current_set_smooth",
    "osgpu_current_smooth", "osgpu_current_zero", "osgpu_species_push", "osgpu_species_update_boundary",
    "osgpu_species_download_async", "osgpu_species_upload_begin", "osgpu_species_upload_end",
    "osgpu_emf_set_smooth", "osgpu_emf_update_particle_fld", "osgpu_move_window", "osgpu_species_inject",
    "osgpu_exchange_ranges_move", "osgpu_vadvance_deposit_release", "osgpu_species_accelerate",
    "osgpu_species_energy", "osgpu_species_deposit_charge", "osgpu_species_phasespace",
    "osgpu_emf_set_solver", "osgpu_field_wall_op", "osgpu_current_wall_op",
] + [f"osgpu_vadvance_deposit_{d}d_s{o}{s}" for d in (2, 3) for o in (1, 2, 3, 4) for s in ("", "_d")]
