"""ctypes binding of ``libmicmec_b200.so`` (the C ABI declared in ``include/micmec_b200.h``).

There is no CPU fallback: if the shared library is missing or no sm_100 device is present, every compute
entry point raises.  The library is built in-tree by ``micmec_b200/build.py`` (``__graft_entry__.build()``).
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# MICMEC_B200_LIB: load another build of the same sources (used only by the ablation timings under profiles/)
LIBPATH = os.environ.get("MICMEC_B200_LIB") or os.path.join(HERE, "libmicmec_b200.so")

MM_OK, MM_ERR_INVALID, MM_ERR_CUDA, MM_ERR_NAN, MM_ERR_STATE = 0, -1, -2, -3, -4
MM_HOST, MM_DEVICE = 0, 1
MM_MAX_TYPES, MM_MAX_STATES, MM_MAX_CHAIN = 8, 16, 8
MODELS = {"original": 0, "default": 1}

# indices into the array filled by mm_md_scalars
S_EPOT, S_EKIN, S_TEMP, S_ETOT, S_ECONS, S_CONS_ERR, S_PRESS, S_RMSD_GPOS, S_RMSD_DELTA = range(9)
S_TIME, S_COUNTER, S_VOLUME, S_NDOF, S_ECONS_CORR = 9, 10, 11, 12, 13
S_VTENS, S_PTENS, S_NFORCE, S_CE_N, S_COUNT = 16, 25, 34, 35, 40

EXPORTS = [
    "mm_create", "mm_destroy", "mm_last_error", "mm_version", "mm_device_ok", "mm_set_pos", "mm_set_rvecs",
    "mm_compute", "mm_get_cell_cache", "mm_launch_count", "mm_device_ptr", "mm_set_stream", "mm_synchronize",
    "mm_set_option", "mm_get_option", "mm_plan_schedule", "mm_profile", "mm_batched_eigh", "mm_qn_create", "mm_qn_destroy", "mm_qn_sweep", "mm_qn_get", "mm_set_rvecs_batch", "mm_get_replica_results", "mm_comm_unique_id", "mm_comm_init", "mm_comm_destroy", "mm_comm_mode", "mm_domain", "mm_md_create", "mm_md_destroy", "mm_md_init", "mm_md_set_state", "mm_md_run", "mm_md_get_state",
    "mm_md_scalars",
]


class Desc(ctypes.Structure):
    _fields_ = [
        ("nnodes", ctypes.c_int64),
        ("ncells", ctypes.c_int64),
        ("surrounding_nodes", ctypes.c_void_p),
        ("surrounding_cells", ctypes.c_void_p),
        ("shift", ctypes.c_void_p),
        ("cell_type", ctypes.c_void_p),
        ("ntypes", ctypes.c_int32),
        ("type_nstates", ctypes.c_void_p),
        ("h0", ctypes.c_void_p),
        ("elasticity", ctypes.c_void_p),
        ("free_energy", ctypes.c_void_p),
        ("effective_temp", ctypes.c_void_p),
        ("boltzmann", ctypes.c_double),
        ("model", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("nx", ctypes.c_int32),
        ("ny", ctypes.c_int32),
        ("nz", ctypes.c_int32),
        ("slab_rank", ctypes.c_int32),
        ("slab_count", ctypes.c_int32),
        ("nnodes_global", ctypes.c_int64),
        ("nreplicas", ctypes.c_int64),
    ]


class MDDesc(ctypes.Structure):
    _fields_ = [
        ("timestep", ctypes.c_double),
        ("ndof", ctypes.c_double),
        ("has_thermo", ctypes.c_int32),
        ("chain_length", ctypes.c_int32),
        ("thermo_temp", ctypes.c_double),
        ("thermo_timecon", ctypes.c_double),
        ("has_baro", ctypes.c_int32),
        ("anisotropic", ctypes.c_int32),
        ("vol_constraint", ctypes.c_int32),
        ("baro_temp", ctypes.c_double),
        ("baro_press", ctypes.c_double),
        ("baro_timecon", ctypes.c_double),
        ("has_langevin", ctypes.c_int32),
        ("thermo_kind", ctypes.c_int32),
        ("langevin_temp", ctypes.c_double),
        ("langevin_timecon", ctypes.c_double),
        ("langevin_seed", ctypes.c_uint64),
        ("time0", ctypes.c_double),
        ("counter0", ctypes.c_int64),
    ]


_lib = None


def load():
    """Load the shared library (raises ``RuntimeError`` with build instructions when it is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise RuntimeError(
            "libmicmec_b200.so has not been built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python micmec_b200/build.py` (nvcc, sm_100a).  There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIBPATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    lib.mm_last_error.restype = ctypes.c_char_p
    lib.mm_create.argtypes = [ctypes.POINTER(Desc), ctypes.POINTER(vp)]
    lib.mm_destroy.argtypes = [vp]
    lib.mm_device_ok.argtypes = [i32]
    lib.mm_set_pos.argtypes = [vp, vp, i32]
    lib.mm_set_rvecs.argtypes = [vp, vp]
    lib.mm_compute.argtypes = [vp, ctypes.POINTER(dbl), vp, i32, vp]
    lib.mm_get_cell_cache.argtypes = [vp, vp, vp]
    lib.mm_launch_count.argtypes = [vp]
    lib.mm_launch_count.restype = i64
    lib.mm_device_ptr.argtypes = [vp, i32]
    lib.mm_device_ptr.restype = vp
    lib.mm_set_stream.argtypes = [vp, vp]
    lib.mm_synchronize.argtypes = [vp]
    lib.mm_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    lib.mm_get_option.argtypes = [vp, ctypes.c_char_p]
    lib.mm_get_option.restype = i64
    lib.mm_plan_schedule.argtypes = [ctypes.c_int] * 6 + [vp, i64, ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.mm_profile.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(dbl)]  # int64[2], double[2]
    lib.mm_set_rvecs_batch.argtypes = [vp, vp]
    lib.mm_get_replica_results.argtypes = [vp, vp, vp]
    lib.mm_comm_unique_id.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    lib.mm_comm_init.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p]
    lib.mm_comm_destroy.argtypes = [vp]
    lib.mm_comm_mode.argtypes = [vp]
    lib.mm_batched_eigh.argtypes = [i32, i64, i32, vp, i32, vp, vp, ctypes.POINTER(i32)]
    lib.mm_qn_create.argtypes = [vp, i32, vp, vp, vp, vp, dbl, dbl, dbl, dbl, dbl, dbl, dbl, ctypes.POINTER(vp)]
    lib.mm_qn_destroy.argtypes = [vp]
    lib.mm_qn_sweep.argtypes = [vp, i32, ctypes.POINTER(i32)]
    lib.mm_qn_get.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.POINTER(i64)]
    lib.mm_domain.argtypes = [vp, i32, ctypes.POINTER(dbl), vp]
    lib.mm_md_create.argtypes = [vp, ctypes.POINTER(MDDesc), ctypes.POINTER(vp)]
    lib.mm_md_destroy.argtypes = [vp]
    lib.mm_md_init.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp]
    lib.mm_md_set_state.argtypes = [vp, vp, vp, i32]
    lib.mm_md_run.argtypes = [vp, i64]
    lib.mm_md_get_state.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp]
    lib.mm_md_scalars.argtypes = [vp, vp]
    _lib = lib
    return lib


def last_error():
    return load().mm_last_error().decode("utf-8", "replace")


def check(rc):
    """Translate a C status into the exception type the reference raises on the same condition."""
    if rc == MM_OK:
        return
    msg = last_error()
    if rc in (MM_ERR_INVALID, MM_ERR_NAN):
        raise ValueError(msg)  # mmff.py:135-147, 251
    raise RuntimeError(msg)


def ptr(arr):
    """Pointer to a NumPy array (host) or the ``data_ptr`` of a torch tensor (device); ``None`` -> NULL."""
    if arr is None:
        return None
    if isinstance(arr, np.ndarray):
        return arr.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(arr.data_ptr())


def where(arr):
    if arr is None or isinstance(arr, np.ndarray):
        return MM_HOST
    return MM_DEVICE if arr.is_cuda else MM_HOST
