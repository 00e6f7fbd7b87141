"""ctypes binding of libscb200.so (the C ABI declared in include/scb200.h).

The product path has NO CPU fallback: if the shared library is missing, or no
CUDA device is visible, every compute entry point raises.  PyTorch is used only
for device memory and streams.
"""

import ctypes as C
import os
from os.path import dirname, join, realpath

import numpy as np

__all__ = ["lib", "check", "device_available", "require_device", "FFDesc", "Patch", "stream_ptr",
           "SCB_FF", "LIB_PATH"]

LIB_PATH = join(dirname(realpath(__file__)), "lib", "libscb200.so")

SCB_OK = 0
SCB_ERR_INVALID = -1
SCB_ERR_CUDA = -2
SCB_ERR_ABOVE_CUTOFF = -3
SCB_ERR_WORKSPACE = -4
SCB_ERR_NOT_CONVERGED = -5
SCB_ERR_UNSUPPORTED = -6


class SCB_FF:
    INVARIANT = 0
    HINSEN = 1
    PFREE = 2
    TABULATED = 3
    TABULATED_DENSE = 4
    EXTERNAL = 5


class FFDesc(C.Structure):
    """struct scb_ff_desc (include/scb200.h)."""
    _fields_ = [
        ("kind", C.c_int32), ("nbins", C.c_int32), ("patched", C.c_int32), ("n_pair_on", C.c_int32),
        ("cutoff_sq", C.c_double),
        ("bonded", C.c_void_p), ("intra", C.c_void_p), ("inter", C.c_void_p), ("edges_sq", C.c_void_p),
        ("res_type", C.c_void_p), ("chain", C.c_void_p), ("bonded_next", C.c_void_p),
        ("dense_table", C.c_void_p), ("external_fc", C.c_void_p),
        ("pair_on", C.c_void_p), ("pair_on_fc", C.c_void_p),
    ]


class Patch(C.Structure):
    """struct scb_patch (include/scb200.h)."""
    _fields_ = [
        ("n_pair_off", C.c_int32), ("n_pair_on", C.c_int32),
        ("dead", C.c_void_p), ("pair_off", C.c_void_p), ("pair_on", C.c_void_p),
    ]


_P = C.c_void_p
_I = C.c_int
_D = C.c_double
_SZ = C.c_size_t
_I64 = C.c_int64
_U64 = C.c_uint64

# name -> (restype, argtypes); must list EVERY symbol declared in include/scb200.h
SIGNATURES = {
    "scb_version": (_I, []),
    "scb_status_string": (C.c_char_p, [_I]),
    "scb_last_cuda_error": (C.c_char_p, []),
    "scb_launch_count": (_U64, []),
    "scb_profile": (_I, [_I, C.POINTER(_D), C.POINTER(_I64), C.POINTER(_I64)]),
    "scb_contacts_count": (_I, [_P, _I, _I, _D, C.POINTER(Patch), _I, _P, _P]),
    "scb_scan_scratch_bytes": (_SZ, [_I64]),
    "scb_contacts_scan": (_I, [_P, _I64, _P, _P, _P]),
    "scb_contacts_fill": (_I, [_P, _I, _I, _D, C.POINTER(Patch), _I, _P, _P, _P]),
    "scb_pairs_materialize": (_I, [_P, _P, _I, _I, _P, _P]),
    "scb_assemble": (_I, [_I, _P, _I, _I, C.POINTER(FFDesc), _P, _P, _P, _P, _P, _P, _P, _P]),
    "scb_force_constant": (_I, [C.POINTER(FFDesc), _I, _P, _P, _P, _I64, _P, _P, _P]),
    "scb_pair_geometry": (_I, [_P, _I, _I, _P, _P, _P, _P, _P]),
    "scb_densify": (_I, [_I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "scb_assemble_dense_allpairs": (_I, [_I, _P, _I, C.POINTER(FFDesc), _P, _I, _I, _P, _P]),
    "scb_spmm": (_I, [_I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P]),
    "scb_paired_bytes": (_SZ, [_I, _I, _I, _I64, _P, _P]),
    "scb_paired_build": (_I, [_I, _I, _I, _I64, _P, _P, _P, _P, _P, _P]),
    "scb_spmm_paired": (_I, [_I, _I, _I, _I64, _P, _P, _P, _P, _I, _P]),
    "scb_rigid_basis": (_I, [_I, _P, _I, _I, _P, _P, _P]),
    "scb_eig_lowest_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I, _I64]),
    "scb_eig_lowest": (_I, [_I, _I, _I, _I64, _P, _P, _P, _P, _P, _P, _I, _I, _I, _D, _I, _I, _U64,
                            _P, _P, _P, _P, _P, _SZ, _P]),
    "scb_dense_slab_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I]),
    "scb_dense_slab_apply": (_I, [_I64, _I64, _I64, _P, _P, _P, _P, _I, _I, _D, _D, _D, _P, _SZ, _P]),
    "scb_dense_slab_apply_allgather": (_I, [_I64, _I64, _I64, _P, _P, _P, _P, _I, _I, _I, _D, _D, _D, _P, _SZ, _P]),
    "scb_peer_alloc": (_I, [_SZ, _P]),
    "scb_peer_free": (_I, [_P]),
    "scb_peer_export": (_I, [_P, _P]),
    "scb_peer_open": (_I, [_P, _P]),
    "scb_peer_close": (_I, [_P]),
    "scb_dense_gershgorin": (_I, [_I64, _I64, _P, _P, _P]),
    "scb_gram": (_I, [_I, _I64, _I, _P, _P, _P, _P]),
    "scb_chol_orth": (_I, [_I, _I, _P, _P, _P]),
    "scb_rotate": (_I, [_I, _I64, _I, _P, _P, _P, _P, _P, _P]),
    "scb_deflate": (_I, [_I, _I64, _I, _I, _P, _P, _P, _P]),
    "scb_residual_norms": (_I, [_I, _I64, _I, _P, _P, _P, _P, _P]),
    "scb_transpose_small": (_I, [_I, _P, _P, _P]),
    "scb_tf32_ld": (_I64, [_I64]),
    "scb_dense_slab_to_f32": (_I, [_I64, _I64, _P, _P, _P, _P]),
    "scb_resform_prepare": (_I, [_I64, _I, _I, _P, _P, _P, _P, _D, _D, _P, _P, _P, _P, _P, _I, _P]),
    "scb_dense_slab_tf32_apply": (_I, [_I64, _I64, _I64, _P, _P, _I, _P, _P, _P, _P, _P, _P, _D, _I, _P]),
    "scb_dense_slab_tf32_apply_allgather": (_I, [_I64, _I64, _I64, _P, _P, _I, _P, _P, _P, _P, _I, _P, _P, _D, _I, _P]),
    "scb_resform_finish": (_I, [_I64, _I, _P, _P, _P, _I, _P]),
    "scb_coldot": (_I, [_I, _I64, _I, _P, _P, _P, _P]),
    "scb_lanczos_axpy": (_I, [_I, _I64, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "scb_lanczos_bound": (_I, [_I, _I, _I, _P, _P, _D, _P, _P]),
    "scb_rand_block": (_I, [_I64, _U64, _P, _P]),
    "scb_eig_full_workspace_bytes": (_SZ, [_I, _I]),
    "scb_eig_full": (_I, [_I, _I, _P, _P, _P, _P, _SZ, _P]),
    "scb_eig_full_workspace_bytes_ex": (_SZ, [_I, _I, _I]),
    "scb_eig_full_ex": (_I, [_I, _I, _I, _P, _P, _P, _P, _SZ, _P]),
    "scb_msf": (_I, [_I, _I, _I, _I, _P, _P, _D, _P, _P]),
    "scb_msf_cols": (_I, [_I, _I, _I, _I, _I, _I, _P, _P, _D, _P, _P]),
    "scb_dcc": (_I, [_I, _I, _I, _P, _P, _I, _D, _I, _I, _P, _P, _SZ, _P]),
    "scb_dcc_workspace_bytes": (_SZ, [_I, _I, _I]),
    "scb_covariance": (_I, [_I, _I, _P, _P, _I, _I, _P, _P, _SZ, _P]),
    "scb_linear_response": (_I, [_I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "scb_dcc_from_covariance": (_I, [_I, _I, _P, _I, _D, _P, _P, _P]),
    "scb_symv": (_I, [_I, _P, _P, _P, _P]),
    "scb_prs": (_I, [_I, _P, _I, _P, _P]),
    "scb_normal_mode": (_I, [_I, _I, _P, _D, _I, _P, _P]),
    "scb_export_modes": (_I, [_I, _I, _I, _I, _I, _P, _P, _P]),
    "scb_enm_ensemble": (_I, [_I, _P, _I, _I, C.POINTER(FFDesc), C.POINTER(Patch), _P, _I, _D,
                              _P, _P, _P, _P, C.POINTER(_I64), _P]),
    "scb_coords_to_soa": (_I, [_P, _I, _I, _P, _P]),
    "scb_enm_ensemble_host": (_I, [_I, _P, _I, _I, C.POINTER(FFDesc), C.POINTER(Patch), _P, _I, _D,
                                   _P, _P, _P, _P, C.POINTER(_I64), _P]),
    "scb_trim_pool": (_I, []),
}

_lib = None


def lib():
    """Load libscb200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
                "springcraft_b200 has no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def device_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def require_device():
    """The compute path needs the library AND a CUDA device; fail loudly otherwise."""
    handle = lib()
    if not device_available():
        raise RuntimeError("springcraft_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return handle


_EXC = {
    SCB_ERR_INVALID: ValueError,
    SCB_ERR_CUDA: RuntimeError,
    SCB_ERR_ABOVE_CUTOFF: ValueError,
    SCB_ERR_WORKSPACE: RuntimeError,
    SCB_ERR_NOT_CONVERGED: RuntimeError,
    SCB_ERR_UNSUPPORTED: NotImplementedError,
}


def check(status, allow=()):
    """Map a scb_status to the exception type the reference raises."""
    if status == SCB_OK or status in allow:
        return status
    handle = lib()
    msg = handle.scb_status_string(status).decode()
    if status == SCB_ERR_CUDA:
        msg += ": " + handle.scb_last_cuda_error().decode()
    raise _EXC.get(status, RuntimeError)(msg)


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def to_device(array, dtype):
    """NumPy -> contiguous device tensor of the given torch dtype."""
    import torch
    a = np.ascontiguousarray(array)
    if not a.flags.writeable:        # e.g. arrays out of np.load: torch wants a writable source buffer
        a = a.copy()
    return torch.from_numpy(a).to(device="cuda", dtype=dtype)
