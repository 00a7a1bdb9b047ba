"""ctypes binding of libeo_b200.so (C ABI declared in include/eo_b200.h).

There is no fallback: if the shared library is missing, or no sm_100 GPU is
visible when a context is created, an `EOError` is raised.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeo_b200.so")

EO_OK = 0
EO_LAYOUT_AOS = 0
EO_LAYOUT_SOA = 1
EO_NITER_BINS = 208
EO_OPERAND_VALUE, EO_OPERAND_GRAD, EO_OPERAND_MANDEL_STRAIN, EO_OPERAND_DEF_GRAD = 0, 1, 2, 3

STATUS_NAMES = {
    0: "EO_OK",
    -1: "EO_ERR_INVALID",
    -2: "EO_ERR_CUDA",
    -3: "EO_ERR_NOMEM",
    -4: "EO_ERR_UNSUPPORTED",
    -5: "EO_ERR_NO_DEVICE",
}


class EOError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {text}")
        self.code = code


class VmParams(C.Structure):
    _fields_ = [("lmbda", C.c_double), ("mu", C.c_double), ("H", C.c_double), ("sigma_0", C.c_double)]


class McParams(C.Structure):
    _fields_ = [("E", C.c_double), ("nu", C.c_double), ("c", C.c_double), ("phi", C.c_double), ("psi", C.c_double),
                ("theta_T", C.c_double), ("a", C.c_double), ("tol", C.c_double), ("Nitermax", C.c_int32)]


class TabDesc(C.Structure):
    _fields_ = [("gdim", C.c_int32), ("bs", C.c_int32), ("nb", C.c_int32), ("nq", C.c_int32),
                ("n_cells", C.c_int64), ("n_dofs", C.c_int64), ("n_nodes", C.c_int64),
                ("dofmap", C.c_void_p), ("x_dofmap", C.c_void_p), ("x", C.c_void_p),
                ("phi", C.c_void_p), ("dphi", C.c_void_p), ("dpsi", C.c_void_p)]


class GTabDesc(C.Structure):
    _fields_ = [("gdim", C.c_int32), ("bs", C.c_int32), ("nb", C.c_int32), ("nq", C.c_int32), ("ng", C.c_int32),
                ("n_sets", C.c_int32), ("n_cells", C.c_int64), ("n_dofs", C.c_int64), ("n_nodes", C.c_int64),
                ("dofmap", C.c_void_p), ("x_dofmap", C.c_void_p), ("x", C.c_void_p),
                ("phi", C.c_void_p), ("dphi", C.c_void_p), ("dgeo", C.c_void_p)]


class IsiharaWeights(C.Structure):
    _fields_ = [("A1", C.c_float * 4 * 64), ("S2", C.c_float * 4 * 64), ("W2", C.c_float * 64 * 64),
                ("W2T", C.c_float * 64 * 64), ("w3", C.c_float * 64), ("s3", C.c_float * 4), ("H", C.c_double * 4)]


EO_JIT_MAX_ARGS = 8
EO_JIT_MAX_PARAMS = 32


class JitDesc(C.Structure):
    _fields_ = [("source", C.c_char_p), ("entry", C.c_char_p),
                ("n_operands", C.c_int32), ("operand_size", C.c_int32 * EO_JIT_MAX_ARGS),
                ("n_state", C.c_int32), ("state_size", C.c_int32 * EO_JIT_MAX_ARGS),
                ("out_size", C.c_int32),
                ("n_aux", C.c_int32), ("aux_size", C.c_int32 * EO_JIT_MAX_ARGS),
                ("n_params", C.c_int32), ("fmad", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [
        ("n_points", C.c_int64),
        ("n_plastic", C.c_int64),
        ("n_nonconverged", C.c_int64),
        ("n_nonfinite", C.c_int64),
        ("niter_hist", C.c_int64 * EO_NITER_BINS),
        ("niter_max", C.c_double),
        ("f_max", C.c_double),
        ("res_max", C.c_double),
        ("reserved", C.c_double),
    ]


_vp = C.c_void_p
_i64 = C.c_int64
_dbl = C.c_double

# name -> (restype, argtypes); kept in one table so the CPU test-suite can check
# that every symbol declared in include/eo_b200.h is exported and bound.
PROTOTYPES = {
    "eo_version": (C.c_int, []),
    "eo_device_count": (C.c_int, []),
    "eo_device_pci_bus_id": (C.c_int, [C.c_int, C.c_char_p, C.c_int]),
    "eo_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "eo_destroy": (C.c_int, [_vp]),
    "eo_last_error": (C.c_char_p, [_vp]),
    "eo_sync": (C.c_int, [_vp]),
    "eo_stream": (_vp, [_vp]),
    "eo_set_chunk": (C.c_int, [_vp, _i64]),
    "eo_launch_count": (_i64, [_vp]),
    "eo_dev_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "eo_dev_free": (C.c_int, [_vp, _vp]),
    "eo_dev_memset": (C.c_int, [_vp, _vp, C.c_int, C.c_size_t]),
    "eo_copy": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "eo_host_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "eo_host_free": (C.c_int, [_vp, _vp]),
    "eo_host_register": (C.c_int, [_vp, _vp, C.c_size_t]),
    "eo_host_unregister": (C.c_int, [_vp, _vp]),
    "eo_event_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "eo_event_destroy": (C.c_int, [_vp, _vp]),
    "eo_event_record": (C.c_int, [_vp, _vp]),
    "eo_event_elapsed_ms": (C.c_int, [_vp, _vp, _vp, C.POINTER(C.c_float)]),
    "eo_flush_l2": (C.c_int, [_vp, C.c_size_t]),
    "eo_assign_gather": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp]),
    "eo_debug_counters": (C.c_int, [_vp, _vp]),
    "eo_fp64_peak": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double)]),
    "eo_fp32_peak": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double)]),
    "eo_fp32_peak_variant": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "eo_stats_reset": (C.c_int, [_vp]),
    "eo_stats_read": (C.c_int, [_vp, C.POINTER(Stats)]),
    "eo_stats_device_ptr": (_vp, [_vp]),
    "eo_allreduce_stats": (C.c_int, [_vp, _vp]),
    "eo_stats_read_global": (C.c_int, [_vp, C.POINTER(Stats)]),
    "eo_stats_collective_begin": (C.c_int, [_vp, C.c_int, C.POINTER(Stats), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "eo_stats_collective_end": (C.c_int, [_vp, C.c_int]),
    "eo_vm_eval": (C.c_int, [_vp, C.POINTER(VmParams), _vp, _vp, _vp, _vp, _vp, _vp, _i64]),
    "eo_vm_eval_resident": (C.c_int, [_vp, C.POINTER(VmParams), _vp, _vp, _vp, _vp, _vp, _vp, _i64, C.c_int]),
    "eo_commit_history": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, C.c_int]),
    "eo_heat_eval": (C.c_int, [_vp, _dbl, _dbl, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64]),
    "eo_tab_create": (C.c_int, [_vp, C.POINTER(TabDesc), C.POINTER(_vp)]),
    "eo_tab_destroy": (C.c_int, [_vp]),
    "eo_tab_ncomp": (C.c_int, [_vp, C.c_int]),
    "eo_tabulate": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _vp]),
    "eo_tab_vm_fused": (C.c_int, [_vp, C.POINTER(VmParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "eo_form_create": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "eo_form_destroy": (C.c_int, [_vp]),
    "eo_form_vector": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp, C.c_int]),
    "eo_form_action": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _i64, _vp, C.c_int]),
    "eo_form_vm_step": (C.c_int, [_vp, C.POINTER(VmParams), _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, C.c_int, C.c_int]),
    "eo_form_vm_step_factored": (C.c_int, [_vp, C.POINTER(VmParams), _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, C.c_int, C.c_int]),
    "eo_form_action_vm_factored": (C.c_int, [_vp, C.POINTER(VmParams), _vp, _vp, _i64, _vp, C.c_int]),
    "eo_vm_expand_tangent": (C.c_int, [_vp, C.POINTER(VmParams), _vp, _vp, _i64, C.c_int]),
    "eo_form_set_pattern": (C.c_int, [_vp, _vp, _vp, _i64]),
    "eo_form_nnz": (_i64, [_vp]),
    "eo_form_matrix": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _i64, _vp, C.c_int]),
    "eo_gtab_create": (C.c_int, [_vp, C.POINTER(GTabDesc), C.POINTER(_vp)]),
    "eo_gtab_destroy": (C.c_int, [_vp]),
    "eo_gtab_ncomp": (C.c_int, [_vp, C.c_int]),
    "eo_gtab_tabulate": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _i64, _vp]),
    "eo_isihara_create": (C.c_int, [_vp, C.POINTER(IsiharaWeights), C.POINTER(_vp)]),
    "eo_isihara_destroy": (C.c_int, [_vp]),
    "eo_isihara_set_correction": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "eo_isihara_eval": (C.c_int, [_vp, _vp, _vp, _vp, _i64]),
    "eo_isihara_eval_on_stream": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "eo_jit_create": (C.c_int, [_vp, C.POINTER(JitDesc), C.POINTER(_vp)]),
    "eo_jit_destroy": (C.c_int, [_vp]),
    "eo_jit_compile": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "eo_jit_nvrtc_version": (C.c_int, []),
    "eo_jit_compile_staged": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "eo_jit_compile_ppt": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "eo_jit_cubin": (C.c_int, [_vp, C.POINTER(C.c_int), _vp, C.c_size_t]),
    "eo_jit_log": (C.c_char_p, [_vp]),
    "eo_jit_last_error": (C.c_char_p, [_vp]),
    "eo_jit_out_width": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "eo_jit_eval": (C.c_int, [_vp, C.POINTER(C.c_int), _vp, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, C.POINTER(_vp), _i64]),
    "eo_jit_eval_tabulated": (C.c_int, [_vp, C.POINTER(C.c_int), _vp, C.POINTER(_vp), C.POINTER(C.c_int), C.POINTER(_vp),
                                        C.POINTER(_vp), _vp, _vp, C.POINTER(_vp)]),
    "eo_jit_compile_tabulated": (C.c_int, [_vp, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                           C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "eo_mc_eval": (C.c_int, [_vp, C.POINTER(McParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64]),
    "eo_mc_eval_tabulated": (C.c_int, [_vp, C.POINTER(McParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "eo_mc_eval_scheme": (C.c_int, [_vp, C.POINTER(McParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, C.c_int]),
}

_lib = None


def load():
    """Load libeo_b200.so and bind every prototype.  Raises EOError if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EOError(
            -5,
            f"{LIB_PATH} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)",
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
