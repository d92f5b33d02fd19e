"""ctypes binding of libmwx_b200.so (the C ABI declared in include/mwx_b200.h).

There is no CPU fallback: if the shared library is missing, or a call returns a non-zero status,
this module raises.  ``python __graft_entry__.py build`` (or ``make -C fem_forth_perturb_b200/csrc``)
produces the library in-tree.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MWX_LIB_PATH') or os.path.join(_HERE, 'libmwx_b200.so')     # MWX_LIB_PATH: a variant build (kernel tuning runs)

_lib = None

i32, i64, f64, vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
_pi64 = ctypes.POINTER(ctypes.c_int64)
_pi32 = ctypes.POINTER(ctypes.c_int32)
_pf64 = ctypes.POINTER(ctypes.c_double)

# name -> argtypes (device pointers and host arrays are passed as void*; host scalars by pointer)
_SIGNATURES = {
    'mwx_version': [],
    'mwx_mesh_sort_columns': [i64, vp, vp],
    'mwx_incidence': [i64, i64, i32, vp, vp, vp, vp],
    'mwx_mesh_facets_count': [i64, i64, vp, vp, vp, vp, _pi64, vp],
    'mwx_mesh_facets_fill': [i64, i64, vp, vp, vp, vp, i64, vp, vp, vp],
    'mwx_mesh_refine': [i64, i64, i64, vp, vp, vp, vp, vp, vp, vp],
    'mwx_element_dofs': [i64, i64, vp, vp, vp, vp],
    'mwx_facet_element_count': [i64, i64, vp, vp, vp],
    'mwx_pattern_count': [i64, i64, i32, vp, vp, vp, vp, _pi64, _pi32, vp],
    'mwx_pattern_fill': [i64, i64, i32, vp, vp, vp, vp, vp, vp, vp],
    'mwx_assemble_lagrange_laplace': [i64, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp],
    'mwx_load_vector': [i64, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp],
    'mwx_load_vector_problem': [i32, f64, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp],
    'mwx_morley_error_sums_problem': [i32, i64, i32, vp, vp, vp, vp, vp, vp, vp],
    'mwx_wv_action': [i64, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp],
    'mwx_morley_error_sums': [i64, i32, vp, vp, vp, vp, vp, vp, vp, vp],
    'mwx_morley_sample_nested': [i64, i32, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp],
    'mwx_assemble_morley_mass': [i64, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    'mwx_morley_load_vector': [i64, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp],
    'mwx_morley_facet_pnv': [i64, i64, vp, vp, vp, vp, vp, vp, vp, vp],
    'mwx_assemble_morley': [i64, i64, vp, vp, vp, i32, vp, vp, vp, vp, f64, f64, i32, vp, vp],
    'mwx_assemble_penalty': [i64, i64, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, f64, f64, f64, f64, vp, vp],
    'mwx_assemble_morley_rows': [i64, vp, vp, i64, vp, vp, vp, i32, vp, vp, vp, vp, f64, f64, vp, vp],
    'mwx_asm_plan_create': [i64, vp, vp, vp, vp, i64, vp, vp, vp, vp, ctypes.POINTER(vp), vp],
    'mwx_asm_plan_destroy': [vp],
    'mwx_asm_plan_info': [vp, vp],
    'mwx_assemble_morley_plan': [vp, vp, f64, f64, i32, vp, vp],
    'mwx_axpby': [i64, f64, vp, f64, vp, vp, vp],
    'mwx_dist_rows_count': [i64, vp, vp, vp, vp, vp, vp, vp],
    'mwx_dist_rows_fill': [i64, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    'mwx_dist_localize': [i64, vp, i32, i32, vp, i32, i32, vp, vp],
    'mwx_gather_f64_i64': [i64, vp, vp, vp, vp],
    'mwx_halo_push': [i64, i32, vp, vp, vp, vp, vp],
    'mwx_bsr_create': [i64, i64, i64, vp, vp, vp, i32, i32, i32, ctypes.POINTER(vp), vp],
    'mwx_amg_set_storage': [vp, i32, vp],
    'mwx_residual_dd': [i64, vp, vp, vp, vp, vp, vp, vp],
    'mwx_spmv_lx_create': [i64, vp, vp, i64, ctypes.POINTER(vp), vp],
    'mwx_spmv_lx_destroy': [vp],
    'mwx_spmv_lx': [i64, i64, vp, vp, vp, vp, vp, vp, vp],
    'mwx_spmv_axpy_lx': [i64, i64, vp, vp, vp, vp, vp, vp, i32, vp, vp],
    'mwx_smoother_step_lx': [i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, f64, f64, vp],
    'mwx_dcg_spmv_dot_lx': [i64, i64, vp, vp, vp, vp, vp, vp, vp, vp],
    'mwx_csr_pack': [i64, vp, vp, vp, vp],
    'mwx_spmv_pk': [i64, i64, vp, vp, vp, vp, vp],
    'mwx_spmv_axpy_pk': [i64, i64, vp, vp, vp, vp, i32, vp, vp],
    'mwx_smoother_step_pk': [i64, i64, vp, vp, vp, vp, vp, vp, vp, f64, f64, vp],
    'mwx_bsr_destroy': [vp],
    'mwx_bsr_spmv': [vp, vp, vp, vp],
    'mwx_bsr_spmv_axpy': [vp, vp, vp, i32, vp, vp],
    'mwx_bsr_smoother_step': [vp, vp, vp, vp, vp, vp, f64, f64, vp],
    'mwx_peer_ctl_bytes': [],
    'mwx_peer_barrier': [vp, vp, i32, i32, i32, vp],
    'mwx_peer_push': [i64, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
    'mwx_peer_allreduce': [vp, i32, vp, vp, i32, i32, i32, vp],
    'mwx_dcg_workspace_doubles': [],
    'mwx_dcg_dot': [i64, vp, vp, vp, i32, vp],
    'mwx_dcg_p_update': [i64, vp, vp, vp, vp, vp, i32, vp],
    'mwx_dcg_spmv_dot': [i64, i64, vp, vp, vp, vp, vp, vp, vp],
    'mwx_dcg_xr_update': [i64, vp, vp, vp, vp, vp, vp, i32, vp],
    'mwx_dcg_residual': [i64, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    'mwx_condense_count': [i64, vp, vp, vp, vp, vp, _pi64, _pi64, vp],
    'mwx_condense_fill': [i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    'mwx_hilbert_keys': [i64, vp, f64, f64, f64, vp, vp],
    'mwx_csr_permute': [i64, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    'mwx_gather_f64': [i64, vp, vp, vp, vp],
    'mwx_spmv': [i64, i64, vp, vp, vp, vp, vp, vp],
    'mwx_spmv_axpy': [i64, i64, vp, vp, vp, vp, vp, i32, vp, vp],
    'mwx_smoother_step': [i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, f64, f64, vp],
    'mwx_smoother_first': [i64, vp, vp, f64, vp, vp, vp],
    'mwx_dense_gemv': [i64, vp, vp, vp, vp],
    'mwx_spmv_plan': [i64, vp, _pi32, vp],
    'mwx_csr_diagonal': [i64, vp, vp, vp, vp, vp],
    'mwx_pcg_jacobi': [i64, vp, vp, vp, vp, vp, f64, i64, i32, vp, _pi64, _pi32, _pf64, vp],
    'mwx_amg_setup_host': [i64, vp, vp, vp, f64, i32, i32, ctypes.POINTER(vp)],
    'mwx_amg_setup_sa_host': [i64, vp, vp, vp, vp, i32, i32, i32, f64, i32, i32, ctypes.POINTER(vp)],
    'mwx_amg_level_candidates': [vp, i32, vp, vp],
    'mwx_amg_num_levels': [vp],
    'mwx_set_host_threads': [i32],
    'mwx_amg_level_dims': [vp, i32, i32, vp],
    'mwx_amg_level_copy': [vp, i32, i32, vp, vp, vp],
    'mwx_amg_level_view': [vp, i32, i32, vp, vp, vp],
    'mwx_amg_level_splitting': [vp, i32, vp],
    'mwx_amg_level_roots': [vp, i32, vp, vp],
    'mwx_amg_destroy_host': [vp],
    'mwx_gs_wavefronts_host': [i64, vp, vp, vp, vp, _pi64],
    'mwx_gs_wavefronts_backward_host': [i64, vp, vp, vp, vp, _pi64],
    'mwx_amg_upload': [vp, i32, i32, vp, ctypes.POINTER(vp), vp],
    'mwx_amg_setup_sa_dev': [i64, vp, vp, vp, vp, i32, i32, i32, f64, i32, i32, i32, i32, ctypes.POINTER(vp), vp],
    'mwx_amg_dev_finalize': [vp, vp, vp],
    'mwx_amg_dev_num_levels': [vp],
    'mwx_amg_dev_level_dims': [vp, i32, i32, vp],
    'mwx_amg_dev_level_copy': [vp, i32, i32, vp, vp, vp, vp],
    'mwx_amg_dev_level_aux': [vp, i32, vp, _pi64, _pi32, vp, _pi32, vp],
    'mwx_amg_level_rho': [vp, i32, _pf64],
    'mwx_amg_set_chebyshev': [vp, i32, f64],
    'mwx_amg_set_cycle': [vp, i32, i32, i32],
    'mwx_amg_destroy_dev': [vp],
    'mwx_amg_vcycle': [vp, vp, vp, vp],
    'mwx_pcg_amg_monitored': [vp, i64, vp, vp, vp, vp, vp, f64, i64, i32, vp, _pi64, _pi32, _pf64, _pi64, vp],
    'mwx_pcg_amg_history': [vp, i64, vp, vp, vp, vp, vp, f64, i64, vp, _pi64, _pi32, _pf64, vp, i64, vp],
    'mwx_pcg_jacobi_history': [i64, vp, vp, vp, vp, vp, f64, i64, i32, vp, _pi64, _pi32, _pf64, vp, i64, vp],
    'mwx_pcg_amg': [vp, i64, vp, vp, vp, vp, vp, f64, i64, vp, _pi64, _pi32, _pf64, vp],
}
_VOID = {'mwx_amg_destroy_host', 'mwx_amg_destroy_dev', 'mwx_asm_plan_destroy', 'mwx_bsr_destroy', 'mwx_spmv_lx_destroy'}

ERRORS = {-1: 'MWX_E_BADARG', -2: 'MWX_E_CAPACITY', -3: 'MWX_E_NOCONV', -4: 'MWX_E_NOMEM'}


class MwxError(RuntimeError):
    pass


def lib():
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MwxError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a). fem_forth_perturb_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = None if name in _VOID else ctypes.c_int
        _lib = L
    return _lib


def check(rc, what=''):
    if rc == 0:
        return
    if rc < 0:
        raise MwxError(f"{what}: {ERRORS.get(rc, rc)}")
    raise MwxError(f"{what}: cudaError {rc}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise MwxError("fem_forth_perturb_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch


def ptr(t):
    """Device (or host) address of a torch tensor / numpy array; None -> NULL."""
    if t is None:
        return None
    if hasattr(t, 'data_ptr'):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


def stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
