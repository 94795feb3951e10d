"""ctypes binding of libpn2_b200.so (C ABI: include/pn2_b200.h).

The library is CUDA-only.  There is no CPU implementation behind it and no
fallback: if the shared object is missing, loading raises ImportError; if a call
is made without a CUDA device the wrappers in ``ext.py`` raise RuntimeError.
"""
import ctypes
import os

import torch  # noqa: F401  -- loads the CUDA runtime (libcudart.so.12) the library links to

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libpn2_b200.so")

_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float

# name -> argtypes; every data pointer is passed as a raw device address
SIGNATURES = {
    "pn2_furthest_point_sampling": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "pn2_furthest_point_sampling_pruned": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "pn2_gather_points": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "pn2_gather_points_grad": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "pn2_query_ball_point": [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp],
    "pn2_group_points": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "pn2_group_points_grad": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "pn2_three_nn": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "pn2_three_interpolate": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "pn2_three_interpolate_grad": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    # shared-MLP block, rows layout
    "pn2_bn_stats_workspace_floats": [_i],
    "pn2_bn_stats": [_i, _i, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pn2_bn_fold_eval": [_i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp],
    "pn2_bn_relu_apply": [_i, _i, _vp, _vp, _vp, _vp, _i, _vp],
    "pn2_bn_relu_groupmax": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "pn2_pool_bn_bwd": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pn2_relu_bn_bwd": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pn2_sa_group_rows": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "pn2_sa_group_rows_grad": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "pn2_fp_interp_rows": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "pn2_fp_interp_rows_grad": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "pn2_inverse_map_supported": [_i, _i],
    "pn2_inverse_map": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "pn2_sa_group_rows_grad_inv": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "pn2_fp_interp_rows_grad_inv": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    # shared-MLP layer on the tensor cores
    "pn2_mlp_gemm_bn": [_i, _i],
    "pn2_mlp_gemm_n_tiles": [_i, _i],
    "pn2_mlp_gemm_k_blocks": [_i],
    "pn2_mlp_gemm_grid": [_i, _i],
    "pn2_mlp_pack_weights": [_i, _i, _i, _i, _vp, _vp, _vp],
    "pn2_mlp_gemm_pool_supported": [_i],
    "pn2_mlp_gemm_fwd": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _f,
                         _i, _vp, _vp, _vp, _vp, _vp],
    "pn2_groupmax_from_extrema": [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pn2_bn_finalize": [_i, _i, _i, _i, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    # whole SA block in one kernel (inference)
    "pn2_sa_chain_supported": [_i, _i, _i, _i, _i],
    "pn2_sa_chain_weight_floats": [_i, _i],
    "pn2_mlp_pack_weights_bn": [_i, _i, _i, _vp, _vp, _vp],
    "pn2_sa_chain_eval": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    # fused backward layer on the tensor cores
    "pn2_mlp_bwd_supported": [_i, _i, _i, _i, _i],
    "pn2_mlp_bwd_grid": [_i],
    "pn2_mlp_bwd_nb": [_i],
    "pn2_mlp_bwd_kb": [_i],
    "pn2_mlp_bwd_kpad": [_i],
    "pn2_mlp_pack_weights_t": [_i, _i, _i, _i, _vp, _vp, _vp],
    "pn2_mlp_bwd_layer": [_i] * 7 + [_vp] * 20 + [_vp, _f, _vp, _f, _vp],
    "pn2_mlp_bwd_finalize_scratch_floats": [_i, _i],
    "pn2_mlp_bwd_finalize": [_i, _i, _i, _i] + [_vp] * 7 + [_i, _i] + [_vp] * 7,
    "pn2_pool_bn_bwd_stats": [_i, _i, _i, _i] + [_vp] * 13,
    "pn2_relu_bn_bwd_stats": [_i, _i, _i, _i] + [_vp] * 11,
    "pn2_bwd_stats_finalize": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    # wide layers: data-gradient weight tiles, weight-gradient product
    "pn2_mlp_pack_weights_dgrad": [_i, _i, _i, _i, _i, _vp, _vp, _vp],
    "pn2_mlp_dgrad_stats": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp],
    "pn2_relu_bn_bwd_dz": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pn2_mlp_wgrad_splits": [_i, _i, _i],
    "pn2_mlp_wgrad_rows": [_i],
    "pn2_mlp_wgrad_cols": [_i],
    "pn2_mlp_wgrad": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    # first layer of a block applied before the gather (csrc/mlp_lead.cu)
    "pn2_pad_rows": [ctypes.c_longlong, _i, _i, _vp, _vp, _vp],
    "pn2_lead_blocks": [_i, _i, _i],
    "pn2_fp_lead_max_skip": [],
    "pn2_mlp_pack_weights_cols": [_i, _i, _i, _i, _i, _vp, _vp, _vp],
    "pn2_mlp_pack_weights_dgrad_cols": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "pn2_sa_lead_fwd": [_i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp],
    "pn2_sa_lead_bwd": [_i, _i, _i, _i, _i, _i] + [_vp] * 8 + [_i] + [_vp] * 3 + [_i, _vp, _vp],
    "pn2_fp_lead_fwd": [_i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp],
    "pn2_fp_lead_bwd": [_i, _i, _i, _i, _i, _i] + [_vp] * 8 + [_i] + [_vp] * 3 + [_i, _vp, _vp],
    # flat-buffer training-step utilities
    "pn2_sum_partials": [_i, _i, _vp, _vp, _vp],
    "pn2_sum_partials_block": [_i, _i, _i, _i, _i, _vp, _vp, _vp],
    "pn2_dropout_mask": [ctypes.c_longlong, _f, _vp, _vp, _vp],
    "pn2_adam_flat": [ctypes.c_longlong, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _f, _vp],
    # weighted cross-entropy of the training step
    "pn2_weighted_ce_partials": [_i],
    "pn2_weighted_ce_fwd": [_i, _i, ctypes.c_longlong, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pn2_weighted_ce_bwd": [_i, _i, ctypes.c_longlong, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    # per-iteration metrics on the device
    "pn2_metrics_table_ints": [_i],
    "pn2_first_occurrence_xyz": [_i, _vp, _vp, _vp, _vp],
    "pn2_first_occurrence_voxel": [_i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp],
    "pn2_confusion": [_i, _vp, _vp, _vp, _i, _vp, _vp, _vp],
    # chunk generation (scene resident in HBM)
    "pn2_crop_try": [_i, _i, _vp, _i, _i, ctypes.c_double, ctypes.c_double, _vp, _vp, _vp],
    "pn2_box_select": [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "pn2_chunk_desc_bytes": [],
    "pn2_chunk_batch": [_i, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp],
    "pn2_chunk_centre": [_i, _i, _vp, _vp, _vp, _vp, _vp],
    "pn2_chunk_assemble": [_i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _i,
                           _vp, _vp, _vp, _vp],
}
LIBRARY_SYMBOLS = ("pn2_abi_version", "pn2_launch_count", "pn2_reset_launch_count",
                   "pn2_status_string")
ABI_VERSION = 1

_lib = None


def load():
    """Return the loaded library, loading it on first use."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: build it with `python -m pointnet2.scannet_b200.build` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = _i
        L.pn2_abi_version.restype = _i
        L.pn2_launch_count.restype = ctypes.c_ulonglong
        L.pn2_reset_launch_count.restype = None
        L.pn2_status_string.argtypes = [_i]
        L.pn2_status_string.restype = ctypes.c_char_p
        if L.pn2_abi_version() != ABI_VERSION:
            raise ImportError("libpn2_b200.so ABI %d != expected %d; rebuild"
                              % (L.pn2_abi_version(), ABI_VERSION))
        _lib = L
    return _lib


def check(status, what):
    if status != 0:
        raise RuntimeError("%s failed: %s (status %d)"
                           % (what, load().pn2_status_string(status).decode(), status))


def launch_count():
    return int(load().pn2_launch_count())


def reset_launch_count():
    load().pn2_reset_launch_count()
