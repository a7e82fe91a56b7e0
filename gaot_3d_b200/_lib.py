"""ctypes binding of libgaot_b200.so -- the binding a maintainer of the reference would add
(see INTEGRATION.md).  Declares every symbol of include/gaot_b200.h.  There is NO fallback:
if the library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_int32, c_int64, c_uint64, c_double, c_float, c_size_t, c_void_p, c_char_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgaot_b200.so")


class GaotError(RuntimeError):
    pass


class MlpDesc(ctypes.Structure):
    _fields_ = [("n_layers", c_int32), ("dims", c_int32 * 8)]


P = c_void_p
_SIGS = {
    "gaot_last_error": (c_char_p, []),
    "gaot_abi_version": (c_int, []),
    "gaot_launch_count": (c_int64, []),
    "gaot_launch_count_reset": (None, []),
    "gaot_launch_count_add": (None, [c_int64]),
    "gaot_profile_enable": (None, [c_int]),
    "gaot_profile_summary": (c_int, [c_char_p, c_size_t]),
    "gaot_radius_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "gaot_radius_count": (c_int, [P, c_int64, P, c_int64, c_double, c_int, P, c_size_t, P, POINTER(c_int64), P]),
    "gaot_radius_emit": (c_int, [P, c_int64, P, c_int64, c_double, c_int, P, c_size_t, P, P, P, P]),
    "gaot_knn_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "gaot_knn": (c_int, [P, c_int64, P, c_int64, c_int, P, c_size_t, P, P, P]),
    "gaot_coalesce_workspace_bytes": (c_size_t, [c_int64]),
    "gaot_coalesce": (c_int, [P, P, c_int64, c_int64, c_int64, P, c_size_t, P, P, P, POINTER(c_int64), P]),
    "gaot_edge_mask_workspace_bytes": (c_size_t, [c_int64]),
    "gaot_edge_mask": (c_int, [P, P, c_int64, c_double, c_uint64, c_uint64, P, c_size_t, P, P, P, POINTER(c_int64), P]),
    "gaot_csr_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "gaot_csr_from_edges": (c_int, [P, P, c_int64, c_int64, c_int64, c_int, P, c_size_t, P, P, P, P, P]),
    "gaot_gno_workspace_bytes": (c_size_t, [c_int64, c_int64, POINTER(MlpDesc)]),
    "gaot_gno_forward": (c_int, [P, c_int64, P, c_int64, P, c_int32, P, P, P, c_int64, POINTER(MlpDesc), P,
                                 c_int, c_int, c_int, P, c_size_t, P, P]),
    "gaot_gno_backward": (c_int, [P, c_int64, P, c_int64, P, c_int32, P, P, P, c_int64, POINTER(MlpDesc), P,
                                  c_int, c_int, c_int, P, P, c_size_t, P, P, P]),
    "gaot_gno_forward_weighted": (c_int, [P, c_int64, P, c_int64, P, c_int32, P, P, P, c_int64, POINTER(MlpDesc), P,
                                          c_int, c_int, c_int, P, P, c_size_t, P, P]),
    "gaot_gno_backward_weighted": (c_int, [P, c_int64, P, c_int64, P, c_int32, P, P, P, c_int64, POINTER(MlpDesc), P,
                                           c_int, c_int, c_int, P, P, P, c_size_t, P, P, P, P]),
    "gaot_geo_stats": (c_int, [P, c_int64, P, c_int64, P, P, P, P]),
    "gaot_geo_moments": (c_int, [P, c_int64, P, c_int64, P, P, P, P]),
    "gaot_geo_from_moments": (c_int, [P, c_int64, P, P]),
    "gaot_geo_zscore_workspace_bytes": (c_size_t, [c_int64]),
    "gaot_pointnet_workspace_bytes": (c_size_t, []),
    "gaot_pointnet_forward": (c_int, [P, c_int64, P, c_int64, P, P, P, c_int, P, P, P]),
    "gaot_pointnet_backward": (c_int, [P, c_int64, P, c_int64, P, P, P, c_int, P, P, P, c_size_t, P, P]),
    "gaot_a2a_put": (c_int, [P, P, c_int32, c_int32, c_int64, P]),
    "gaot_p2p_put": (c_int, [P, P, c_int32, c_int32, c_int64, c_int32, P]),
    "gaot_p2p_reduce": (c_int, [P, c_int32, c_int64, c_int64, P, P]),
    "gaot_node_linear_supported": (c_int, [c_int32, c_int32]),
    "gaot_node_linear_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "gaot_node_linear_forward": (c_int, [P, c_int64, c_int32, c_int32, P, P, P, P]),
    "gaot_node_linear_backward": (c_int, [P, P, c_int64, c_int32, c_int32, P, P, c_size_t, P, P, P, P]),
    "gaot_node_mlp2_supported": (c_int, [c_int32, c_int32, c_int32]),
    "gaot_node_mlp2_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "gaot_node_mlp2_forward": (c_int, [P, c_int64, c_int32, c_int32, c_int32, P, P, P, P, P, P]),
    "gaot_node_mlp2_backward": (c_int, [P, P, c_int64, c_int32, c_int32, c_int32, P, P, P, P, c_size_t, P, P, P]),
    "gaot_geo_zscore": (c_int, [P, c_int64, c_int32, P, c_size_t, P]),
    "gaot_attn_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32, c_int32, c_int32]),
    "gaot_attn_forward": (c_int, [P, P, P, c_int64, c_int64, c_int32, c_int32, c_int32, P, c_float, c_uint64, P, c_size_t, P, P, P]),
    "gaot_attn_backward": (c_int, [P, P, P, P, P, P, c_int64, c_int64, c_int32, c_int32, c_int32, P, c_float, c_uint64,
                                   P, c_size_t, P, P, P, P]),
    "gaot_linear_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "gaot_linear_forward": (c_int, [P, c_int, c_int64, P, c_int64, c_int64, P, c_int, c_int64, c_int64, c_int64,
                                    P, P, c_int64, P, c_int, c_int64, P, c_size_t, P]),
    "gaot_linear_backward_input": (c_int, [P, c_int, c_int64, P, c_int, c_int64, c_int64, c_int64, P, c_int64,
                                           P, c_int, c_int64, c_int, P, c_size_t, P]),
    "gaot_linear_backward_weight": (c_int, [P, c_int, c_int64, P, c_int, c_int64, c_int64, c_int64, c_int64,
                                            P, c_int64, c_int, P, c_size_t, P]),
    "gaot_cast_bf16": (c_int, [P, P, c_int64, P]),
    "gaot_cast_bf16_batch": (c_int, [P, P, P, c_int32, P]),
    "gaot_rmsnorm_forward": (c_int, [P, P, c_int64, c_int32, c_float, P, P, P, P]),
    "gaot_rmsnorm_backward_workspace_bytes": (c_size_t, [c_int32]),
    "gaot_rmsnorm_backward": (c_int, [P, P, P, P, P, c_int64, c_int32, P, P, P, c_size_t, P]),
    "gaot_colsum_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "gaot_colsum": (c_int, [P, c_int64, c_int64, P, P, c_size_t, P]),
    "gaot_swiglu_forward": (c_int, [P, c_int64, c_int32, P, P]),
    "gaot_swiglu_backward": (c_int, [P, P, c_int64, c_int32, P, P]),
    "gaot_attn_packed_bytes": (c_size_t, [c_int64, c_int64, c_int32, c_int32, c_int32]),
    "gaot_attn_fused_forward": (c_int, [P, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, P, c_float, c_uint64, P, P, P, P, P]),
    "gaot_attn_fused_backward_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32, c_int32]),
    "gaot_attn_fused_backward": (c_int, [P, P, P, P, c_int64, c_int64, c_int32, c_int32, c_int32, P, c_float, c_uint64,
                                         P, c_size_t, P, c_int64, P]),
    "gaot_radius_host": (c_int, [P, c_int64, P, c_int64, c_double, c_int, P, P, POINTER(c_int64)]),
    "gaot_knn_host": (c_int, [P, c_int64, P, c_int64, c_int, P, P, POINTER(c_int64)]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)
_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GaotError(
            f"{LIB_PATH} not found: build it with `python -m gaot_3d_b200.build` "
            "(there is no CPU / PyTorch fallback for the B200 hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)           # AttributeError if the ABI and the header disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gaot_last_error()
        msg = msg.decode() if msg else ""
        if rc == 4:
            raise NotImplementedError(f"{what}: {msg}")
        if rc == 1:
            raise ValueError(f"{what}: {msg}")
        raise GaotError(f"{what} failed (status {rc}): {msg}")
