"""ctypes binding of the C-ABI CUDA library (include/s2v_b200.h).

There is NO CPU fallback: if libs2v_b200.so is missing or cannot be loaded, every
op raises.  Build it with ``python -m simmodel_b200.build`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libs2v_b200.so")

i32, i64, vp = C.c_int32, C.c_int64, C.c_void_p


class Graph(C.Structure):
    _fields_ = [("num_nodes", i32), ("num_graphs", i32), ("period", i32), ("num_edges", i32),
                ("rowptr", vp), ("col", vp), ("rowptr_t", vp), ("col_t", vp), ("indeg", vp),
                ("ptr", vp), ("gid", vp), ("npad", vp), ("npad_scalar", i32), ("reserved", i32)]


class EmbedParams(C.Structure):
    _fields_ = [("node_dim", i32), ("emb_dim", i32), ("rounds", i32), ("flags", i32),
                ("theta1a_w", vp), ("theta1a_b", vp), ("theta1b_w", vp), ("theta1b_b", vp),
                ("theta2_w", vp), ("theta2_b", vp), ("theta3_w", vp), ("theta3_b", vp),
                ("theta4_w", vp), ("theta4_b", vp)]


class HeadParams(C.Structure):
    _fields_ = [("emb_dim", i32), ("norm_agg", i32),
                ("theta5_w", vp), ("theta5_b", vp), ("theta6_w", vp), ("theta6_b", vp),
                ("theta7_w", vp), ("theta7_b", vp), ("flags", i32), ("reserved", i32)]


GP, EP, HP = C.POINTER(Graph), C.POINTER(EmbedParams), C.POINTER(HeadParams)

# name -> (restype, argtypes); must list every symbol include/s2v_b200.h declares
SIGNATURES = {
    "s2v_abi_version": (C.c_int, []),
    "s2v_error_string": (C.c_char_p, [C.c_int]),
    "s2v_embed_grad_floats": (i64, [i32, i32]),
    "s2v_head_grad_floats": (i64, [i32]),
    "s2v_csr_workspace_bytes": (i64, [i64, i64]),
    "s2v_csr_build": (C.c_int, [vp, i64, i64, vp, vp, vp, vp, vp, vp, i64, vp]),
    "s2v_csr_build_local": (C.c_int, [vp, i32, i64, i64, vp, vp, i32, vp, vp, vp, vp, vp, vp, i64, vp]),
    "s2v_csr_row_to_col": (C.c_int, [vp, i64, i64, vp, vp, vp]),
    "s2v_embed_forward": (C.c_int, [GP, EP, vp, vp, vp, vp]),
    "s2v_embed_backward_workspace_bytes": (i64, [i32, i32]),
    "s2v_embed_backward": (C.c_int, [GP, EP, vp, vp, vp, vp, i32, vp, vp, i64, vp, vp]),
    "s2v_embed_backward_dx": (C.c_int, [GP, EP, vp, vp, vp, vp, i32, vp, vp, i64, vp, vp, vp]),
    "s2v_embed_round_forward": (C.c_int, [GP, i32, i32, vp, vp, vp, vp, vp]),
    "s2v_embed_round_backward": (C.c_int, [GP, i32, i32, i32, vp, vp, vp, vp, vp, i64, vp]),
    "s2v_debug_tc_gemm": (C.c_int, [i32, vp, vp, vp, i64, vp]),
    "s2v_head_forward": (C.c_int, [GP, HP, vp, vp, vp, vp, vp]),
    "s2v_head_fused_forward": (C.c_int, [GP, HP, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "s2v_head_fused_eligible": (C.c_int, [i32, i32, i64, i64]),
    "s2v_qnet_small_eligible": (C.c_int, [i32, i32, i64, i64]),
    "s2v_qnet_small_forward": (C.c_int, [GP, EP, HP, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "s2v_qnet_small_workspace_floats": (i64, [i32, i64, i64]),
    "s2v_qnet_small_forward_ws": (C.c_int, [GP, EP, HP, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    "s2v_head_backward_workspace_bytes": (i64, [i32, i32]),
    "s2v_head_backward": (C.c_int, [GP, HP, vp, vp, vp, vp, i32, vp, vp, i64, vp, vp]),
    "s2v_pool_forward": (C.c_int, [GP, i32, i32, vp, vp, vp]),
    "s2v_pool_backward": (C.c_int, [GP, i32, i32, vp, vp, vp]),
    "s2v_relu_mask": (C.c_int, [vp, vp, vp, i64, vp]),
    "s2v_list_argmax": (C.c_int, [GP, vp, vp, vp, vp, vp, vp, vp]),
    "s2v_masked_argmax": (C.c_int, [GP, vp, vp, vp, vp, vp]),
    "s2v_masked_softmax_forward": (C.c_int, [GP, vp, vp, vp, vp]),
    "s2v_masked_softmax_backward": (C.c_int, [GP, vp, vp, vp, vp]),
    "s2v_scatter_rows": (C.c_int, [GP, vp, vp, vp, vp]),
    "s2v_gat_attn_forward": (C.c_int, [GP, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp]),
    "s2v_gat_edge_floats": (i64, [GP, i32]),
    "s2v_gat_uses_edge_alpha": (C.c_int, [i32, i32, i32]),
    "s2v_gat_attn_backward_workspace_bytes": (i64, [i32, i32, i64]),
    "s2v_rows_linear": (C.c_int, [vp, i32, i32, vp, i32, i32, i32, vp, i32, vp, i32, i64, vp]),
    "s2v_rows_outer": (C.c_int, [vp, i32, i32, vp, i32, i32, vp, i32, vp, i64, i64, vp]),
    "s2v_rows_gemm64": (C.c_int, [vp, i32, i32, vp, i32, i32, vp, i32, vp, i32, i64, vp]),
    "s2v_rows_reduce64_workspace_bytes": (i64, [i32]),
    "s2v_rows_reduce64": (C.c_int, [vp, i32, i32, vp, i32, vp, vp, i64, i64, vp]),
    "s2v_gat_attn_backward": (C.c_int, [GP, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
}

_lib = None
_lock = threading.Lock()


class S2VError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it is not built (no fallback path exists)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise S2VError("simmodel_b200: %s not found -- build it with `python -m simmodel_b200.build`; "
                                   "there is no CPU or PyTorch fallback for this path" % LIB_PATH)
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                if handle.s2v_abi_version() != 5:
                    raise S2VError("simmodel_b200: ABI version mismatch")
                _lib = handle
    return _lib


def check(code: int, what: str):
    if code != 0:
        raise S2VError("%s failed: %s (code %d)" % (what, lib().s2v_error_string(code).decode(), code))


PATHS = {"auto": 0, "ffma": 1, "tc": 2, "tcws": 3, "tc16": 4}


def path_flag() -> int:
    """Kernel path selector from $S2V_B200_PATH (auto | ffma | tc | tcws | tc16); see s2v_embed_params_t.flags."""
    return PATHS[os.environ.get("S2V_B200_PATH", "auto").lower()]


launch_count = 0   # number of C-ABI calls that enqueue kernels (bench.py reports kernel launches from this)
