"""ctypes binding of libegnn_b200.so (the C-ABI declared in include/egnn_b200.h).

There is no CPU fallback and no alternative backend: if the shared object is missing the
import of any op fails loudly with instructions to build it (`python -c "import
__graft_entry__ as g; g.build()"` or `make -C elliptic-gnn-project_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libegnn_b200.so")
CSRC = os.path.join(_HERE, "csrc")

ABI_VERSION = 2  # EGNN_ABI_VERSION of include/egnn_b200.h
F32, BF16, F64 = 0, 1, 2
G_SYMMETRIZE, G_SELF_LOOPS = 1, 2
SPMM_SUM, SPMM_MEAN, SPMM_DIV_NBR, SPMM_WEIGHTED = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_ELU = 0, 1, 2

_i64, _i32, _u64, _u32, _f32, _f64, _vp, _sz = (C.c_int64, C.c_int, C.c_uint64, C.c_uint32, C.c_float,
                                                C.c_double, C.c_void_p, C.c_size_t)

# name -> (restype, argtypes); must list every symbol include/egnn_b200.h declares
SIGNATURES = {
    "egnn_abi_version": (_i32, []),
    "egnn_last_error": (C.c_char_p, []),
    "egnn_launch_count": (_u64, []),
    "egnn_set_f32_tc_exact": (_i32, [_i32]),
    "egnn_counter_add": (_i32, [_vp, _i64, _vp]),
    "egnn_graph_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "egnn_graph_build": (_i32, [_vp, _i64, _i64, _i32, _i32] + [_vp] * 16 + [_vp, _sz, _vp]),
    "egnn_hub_ablation_workspace_bytes": (_sz, [_i64, _i64]),
    "egnn_hub_ablation": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "egnn_edge_gather": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    "egnn_neighbor_sample_caps": (_i32, [_i64, _i64, _i64, _vp, _i32, _vp, _vp]),
    "egnn_neighbor_sample_workspace_bytes": (_sz, [_i64, _i64, _i64, _vp, _i32]),
    "egnn_neighbor_sample_state_init": (_i32, [_vp, _i64, _vp]),
    "egnn_neighbor_sample": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _i32, _u64, _u64, _vp, _vp, _i64, _vp, _vp,
                                    _i64, _vp, _vp, _vp, _sz, _vp]),
    "egnn_gather_rows": (_i32, [_vp, _i64, _vp, _i64, _i64, _vp, _i64, _vp]),
    "egnn_txid_join_workspace_bytes": (_sz, [_i64, _i64]),
    "egnn_txid_join": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "egnn_temporal_masks": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "egnn_buffers_differ": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "egnn_spmm": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _i64, _vp, _i32, _i64, _i64,
                         _i64, _vp, _i32, _i32, _vp, _i64, _vp]),
    "egnn_ap_workspace_bytes": (_sz, [_i64]),
    "egnn_average_precision": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "egnn_ranking_metrics": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _f64, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "egnn_temperature_fit": (_i32, [_vp, _i64, _vp, _vp, _i64, _i32, _vp, _vp]),
    "egnn_early_stop_update": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "egnn_snapshot_if_improved": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "egnn_spmm_partition_tasks": (_i64, [_i64, _i64]),
    "egnn_spmm_partition": (_i32, [_vp, _i64, _vp, _i64, _vp]),
    "egnn_gemm_workspace_floats": (_sz, [_i64, _i64, _i64, _i32]),
    "egnn_gemm": (_i32, [_vp, _i32, _i64, _i64, _vp, _i32, _i64, _i64, _vp, _i32, _i64, _i64, _i64, _i64,
                         _vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp]),
    "egnn_linear_stats_parts": (_i64, [_i64]),
    "egnn_linear_tc_workspace_floats": (_sz, [_i32, _i64, _i64]),
    "egnn_linear_tc": (_i32, [_vp, _i64, _vp, _i64, _vp, _i32, _i64, _i64, _i64, _i64, _vp, _vp, _i64, _i32, _vp, _i64,
                              _i64, _vp, _i64, _i32, _vp, _vp]),
    "egnn_wgrad_tc_workspace_floats": (_sz, [_i64, _i64]),
    "egnn_wgrad_tc": (_i32, [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i32, _vp,
                             _vp]),
    "egnn_colstats_reduce": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "egnn_bn_finalize_parts": (_i32, [_vp, _i64, _i64, _f64, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "egnn_pack_sage_weights": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _i32, _vp]),
    "egnn_f64_to_f32": (_i32, [_vp, _vp, _i64, _vp]),
    "egnn_cast": (_i32, [_vp, _i32, _i64, _vp, _i32, _i64, _i64, _i64, _vp]),
    "egnn_inject_time": (_i32, [_vp, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "egnn_embed_grad_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "egnn_embed_grad": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp]),
    "egnn_colreduce_workspace_bytes": (_sz, [_i64]),
    "egnn_colreduce": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "egnn_bn_finalize": (_i32, [_vp, _vp, _f64, _i64, _f32, _f32, _vp, _vp, _vp, _vp, _vp]),
    "egnn_bn_act_dropout_res_fwd": (_i32, [_vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i32,
                                           _f32, _u64, _vp, _u32, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "egnn_bn_act_dropout_bwd_reduce": (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i32,
                                              _f32, _u64, _vp, _u32, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp,
                                              _vp]),
    "egnn_bn_act_dropout_bwd_apply": (_i32, [_vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i32,
                                             _f32, _u64, _vp, _u32, _i64, _vp, _vp, _f64, _vp, _vp, _i64, _vp, _vp]),
    "egnn_dropout_mask": (_i32, [_vp, _i64, _i64, _f32, _u64, _vp, _u32, _i64, _vp]),
    "egnn_gat_scores": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "egnn_gat_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _f32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "egnn_gat_bwd_dst": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _i32, _i32, _i32, _vp, _vp, _i64,
                                _vp]),
    "egnn_gat_bwd_src": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp,
                                _i64, _vp]),
    "egnn_gat_att_grad": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "egnn_ce_workspace_floats": (_sz, [_i64]),
    "egnn_skinny_project": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _i32, _vp, _vp]),
    "egnn_sage_out_fwd": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp, _i64, _vp]),
    "egnn_sage_out_bwd": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _vp, _i64, _vp]),
    "egnn_gcn_out_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp, _i64, _vp]),
    "egnn_gcn_out_bwd": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _vp, _i64, _vp]),
    "egnn_skinny_wgrad_workspace_floats": (_sz, [_i64, _i64, _i32]),
    "egnn_skinny_wgrad": (_i32, [_vp, _i32, _i64, _vp, _i32, _i64, _i64, _vp, _vp, _vp, _vp]),
    "egnn_skinny_wgrad_split": (_i32, [_vp, _i32, _i64, _vp, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "egnn_concat2_f32": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "egnn_skinny_dgrad": (_i32, [_vp, _vp, _i32, _vp, _i32, _i64, _i64, _i64, _vp]),
    "egnn_p2p_allreduce_buffer_bytes": (_sz, [_i32, _i64, _i32]),
    "egnn_p2p_allreduce": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _i32, _i32, _vp, _vp, _i64, _vp]),
    "egnn_bn_stats_exchange": (_i32, [_vp, _i64, _i64, _f64, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32,
                                      _vp, _vp, _i64, _vp]),
    "egnn_bn_bwd_sums_exchange": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _i64, _vp]),
    "egnn_bn_bwd_reduce_parts": (_i64, [_i64, _i64]),
    "egnn_masked_ce": (_i32, [_vp, _i32, _i64, _vp, _vp, _i64, _vp, _f64, _vp, _vp, _vp, _vp]),
    "egnn_masked_loss": (_i32, [_vp, _i32, _i64, _vp, _vp, _i64, _vp, _f64, _f64, _vp, _f64, _f64, _i32, _vp, _vp, _vp,
                                _vp]),
    "egnn_l2_mean_penalty": (_i32, [_vp, _i64, _f64, _vp, _vp, _vp]),
    "egnn_adam_workspace_floats": (_sz, [_i64]),
    "egnn_clip_adam_step": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _vp, _vp,
                                   _vp, _vp]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into the in-tree shared object (nvcc, no GPU needed)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libegnn_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension must be built (make -C {CSRC}); "
                "egnn_b200 has no CPU or PyTorch fallback")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        if _lib.egnn_abi_version() != ABI_VERSION:
            raise RuntimeError("libegnn_b200.so ABI version mismatch; rebuild it")
    return _lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError(lib().egnn_last_error().decode() or f"egnn_b200 error {rc}")


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"egnn_b200 kernels take float32 or bfloat16 tensors, got {t.dtype} "
                    "(there are no fp16 kernels: the convs / nets widen fp16 inputs to fp32, ops.widen_fp16)")


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL).  CPU tensors are an error, not a fallback."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("egnn_b200 ops take CUDA tensors only (no CPU fallback)")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib().egnn_launch_count())
