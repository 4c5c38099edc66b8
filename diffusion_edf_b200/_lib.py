"""ctypes binding of libdedf.so (the C ABI declared in include/dedf.h).

There is NO CPU or pure-PyTorch fallback: if the shared library is missing, or an
op is called with tensors that are not CUDA fp32, the call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

MAX_SCALES = 8
MLP_MAX_LAYERS = 4
MLP_IN_ROWS, MLP_IN_RBF, MLP_IN_FIELD = 0, 1, 2
EPI_ACT, EPI_LIN = 0, 1

# DEDF_LIB: another build of the SAME library (A/B runs of a kernel change under profiles/); never a fallback
_LIB_PATH = os.environ.get("DEDF_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdedf.so")
_lib: Optional[C.CDLL] = None

c_fp = C.c_void_p
c_int = C.c_int
c_f = C.c_float
c_d = C.c_double
c_ll = C.c_longlong
c_ull = C.c_ulonglong


class MlpDesc(C.Structure):
    _fields_ = [
        ("mode", c_int), ("n_edges_dev", c_fp), ("x_in", c_fp), ("length", c_fp),
        ("rbf_mean", c_fp), ("rbf_std_logit", c_fp), ("rbf_weight_logit", c_fp),
        ("rbf_cutoff", c_f), ("rbf_offset", c_f),
        ("n_scales", c_int), ("n_dst", c_int), ("row_ptr", c_fp), ("edge_dst", c_fp),
        ("enc_mean", c_fp * MAX_SCALES), ("enc_std_logit", c_fp * MAX_SCALES), ("enc_weight_logit", c_fp * MAX_SCALES),
        ("enc_r", c_f * MAX_SCALES), ("enc_max_r", c_f), ("enc_n", c_f), ("enc_freq", c_fp),
        ("pre_w", c_fp * MAX_SCALES), ("row_bias", c_fp), ("n_rb", c_int), ("rb_div", c_int),
        ("n_layers", c_int), ("dims", c_int * (MLP_MAX_LAYERS + 1)),
        ("W", c_fp * MLP_MAX_LAYERS), ("b", c_fp * MLP_MAX_LAYERS),
        ("ln_g", c_fp * MLP_MAX_LAYERS), ("ln_b", c_fp * MLP_MAX_LAYERS),
        ("flags", c_int * MLP_MAX_LAYERS), ("out_offset", c_fp), ("out", c_fp),
        ("W_tc", c_fp * MLP_MAX_LAYERS), ("pre_w_tc", c_fp), ("tc_f16", c_int),
    ]


class NodeChainDesc(C.Structure):
    _fields_ = [
        ("x", c_fp), ("n", c_int), ("irr_emb", c_int * 3), ("irr_pre", c_int * 3),
        ("P0", c_fp), ("P1", c_fp), ("P2", c_fp), ("pb", c_fp), ("res1", c_fp),
        ("ln_w", c_fp), ("ln_b", c_fp), ("ln_eps", c_f),
        ("A0", c_fp), ("A1", c_fp), ("A2", c_fp), ("ab", c_fp),
        ("B0", c_fp), ("B1", c_fp), ("B2", c_fp), ("bb", c_fp), ("y", c_fp),
    ]


class HeadFrontDesc(C.Structure):
    _fields_ = [
        ("Ts", c_fp), ("n_t", c_int), ("qx", c_fp), ("n_q", c_int),
        ("x_src", c_fp), ("b_src", c_fp), ("b_q", c_fp),
        ("n_scales", c_int), ("src_off", c_int * (MAX_SCALES + 1)), ("r", c_f * MAX_SCALES),
        ("max_num_neighbors", c_int), ("ns_lo", c_f), ("ns_hi", c_f), ("capacity", c_int),
        ("x_dst", c_fp), ("row_ptr", c_fp), ("counts", c_fp), ("edge_src", c_fp), ("edge_dst", c_fp),
        ("length", c_fp), ("sh", c_fp), ("logit", c_fp), ("n_edges", c_fp), ("overflow", c_fp),
        ("cta_sum", c_fp), ("barrier", c_fp),
        ("step", c_fp), ("n_steps", c_int), ("rows_all", c_fp), ("rows_cur", c_fp), ("rows_k", c_int),
        ("stage_early", c_int),
    ]


class ScoreStepDesc(C.Structure):
    _fields_ = [
        ("Ts", c_fp), ("n_t", c_int), ("qf", c_fp), ("key_f", c_fp), ("qx", c_fp), ("qw", c_fp), ("n_q", c_int),
        ("irr", c_int * 3), ("Wd", c_fp * 2), ("Wl0", c_fp * 2), ("Wl1", c_fp * 2), ("bl", c_fp * 2),
        ("n_vec", c_int), ("lin_mult", c_f), ("ang_out", c_fp), ("lin_out", c_fp),
        ("T64", c_fp), ("sched", c_fp), ("n_steps", c_int), ("counter", c_fp), ("noise", c_fp), ("seed", c_ull), ("seed_dev", c_fp),
        ("ang_mult", c_d), ("lin_mult_d", c_d), ("traj", c_fp), ("T32", c_fp), ("ticket", c_fp),
    ]


class TimeDesc(C.Structure):
    _fields_ = [
        ("max_time", c_f), ("enc_n", c_f), ("enc_freq", c_fp),
        ("enc_dim", c_int), ("h_dim", c_int), ("e_dim", c_int), ("out_dim", c_int), ("n_scales", c_int),
        ("W1", c_fp * MAX_SCALES), ("b1", c_fp * MAX_SCALES), ("W2", c_fp * MAX_SCALES), ("b2", c_fp * MAX_SCALES),
        ("Wp", c_fp * MAX_SCALES), ("bp", c_fp * MAX_SCALES),
    ]


_PROTOS = {
    "dedf_fps": [c_fp, c_int, c_int, c_int, c_fp, c_int, c_fp, c_fp, c_fp],
    "dedf_radius_count": [c_fp, c_fp, c_int, c_int, C.POINTER(c_int), C.POINTER(c_f), c_fp, c_fp, c_int, c_fp, c_int, c_fp, c_fp, c_int, c_fp, c_fp, c_fp],
    "dedf_radius_fill": [c_fp, c_fp, c_int, c_int, C.POINTER(c_int), C.POINTER(c_f), c_fp, c_fp, c_int, c_fp, c_int, c_fp, c_fp, c_fp, c_fp],
    "dedf_grid_build": [c_fp, c_int, c_f, c_int, c_fp, c_fp, c_fp, c_fp, c_fp],
    "dedf_radius_grid_count": [c_fp, c_int, c_fp, c_int, c_f, c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_fp, c_int, c_fp, c_fp, c_int, c_fp, c_fp, c_fp],
    "dedf_radius_grid_fill": [c_fp, c_int, c_fp, c_int, c_f, c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_fp, c_int, c_fp, c_fp, c_fp, c_fp],
    "dedf_edge_geom": [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, C.POINTER(c_int), C.POINTER(c_f), c_f, c_f, c_fp, c_fp, c_fp, c_fp],
    "dedf_edge_mlp": [C.POINTER(MlpDesc), c_int, c_fp],
    "dedf_edge_mlp_tc": [C.POINTER(MlpDesc), c_int, c_fp],
    "dedf_edge_tp_lin": [c_int, c_int, c_fp, c_fp, c_int, c_fp, c_fp, c_fp, c_int, c_fp, c_fp, c_ll, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp],
    "dedf_l2_persist": [c_fp, c_ll, c_fp],
    "dedf_edge_tp_act_tc": [c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_fp, c_fp, c_ll, c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp],
    "dedf_segment_softmax_reduce": [c_fp, c_int, c_int, c_fp, c_fp, c_int, c_int, c_int, c_fp, c_fp],
    "dedf_edge_tp_reduce": [c_int, c_fp, c_fp, c_fp, c_fp, c_int, c_fp, c_fp, c_int, c_fp, c_fp],
    "dedf_node_linear": [c_fp, c_int, C.POINTER(c_int), C.POINTER(c_int), c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_f, c_int, c_fp, c_f, c_fp, c_fp],
    "dedf_head_front": [C.POINTER(HeadFrontDesc), c_fp],
    "dedf_node_chain": [C.POINTER(NodeChainDesc), c_fp],
    "dedf_node_linear_pair": [c_fp, c_int, C.POINTER(c_int), C.POINTER(c_fp), c_fp, c_fp, c_fp, c_int, C.POINTER(c_int), C.POINTER(c_fp), c_fp, c_fp,
                              C.POINTER(c_int), c_fp],
    "dedf_weight_post": [c_fp, c_int, c_int, c_fp, c_fp, c_fp, c_fp, c_int, c_fp, c_fp, c_fp],
    "dedf_gather_rows": [c_fp, c_fp, c_int, c_int, c_fp, c_fp],
    "dedf_add_scale": [c_fp, c_fp, c_f, c_ll, c_fp, c_fp],
    "dedf_time_embed": [C.POINTER(TimeDesc), c_fp, c_int, c_fp, c_fp],
    "dedf_query_transform": [c_fp, c_int, c_fp, c_fp, c_int, C.POINTER(c_int), c_fp, c_fp, c_fp],
    "dedf_score_tp": [c_fp, c_int, c_fp, c_fp, c_fp, c_fp, c_int, C.POINTER(c_int), C.POINTER(c_fp), C.POINTER(c_fp),
                      C.POINTER(c_fp), C.POINTER(c_fp), c_int, c_f, c_fp, c_fp, c_fp],
    "dedf_score_tp_step": [C.POINTER(ScoreStepDesc), c_fp],
    "dedf_pose_update": [c_fp, c_int, c_fp, c_fp, c_fp, c_ull, c_ull, c_d, c_d, c_d, c_d, c_d, c_d, c_fp, c_fp, c_fp, c_fp, c_fp],
    "dedf_sample_advance": [c_fp, c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_fp],
    "dedf_flag_if_differs": [c_fp, c_fp, c_ll, c_fp, c_fp],
    "dedf_prefetch_l2": [c_fp, c_fp, c_int, c_fp],
    "dedf_tc_selftest": [c_fp, c_fp, c_int, c_int, c_int, c_fp, c_fp],
    "dedf_lin_wgrad": [c_fp, c_fp, c_int, C.POINTER(c_int), C.POINTER(c_int), c_fp, c_fp, c_fp, c_fp, c_fp],
    "dedf_ln_fwd": [c_fp, c_int, C.POINTER(c_int), c_fp, c_fp, c_f, c_fp, c_fp],
    "dedf_ln_bwd": [c_fp, c_fp, c_int, C.POINTER(c_int), c_fp, c_f, c_fp, c_fp, c_fp, c_fp],
    "dedf_gate_fwd": [c_fp, c_int, C.POINTER(c_int), c_fp, c_fp],
    "dedf_gate_bwd": [c_fp, c_fp, c_int, C.POINTER(c_int), c_fp, c_fp],
    "dedf_act_fwd": [c_fp, c_ll, c_int, c_fp, c_fp],
    "dedf_act_bwd": [c_fp, c_fp, c_ll, c_int, c_fp, c_fp],
    "dedf_dtp_fwd": [c_int, c_fp, c_fp, c_fp, c_ll, c_int, c_fp, c_fp],
    "dedf_dtp_bwd": [c_int, c_fp, c_fp, c_fp, c_ll, c_fp, c_int, c_fp, c_fp, c_fp],
    "dedf_gather_rows_i32": [c_fp, c_fp, c_int, c_int, c_fp, c_fp],
    "dedf_scatter_add_rows": [c_fp, c_fp, c_int, c_int, c_int, c_fp, c_fp],
    "dedf_alpha_fwd": [c_fp, c_int, c_int, c_fp, c_fp, c_fp, c_fp],
    "dedf_alpha_bwd": [c_fp, c_int, c_int, c_fp, c_fp, c_fp, c_fp, c_fp],
    "dedf_softmax_reduce_bwd": [c_fp, c_int, c_int, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_fp, c_fp, c_fp],
    "dedf_rbf_fwd": [c_fp, c_int, c_int, c_fp, c_fp, c_fp, c_f, c_f, c_int, c_fp, c_fp],
    "dedf_rbf_bwd": [c_fp, c_int, c_int, c_fp, c_fp, c_fp, c_f, c_f, c_int, c_fp, c_fp, c_fp, c_fp, c_fp],
    "dedf_sinusoid": [c_fp, c_int, c_int, c_fp, c_f, c_fp, c_fp],
    "dedf_score_tp_fwd": [c_fp, c_fp, c_fp, c_int, C.POINTER(c_int), c_fp, c_fp],
    "dedf_score_tp_bwd": [c_fp, c_fp, c_fp, c_int, C.POINTER(c_int), c_fp, c_fp, c_fp, c_fp, c_fp],
    "dedf_query_transform_bwd": [c_fp, c_int, c_int, C.POINTER(c_int), c_fp, c_fp, c_fp],
    "dedf_assemble_fwd": [c_fp, c_int, c_int, c_int, c_fp, c_fp, c_fp, c_fp, c_f, c_fp, c_fp, c_fp],
    "dedf_assemble_bwd": [c_fp, c_int, c_int, c_int, c_fp, c_fp, c_fp, c_fp, c_f, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp],
    "dedf_dropout_mask": [c_ull, c_ull, c_ll, c_f, c_fp, c_fp],
    "dedf_group_scale": [c_fp, c_fp, c_int, C.POINTER(c_int), c_int, c_fp, c_fp],
    "dedf_ebm_energy": [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_f, c_fp, c_fp],
    "dedf_bbox": [c_fp, c_int, c_fp, c_fp, c_fp],
    "dedf_voxel_count": [c_fp, c_int, c_fp, c_f, c_int, c_int, c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp],
    "dedf_voxel_reduce": [c_fp, c_fp, c_int, c_int, c_fp, c_f, c_int, c_int, c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_fp, c_fp, c_fp],
    "dedf_edge_gather_scalar": [c_fp, c_fp, c_fp, c_int, c_fp, c_fp],
    "dedf_rowdot": [c_fp, c_fp, c_int, c_int, c_fp, c_fp],
    "dedf_value_reduce": [c_int, c_fp, c_int, c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp],
    "dedf_dtp_bwd_sh": [c_int, c_fp, c_fp, c_ll, c_fp, c_int, c_fp, c_fp],
    "dedf_rbf_bwd_len": [c_fp, c_int, c_int, c_fp, c_fp, c_fp, c_f, c_f, c_int, c_fp, c_fp, c_fp],
    "dedf_sinusoid_bwd": [c_fp, c_int, c_int, c_fp, c_f, c_fp, c_fp, c_fp],
    "dedf_edge_geom_bwd": [c_fp, c_fp, c_fp, c_fp, c_int, c_int, C.POINTER(c_int), C.POINTER(c_f), c_f, c_f, c_fp, c_fp, c_fp, c_fp, c_fp],
    "dedf_ebm_energy_bwd": [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_f, c_fp, c_fp, c_fp],
    "dedf_ebm_pose_grad": [c_fp, c_int, c_int, C.POINTER(c_int), c_fp, c_fp, c_fp, c_fp, c_f, c_f, c_fp, c_fp, c_fp],
    "dedf_collision_check": [c_fp, c_int, c_fp, c_ll, c_fp, c_int, c_int, c_f, c_fp, c_fp],
    "dedf_collision_energy": [c_fp, c_int, c_fp, c_ll, c_fp, c_int, c_int, c_f, c_int, c_f, c_int, c_fp, c_fp, c_fp],
    "dedf_collision_step": [c_fp, c_fp, c_int, c_f, c_f, c_fp, c_fp],
    "dedf_dtp_generic_fwd": [c_fp, C.POINTER(c_int), c_fp, c_fp, c_ll, c_int, C.POINTER(c_int), C.POINTER(c_int), c_int, c_fp, c_fp],
    "dedf_stamp": [c_fp, c_fp],
    "dedf_build_arch": [],
}

EXPORTED_SYMBOLS = tuple(_PROTOS.keys())


class DedfError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load libdedf.so (built by ``__graft_entry__.build()`` / ``make -C diffusion_edf_b200/csrc``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise DedfError(
                f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  diffusion_edf_b200 has no CPU / PyTorch fallback.")
        lib = C.CDLL(_LIB_PATH)
        for name, args in _PROTOS.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = c_int
        _lib = lib
    return _lib


_ERR = {-1: "bad argument", -2: "kernel launch failed", -3: "unsupported configuration"}


def check(rc: int, what: str) -> None:
    if rc != 0:
        extra = ""
        if rc == -2 and torch.cuda.is_available():
            extra = " (" + str(torch.cuda.get_device_name()) + ")"
        raise DedfError(f"{what}: {_ERR.get(rc, 'error')} (code {rc}){extra}")


def ptr(t: Optional[torch.Tensor], dtype=torch.float32) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor of the expected dtype (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise DedfError("diffusion_edf_b200 ops need CUDA tensors (no CPU fallback); got a CPU tensor")
    if t.dtype != dtype:
        raise DedfError(f"expected dtype {dtype}, got {t.dtype} (the kernels are fp32; .half() is not supported)")
    if not t.is_contiguous():
        raise DedfError("expected a contiguous tensor")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def int_array(vals: Sequence[int]):
    return (c_int * len(vals))(*[int(v) for v in vals])


def float_array(vals: Sequence[float]):
    return (c_f * len(vals))(*[float(v) for v in vals])
