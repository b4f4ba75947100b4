"""ctypes binding of libunib200.so (include/unib200.h).  The product path FAILS LOUDLY when the library is missing;
there is no CPU / PyTorch fallback for any op."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# UNIB200_LIB selects another build of the SAME library (A/B measurements of kernel variants); never a fallback
LIB_PATH = os.environ.get("UNIB200_LIB") or os.path.join(HERE, "libunib200.so")

SEG_1x1, SEG_3x3, SEG_3x3_S2, SEG_3x3_S2P0, SEG_UP2x2 = 0, 1, 2, 3, 4
OP_OTHER, OP_GEMM, OP_ATTENTION, OP_GROUPNORM, OP_LAYERNORM = 0, 1, 2, 3, 4
OP_NAMES = {OP_OTHER: "other", OP_GEMM: "conv_gemm", OP_ATTENTION: "attention", OP_GROUPNORM: "groupnorm",
            OP_LAYERNORM: "layernorm"}
EPI_GEGLU, EPI_OUT_NCHW, EPI_OUT_F32, EPI_SILU, EPI_AXPBY = 1, 2, 4, 8, 16


class Seg(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int), ("ld", C.c_int), ("kind", C.c_int)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("nseg", C.c_int),
        ("seg", Seg * 4),
        ("weight", C.c_void_p), ("bias", C.c_void_p), ("bias_bstride", C.c_int),
        ("bias_step", C.c_void_p), ("bias_step_stride", C.c_int64),
        ("res", C.c_void_p), ("ldr", C.c_int),
        ("out", C.c_void_p), ("ldc", C.c_int),
        ("flags", C.c_int), ("splits", C.c_int),
        ("partial", C.c_void_p), ("partial_bytes", C.c_size_t),
        ("axpby", C.c_void_p), ("axpby_step", C.c_void_p), ("aux", C.c_void_p), ("aux_out", C.c_void_p),
        ("axpby_first_channel", C.c_int),
        ("rowstats_out", C.c_void_p), ("ln_rowstats", C.c_void_p), ("ln_parts", C.c_int), ("ln_wsum", C.c_void_p),
        ("ln_eps", C.c_float), ("ln_C", C.c_int),
        ("gn_part", C.c_void_p), ("gn_gran", C.c_int), ("gn_rows", C.c_int),
    ]


class WgradDesc(C.Structure):
    _fields_ = [("x", C.c_void_p), ("C", C.c_int), ("ldx", C.c_int), ("dy", C.c_void_p), ("N", C.c_int), ("lddy", C.c_int),
                ("M", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("taps", C.c_int),
                ("dw", C.c_void_p), ("db", C.c_void_p), ("partial", C.c_void_p), ("partial_bytes", C.c_size_t)]


class GnBwdDesc(C.Structure):
    _fields_ = [("x", C.c_void_p), ("ldx", C.c_int), ("dz", C.c_void_p), ("ldz", C.c_int), ("dx", C.c_void_p),
                ("lddx", C.c_int), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("dgamma", C.c_void_p),
                ("dbeta", C.c_void_p), ("scratch", C.c_void_p), ("B", C.c_int), ("HW", C.c_int), ("C", C.c_int),
                ("groups", C.c_int), ("silu", C.c_int), ("eps", C.c_float)]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("ldq", C.c_int), ("k", C.c_void_p), ("ldk", C.c_int), ("v", C.c_void_p), ("ldv", C.c_int),
        ("out", C.c_void_p), ("ldo", C.c_int),
        ("B", C.c_int), ("heads", C.c_int), ("Nq", C.c_int), ("Nk", C.c_int), ("d", C.c_int), ("scale", C.c_float),
        ("lse2", C.c_void_p),
    ]


class AttnBwdDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("ldq", C.c_int), ("k", C.c_void_p), ("ldk", C.c_int), ("v", C.c_void_p), ("ldv", C.c_int),
        ("o", C.c_void_p), ("ldo", C.c_int), ("dout", C.c_void_p), ("lddo", C.c_int),
        ("lse2", C.c_void_p), ("D", C.c_void_p), ("dq_acc", C.c_void_p), ("ld_dq", C.c_int),
        ("dk", C.c_void_p), ("ld_dk", C.c_int), ("dv", C.c_void_p), ("ld_dv", C.c_int),
        ("B", C.c_int), ("heads", C.c_int), ("Nq", C.c_int), ("Nk", C.c_int), ("d", C.c_int), ("scale", C.c_float),
    ]


class GnDesc(C.Structure):
    _fields_ = [
        ("x1", C.c_void_p), ("ld1", C.c_int), ("C1", C.c_int),
        ("x2", C.c_void_p), ("ld2", C.c_int), ("C2", C.c_int),
        ("B", C.c_int), ("HW", C.c_int), ("groups", C.c_int), ("eps", C.c_float),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("out", C.c_void_p), ("silu", C.c_int),
        ("scratch", C.c_void_p), ("scratch_floats", C.c_size_t),
        ("part1", C.c_void_p), ("part2", C.c_void_p), ("part_gran", C.c_int), ("part_rows", C.c_int),
    ]


# every symbol include/unib200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "unib200_version", "unib200_last_error", "unib200_device_info", "unib200_set_pdl",
    "unib200_program_create", "unib200_program_destroy", "unib200_program_num_launches", "unib200_program_run",
    "unib200_program_graph_instantiate", "unib200_program_graph_launch", "unib200_program_set_lane",
    "unib200_program_barrier",
    "unib200_program_num_ops", "unib200_program_op_info", "unib200_program_op_desc", "unib200_program_profile",
    "unib200_conv_gemm", "unib200_conv_gemm_dual", "unib200_packed_k", "unib200_pick_bn", "unib200_debug_set_trace", "unib200_attention", "unib200_groupnorm", "unib200_layernorm",
    "unib200_to_nhwc", "unib200_from_nhwc", "unib200_upsample2x", "unib200_timestep_sinusoid", "unib200_gemv",
    "unib200_axpby", "unib200_add_int", "unib200_add_f16", "unib200_unipc_step",
    "unib200_softmax_rows", "unib200_gaussian_sample",
    "unib200_conv_wgrad", "unib200_groupnorm_backward", "unib200_colsum", "unib200_layernorm_backward", "unib200_geglu",
    "unib200_softmax_backward", "unib200_cvt_f32_f16", "unib200_silu_f16", "unib200_pool2x2_sum", "unib200_scatter2x",
    "unib200_adamw_step", "unib200_attention_backward", "unib200_pack_master_weight", "unib200_wgrad_scatter_add",
    "unib200_create", "unib200_destroy", "unib200_load_weight", "unib200_alloc", "unib200_bind", "unib200_buffer",
    "unib200_ctx_attach", "unib200_ctx_run", "unib200_unet_forward", "unib200_attr_enc_forward",
    "unib200_attr_dec_forward", "unib200_dual_step", "unib200_sample_loop",
]

_lib = None


class Unib200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libunib200.so or raise -- never falls back to another implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Unib200Error(f"{LIB_PATH} is missing: build it with `python -m uni_renderer_b200.build` "
                           "(or __graft_entry__.build()); there is no fallback path")
    lib = C.CDLL(LIB_PATH)
    vp, ci, cf, i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
    lib.unib200_version.restype = ci
    lib.unib200_last_error.restype = C.c_char_p
    lib.unib200_device_info.argtypes = [C.POINTER(ci)] * 3
    lib.unib200_set_pdl.argtypes = [ci]
    lib.unib200_set_pdl.restype = None
    if os.environ.get("UNIB200_PDL") is not None:      # A/B runs: 0 off, 1 on (trigger at CTA end), 2 early trigger
        lib.unib200_set_pdl(int(os.environ["UNIB200_PDL"]))
    lib.unib200_program_create.restype = vp
    lib.unib200_program_destroy.argtypes = [vp]
    lib.unib200_program_destroy.restype = None
    lib.unib200_program_num_launches.argtypes = [vp]
    lib.unib200_program_run.argtypes = [vp, vp]
    lib.unib200_program_graph_instantiate.argtypes = [vp, vp]
    lib.unib200_program_graph_launch.argtypes = [vp, vp]
    lib.unib200_program_set_lane.argtypes = [vp, ci]
    lib.unib200_program_barrier.argtypes = [vp]
    lib.unib200_program_num_ops.argtypes = [vp]
    lib.unib200_program_op_info.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                            C.POINTER(ci)]
    lib.unib200_program_op_desc.argtypes = [vp, ci]
    lib.unib200_program_op_desc.restype = C.c_char_p
    lib.unib200_program_profile.argtypes = [vp, vp, ci, C.POINTER(cf)]
    lib.unib200_conv_gemm.argtypes = [vp, C.POINTER(GemmDesc), vp]
    lib.unib200_conv_gemm_dual.argtypes = [vp, C.POINTER(GemmDesc), C.POINTER(GemmDesc), vp]
    lib.unib200_packed_k.argtypes = [ci, C.POINTER(Seg)]
    lib.unib200_packed_k.restype = C.c_size_t
    lib.unib200_pick_bn.argtypes = [ci, ci]
    lib.unib200_debug_set_trace.argtypes = [vp]
    lib.unib200_debug_set_trace.restype = None
    lib.unib200_attention.argtypes = [vp, C.POINTER(AttnDesc), vp]
    lib.unib200_groupnorm.argtypes = [vp, C.POINTER(GnDesc), vp]
    lib.unib200_layernorm.argtypes = [vp, vp, vp, vp, vp, ci, ci, cf, vp]
    lib.unib200_to_nhwc.argtypes = [vp, vp, ci, vp, ci, ci, ci, ci, i64, i64, i64, i64, ci, vp]
    lib.unib200_from_nhwc.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, vp]
    lib.unib200_upsample2x.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp]
    lib.unib200_timestep_sinusoid.argtypes = [vp, vp, vp, ci, vp, ci, ci, vp]
    lib.unib200_gemv.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, vp]
    lib.unib200_axpby.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp]
    lib.unib200_add_int.argtypes = [vp, vp, ci, vp]
    lib.unib200_add_f16.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.unib200_unipc_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp]
    lib.unib200_softmax_rows.argtypes = [vp, vp, ci, ci, ci, cf, vp]
    lib.unib200_gaussian_sample.argtypes = [vp, vp, vp, vp, ci, ci, ci, cf, vp]
    lib.unib200_conv_wgrad.argtypes = [vp, C.POINTER(WgradDesc), vp]
    lib.unib200_groupnorm_backward.argtypes = [vp, C.POINTER(GnBwdDesc), vp]
    lib.unib200_colsum.argtypes = [vp, vp, ci, ci, ci, vp, vp]
    lib.unib200_pack_master_weight.argtypes = [vp, vp, ci, ci, ci, vp, ci, vp]
    lib.unib200_wgrad_scatter_add.argtypes = [vp, vp, vp, ci, ci, ci, vp]
    lib.unib200_layernorm_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_size_t, ci, ci, cf, vp]
    lib.unib200_geglu.argtypes = [vp, vp, vp, vp, i64, ci, vp]
    lib.unib200_softmax_backward.argtypes = [vp, vp, vp, ci, ci, ci, cf, vp]
    lib.unib200_cvt_f32_f16.argtypes = [vp, vp, vp, i64, ci, ci, vp]
    lib.unib200_silu_f16.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.unib200_pool2x2_sum.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp]
    lib.unib200_scatter2x.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp]
    lib.unib200_attention_backward.argtypes = [vp, C.POINTER(AttnBwdDesc), vp]
    lib.unib200_adamw_step.argtypes = [vp, vp, vp, vp, vp, i64, cf, cf, cf, cf, cf, ci, cf, vp]
    lib.unib200_create.argtypes = [ci, vp]
    lib.unib200_create.restype = vp
    lib.unib200_destroy.argtypes = [vp]
    lib.unib200_destroy.restype = None
    lib.unib200_load_weight.argtypes = [vp, C.c_char_p, vp, ci, C.POINTER(i64), ci]
    lib.unib200_alloc.argtypes = [vp, C.c_char_p, C.c_size_t, ci]
    lib.unib200_bind.argtypes = [vp, C.c_char_p, vp, C.c_size_t]
    lib.unib200_buffer.argtypes = [vp, C.c_char_p, C.POINTER(C.c_size_t)]
    lib.unib200_buffer.restype = vp
    lib.unib200_ctx_attach.argtypes = [vp, C.c_char_p, vp]
    lib.unib200_ctx_run.argtypes = [vp, C.c_char_p, vp]
    for fn in (lib.unib200_unet_forward, lib.unib200_attr_enc_forward, lib.unib200_attr_dec_forward, lib.unib200_dual_step):
        fn.argtypes = [vp, vp]
    lib.unib200_sample_loop.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp]
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().unib200_last_error()
        raise Unib200Error(f"{what}: {msg.decode() if msg else rc}")
