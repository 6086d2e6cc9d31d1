"""ctypes binding of ``libonetrans_sm100.so`` (C-ABI declared in ``include/onetrans_b200.h``).

The structures below mirror the header field by field.  There is no fallback: if the shared library is
missing or a call fails, the error is raised (north_star: no CPU fallback, no dispatch)."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# OT_LIB_PATH: an alternative build of the same library (A/B experiments with compile-time switches); never another implementation
LIB_PATH = os.environ.get('OT_LIB_PATH') or os.path.join(_HERE, 'lib', 'libonetrans_sm100.so')

i32, i64, vp, fp = C.c_int32, C.c_int64, C.c_void_p, C.c_void_p  # float* passed as raw address

OT_EPI_BIAS, OT_EPI_GELU, OT_EPI_RESIDUAL, OT_EPI_GELU_GRAD, OT_EPI_ROW_SCALE, OT_EPI_DROPOUT, OT_EPI_NORM = 1, 2, 4, 8, 16, 32, 64


class GemmSeg(C.Structure):
    _fields_ = [('row_start', i32), ('n_units', i32), ('rows_per_unit', i32), ('group_start', i32),
                ('group_stride', i32), ('a_row_start', i32)]


class GemmParams(C.Structure):
    _fields_ = [('A', vp), ('a_dim1', i64), ('a_dim2', i64), ('a_stride1', i64), ('a_stride2', i64),
                ('a_transposed', i32), ('n_groups', i32), ('W', vp), ('ldw', i64), ('N', i32), ('K', i32),
                ('n_segs', i32), ('flags', i32), ('segs', GemmSeg * 3),
                ('out', vp), ('ldo', i64), ('out2', vp), ('ldo2', i64), ('res', vp), ('ldr', i64),
                ('aux', vp), ('ldaux', i64), ('bias', fp), ('bias_group_stride', i64), ('row_scale', fp),
                ('block_n', i32), ('swizzle', i32), ('res_hp', fp), ('out_hp', fp), ('ld_hp', i64), ('hp_row0', i64), ('drop_seed', C.c_uint32), ('drop_rate', C.c_float),
                ('norm_out', vp), ('ld_norm', i64), ('norm_gain', fp), ('norm_rstd', fp), ('norm_eps', C.c_float)]


class FfnParams(C.Structure):
    _fields_ = [('zn', vp), ('ldzn', i64), ('W1', vp), ('ldw1', i64), ('W2', vp), ('ldw2', i64),
                ('b1', fp), ('b1_group_stride', i64), ('b2', fp), ('b2_group_stride', i64),
                ('n_groups', i32), ('d', i32), ('F', i32), ('n_segs', i32), ('flags', i32), ('segs', GemmSeg * 3),
                ('out', vp), ('ldo', i64), ('pre', vp), ('ldpre', i64), ('h', vp), ('ldh', i64), ('res', vp), ('ldr', i64),
                ('res_hp', fp), ('out_hp', fp), ('ld_hp', i64), ('hp_row0', i64), ('drop_seed', C.c_uint32), ('drop_rate', C.c_float),
                ('norm_out', vp), ('ld_norm', i64), ('norm_gain', fp), ('norm_rstd', fp), ('norm_eps', C.c_float)]


class WgradSeg(C.Structure):
    _fields_ = [('P', vp), ('p_stride_row', i64), ('p_stride_unit', i64), ('Q', vp), ('q_stride_row', i64),
                ('q_stride_unit', i64), ('n_units', i32), ('rows_per_unit', i32), ('group_start', i32),
                ('group_stride', i32)]


class WgradParams(C.Structure):
    _fields_ = [('Mdim', i32), ('Ndim', i32), ('n_segs', i32), ('swizzle', i32), ('segs', WgradSeg * 2),
                ('C', fp), ('c_group_stride', i64), ('c_stride_m', i64), ('c_stride_n', i64),
                ('p_row_scale', fp), ('block_n', i32), ('target_ctas', i32), ('q_colsum', fp), ('q_colsum_group_stride', i64),
                ('p_gelu', i32)]


class AttnParams(C.Structure):
    _fields_ = [('q', vp), ('ldq', i64), ('k', vp), ('ldk', i64), ('v', vp), ('ldv', i64), ('o', vp), ('ldo', i64),
                ('lse', fp), ('d_o', vp), ('lddo', i64), ('dq', vp), ('lddq', i64), ('dk', vp), ('lddk', i64),
                ('dv', vp), ('lddv', i64), ('delta', fp), ('B', i32), ('H', i32), ('Lq', i32), ('Lk', i32),
                ('head_dim', i32), ('swizzle', i32)]


class AttnCachedParams(C.Structure):
    _fields_ = [('q', vp), ('ldq', i64), ('k_own', vp), ('ld_own_k', i64), ('v_own', vp), ('ld_own_v', i64),
                ('k_shared', vp), ('ld_shared_k', i64), ('v_shared', vp), ('ld_shared_v', i64), ('o', vp), ('ldo', i64),
                ('C', i32), ('H', i32), ('Tq', i32), ('Tn', i32), ('Ls', i32), ('head_dim', i32)]


class RmsnormParams(C.Structure):
    _fields_ = [('x', vp), ('ldx', i64), ('y', vp), ('ldy', i64), ('gain', fp), ('rstd', fp),
                ('dy', vp), ('lddy', i64), ('dres', vp), ('lddres', i64), ('dx', vp), ('lddx', i64),
                ('dgain', fp), ('rows', i64), ('d', i32), ('eps', C.c_float), ('x_hp', fp), ('hp_row0', i64),
                ('dx_drop', vp), ('lddx_drop', i64), ('drop_row0', i64), ('drop_seed', C.c_uint32), ('drop_rate', C.c_float)]


class NsTokenizerParams(C.Structure):
    _fields_ = [('x', fp), ('W', fp), ('bias', fp), ('out', vp), ('dout', vp), ('ldo', i64), ('dW', fp),
                ('dbias', fp), ('row0', i64), ('B', i32), ('L_ns', i32), ('d', i32), ('n_feat', i32), ('out_hp', fp)]


class ColsumParams(C.Structure):
    _fields_ = [('in_', vp), ('ld', i64), ('row_start', i64), ('n_units', i32), ('rows_per_unit', i32),
                ('group_start', i32), ('group_stride', i32), ('out', fp), ('out_group_stride', i64), ('N', i32)]


class RmspropParams(C.Structure):
    _fields_ = [('param_ptrs', vp), ('seg_off', vp), ('seg_numel', vp), ('n_seg', i32), ('n_flat', i64),
                ('grad', fp), ('rms', fp), ('mom', fp), ('sqnorm', fp),
                ('lr', C.c_float), ('rho', C.c_float), ('momentum', C.c_float), ('eps', C.c_float),
                ('clip_norm', C.c_float), ('grad_scale', C.c_float), ('zero_grad', i32), ('seg_slot', vp), ('n_slots', i32)]


class EmbedParams(C.Structure):
    _fields_ = [('table', fp), ('field_off', vp), ('field_rows', vp), ('ids', vp), ('events', vp), ('ld_events', i64),
                ('n_events', i64), ('n_fields', i32), ('ef', i32), ('bad_ids', vp), ('grad', fp), ('acc', fp), ('stamp', vp),
                ('step_id', i32), ('lr', C.c_float), ('eps', C.c_float)]


MAX_TASKS = 4  # OT_MAX_TASKS


class HeadsParams(C.Structure):
    _fields_ = [('x', fp), ('ldx', i64), ('gain', fp), ('eps', C.c_float), ('B', i32), ('d', i32), ('hidden', i32), ('n_tasks', i32),
                ('W0', fp * MAX_TASKS), ('b0', fp * MAX_TASKS), ('W1', fp * MAX_TASKS), ('b1', fp * MAX_TASKS),
                ('xn', fp), ('rstd', fp), ('pre', fp), ('logits', fp), ('probs', fp), ('labels', fp), ('loss', fp), ('g_bce', fp),
                ('dlogit', fp), ('dpre', fp),
                ('dW0', fp * MAX_TASKS), ('db0', fp * MAX_TASKS), ('dW1', fp * MAX_TASKS), ('db1', fp * MAX_TASKS),
                ('dgain', fp), ('dx', fp), ('lddx', i64)]


class MetricsParams(C.Structure):
    _fields_ = [('probs', fp), ('labels', fp), ('ld', i64), ('B', i64), ('n_tasks', i32), ('num_thresholds', i32),
                ('threshold', C.c_float), ('state', vp), ('state_stride', i64), ('result', vp)]


class AucParams(C.Structure):
    _fields_ = [('probs', fp), ('labels', fp), ('segment_ids', vp), ('n', i64), ('n_segments', i32), ('keys', vp),
                ('rejected', vp), ('seg_count', vp), ('seg_pos', vp), ('seg_sum2', vp)]


METRICS_MAX_THRESHOLDS, METRICS_TAIL_WORDS, METRICS_RESULT_WORDS = 512, 8, 8  # OT_METRICS_*
OPT_CHUNK = 1024  # OT_OPT_CHUNK
ABI_VERSION = 14  # OT_ABI_VERSION of include/onetrans_b200.h this binding was written against

# every symbol include/onetrans_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    'ot_version', 'ot_last_error_string', 'ot_num_sms', 'ot_mixed_gemm', 'ot_ffn_fwd', 'ot_ffn_bwd', 'ot_wgrad', 'ot_attn_fwd', 'ot_attn_bwd', 'ot_attn_ns_cached_fwd',
    'ot_rmsnorm_fwd', 'ot_rmsnorm_bwd', 'ot_ns_tokenizer_fwd', 'ot_ns_tokenizer_bwd', 'ot_fill_rows', 'ot_colsum', 'ot_dropout_mask',
    'ot_clip_rmsprop_step', 'ot_embed_gather_fwd', 'ot_embed_scatter_bwd', 'ot_embed_adagrad_step',
    'ot_heads_fwd', 'ot_heads_bwd', 'ot_metrics_update', 'ot_metrics_result', 'ot_auc_pack_keys', 'ot_auc_ranksum',
]

_lib = None
_lock = threading.Lock()


class OneTransLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA extension, or raise.  Never substitutes anything else."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise OneTransLibraryError(
                f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                f'or recommend_b200/csrc/build.sh — there is no CPU fallback')
        lib = C.CDLL(LIB_PATH)
        lib.ot_version.restype = C.c_int
        if lib.ot_version() != ABI_VERSION:
            raise OneTransLibraryError(f'{LIB_PATH} has ABI version {lib.ot_version()}, this binding needs {ABI_VERSION}: rebuild it '
                                       f'(recommend_b200/csrc/build.sh)')
        lib.ot_last_error_string.restype = C.c_char_p
        lib.ot_num_sms.restype = C.c_int
        for name, st in [('ot_mixed_gemm', GemmParams), ('ot_ffn_fwd', FfnParams), ('ot_ffn_bwd', FfnParams), ('ot_wgrad', WgradParams), ('ot_attn_fwd', AttnParams),
                         ('ot_attn_bwd', AttnParams), ('ot_attn_ns_cached_fwd', AttnCachedParams), ('ot_rmsnorm_fwd', RmsnormParams), ('ot_rmsnorm_bwd', RmsnormParams),
                         ('ot_ns_tokenizer_fwd', NsTokenizerParams), ('ot_ns_tokenizer_bwd', NsTokenizerParams),
                         ('ot_colsum', ColsumParams), ('ot_clip_rmsprop_step', RmspropParams), ('ot_embed_gather_fwd', EmbedParams),
                         ('ot_embed_scatter_bwd', EmbedParams), ('ot_embed_adagrad_step', EmbedParams),
                         ('ot_heads_fwd', HeadsParams), ('ot_heads_bwd', HeadsParams),
                         ('ot_metrics_update', MetricsParams), ('ot_metrics_result', MetricsParams),
                         ('ot_auc_pack_keys', AucParams), ('ot_auc_ranksum', AucParams)]:
            fn = getattr(lib, name)
            fn.argtypes = [C.POINTER(st), C.c_void_p]
            fn.restype = C.c_int
        lib.ot_fill_rows.argtypes = [C.c_void_p, C.c_void_p, i64, i64, i64, i32, C.c_void_p]
        lib.ot_fill_rows.restype = C.c_int
        lib.ot_dropout_mask.argtypes = [C.c_void_p, i64, C.c_void_p, i64, i64, i32, C.c_uint32, C.c_float, C.c_void_p]
        lib.ot_dropout_mask.restype = C.c_int
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().ot_last_error_string()
        raise OneTransLibraryError(f'{what} failed (rc={rc}): {msg.decode() if msg else ""}')


# number of kernels this process has enqueued through the library (bench.py reports it as gpu_launches)
launch_count = 0


def count_launch(n: int = 1) -> None:
    global launch_count
    launch_count += n
