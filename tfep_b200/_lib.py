"""ctypes binding of the C ABI declared in include/tfep_b200.h.

There is no CPU or PyTorch fallback: if the CUDA library is missing, or a tensor is not on a CUDA
device, the calls below raise.  The library is built in-tree by ``tfep_b200._build.build()``.
"""

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_uint32, c_uint64, c_void_p

import torch

from . import _build

F32, F64 = 0, 1
ACT_NONE, ACT_ELU = 0, 1
GEMM_TILE_N = 64
ABI_VERSION = 5


class TfepB200Error(RuntimeError):
    pass


class LinearFwdArgs(Structure):
    _fields_ = [('dtype', c_int32), ('batch', c_int32), ('in_features', c_int32), ('out_features', c_int32),
                ('x', c_void_p), ('ldx', c_int64), ('w', c_void_p), ('ldw', c_int64), ('bias', c_void_p),
                ('y', c_void_p), ('ldy', c_int64), ('activation', c_int32), ('reserved', c_int32),
                ('k_ranges', c_void_p)]


class LinearBwdInputArgs(Structure):
    _fields_ = [('dtype', c_int32), ('batch', c_int32), ('in_features', c_int32), ('out_features', c_int32),
                ('grad_y', c_void_p), ('ldgy', c_int64), ('w', c_void_p), ('ldw', c_int64),
                ('grad_x', c_void_p), ('ldgx', c_int64), ('act_out', c_void_p), ('ldact', c_int64),
                ('accumulate', c_int32), ('reserved', c_int32), ('n_ranges', c_void_p)]


class LinearBwdWeightArgs(Structure):
    _fields_ = [('dtype', c_int32), ('batch', c_int32), ('in_features', c_int32), ('out_features', c_int32),
                ('grad_y', c_void_p), ('ldgy', c_int64), ('x', c_void_p), ('ldx', c_int64),
                ('grad_w', c_void_p), ('ldgw', c_int64), ('grad_bias', c_void_p), ('n_ranges', c_void_p)]


class TxIo(Structure):
    _fields_ = [('dtype', c_int32), ('batch', c_int32), ('n_features', c_int32), ('inverse', c_int32),
                ('x', c_void_p), ('ldx', c_int64), ('y', c_void_p), ('ldy', c_int64),
                ('par', c_void_p), ('ldp', c_int64),
                ('par_offset', c_int64), ('par_stride_p', c_int64), ('par_stride_f', c_int64),
                ('par_base', c_void_p), ('cols', c_void_p), ('feat_ids', c_void_p), ('logdet', c_void_p),
                ('accumulate_logdet', c_int32), ('reserved', c_int32)]


class SplineCfg(Structure):
    _fields_ = [('n_bins', c_int32), ('circular', c_int32), ('identity_boundary_slopes', c_int32),
                ('learn_lower_bound', c_int32), ('learn_upper_bound', c_int32), ('reserved', c_int32),
                ('x0', c_void_p), ('xf', c_void_p), ('y0', c_void_p), ('yf', c_void_p),
                ('min_bin_size', c_double), ('min_slope', c_double),
                ('bins_out', c_void_p), ('ldbins', c_int64)]


class TxGrads(Structure):
    _fields_ = [('grad_y', c_void_p), ('ldgy', c_int64), ('grad_logdet', c_void_p),
                ('grad_x', c_void_p), ('ldgx', c_int64), ('grad_par', c_void_p)]


class FusedLayer(Structure):
    _fields_ = [('ops', c_void_p), ('n_ops', c_int32), ('n_chunks', c_int32), ('weights', c_void_p), ('feats', c_void_p),
                ('min_bin_size', c_float), ('min_slope', c_float), ('slope_offset', c_float), ('reserved', c_int32),
                ('input_map', c_void_p), ('emb_lower', c_float), ('emb_scale', c_float), ('hidden_split', c_int32 * 2)]


class FusedArgs(Structure):
    _fields_ = [('x', c_void_p), ('y', c_void_p), ('logdet', c_void_p), ('batch', c_int32), ('n_features', c_int32),
                ('k1', c_int32), ('hidden_padded', c_int32), ('n_layers', c_int32), ('hidden_halves', c_int32), ('hidden_split', c_int32 * 2),
                ('layers', POINTER(FusedLayer)), ('tile_flags', c_void_p), ('epoch', c_uint32), ('debug_mode', c_int32),
                ('mixed_splines', c_int32), ('n_inputs', c_int32), ('x_operand_column', c_int32), ('reserved3', c_int32),
                ('error_flag', c_void_p), ('debug_params', c_void_p)]


class FusedInvLayer(Structure):
    _fields_ = [('ops', c_void_p), ('steps', c_void_p), ('n_ops', c_int32), ('n_steps', c_int32), ('weights', c_void_p),
                ('min_bin_size', c_float), ('min_slope', c_float), ('slope_offset', c_float), ('reserved', c_int32),
                ('emb_lower', c_float), ('emb_scale', c_float), ('init_map', c_void_p)]


class FusedInvArgs(Structure):
    _fields_ = [('y', c_void_p), ('x', c_void_p), ('logdet', c_void_p), ('batch', c_int32), ('n_features', c_int32),
                ('k1', c_int32), ('hidden_padded', c_int32), ('n_layers', c_int32), ('mixed_splines', c_int32),
                ('layers', POINTER(FusedInvLayer)), ('tile_flags', c_void_p), ('epoch', c_uint32), ('n_inputs', c_int32),
                ('error_flag', c_void_p)]


class TcTx(Structure):
    _fields_ = [('kind', c_int32), ('backward', c_int32), ('n_units', c_int32), ('unit_sphere', c_int32),
                ('max_radius', c_double), ('cols', c_void_p), ('x', c_void_p), ('ldx', c_int64), ('y', c_void_p),
                ('ldy', c_int64), ('logdet', c_void_p), ('grad_y', c_void_p), ('ldgy', c_int64), ('grad_logdet', c_void_p),
                ('grad_x', c_void_p), ('ldgx', c_int64), ('spline_table', c_void_p)]


class TcGemmArgs(Structure):
    _fields_ = [('a_image', c_void_p), ('b_image', c_void_p), ('m', c_int32), ('n', c_int32), ('k', c_int32),
                ('activation', c_int32), ('c', c_void_p), ('ldc', c_int64), ('bias', c_void_p), ('aux', c_void_p),
                ('ldaux', c_int64), ('out_image', c_void_p), ('k_block_ranges', c_void_p), ('split_k', c_int32),
                ('out_image_t_rows', c_int32), ('error_flag', c_void_p), ('row_ranges', c_void_p),
                ('out_image_t', c_void_p), ('column_sums', c_void_p), ('n_split', c_int32), ('c_accumulate', c_int32),
                ('tx', POINTER(TcTx)), ('mn_major', c_int32), ('cluster', c_int32), ('tile_list', c_void_p),
                ('n_tile_list', c_int32), ('reserved3', c_int32), ('aux_image', c_void_p)]


class CentroidArgs(Structure):
    _fields_ = [('dtype', c_int32), ('batch', c_int32), ('n_features', c_int32), ('space_dimension', c_int32),
                ('x', c_void_p), ('ldx', c_int64), ('y_propagated', c_void_p), ('ldy', c_int64), ('out', c_void_p),
                ('ldout', c_int64), ('shift', c_void_p), ('n_propagated', c_int32), ('propagated_columns', c_void_p),
                ('column_to_propagated', c_void_p), ('centroid_points', c_void_p), ('n_centroid_points', c_int32),
                ('weights', c_void_p), ('origin', c_double * 4), ('fixed_point', c_int32), ('fixed_slot', c_int32),
                ('restore_fixed_point', c_int32), ('translate_back', c_int32)]


class OrientedArgs(Structure):
    _fields_ = [('dtype', c_int32), ('batch', c_int32), ('n_features', c_int32), ('n_propagated', c_int32),
                ('x', c_void_p), ('ldx', c_int64), ('y_propagated', c_void_p), ('ldy', c_int64), ('out', c_void_p),
                ('ldout', c_int64), ('rotation', c_void_p), ('propagated_columns', c_void_p),
                ('column_to_propagated', c_void_p), ('axis_point', c_int32), ('plane_point', c_int32), ('axis', c_int32),
                ('plane_axis', c_int32), ('round_off_imprecisions', c_int32), ('rotate_back', c_int32)]


class SweepArgs(Structure):
    _fields_ = [('dtype', c_int32), ('batch', c_int32), ('n_features', c_int32), ('n_linear', c_int32),
                ('y', c_void_p), ('ldy', c_int64), ('x', c_void_p), ('ldx', c_int64), ('logdet', c_void_p),
                ('w', c_void_p * 5), ('b', c_void_p * 5), ('n_out', c_int32 * 5), ('ldw', c_int32 * 5),
                ('groups', c_void_p), ('n_groups', c_int32), ('max_params', c_int32),
                ('parts', c_void_p), ('group_parts', c_void_p), ('ids', c_void_p), ('fixed_cols', c_void_p),
                ('n_fixed', c_int32), ('n_embedded', c_int32), ('emb_out_col', c_void_p), ('emb_periodic', c_void_p),
                ('emb_lower', c_double), ('emb_scale', c_double), ('reserved', c_int32), ('max_group_weight_elems', c_int32),
                ('extra', c_void_p * 5), ('ldextra', c_int64 * 5), ('act_out', c_void_p * 5), ('ldact_out', c_int64 * 5)]


# every symbol include/tfep_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    'tfepb_abi_version': (c_int32, []),
    'tfepb_last_error': (c_char_p, []),
    'tfepb_device_info': (c_int32, [POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    'tfepb_masked_linear_forward': (c_int32, [POINTER(LinearFwdArgs), c_void_p]),
    'tfepb_masked_linear_backward_input': (c_int32, [POINTER(LinearBwdInputArgs), c_void_p]),
    'tfepb_masked_linear_backward_weight': (c_int32, [POINTER(LinearBwdWeightArgs), c_void_p]),
    'tfepb_affine': (c_int32, [POINTER(TxIo), c_void_p]),
    'tfepb_spline': (c_int32, [POINTER(TxIo), POINTER(SplineCfg), c_void_p]),
    'tfepb_shift': (c_int32, [POINTER(TxIo), c_void_p, c_void_p, c_void_p]),
    'tfepb_shift_backward': (c_int32, [POINTER(TxIo), c_void_p, c_void_p, POINTER(TxGrads), c_void_p]),
    'tfepb_sos': (c_int32, [POINTER(TxIo), c_int32, c_void_p]),
    'tfepb_moebius': (c_int32, [POINTER(TxIo), c_int32, c_double, c_int32, c_void_p]),
    'tfepb_affine_backward': (c_int32, [POINTER(TxIo), POINTER(TxGrads), c_void_p]),
    'tfepb_spline_backward': (c_int32, [POINTER(TxIo), POINTER(SplineCfg), POINTER(TxGrads), c_void_p]),
    'tfepb_sos_backward': (c_int32, [POINTER(TxIo), c_int32, POINTER(TxGrads), c_void_p]),
    'tfepb_moebius_backward': (c_int32, [POINTER(TxIo), c_int32, c_double, c_int32, POINTER(TxGrads), c_void_p]),
    'tfepb_tc_image_bytes': (c_int64, [c_int64, c_int64, c_int32]),
    'tfepb_tc_pack': (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    'tfepb_tc_pack_split': (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    'tfepb_tc_pack_dual': (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p,
                                     c_void_p]),
    'tfepb_tc_gemm': (c_int32, [POINTER(TcGemmArgs), c_void_p]),
    'tfepb_wn_pack': (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_void_p, c_int32,
                                c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    'tfepb_wn_pack_backward': (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int32,
                                         c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    'tfepb_centroid_pre': (c_int32, [POINTER(CentroidArgs), c_void_p]),
    'tfepb_centroid_post': (c_int32, [POINTER(CentroidArgs), c_void_p]),
    'tfepb_oriented_pre': (c_int32, [POINTER(OrientedArgs), c_void_p]),
    'tfepb_oriented_post': (c_int32, [POINTER(OrientedArgs), c_void_p]),
    'tfepb_periodic_embedding': (c_int32, [c_int32, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_double, c_double,
                                           c_void_p, c_int64, c_void_p]),
    'tfepb_periodic_embedding_backward': (c_int32, [c_int32, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_double,
                                                    c_double, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    'tfepb_maf_spline_forward_bf16': (c_int32, [POINTER(FusedArgs), c_void_p]),
    'tfepb_maf_spline_inverse_bf16': (c_int32, [POINTER(FusedInvArgs), c_void_p]),
    'tfepb_maf_inverse_sweep': (c_int32, [POINTER(SweepArgs), c_void_p]),
    'tfepb_lse_workspace_bytes': (c_int64, []),
    'tfepb_lse': (c_int32, [c_int32, c_void_p, c_void_p, c_int64, c_double, c_void_p, c_void_p, c_void_p]),
    'tfepb_fep_estimate': (c_int32, [c_int32, c_void_p, c_void_p, c_int64, c_double, c_double, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    'tfepb_kl_loss_workspace_bytes': (c_int64, []),
    'tfepb_kl_loss': (c_int32, [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    'tfepb_kl_loss_backward': (c_int32, [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'tfepb_mt19937_seed': (c_int32, [c_uint32, c_void_p]),
    'tfepb_mt19937_indices': (c_int32, [c_void_p, c_int64, c_uint32, c_void_p, c_void_p]),
    'tfepb_mt19937_parallel_workspace_bytes': (c_int64, [c_int32]),
    'tfepb_mt19937_indices_parallel': (c_int32, [c_void_p, c_int64, c_uint32, c_void_p, c_int32, c_void_p, c_void_p]),
    'tfepb_mt19937_jump_polynomial': (c_int32, [c_uint64, c_void_p]),
    'tfepb_exp_table': (c_int32, [c_int32, c_void_p, c_int64, c_double, c_void_p, c_void_p, c_void_p]),
    'tfepb_bayesian_bootstrap_sums': (c_int32, [c_void_p, c_int64, c_int32, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p]),
    'tfepb_bootstrap_sums': (c_int32, [c_void_p, c_int64, c_int64, c_uint32, c_void_p, c_int64, c_int32, c_int64, c_uint64,
                                       c_uint64, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def library_path():
    return _build.LIBPATH


def load():
    """Load libtfep_b200.so (building it is the job of __graft_entry__.build / tfep_b200._build)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise TfepB200Error(
                f'{path} not found: build it with `python -m tfep_b200._build` (needs nvcc). '
                'tfep_b200 has no CPU / PyTorch fallback.')
        lib = ctypes.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)      # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        if lib.tfepb_abi_version() != ABI_VERSION:
            raise TfepB200Error('libtfep_b200.so ABI version mismatch; rebuild the library')
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().tfepb_last_error()
        raise TfepB200Error(f'tfep_b200 error {rc}: {msg.decode() if msg else "unknown"}')


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float64:
        return F64
    raise TfepB200Error(f'unsupported dtype {t.dtype}: tfep_b200 kernels compute in float32 or float64')


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise TfepB200Error('tfep_b200 runs on CUDA tensors only (sm_100a); there is no CPU fallback. '
                                'Move the module and its inputs to a B200 device.')


def stream_ptr(t: torch.Tensor):
    return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def ptr(t):
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())
