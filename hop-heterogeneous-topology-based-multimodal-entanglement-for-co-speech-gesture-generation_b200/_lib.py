"""ctypes binding of libhopk.so -- the only door between the Python mirror of the reference's
module interface and the sm_100a kernels (C ABI declared in include/hopk.h).

There is deliberately no fallback: if the shared library is missing or a kernel reports an error
the call raises, so a silent eager/CPU path can never stand in for the CUDA one.
"""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'libhopk.so')
MAX_LAYERS = 16
GEMM_A_MN, GEMM_B_MN, GEMM_OUT_BF16, GEMM_ACCUMULATE, GEMM_RELU, GEMM_LEAKY, GEMM_GELU = 1, 2, 4, 8, 16, 32, 64
GEMM_BIAS_ROW, GEMM_MASK_BF16, GEMM_MASK_GELU = 128, 256, 512

_vp = C.c_void_p
_LAYER_ARR = _vp * MAX_LAYERS


class GwnetShape(C.Structure):
    _fields_ = [('B', C.c_int), ('V', C.c_int), ('T', C.c_int), ('in_dim', C.c_int), ('out_dim', C.c_int),
                ('C', C.c_int), ('S', C.c_int), ('E', C.c_int), ('L', C.c_int), ('dil', C.c_int * MAX_LAYERS),
                ('rank', C.c_int), ('training', C.c_int), ('dtype', C.c_int),
                ('bn_momentum', C.c_float), ('bn_eps', C.c_float)]


class GwnetParams(C.Structure):
    _fields_ = [('nodevec1', _vp), ('nodevec2', _vp), ('start_w', _vp), ('start_b', _vp),
                ('filter_w', _LAYER_ARR), ('filter_b', _LAYER_ARR), ('gate_w', _LAYER_ARR), ('gate_b', _LAYER_ARR),
                ('skip_w', _LAYER_ARR), ('skip_b', _LAYER_ARR), ('mlp_w', _LAYER_ARR), ('mlp_b', _LAYER_ARR),
                ('bn_w', _LAYER_ARR), ('bn_b', _LAYER_ARR), ('bn_mean', _LAYER_ARR), ('bn_var', _LAYER_ARR),
                ('bn_nbt', _LAYER_ARR), ('end1_w', _vp), ('end1_b', _vp), ('end2_w', _vp), ('end2_b', _vp)]


class GwnetGrads(C.Structure):
    _fields_ = [('nodevec1', _vp), ('nodevec2', _vp), ('start_w', _vp), ('start_b', _vp),
                ('filter_w', _LAYER_ARR), ('filter_b', _LAYER_ARR), ('gate_w', _LAYER_ARR), ('gate_b', _LAYER_ARR),
                ('skip_w', _LAYER_ARR), ('skip_b', _LAYER_ARR), ('mlp_w', _LAYER_ARR), ('mlp_b', _LAYER_ARR),
                ('bn_w', _LAYER_ARR), ('bn_b', _LAYER_ARR),
                ('end1_w', _vp), ('end1_b', _vp), ('end2_w', _vp), ('end2_b', _vp),
                ('flat', _vp), ('flat_bytes', C.c_size_t)]


GRU_MAX_LAYERS = 8
_GRU_ARR = (_vp * 2) * GRU_MAX_LAYERS


class GruShape(C.Structure):
    _fields_ = [('B', C.c_int), ('T', C.c_int), ('I', C.c_int), ('H', C.c_int), ('L', C.c_int), ('save', C.c_int)]


class GruParams(C.Structure):
    _fields_ = [('w_ih', _GRU_ARR), ('w_hh', _GRU_ARR), ('b_ih', _GRU_ARR), ('b_hh', _GRU_ARR)]


GruGrads = GruParams            # same layout: pointers indexed [layer][direction]


# every symbol include/hopk.h declares, with its ctypes signature (restype int unless noted)
_i, _f, _u64, _sz = C.c_int, C.c_float, C.c_uint64, C.c_size_t
_SHP, _PRM, _GRD = C.POINTER(GwnetShape), C.POINTER(GwnetParams), C.POINTER(GwnetGrads)
_I64x4 = C.c_int64 * 4
SIGNATURES = {
    'hopk_last_error': (C.c_char_p, []),
    'hopk_version': (_i, []),
    'hopk_launch_count': (C.c_longlong, []),
    'hopk_gwnet_workspace_bytes': (_sz, [_SHP]),
    'hopk_gwnet_scratch_bytes': (_sz, [_SHP]),
    'hopk_gwnet_out_steps': (_i, [_SHP]),
    'hopk_gwnet_ws_field': (_i, [_SHP, C.c_char_p, _i, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_i)]),
    'hopk_gwnet_forward': (_i, [_SHP, _PRM, _vp, _I64x4, _vp, _vp, _vp]),
    'hopk_gwnet_backward': (_i, [_SHP, _PRM, _vp, _I64x4, _vp, _vp, _vp, _GRD, _vp, _vp]),
    'hopk_adp_softmax_fwd': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'hopk_adp_softmax_bwd': (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    'hopk_gated_tcn_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'hopk_gated_tcn_bwd_scratch_bytes': (_sz, [_i, _i, _i, _i, _i]),
    'hopk_gated_tcn_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'hopk_gcn_scratch_bytes': (_sz, [_i, _i, _i, _i]),
    'hopk_gcn_diffuse_mlp_res_bnstat_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'hopk_gcn_diffuse_mlp_res_bnstat_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'hopk_bn_finalize': (_i, [_vp, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _vp]),
    'hopk_nconv_fwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'hopk_nconv_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'hopk_linear_fwd': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'hopk_linear_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'hopk_conv1x1_nchw_fwd': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'hopk_conv1x1_nchw_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'hopk_gemm_bf16': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, C.c_long, C.c_long, C.c_long, _i, _f, _i, _vp]),
    'hopk_cast_bf16': (_i, [_vp, _vp, C.c_long, _i, C.c_long, _i, C.c_long, _i, _vp]),
    'hopk_unfold_bf16': (_i, [_vp, _vp, _i, _i, _i, _i, C.c_long, C.c_long, _vp]),
    'hopk_colsum': (_i, [_vp, _vp, C.c_long, _i, C.c_long, _i, _vp]),
    'hopk_ln_fwd': (_i, [_vp, _vp, _i, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _vp]),
    'hopk_ln_bwd': (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    'hopk_gelu_bf16': (_i, [_vp, _vp, C.c_long, _vp]),
    'hopk_bert_attn_fwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'hopk_bert_attn_bwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'hopk_beat_rows_fwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    'hopk_beat_rows_bwd': (_i, [_vp, _vp, _vp, _i, _i, _i, C.c_long, _vp]),
    'hopk_step_losses_fwd': (_i, [_vp] * 7 + [_i, _i, _i, C.c_float, C.c_float, C.c_float, _vp, _vp, _vp, _vp]),
    'hopk_step_losses_bwd': (_i, [_vp] * 7 + [_i, _i, _i, C.c_float, C.c_float, C.c_float, _vp, _vp, _vp, _vp]),
    'hopk_gru_workspace_bytes': (_sz, [C.POINTER(GruShape)]),
    'hopk_gru_scratch_bytes': (_sz, [C.POINTER(GruShape)]),
    'hopk_gru_forward': (_i, [C.POINTER(GruShape), C.POINTER(GruParams), _vp, _vp, _vp, _vp]),
    'hopk_gru_backward': (_i, [C.POINTER(GruShape), C.POINTER(GruParams), _vp, _vp, _vp, C.POINTER(GruParams), _vp, _vp]),
    'hopk_xattn_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _u64, _vp]),
    'hopk_dropout_epoch_advance': (_i, [_i, _vp]),
    'hopk_xattn_pack_bytes': (_sz, [_i, _i]),
    'hopk_xattn_fwd_tc': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _u64, _vp]),
    'hopk_xattn_bwd_scratch_bytes': (_sz, [_i, _i, _i]),
    'hopk_xattn_bwd_tc': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _u64, _vp]),
    'hopk_xattn_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _u64, _vp]),
}

_lib = None


def lib():
    """Load (once) and return the CDLL; raises if libhopk.so has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python __graft_entry__.py build` '
                '(nvcc, sm_100a). hop_b200 has no CPU / eager fallback.')
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError here == header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(lib().hopk_last_error().decode() + f' (code {rc})')


def stream_ptr():
    return _vp(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL); refuses host tensors so nothing can run on the CPU."""
    if t is None:
        return _vp(0)
    if not t.is_cuda:
        raise RuntimeError('hop_b200 kernels need CUDA tensors (no CPU fallback)')
    return _vp(t.data_ptr())


def f32c(t):
    """fp32 + contiguous view/copy of a CUDA tensor."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
