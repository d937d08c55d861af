"""
ctypes binding of libdlwp_b200.so (include/dlwp_b200.h).  This is the ONLY compute backend of the package: if the
library is missing or cannot be loaded the import fails loudly -- there is no CPU / PyTorch fallback path.
"""

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libdlwp_b200.so')

DLWP_OK = 0
ABI_VERSION = 3
PAD_ZERO, PAD_PERIODIC, PAD_EDGE, PAD_REFLECT, PAD_SYMMETRIC = 0, 1, 2, 3, 4
ACT_LINEAR, ACT_TANH, ACT_RELU = 0, 1, 2
IMPL_AUTO, IMPL_DIRECT, IMPL_FFMA, IMPL_FFMA_TMA, IMPL_TC = 0, 1, 2, 3, 4
BUF_INTERNAL, BUF_INPUT, BUF_OUTPUT = 0, 1, 2
OP_CONV, OP_PAD, OP_MAXPOOL, OP_UPSAMPLE, OP_COPY, OP_LSTM, OP_TO_NCHW, OP_TO_NHWC = 0, 1, 2, 3, 4, 5, 6, 7
RECURRENT_ACTIVATIONS = {'hard_sigmoid': 0, 'sigmoid': 1}
PRECISION_FP32, PRECISION_BF16 = 0, 1
ACTIVATIONS = {None: ACT_LINEAR, 'linear': ACT_LINEAR, 'tanh': ACT_TANH, 'relu': ACT_RELU}
IMPLS = {'auto': IMPL_AUTO, 'direct': IMPL_DIRECT, 'ffma': IMPL_FFMA, 'ffma_tma': IMPL_FFMA_TMA, 'tc': IMPL_TC}

i32, i64 = ctypes.c_int32, ctypes.c_int64
fptr = ctypes.c_void_p


class ConvDesc(ctypes.Structure):
    _fields_ = [(n, i32) for n in ('N', 'Cin', 'H', 'W', 'Cout', 'kh', 'kw', 'dil_h', 'dil_w', 'pad_t', 'pad_b',
                                   'pad_l', 'pad_r', 'pad_mode_h', 'pad_mode_w', 'act', 'pre_op', 'rowwise', 'impl',
                                   'reserved', 'row_begin', 'row_end')] + \
               [(n, i64) for n in ('x_stride_n', 'x_stride_c', 'x_stride_h', 'y_stride_n', 'y_stride_c',
                                   'y_stride_h')]


class BufferDesc(ctypes.Structure):
    _fields_ = [(n, i32) for n in ('kind', 'C', 'H', 'W', 'output_index', 'reserved')]


class OpDesc(ctypes.Structure):
    _fields_ = [(n, i32) for n in ('kind', 'src', 'src_c0', 'src_c', 'dst', 'dst_c0', 'weight_id', 'pad_t', 'pad_b',
                                   'pad_l', 'pad_r', 'pad_mode_h', 'pad_mode_w', 'Cout', 'kh', 'kw', 'dil_h', 'dil_w',
                                   'act', 'pre_op', 'rowwise', 'impl', 'row_begin', 'row_end', 'aux', 'aux_c0', 'act2')]


class BandInfo(ctypes.Structure):
    _fields_ = [(n, i32) for n in ('rank', 'world', 'band_lo', 'band_hi', 'recv_top', 'recv_bot', 'send_up', 'send_down')]


class PlanOptions(ctypes.Structure):
    """DlwpPlanOptions (include/dlwp_b200.h): all zeros = defaults; tc_taps_in_k = -1 leaves the choice to the planner."""
    _fields_ = [(n, i32) for n in ('math', 'fuse', 'tc_generic', 'tc_bands', 'tc_no_tma', 'tc_taps_in_k', 'tc_debug',
                                   'precision', 'latband_spare_sms')] + \
               [('reserved', i32 * 7)]

    def __init__(self, **kw):
        super(PlanOptions, self).__init__()
        self.tc_taps_in_k = -1
        for k, v in kw.items():
            if k not in [f[0] for f in self._fields_]:
                raise TypeError('unknown plan option %r' % k)
            setattr(self, k, int(v))


MATH_AUTO, MATH_FFMA = 0, 1
FLAG_FFMA_TIMEOUT, FLAG_TC_TIMEOUT, FLAG_TC_RANGE, FLAG_TC_UNDERFLOW = 1, 2, 4, 8


class NetDesc(ctypes.Structure):
    _fields_ = [('n_buffers', i32), ('n_ops', i32), ('n_weights', i32), ('max_batch', i32),
                ('buffers', ctypes.POINTER(BufferDesc)), ('ops', ctypes.POINTER(OpDesc))]


# every symbol include/dlwp_b200.h declares: (restype, argtypes)
SYMBOLS = {
    'dlwp_conv2d_fwd': (ctypes.c_int, [ctypes.POINTER(ConvDesc), fptr, fptr, fptr, fptr, ctypes.c_void_p]),
    'dlwp_pad2d': (ctypes.c_int, [fptr, fptr] + [i32] * 10 + [i64] * 6 + [ctypes.c_void_p]),
    'dlwp_gather_series': (ctypes.c_int, [fptr] * 4 + [i32] * 9 + [ctypes.c_void_p]),
    'dlwp_layout2d': (ctypes.c_int, [fptr, fptr] + [i32] * 5 + [ctypes.c_void_p]),
    'dlwp_maxpool2d': (ctypes.c_int, [fptr, fptr] + [i32] * 4 + [i64] * 6 + [ctypes.c_void_p]),
    'dlwp_upsample2d': (ctypes.c_int, [fptr, fptr] + [i32] * 4 + [i64] * 6 + [ctypes.c_void_p]),
    'dlwp_copy4d': (ctypes.c_int, [fptr, fptr] + [i32] * 4 + [i64] * 6 + [ctypes.c_void_p]),
    'dlwp_conv2d_bwd_input': (ctypes.c_int, [ctypes.POINTER(ConvDesc), fptr, fptr, fptr, ctypes.c_void_p]),
    'dlwp_conv2d_bwd_weight': (ctypes.c_int, [ctypes.POINTER(ConvDesc), fptr, fptr, fptr, fptr, ctypes.c_void_p]),
    'dlwp_convlstm_gates': (ctypes.c_int, [fptr] * 5 + [i32] * 4 + [i64] * 4 + [i32] * 4 + [ctypes.c_void_p]),
    'dlwp_rows_op': (ctypes.c_int, [i32, fptr, fptr] + [i32] * 10 + [i64] * 6 + [i32, i32, ctypes.c_void_p]),
    'dlwp_plan_create': (ctypes.c_int, [ctypes.POINTER(NetDesc), ctypes.POINTER(ctypes.c_void_p)]),
    'dlwp_plan_create_opts': (ctypes.c_int, [ctypes.POINTER(NetDesc), ctypes.POINTER(PlanOptions),
                                              ctypes.POINTER(ctypes.c_void_p)]),
    'dlwp_plan_destroy': (None, [ctypes.c_void_p]),
    'dlwp_plan_set_weights': (ctypes.c_int, [ctypes.c_void_p, i32, fptr, i64, fptr, i64]),
    'dlwp_plan_get_weights': (ctypes.c_int, [ctypes.c_void_p, i32, fptr, i64, fptr, i64]),
    'dlwp_plan_forward': (ctypes.c_int, [ctypes.c_void_p, i32, fptr, ctypes.POINTER(ctypes.c_void_p),
                                         ctypes.c_void_p]),
    'dlwp_rollout': (ctypes.c_int, [ctypes.c_void_p, i32, fptr, fptr, i32, i32, ctypes.c_void_p]),
    'dlwp_rollout_host': (ctypes.c_int, [ctypes.c_void_p, i32, fptr, fptr, i32, i32]),
    'dlwp_comm_unique_id': (ctypes.c_int, [ctypes.c_char_p, ctypes.c_void_p]),
    'dlwp_comm_create': (ctypes.c_int, [ctypes.c_char_p, i32, i32, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    'dlwp_comm_destroy': (None, [ctypes.c_void_p]),
    'dlwp_rollout_latband': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, i32, fptr, fptr, i32,
                                            ctypes.POINTER(BandInfo), i32, ctypes.c_void_p]),
    'dlwp_rollout_latband_host': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, i32, fptr, fptr, i32,
                                                 ctypes.POINTER(BandInfo), i32]),
    'dlwp_train_step': (ctypes.c_int, [ctypes.c_void_p, i32, fptr, ctypes.POINTER(ctypes.c_void_p),
                                       ctypes.POINTER(ctypes.c_float), i32, i32, ctypes.POINTER(ctypes.c_float),
                                       ctypes.POINTER(ctypes.c_float), ctypes.c_void_p]),
    'dlwp_train_loss_weights': (ctypes.c_int, [ctypes.c_void_p, fptr, i64]),
    'dlwp_train_allreduce': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'dlwp_train_loss_kind': (ctypes.c_int, [ctypes.c_void_p, i32, i32, i32, fptr, i64]),
    'dlwp_train_buffers': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(i64),
                                          ctypes.POINTER(ctypes.c_void_p)]),
    'dlwp_train_weight_offsets': (ctypes.c_int, [ctypes.c_void_p, i32, ctypes.POINTER(i64), ctypes.POINTER(i64)]),
    'dlwp_train_regularize': (ctypes.c_int, [ctypes.c_void_p, i32, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                             ctypes.c_float, ctypes.POINTER(ctypes.c_float), ctypes.c_void_p]),
    'dlwp_train_adam_state': (ctypes.c_int, [ctypes.c_void_p, fptr, fptr, i64, ctypes.POINTER(i64), i32]),
    'dlwp_train_adam': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                       ctypes.c_void_p]),
    'dlwp_rollout_step_sequence': (ctypes.c_int, [ctypes.c_void_p, i32, fptr, fptr, i32, i32, ctypes.c_void_p]),
    'dlwp_plan_halo_enable': (ctypes.c_int, [ctypes.c_void_p]),
    'dlwp_plan_halo_export': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'dlwp_plan_halo_import': (ctypes.c_int, [ctypes.c_void_p, i32, ctypes.c_void_p]),
    'dlwp_plan_halo_connect': (ctypes.c_int, [ctypes.c_void_p, i32, ctypes.c_void_p]),
    'dlwp_plan_profile_op': (ctypes.c_int, [ctypes.c_void_p, i32, i32, i32, ctypes.POINTER(ctypes.c_float),
                                            ctypes.c_void_p]),
    'dlwp_plan_uses_tensor_cores': (ctypes.c_int, [ctypes.c_void_p]),
    'dlwp_plan_fused_pair': (ctypes.c_int, [ctypes.c_void_p]),
    'dlwp_last_error_string': (ctypes.c_char_p, []),
    'dlwp_abi_version': (ctypes.c_int, []),
    'dlwp_kernel_launch_count': (ctypes.c_int64, []),
    'dlwp_debug_flags': (ctypes.c_int, []),
    'dlwp_conv2d_impl_name': (ctypes.c_char_p, [ctypes.POINTER(ConvDesc)]),
    'dlwp_debug_tc_plan': (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(i32), i32]),
    'dlwp_debug_sw_cover': (ctypes.c_int, [ctypes.POINTER(ConvDesc), i32, ctypes.POINTER(i32), ctypes.c_int64,
                                           ctypes.POINTER(i32), i32]),
    'dlwp_debug_tc_pack': (ctypes.c_int64, [ctypes.POINTER(ConvDesc), ctypes.POINTER(ctypes.c_float),
                                            ctypes.POINTER(ctypes.c_uint16), ctypes.c_int64,
                                            ctypes.POINTER(ctypes.c_uint32), i32, ctypes.POINTER(i32),
                                            ctypes.POINTER(ctypes.c_float)]),
    'dlwp_debug_tc_folded': (ctypes.c_char_p, [ctypes.POINTER(ConvDesc), i32]),
    'dlwp_debug_counters': (ctypes.c_int, [ctypes.POINTER(ctypes.c_int64), i32]),
    'dlwp_debug_exp_for_bound': (ctypes.c_int, [ctypes.c_float]),
}


class NativeError(RuntimeError):
    pass


_lib = None


def lib():
    """Load (once) and return the shared library; raises if it is missing -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                'libdlwp_b200.so is not built (%s). Run `python -m dlwp_b200.build` (needs nvcc); dlwp_b200 has no '
                'CPU or PyTorch fallback for its CUDA kernels.' % LIB_PATH)
        # DLWP_B200_LIB: an alternative build of the same ABI (same-box A/B experiments, profiles/r02_variants_ab.txt)
        handle = ctypes.CDLL(os.environ.get('DLWP_B200_LIB') or LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the ABI and the binding drift apart
            fn.restype = res
            fn.argtypes = args
        if handle.dlwp_abi_version() != ABI_VERSION:
            raise ImportError('libdlwp_b200.so ABI version mismatch')
        _lib = handle
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().dlwp_last_error_string().decode('utf-8', 'replace')
        exc = ValueError if rc in (-1, -2) else NativeError
        raise exc('%s failed (code %d): %s' % (what or 'libdlwp_b200 call', rc, msg))


def launch_count():
    return int(lib().dlwp_kernel_launch_count())
