"""ctypes binding of libjcm.so (include/jcm.h).  No fallback: if the library is missing this raises at import of the
first op, telling the user to build it (python joint-cnn-mrf_b200/build.py)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libjcm.so')

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_L = _c.c_long
_F = _c.c_float

# name -> (restype, argtypes); must list every symbol declared in include/jcm.h (tests/test_abi.py checks this)
SIGNATURES = {
    'jcm_last_error': (_c.c_char_p, []),
    'jcm_version': (_I, []),
    'jcm_sm_count': (_I, []),
    'jcm_launch_count': (_L, []),
    'jcm_prep_input': (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    'jcm_pack_weights': (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    'jcm_pack_weights_batch': (_I, [_P, _I, _P]),
    'jcm_pack_weights_s2d': (_I, [_P, _I, _P, _P, _P]),
    'jcm_split_planes': (_I, [_P, _L, _P, _P, _P]),
    'jcm_conv2d_fwd': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'jcm_conv2d_fwd_variant': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'jcm_bn_stats_blocks': (_I, [_L, _I]),
    'jcm_bn_stats': (_I, [_P, _I, _L, _I, _P, _P]),
    'jcm_bn_finalize': (_I, [_P, _L, _I, _P, _P, _P, _P, _F, _F, _I, _I, _P, _P, _P, _P, _P]),
    'jcm_bn_apply_pool': (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    'jcm_upsample_avg3': (_I, [_P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    'jcm_spatial_softmax': (_I, [_P, _I, _I, _I, _P, _P]),
    'jcm_softmax_ce': (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    'jcm_argmax_hw': (_I, [_P, _I, _I, _I, _I, _P, _P]),
    'jcm_spatial_model_workspace': (_L, [_I, _I, _I, _I, _I]),
    'jcm_spatial_model_fwd': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P]),
    'jcm_conv_mrf_fwd': (_I, [_P, _P, _P, _P, _L, _I, _I, _I, _P]),
    'jcm_softmax_ce_bwd': (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _P, _P]),
    'jcm_spatial_softmax_bwd': (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    'jcm_bn_relu_bwd_blocks': (_I, [_L, _I]),
    'jcm_bn_relu_bwd': (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _F, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    'jcm_colsum': (_I, [_P, _L, _I, _P, _P, _P]),
    'jcm_upsample_avg3_bwd': (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    'jcm_pad_planes': (_I, [_P, _L, _I, _I, _P, _P, _P]),
    'jcm_conv2d_wgrad_workspace': (_L, [_I, _I, _I, _I, _I, _I, _I]),
    'jcm_conv2d_wgrad': (_I, [_P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'jcm_unpack_s2d_grad': (_I, [_P, _I, _P, _P]),
    'jcm_spatial_model_bwd_workspace': (_L, [_I, _I, _I, _I, _I]),
    'jcm_spatial_model_tc_workspace': (_L, [_I, _I, _I, _I, _I]),
    'jcm_spatial_model_tc_fwd': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P]),
    'jcm_spatial_model_tc_bwd_workspace': (_L, [_I, _I, _I, _I, _I]),
    'jcm_spatial_model_tc_bwd': (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _L, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'jcm_spatial_model_bwd': (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _L, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'jcm_pack_weights_taps': (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    'jcm_tap_gather': (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'jcm_tap_scatter_planes': (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    'jcm_unpack_tap_grad': (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    'jcm_augment_color': (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    'jcm_augment_flip_channels': (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    'jcm_augment_rotate': (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    'jcm_augment_crop_resize': (_I, [_P, _P, _I, _I, _I, _I, _F, _P, _P]),
    'jcm_augment_hm_renorm': (_I, [_P, _I, _I, _I, _I, _F, _F, _P, _P]),
    'jcm_optim_blocks': (_I, [_L]),
    'jcm_grad_prepare': (_I, [_P, _P, _L, _L, _F, _F, _P, _P, _P]),
    'jcm_clip_adam': (_I, [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _F, _I, _P]),
    'jcm_sumsq': (_I, [_P, _L, _F, _I, _P, _P, _P]),
    'jcm_clip_scale': (_I, [_P, _L, _P, _F, _P, _P]),
    'jcm_tower_mean': (_I, [_P, _I, _L, _P, _P]),
    'jcm_subsample2': (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    'jcm_bias_relu': (_I, [_P, _P, _L, _I, _I, _P, _P]),
    'jcm_debug_set_wgrad_variant': (_I, [_I]),
    'jcm_debug_tile_plan': (_I, [_I, _I, _I, _I, _I, _I, _P, _I]),
    'jcm_fma_peak': (_I, [_P, _I, _I, _I, _P, _P]),
    'jcm_debug_conv2d_naive': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
}

_lib = None


class JcmError(RuntimeError):
    pass


def lib():
    """Loads libjcm.so once and attaches the signatures."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise JcmError('libjcm.so not found at %s - build it with `python joint-cnn-mrf_b200/build.py` '
                           '(there is no CPU / PyTorch fallback for the jcm kernels)' % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if not hasattr(l, name):
                continue  # later-round symbols may be absent from an older build; test_abi checks the header
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what):
    """Turns a jcm return code into a Python exception (ValueError for argument errors, RuntimeError otherwise)."""
    if rc == 0:
        return
    msg = lib().jcm_last_error().decode('utf-8', 'replace')
    if rc < 0:
        raise ValueError('%s failed (%d): %s' % (what, rc, msg))
    raise JcmError('%s failed (cudaError %d): %s' % (what, rc, msg))
