"""ctypes binding of ``csrc/libwmk.so`` (the C ABI of ``include/wmk.h``).

Fails loudly: a missing library or a failing call raises; nothing falls back to the CPU."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libwmk.so")

c_f32p = ctypes.c_void_p
_i, _f, _vp, _u64, _sz = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t
_dp = ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes); must list every symbol include/wmk.h declares
SIGNATURES = {
    "wmk_version": (_i, []),
    "wmk_last_error": (ctypes.c_char_p, []),
    "wmk_launch_count": (_u64, []),
    "wmk_profile_enable": (_i, [_i]),
    "wmk_profile_num_families": (_i, []),
    "wmk_profile_family_name": (ctypes.c_char_p, [_i]),
    "wmk_profile_collect": (_i, [_dp, _dp, _dp, ctypes.POINTER(ctypes.c_uint64)]),
    "wmk_stft_num_frames": (_i, [_i]),
    "wmk_stft_clips_f32": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "wmk_istft_clips_f32": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "wmk_stft256_num_frames": (_i, [_i]),
    "wmk_stft256_clips_f32": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "wmk_minmax_f32": (_i, [_vp, _sz, _vp, _vp, _vp]),
    "wmk_pcm_decode_f32": (_i, [_vp, _i, _sz, _i, _vp, _vp]),
    "wmk_resample_poly_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "wmk_attack_awgn_f32": (_i, [_vp, _vp, _i, _i, _f, _vp, _u64, _vp]),
    "wmk_attack_scale_f32": (_i, [_vp, _vp, _i, _i, _f, _vp]),
    "wmk_attack_echo_f32": (_i, [_vp, _vp, _i, _i, _i, _f, _vp]),
    "wmk_attack_lowpass_f32": (_i, [_vp, _vp, _i, _i, _i, _dp, _dp, _dp, _vp]),
    "wmk_attack_jitter_zero_f32": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "wmk_attack_jitter_delete_f32": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _vp]),
    "wmk_attack_requant8_f32": (_i, [_vp, _vp, _i, _i, _vp]),
    "wmk_attack_resample2_f32": (_i, [_vp, _vp, _i, _i, _dp, _i, _vp]),
    "wmk_wave_stats_f64": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "wmk_wm_stats_f64": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "wmk_wm_stats_mapped_f64": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "wmk_stats_finalize_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "wmk_conv3x3_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "wmk_convT2x2_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "wmk_maxpool2x2_f32": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "wmk_conv3x3_c1_nhwc_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "wmk_conv3x3_nhwc_bf16_tc": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "wmk_maxpool2x2_nhwc_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "wmk_conv3x3_nhwc_to1_f32": (_i, [_vp, _vp, _vp, _f, _f, _f, _i, _i, _i, _i, _vp]),
    "wmk_bn_train_fwd_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _f, _vp]),
    "wmk_bn_train_bwd_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "wmk_maxpool2x2_bwd_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "wmk_mask_scale_f32": (_i, [_vp, _vp, _vp, _sz, _f, _vp]),
    "wmk_conv3x3_wgrad_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "wmk_conv3x3_dgrad_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "wmk_bn_pool_train_fwd_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _i, _f, _vp]),
    "wmk_bn_pool_train_bwd_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "wmk_convT2x2_dgrad_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "wmk_convT2x2_wgrad_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "wmk_affine_f32": (_i, [_vp, _vp, _sz, _f, _f, _vp]),
    "wmk_mse_f32": (_i, [_vp, _vp, _vp, _sz, _f, _vp, _vp]),
    "wmk_adam_step_f32": (_i, [_vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _f, _i, _f, _i, _vp, _vp]),
    "wmk_noise_mix_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "wmk_noise_crop_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "wmk_noise_resize_nearest_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "wmk_noise_quantize_f32": (_i, [_vp, _vp, _sz, _vp]),
    "wmk_noise_jpeg_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "wmk_magphase_split_f32": (_i, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "wmk_magphase_merge_f32": (_i, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "wmk_uformer_plan_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "wmk_plan_destroy": (_i, [_vp]),
    "wmk_plan_set_tensor": (_i, [_vp, ctypes.c_char_p, _vp, ctypes.POINTER(ctypes.c_int64), _i]),
    "wmk_plan_finalize": (_i, [_vp]),
    "wmk_plan_set_chunk": (_i, [_vp, _i]),
    "wmk_plan_workspace_bytes": (_sz, [_vp]),
    "wmk_uformer_forward": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wmk_uformer_forward_mapped": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wmk_uformer_extract": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "wmk_uformer_autoencode": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "wmk_plan_enable_taps": (_i, [_vp, _i]),
    "wmk_plan_get_tap": (_i, [_vp, ctypes.c_char_p, _vp, _sz, ctypes.POINTER(_sz)]),
    "wmk_leff_block_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "wmk_window_attention_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "wmk_linear_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "wmk_lewin_block_train_f32": (_i, [_vp, _vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "wmk_transpose_batched_f32": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "wmk_leaky_relu_f32": (_i, [_vp, _vp, _vp, _sz, _f, _vp]),
    "wmk_sigmoid_f32": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "wmk_downsample_train_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "wmk_extract_head_train_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "wmk_maxpool16x8_f32": (_i, [_vp, _vp, _vp, _i, _vp]),
    "wmk_upsample_train_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "wmk_stft_projection_adjoint_f32": (_i, [_vp, _vp, _i, _vp]),
    "wmk_comm_unique_id": (_i, [_vp]),
    "wmk_comm_create": (_i, [_vp, _i, _i, ctypes.POINTER(_vp)]),
    "wmk_comm_destroy": (_i, [_vp]),
    "wmk_stats_allreduce_f64": (_i, [_vp, _i, _vp, _vp]),
    "wmk_grad_allreduce_f32": (_i, [_vp, _sz, _vp, _vp]),
}

PREC_FP32, PREC_BF16, PREC_MIXED, PREC_F16 = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "mixed": PREC_MIXED, "fp16": PREC_F16}
_lib = None


class WmkError(RuntimeError):
    pass


def load():
    """Load libwmk.so (once).  Raises if it has not been built (``__graft_entry__.build()`` /
    ``make -C image-in-speech-watermarking_b200/csrc``)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise WmkError("libwmk.so not built at %s: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        raise WmkError("libwmk call failed (%d): %s" % (status, load().wmk_last_error().decode()))


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise WmkError("expected a CUDA tensor (the hot path has no CPU implementation)")
    if not t.is_contiguous():
        raise WmkError("expected a contiguous tensor")
    return ctypes.c_void_p(t.data_ptr())


def profile_enable(on=True):
    check(load().wmk_profile_enable(int(on)))


def profile_collect():
    """{family: {"ms", "work", "launches"}} accumulated since the last collect (synchronises)."""
    lib = load()
    n = lib.wmk_profile_num_families()
    ms = (ctypes.c_double * n)()
    work = (ctypes.c_double * n)()
    work2 = (ctypes.c_double * n)()
    cnt = (ctypes.c_uint64 * n)()
    check(lib.wmk_profile_collect(ms, work, work2, cnt))
    return {lib.wmk_profile_family_name(i).decode(): {"ms": ms[i], "work": work[i], "work2": work2[i], "launches": int(cnt[i])}
            for i in range(n)}
