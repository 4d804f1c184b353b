"""CPU-side checks of the boundary: the library loads, exports every symbol include/wmk.h declares,
argument errors are reported through the status / last-error convention, and the product modules
refuse to run without CUDA (no silent CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "wmk.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(wmk_[A-Za-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from image_in_speech_watermarking_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libwmk.so does not export %s" % n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert lib.wmk_version() >= 100


def test_argument_errors_use_status_and_last_error():
    from image_in_speech_watermarking_b200 import _lib
    lib = _lib.load()
    assert lib.wmk_stft_num_frames(16000) == 254
    assert lib.wmk_stft_num_frames(8002) == 128
    assert lib.wmk_stft_num_frames(63 * 127) == 127           # T = 1 + (L-1)//63
    st = lib.wmk_stft_clips_f32(None, 1, 16000, None, 2, None)
    assert st == -1 and b"stft" in lib.wmk_last_error()
    st = lib.wmk_plan_set_tensor(None, b"x", None, None, 0)
    assert st == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from image_in_speech_watermarking_b200 import _lib
    from image_in_speech_watermarking_b200.model import UformerAudio
    from image_in_speech_watermarking_b200 import audio_uformer_stft as FE
    m = UformerAudio()
    with pytest.raises(_lib.WmkError):
        m(torch.zeros(1, 2, 128, 128), torch.zeros(1, 1, 32, 32))
    with pytest.raises(_lib.WmkError):
        FE.stft_clips(torch.zeros(1, 16000))
    handle = ctypes.c_void_p()
    assert _lib.load().wmk_uformer_plan_create(0, ctypes.byref(handle)) == -2     # WMK_ERR_CUDA


def test_drop_in_state_dict_matches_reference_schema():
    from image_in_speech_watermarking_b200.model import UformerAudio
    from oracle import uformer as O
    m = UformerAudio()
    sch = O.state_dict_schema()
    sd = m.state_dict()
    assert list(sd.keys()) == list(sch.keys())
    assert all(tuple(sd[k].shape) == tuple(sch[k][0]) for k in sd)
    assert sum(p.numel() for p in m.parameters()) == 68714978
    m.load_state_dict(sd, strict=True)
