"""Host unit test of the 16 x 16 Cooley-Tukey 256-point DFT (csrc/dft256.cuh) behind the training-time
STFT - torch.stft(x, n_fft=256, hop_length=128, win_length=256) at uformerWM/audio_test.py:465-469 -
run on the CPU and compared with numpy's FFT."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dft256") / "libhost_dft256.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", os.path.join(ROOT, "tests", "host_dft256.cpp"), "-o", out])
    return ctypes.CDLL(out)


def test_frame_matches_fft(host_lib):
    rng = np.random.default_rng(0)
    for _ in range(4):
        frame = rng.standard_normal(256).astype(np.float32)
        out = np.zeros((2, 128), np.float32)
        host_lib.host_fft256_frame(frame.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
        ref = np.fft.fft(frame.astype(np.float64))[:128]
        got = out[0].astype(np.float64) + 1j * out[1]
        assert np.abs(got - ref).max() / np.abs(ref).max() < 2e-6


def test_impulses(host_lib):
    # delta at n0 -> X[k] = exp(-2 pi i k n0 / 256): exercises every twiddle index
    for n0 in (0, 1, 15, 16, 17, 127, 128, 255):
        frame = np.zeros(256, np.float32)
        frame[n0] = 1.0
        out = np.zeros((2, 128), np.float32)
        host_lib.host_fft256_frame(frame.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
        k = np.arange(128)
        ref = np.exp(-2j * np.pi * k * n0 / 256)
        got = out[0].astype(np.float64) + 1j * out[1]
        assert np.abs(got - ref).max() < 2e-6
