"""Host unit test of the prime-factor 255-point DFT index maps and butterflies (csrc/dft255.cuh),
emulated lane by lane on the CPU and compared with numpy's FFT - the transform torch.stft/istft
compute at uformerWM/audio_test.py:315-316,598-600."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dft255") / "libhost_dft255.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", os.path.join(ROOT, "tests", "host_dft255.cpp"), "-o", out])
    return ctypes.CDLL(out)


def test_forward_tile_matches_rfft(host_lib):
    rng = np.random.default_rng(0)
    samp = rng.standard_normal(63 * 31 + 255).astype(np.float32)
    out = np.zeros((256, 32), np.float32)
    host_lib.host_stft_tile(samp.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    frames = np.stack([samp[63 * f:63 * f + 255] for f in range(32)]).astype(np.float64)
    ref = np.fft.fft(frames, axis=1)[:, :128]
    got = out[:128].T.astype(np.float64) + 1j * out[128:].T
    assert np.abs(got - ref).max() / np.abs(ref).max() < 2e-6


def test_inverse_tile_matches_irfft(host_lib):
    rng = np.random.default_rng(1)
    xs = rng.standard_normal((256, 32)).astype(np.float32)      # imag of DC deliberately non-zero
    fr = np.zeros((32, 255), np.float32)
    host_lib.host_istft_tile(xs.ctypes.data_as(ctypes.c_void_p), fr.ctypes.data_as(ctypes.c_void_p))
    spec = xs[:128].T.astype(np.float64) + 1j * xs[128:].T
    spec[:, 0] = spec[:, 0].real
    ref = np.fft.irfft(spec, n=255, axis=1)
    assert np.abs(fr - ref).max() / np.abs(ref).max() < 2e-6
