"""Import the UNMODIFIED reference model in the build container.

Only used by ``oracle/make_golden.py`` and ``tests/test_oracle_vs_reference.py``
(both skip when ``/root/reference`` is absent, e.g. on the GPU box).

The reference (`uformerWM/model.py:4,12`) imports ``timm`` and ``torchsummary``
which are not installed, and calls ``torch.stft`` / ``torch.istft`` with the
torch-1.x real-view convention (`uformerWM/model.py:2458,2463`).  The shim
provides the three timm helpers the model uses, an empty ``torchsummary`` and
legacy real-view wrappers for stft/istft.  Nothing of the reference is copied.
"""
import os
import sys
import types
import contextlib

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("WMK_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "uformerWM", "model.py"))


class _DropPath(nn.Module):
    """timm.models.layers.DropPath: identity in eval, stochastic depth in train."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        return x * mask / keep


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def _install_stubs():
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = _DropPath
        layers.to_2tuple = _to_2tuple
        layers.trunc_normal_ = nn.init.trunc_normal_
        utils = types.ModuleType("timm.utils")
        timm.models = models
        models.layers = layers
        timm.utils = utils
        sys.modules.update({"timm": timm, "timm.models": models,
                            "timm.models.layers": layers, "timm.utils": utils})
    if "torchsummary" not in sys.modules:
        ts = types.ModuleType("torchsummary")
        ts.summary = lambda *a, **k: None
        sys.modules["torchsummary"] = ts


_ORIG_STFT = torch.stft
_ORIG_ISTFT = torch.istft


def _legacy_stft(x, n_fft, *args, return_complex=None, **kw):
    out = _ORIG_STFT(x, n_fft, *args, return_complex=True, **kw)
    return out if return_complex else torch.view_as_real(out)


def _legacy_istft(x, n_fft, *args, return_complex=False, **kw):
    if not torch.is_complex(x):
        x = torch.view_as_complex(x.contiguous())
    return _ORIG_ISTFT(x, n_fft, *args, return_complex=return_complex, **kw)


@contextlib.contextmanager
def legacy_torch_spectral():
    """torch-1.x semantics for torch.stft / torch.istft while the reference runs."""
    torch.stft, torch.istft = _legacy_stft, _legacy_istft
    try:
        yield
    finally:
        torch.stft, torch.istft = _ORIG_STFT, _ORIG_ISTFT


def import_reference_model():
    """Return the reference's ``uformerWM/model.py`` as a module (unmodified)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    path = os.path.join(REFERENCE_ROOT, "uformerWM")
    if path not in sys.path:
        sys.path.insert(0, path)
    import importlib
    with legacy_torch_spectral():
        mod = importlib.import_module("model")
    return mod


def import_reference_hidden(name):
    """Import ``hidden/<name>`` (e.g. 'noise_layers.quantization') unmodified."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    path = os.path.join(REFERENCE_ROOT, "hidden")
    if path not in sys.path:
        sys.path.insert(0, path)
    import importlib
    return importlib.import_module(name)


def build_reference_uformer_audio(seed=0):
    """`uformerWM/utils/model_utils.py:83-85` constructor arguments."""
    ref = import_reference_model()
    torch.manual_seed(seed)
    m = ref.UformerAudio(img_size=128, embed_dim=32, win_size=8, token_projection='linear',
                         token_mlp='leff', depths=[1, 2, 8, 8, 2, 8, 8, 2, 1], modulator=True,
                         dd_in=2, in_chans=2, audio_scale='0')
    return m.eval()


# --------------------------------------------------------------------------- executing reference
# functions whose modules cannot be imported (import-time argparse / datasets / absent
# third-party packages).  The function's own source is parsed from the reference file where it
# lies and executed unmodified in a namespace we provide; nothing is copied into the repo.
def extract_functions(rel_path, names, namespace):
    """Compile the top-level functions ``names`` of ``<reference>/<rel_path>`` into ``namespace``."""
    import ast
    path = os.path.join(REFERENCE_ROOT, rel_path)
    with open(path, "r") as f:
        tree = ast.parse(f.read(), filename=path)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    found = {n.name for n in wanted}
    if found != set(names):
        raise RuntimeError("functions %s not found in %s" % (sorted(set(names) - found), path))
    mod = ast.Module(body=wanted, type_ignores=[])
    exec(compile(mod, path, "exec"), namespace)
    return namespace


class _CpuTorchProxy:
    """``torch`` as seen by the reference driver on a machine without CUDA: ``torch.device('cuda')``
    resolves to the CPU; every other attribute is the real torch (looked up at call time so the
    legacy stft/istft wrappers are honoured)."""

    def __getattr__(self, name):
        if name == "device":
            return lambda *a, **k: torch.device("cpu")
        if name == "from_numpy":
            # scipy.signal.filtfilt returns a negative-stride view, which torch.from_numpy
            # rejects (`uformerWM/audio_test.py:677` with attack 'low_pass'); copy first.
            import numpy as _np
            return lambda a: torch.from_numpy(_np.ascontiguousarray(a))
        return getattr(torch, name)


def reference_attack_functions():
    """The pure numpy/scipy attacks of `uformerWM/audio_attack.py`, executed unmodified."""
    import math as _math
    import random as _random
    import numpy as _np
    from scipy import signal as _signal
    ns = {"np": _np, "signal": _signal, "random": _random, "math": _math}
    names = ["low_pass_filter", "echo_addition", "amplitude_scaling", "closed_loop", "awgn",
             "jittering_2", "jittering"]
    return extract_functions("uformerWM/audio_attack.py", names, ns)


def reference_metric_functions():
    """`cal_snr`, `signaltonoise`, `SNR_singlech` of `uformerWM/evaluate.py`, executed unmodified."""
    import math as _math
    import numpy as _np
    ns = {"np": _np, "math": _math}
    return extract_functions("uformerWM/evaluate.py", ["cal_snr", "signaltonoise", "SNR_singlech"], ns)


def reference_reconstruct_audio():
    """`reconstruct_audio` of `uformerWM/audio_test.py:528-785`, executed unmodified on the CPU."""
    import numpy as _np
    import torch.nn.functional as _F
    ns = dict(reference_attack_functions())
    ns.update({"torch": _CpuTorchProxy(), "np": _np, "F": _F})
    extract_functions("uformerWM/audio_test.py", ["reconstruct_audio", "signaltonoise"], ns)
    fn = ns["reconstruct_audio"]

    def run(*a, **k):
        with legacy_torch_spectral():
            return fn(*a, **k)
    return run


def extract_method(rel_path, cls, name, namespace):
    """Compile method ``cls.name`` of ``<reference>/<rel_path>`` as a plain function into ``namespace``."""
    import ast
    path = os.path.join(REFERENCE_ROOT, rel_path)
    with open(path, "r") as f:
        tree = ast.parse(f.read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == name:
                    exec(compile(ast.Module(body=[item], type_ignores=[]), path, "exec"), namespace)
                    return namespace[name]
    raise RuntimeError("%s.%s not found in %s" % (cls, name, path))


def reference_prepare_data_train():
    """`SpeechDataTrain.prepare_data` + `normalize_batch` of `uformerWM/audio_test.py:33-55,439-502`, executed
    unmodified on the CPU.  Returns run(waves, audio_scale) -> (data, min, max) where `waves` is a list of
    (1, L) tensors standing in for the LibriSpeech items (`self.data_raw[i][0]`)."""
    import types
    import numpy as _np
    import torch.nn.functional as _F
    ns = {"torch": _CpuTorchProxy(), "np": _np, "F": _F}
    extract_functions("uformerWM/audio_test.py", ["normalize_batch"], ns)
    fn = extract_method("uformerWM/audio_test.py", "SpeechDataTrain", "prepare_data", ns)

    def run(waves, audio_scale):
        me = types.SimpleNamespace(data_raw=[(w, 16000) for w in waves], size=len(waves), data_type="train",
                                   frequency=128, len_clip=128, audio_scale=audio_scale)
        with legacy_torch_spectral():
            return fn(me)
    return run


def reference_hidden_modules():
    """The unmodified `hidden/model/decoder.py`, `hidden/options.py` and `hidden/noise_layers/*`.
    `hidden/` uses top-level module names (`model`, `options`, `noise_layers`) that collide with
    `uformerWM/model.py`, so they are imported with sys.modules temporarily swapped."""
    import importlib
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    hp = os.path.join(REFERENCE_ROOT, "hidden")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "model" or k.startswith("model.")
             or k in ("options", "noise_layers") or k.startswith("noise_layers.")}
    up = os.path.join(REFERENCE_ROOT, "uformerWM")
    removed = [q for q in sys.path if q == up]          # `model.py` there would shadow the `hidden/model/` package
    sys.path[:] = [q for q in sys.path if q != up]
    sys.path.insert(0, hp)
    importlib.invalidate_caches()
    try:
        out = {"decoder": importlib.import_module("model.decoder"), "options": importlib.import_module("options")}
        for n in ("identity", "crop", "cropout", "dropout", "resize", "quantization", "jpeg_compression"):
            out[n] = importlib.import_module("noise_layers." + n)
    finally:
        sys.path.remove(hp)
        for q in removed:
            sys.path.insert(0, q)
        for k in list(sys.modules):
            if k == "model" or k.startswith("model.") or k in ("options", "noise_layers") or k.startswith("noise_layers."):
                sys.modules.pop(k)
        sys.modules.update(saved)
    return out
