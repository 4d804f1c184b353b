"""Import alias: the product code lives in ``image-in-speech-watermarking_b200/`` (a directory
name Python cannot import directly); this package forwards to it."""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_real = _os.path.join(_os.path.dirname(_here), "image-in-speech-watermarking_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
