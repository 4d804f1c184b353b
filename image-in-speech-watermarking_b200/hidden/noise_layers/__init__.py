from .identity import Identity          # noqa: F401
from .crop import Crop                  # noqa: F401
from .cropout import Cropout            # noqa: F401
from .dropout import Dropout            # noqa: F401
from .resize import Resize              # noqa: F401
from .quantization import Quantization  # noqa: F401
from .jpeg_compression import JpegCompression  # noqa: F401
from .noiser import Noiser              # noqa: F401
