from .compressor import BaseCompressor, Compressor  # noqa: F401
from .quantizer import UMGMQuantizer  # noqa: F401
