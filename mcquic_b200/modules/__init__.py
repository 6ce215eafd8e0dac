from .compressor import BaseCompressor, Compressor, Neon  # noqa: F401
from .quantizer import ResidualBackwardQuantizer, UMGMQuantizer  # noqa: F401
