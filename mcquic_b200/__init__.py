"""mcquic_b200 -- B200-native (sm_100a) implementation of McQuic's Compressor.encode/decode hot path."""
from .modules.compressor import BaseCompressor, Compressor, Neon  # noqa: F401
from .modules.quantizer import ResidualBackwardQuantizer, UMGMQuantizer  # noqa: F401

__version__ = "0.1.0"
