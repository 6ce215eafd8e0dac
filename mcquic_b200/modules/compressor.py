"""`Compressor` with the reference's constructor, state_dict layout and encode/decode contract
(mcquic/modules/compressor.py:18-117 BaseCompressor, :120-177 Compressor), executed by the CUDA engine.

    Compressor(channel, m, k).to("cuda").eval().load_state_dict(reference_checkpoint["model"])
    codes = model.encode(x)        # x fp32 [n,3,H,W] in [-1,1]  ->  L x int64 [n, m, h_l, w_l]
    xHat  = model.decode(codes)    # fp32 [n, 3, H_pad, W_pad]
"""
from typing import List, Optional, Tuple

import torch
from torch import nn

from ..engine import Engine
from ..nn import AttentionBlock, ResidualBlock, ResidualBlockShuffle, ResidualBlockWithStride, conv3x3, pixelShuffle3x3
from .quantizer import UMGMQuantizer

ALIGN_BASE = 128  # mcquic/data/transforms.py:82


def aligned_pad_amounts(h: int, w: int, base: int = ALIGN_BASE) -> Tuple[int, int, int, int]:
    """(top, left, padded_h, padded_w) exactly as AlignedPadding.forward splits them (transforms.py:86-99)."""
    wPadding = ((w // base + 1) * base - w) % base
    hPadding = ((h // base + 1) * base - h) % base
    return hPadding // 2, wPadding // 2, h + hPadding, w + wPadding


class BaseCompressor(nn.Module):
    def __init__(self, encoder: nn.Module, quantizer: UMGMQuantizer, decoder: nn.Module):
        super().__init__()
        self._encoder = encoder
        self._decoder = decoder
        self._quantizer = quantizer
        self._qp = "-1"
        self._engine: Optional[Engine] = None
        self.encode_passes = 3  # split-fp16 x3: fp32-grade, code indices match the fp32 reference
        self.decode_passes = 1  # single fp16 pass: TF32-grade, pixels within 1e-3

    @property
    def QuantizationParameter(self) -> str:
        return self._qp

    @QuantizationParameter.setter
    def QuantizationParameter(self, qp: str):
        self._qp = qp

    @property
    def Codebooks(self):
        return self._quantizer.Codebooks

    @property
    def NormalizedFreq(self):
        return self._quantizer.NormalizedFreq

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine()
        return self._engine

    def set_impl(self, impl: str):
        """'tcgen05' (default) or 'simt' (fp32 CUDA-core cross-check kernels)."""
        self._engine = Engine(impl)

    def _check_image(self, x: torch.Tensor):
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"expected an image batch [n, 3, h, w], got {tuple(x.shape)}")
        if not x.is_cuda and not self.engine.emulated:
            raise RuntimeError("mcquic_b200 runs on CUDA tensors only (there is no CPU fallback)")

    @torch.no_grad()
    def encode(self, x: torch.Tensor, hist: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        """compressor.py:79-88.  `hist`: optional flat int32 [sum_l m*k_l] code histogram, accumulated in place."""
        self._check_image(x)
        eng = self.engine
        eng.passes = self.encode_passes
        n, _, h, w = x.shape
        y0 = eng.stem(self._encoder[0], x, aligned_pad_amounts(h, w), eng.needs_of(self._encoder[1]))
        y = eng.run_seq(list(self._encoder)[1:], y0, self._quantizer.first_needs(eng))
        return self._quantizer.encode_act(eng, y, hist)

    @torch.no_grad()
    def decode(self, codes: List[torch.Tensor]) -> torch.Tensor:
        """compressor.py:114-117 (no crop; `decompress` crops upstream)."""
        if len(codes) == 0:
            raise RuntimeError("Length of codes is 0.")
        eng = self.engine
        eng.passes = self.decode_passes
        status = torch.zeros(1, dtype=torch.int32, device=codes[0].device)
        yHat = self._quantizer.decode_act(eng, codes, eng.needs_of(self._decoder[0]), status)
        out = eng.run_seq(list(self._decoder), yHat, set()).f32
        if int(status.item()) != 0:
            raise RuntimeError("code index out of range for its codebook")
        return out

    def forward(self, x: torch.Tensor):
        raise NotImplementedError("mcquic_b200 accelerates inference (encode/decode); training forward is out of scope")


class Compressor(BaseCompressor):
    def __init__(self, channel: int, m: int, k: List[int], permutationRate: float = 0.0):
        if channel % 8 != 0:
            raise ValueError("channel must be a multiple of 8")
        RB, RBS, RBU, AB = ResidualBlock, ResidualBlockWithStride, ResidualBlockShuffle, AttentionBlock
        C = channel
        encoder = nn.Sequential(conv3x3(3, C, 2), RB(C, C), RBS(C, C), AB(C), RB(C, C), RBS(C, C), RB(C, C))
        decoder = nn.Sequential(RB(C, C), RBU(C, C), AB(C), RB(C, C), RBU(C, C), RB(C, C), pixelShuffle3x3(C, 3, 2))
        quantizer = UMGMQuantizer(C, m, k, permutationRate, {
            "latentStageEncoder": lambda: nn.Sequential(RBS(C, C), RB(C, C), AB(C)),
            "quantizationHead": lambda: nn.Sequential(RB(C, C), AB(C), conv3x3(C, C)),
            "latentHead": lambda: nn.Sequential(RB(C, C), AB(C), conv3x3(C, C)),
            "restoreHead": lambda: nn.Sequential(AB(C), RB(C, C), RBU(C, C)),
            "dequantizationHead": lambda: nn.Sequential(AB(C), conv3x3(C, C), RB(C, C)),
            "sideHead": lambda: nn.Sequential(AB(C), conv3x3(C, C), RB(C, C)),
        })
        super().__init__(encoder, quantizer, decoder)
