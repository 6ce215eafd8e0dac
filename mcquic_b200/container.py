"""The reference's `.mcq` container (SURVEY.md section 8f NEXT-2): `File{fileHeader, contents}` serialised exactly as
mcquic/utils/specification.py:136-156 does -- `FileSchema().dump(file)` (marshmallow: a plain dict in field
declaration order) -> `msgpack.packb(..., use_bin_type=True)` -- so a file written here opens in the reference CLI and
vice versa.  marshmallow / vlutils are not needed: the schema is six fixed fields.

    fileHeader: {qp: str, version: str, codeSize: {m, heights, widths, k: [int]}, imageSize: {height, width, channel}}
    contents:   [bytes]   one rANS stream per level
"""
import re
import warnings
from dataclasses import dataclass
from typing import List, Union

import msgpack

from .entropy import CodeSize, FileHeader, ImageSize

# the reference release this container / checkpoint layout is compatible with (mcquic/__init__.py:1)
REFERENCE_VERSION = "0.1.40"
MAX_LEVELS = 16          # sanity bounds of header fields read from a file
MAX_SIDE = 1 << 16


def _parse_version(v: str):
    m = re.fullmatch(r"(\d+)\.(\d+)(?:\.(\d+))?", v.strip())
    if not m:
        raise ValueError(f"invalid version number '{v}'")          # distutils.StrictVersion's message
    return int(m.group(1)), int(m.group(2)), int(m.group(3) or 0)


def version_check(version: str, built_in: str = REFERENCE_VERSION) -> bool:
    """mcquic/utils/__init__.py:32-48: newer file -> ValueError, major mismatch -> ValueError, minor mismatch -> warning."""
    given, mine = _parse_version(version), _parse_version(built_in)
    if mine < given:
        raise ValueError(f"Version too new. Given {version}, but I'm {built_in} now.")
    if given[0] != mine[0]:
        raise ValueError(f"Major version mismatch. Given {version}, but I'm {built_in} now.")
    if given[1] != mine[1]:
        warnings.warn(f"Minor version mismatch. Given {version}, but I'm {built_in} now.")
    return True


def readable_size(size: int) -> str:
    """vlutils.logger.readableSize as the reference prints sizes (binary units, two decimals)."""
    value = float(size)
    for unit in ("B", "KiB", "MiB", "GiB", "TiB"):
        if value < 1024.0 or unit == "TiB":
            return f"{int(value)} {unit}" if unit == "B" else f"{value:.2f} {unit}"
        value /= 1024.0
    return f"{size} B"


@dataclass
class File:
    fileHeader: FileHeader
    contents: List[bytes]

    @property
    def FileHeader(self) -> FileHeader:
        return self.fileHeader

    @property
    def Content(self) -> List[bytes]:
        return self.contents

    # ---- specification.py:147-156
    def to_dict(self) -> dict:
        h = self.fileHeader
        for c in self.contents:
            if not isinstance(c, bytes) or c == b"":
                raise ValueError("Invalid value")                     # BytesField._validate (specification.py:15-20)
        return {
            "fileHeader": {
                "qp": str(h.qp),
                "version": str(h.version),
                "codeSize": {"m": [int(v) for v in h.codeSize.m], "heights": [int(v) for v in h.codeSize.heights],
                             "widths": [int(v) for v in h.codeSize.widths], "k": [int(v) for v in h.codeSize.k]},
                "imageSize": {"height": int(h.imageSize.height), "width": int(h.imageSize.width),
                              "channel": int(h.imageSize.channel)},
            },
            "contents": list(self.contents),
        }

    def serialize(self) -> bytes:
        return msgpack.packb(self.to_dict(), use_bin_type=True)

    @staticmethod
    def deserialize(data: bytes) -> "File":
        try:
            d = msgpack.unpackb(data, use_list=False, raw=False)
            fh = d["fileHeader"]
            cs, im = fh["codeSize"], fh["imageSize"]
            extra = (set(d) - {"fileHeader", "contents"}) | (set(fh) - {"qp", "version", "codeSize", "imageSize"})
            if extra:
                raise KeyError(f"unknown field(s) {sorted(extra)}")    # marshmallow: unknown = RAISE
            header = FileHeader(version=str(fh["version"]), qp=str(fh["qp"]),
                                codeSize=CodeSize(m=[int(v) for v in cs["m"]], heights=[int(v) for v in cs["heights"]],
                                                  widths=[int(v) for v in cs["widths"]], k=[int(v) for v in cs["k"]]),
                                imageSize=ImageSize(height=int(im["height"]), width=int(im["width"]),
                                                    channel=int(im["channel"])))
            contents = [bytes(c) for c in d["contents"]]
        except (KeyError, TypeError, ValueError, msgpack.exceptions.ExtraData, msgpack.exceptions.FormatError,
                msgpack.exceptions.StackError) as e:
            raise ValueError(f"not a valid .mcq file: {e}") from e
        version_check(header.version)                                  # FileHeader.__init__ (specification.py:108-113)
        for c in contents:
            if c == b"":
                raise ValueError("not a valid .mcq file: empty stream")
        # plausibility of the numbers everything downstream sizes buffers and crops with (the model-specific checks --
        # m, k, level count against the loaded model -- happen in CodeFrequency.decompress)
        cs, im = header.codeSize, header.imageSize
        levels = len(cs.k)
        if not (1 <= levels <= MAX_LEVELS and len(cs.m) == len(cs.heights) == len(cs.widths) == levels == len(contents)):
            raise ValueError("not a valid .mcq file: codeSize lists and contents must have one entry per level")
        if any(not 1 <= v <= MAX_SIDE for v in (im.height, im.width, *cs.heights, *cs.widths)) or \
                any(not 1 <= v <= 65535 for v in (*cs.m, *cs.k)) or not 1 <= im.channel <= 4:
            raise ValueError("not a valid .mcq file: implausible image / code size")
        return File(header, contents)

    @property
    def BPP(self) -> float:                                            # specification.py:158-160
        return sum(len(x) for x in self.contents) * 8 / self.fileHeader.imageSize.Pixels

    def size(self, human: bool = False) -> Union[int, str]:
        size = sum(len(x) for x in self.contents)
        return readable_size(size) if human else size

    def __str__(self) -> str:
        h = self.fileHeader
        cs = h.codeSize
        seq = ", ".join(f"[{w}x{hh}, {k}]x{m}" for hh, w, k, m in zip(cs.heights, cs.widths, cs.k, cs.m))
        return (f"Header: \n    Version    : {h.version}\n    QP         : {h.qp}\n"
                f"    Image size : [{h.imageSize.width}x{h.imageSize.height}, {h.imageSize.channel}]\n"
                f"    Code size  : \n        {cs.m} code-groups: {seq}\n"
                f"Size  : {self.size(True)}\nBPP   : {self.BPP:.4f}")

    def __hash__(self) -> int:
        return hash(self.serialize())
