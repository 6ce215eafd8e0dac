"""NEXT-2: the `.mcq` container and the CLI helpers (mcquic/utils/specification.py:136-160, mcquic/demo.py,
mcquic/data/transforms.py:60-80, mcquic/utils/vision.py:135-146).

Container bytes are pinned to the reference's OWN `File.serialize` / `File.deserialize` (specification.py:147-156), executed
unmodified through the functional marshmallow stand-in of oracle/ref_import.py where the reference tree is present, and to
its committed outputs (tests/golden/container_reference.json, oracle/gen_golden.py --container) elsewhere; plus a
hand-decoded known answer.  The image helpers are compared with the reference's own classes."""
import warnings

import msgpack
import pytest
import torch

from mcquic_b200 import cli
from mcquic_b200.container import File, REFERENCE_VERSION, readable_size, version_check
from mcquic_b200.entropy import CodeSize, FileHeader, ImageSize
from oracle import ref_import

KAT_HEX = ("82aa66696c6548656164657284a27170ab71705f315f6d737373696da776657273696f6ea6302e312e3430a8636f646553697a6584a1"
           "6d93010101a76865696768747393100804a677696474687393100804a16b93cd2000cd0800cd0200a9696d61676553697a6583a668"
           "6569676874cd0100a57769647468cd0100a76368616e6e656c03a8636f6e74656e747393c4020102c40103c403040506")


def _file():
    return File(FileHeader("0.1.40", "qp_1_msssim", CodeSize([1, 1, 1], [16, 8, 4], [16, 8, 4], [8192, 2048, 512]),
                           ImageSize(256, 256, 3)), [b"\x01\x02", b"\x03", b"\x04\x05\x06"])


def test_known_answer_bytes_and_field_order():
    f = _file()
    data = f.serialize()
    assert data.hex() == KAT_HEX
    d = msgpack.unpackb(data, raw=False)
    assert list(d) == ["fileHeader", "contents"]                                   # FileSchema declaration order
    assert list(d["fileHeader"]) == ["qp", "version", "codeSize", "imageSize"]     # FileHeaderSchema
    assert list(d["fileHeader"]["codeSize"]) == ["m", "heights", "widths", "k"]    # CodeSizeSchema
    assert list(d["fileHeader"]["imageSize"]) == ["height", "width", "channel"]    # ImageSizeSchema
    assert all(isinstance(c, bytes) for c in d["contents"])                        # use_bin_type=True -> bin, not str


def _cases():
    import hashlib
    import json
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    from oracle.gen_golden import CONTAINER_CASES
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "container_reference.json")) as fp:
        golden = json.load(fp)
    return CONTAINER_CASES, golden, hashlib


def test_bytes_equal_the_reference_outputs_committed_as_golden():
    cases, golden, hashlib = _cases()
    assert len(cases) == len(golden)
    for (version, qp, m, hs, ws, k, (ih, iw, ic), contents), g in zip(cases, golden):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            f = File(FileHeader(version, qp, CodeSize(m, hs, ws, k), ImageSize(ih, iw, ic)), list(contents))
            data = f.serialize()
            assert len(data) == g["size"] and hashlib.sha256(data).hexdigest() == g["sha256"]
            assert data[:160].hex() == g["head_hex"]
            assert f.BPP == g["bpp"] and str(f) == g["str"]
            assert File.deserialize(data) == f


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_bytes_equal_reference_file_serialize_and_each_reads_the_other():
    """the reference's File / FileHeader / schemas, unmodified, next to ours: same bytes out, and each side deserialises
    what the other wrote; malformed input is rejected by both"""
    ref_import.load()
    from marshmallow import ValidationError
    from mcquic.utils import specification as R
    cases, _, _ = _cases()
    for version, qp, m, hs, ws, k, (ih, iw, ic), contents in cases:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mine = File(FileHeader(version, qp, CodeSize(m, hs, ws, k), ImageSize(ih, iw, ic)), list(contents))
            ref = R.File(R.FileHeader(version, qp, R.CodeSize(m, hs, ws, k), R.ImageSize(ih, iw, ic)), list(contents))
            a, b = mine.serialize(), ref.serialize()
            assert a == b
            back = R.File.deserialize(a)                      # the reference opens our file
            assert back.fileHeader.qp == qp and back.fileHeader.codeSize.heights == hs and list(back.contents) == contents
            assert File.deserialize(b) == mine                # we open the reference's
            assert mine.BPP == ref.BPP and mine.size() == ref.size() and str(mine) == str(ref)
    d = _file().to_dict()
    d["extra"] = 1
    bad = msgpack.packb(d, use_bin_type=True)
    with pytest.raises(ValidationError):
        R.File.deserialize(bad)
    with pytest.raises(ValueError):
        File.deserialize(bad)
    d = _file().to_dict()
    d["contents"] = [b""]
    bad = msgpack.packb(d, use_bin_type=True)
    with pytest.raises(ValidationError):
        R.File.deserialize(bad)
    with pytest.raises(ValueError):
        File.deserialize(bad)


def test_round_trip_bpp_and_size():
    f = _file()
    g = File.deserialize(f.serialize())
    assert g == f and hash(g) == hash(f)
    assert g.BPP == 6 * 8 / (256 * 256)                                            # specification.py:158-160
    assert g.size() == 6 and g.size(True) == "6 B"
    assert readable_size(424 + 96 + 24) == "544 B" and readable_size(2048) == "2.00 KiB"
    assert "qp_1_msssim" in str(g) and "[16x16, 8192]x1" in str(g)


def test_malformed_files_are_rejected():
    f = _file()
    with pytest.raises(ValueError):
        File.deserialize(b"\x00\x01\x02")
    d = f.to_dict()
    d["extra"] = 1
    with pytest.raises(ValueError):
        File.deserialize(msgpack.packb(d, use_bin_type=True))                      # marshmallow: unknown fields raise
    d = f.to_dict()
    d["contents"] = [b""]
    with pytest.raises(ValueError):
        File.deserialize(msgpack.packb(d, use_bin_type=True))                      # BytesField rejects empty streams
    with pytest.raises(ValueError):
        File(f.fileHeader, [b"ok", b""]).serialize()


def test_header_numbers_from_a_file_are_bounded():
    """everything downstream sizes buffers and crops with these fields: zero / negative / huge sizes, list lengths that do
    not match the number of streams and non-integers are refused at the door (the model-specific checks follow in
    CodeFrequency.decompress)"""
    def broken(edit):
        d = _file().to_dict()
        edit(d)
        with pytest.raises(ValueError, match="not a valid .mcq file"):
            File.deserialize(msgpack.packb(d, use_bin_type=True))

    broken(lambda d: d["fileHeader"]["imageSize"].update(height=0))
    broken(lambda d: d["fileHeader"]["imageSize"].update(width=-256))
    broken(lambda d: d["fileHeader"]["imageSize"].update(channel=0))
    broken(lambda d: d["fileHeader"]["imageSize"].update(height=1 << 20))
    broken(lambda d: d["fileHeader"]["imageSize"].update(height="tall"))
    broken(lambda d: d["fileHeader"]["codeSize"].update(heights=[1 << 30] + list(d["fileHeader"]["codeSize"]["heights"])[1:]))
    broken(lambda d: d["fileHeader"]["codeSize"].update(k=list(d["fileHeader"]["codeSize"]["k"]) + [512]))
    broken(lambda d: d["fileHeader"]["codeSize"].update(m=[0] * len(d["fileHeader"]["codeSize"]["m"])))
    broken(lambda d: d.update(contents=list(d["contents"]) + [b"x"]))
    broken(lambda d: d["fileHeader"].update(codeSize={"m": [], "heights": [], "widths": [], "k": []}) or d.update(contents=[]))


def test_a_checkpoint_path_from_a_file_header_is_never_unpickled(tmp_path):
    """demo.py:77-93 lets the header's qp field name a local checkpoint; such a path is the file's claim, not the user's:
    it goes through the restricted unpickler only (a plain {model, config, version} checkpoint loads, a pickle carrying
    code does not run), while `--local` keeps upstream's trusted-source behaviour"""
    import pickle

    marker = tmp_path / "executed"

    class Evil:
        def __reduce__(self):
            return (open, (str(marker), "w"))

    bad = tmp_path / "bad.mcquic"
    with open(bad, "wb") as fp:
        pickle.dump({"model": Evil(), "config": {}, "version": "0.1.0"}, fp)
    with pytest.raises(RuntimeError, match="named by the file header"):
        cli._load_checkpoint(bad, trusted=False)
    assert not marker.exists()
    good = tmp_path / "good.mcquic"
    torch.save({"model": {"w": torch.ones(2)}, "config": {"model": {"params": {"channel": 128, "m": 1, "k": [8]}}},
                "version": "0.1.0"}, good)
    ck = cli._load_checkpoint(good, trusted=False)
    assert torch.equal(ck["model"]["w"], torch.ones(2)) and ck["version"] == "0.1.0"
    assert cli._load_checkpoint(good, trusted=True)["config"]["model"]["params"]["k"] == [8]


def test_version_check_follows_the_reference():
    assert REFERENCE_VERSION == "0.1.40"
    assert version_check("0.1.40") and version_check("0.1.0")
    with pytest.raises(ValueError, match="too new"):
        version_check("0.1.41")
    with pytest.raises(ValueError, match="Major"):
        version_check("0.0.9", "1.0.0")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert version_check("0.0.9")
        assert any("Minor version mismatch" in str(x.message) for x in w)
    d = _file().to_dict()
    d["fileHeader"]["version"] = "9.9.9"
    with pytest.raises(ValueError):
        File.deserialize(msgpack.packb(d, use_bin_type=True))


def test_cli_helpers():
    assert cli.parse_qp("qp_2_msssim") == (2, False) and cli.parse_qp("qp_13_mse") == (13, True)
    assert cli.parse_qp("/some/ckpt.mcquic") is None and cli.parse_qp("qp_x") is None
    assert cli.model_params_of({"model": {"key": "Compressor", "params": {"channel": 128, "m": 2, "k": [8192, 2048, 512]}},
                                "train": {}}) == {"channel": 128, "m": 2, "k": [8192, 2048, 512]}
    with pytest.raises(RuntimeError):
        cli.model_params_of({"train": {}})
    args = cli.build_parser().parse_args(["-qp", "3", "--mse", "--crop", "in.png", "out.mcq"])
    assert (args.qp, args.mse, args.crop, args.local, str(args.input), str(args.output)) == (3, True, True, None, "in.png", "out.mcq")
    with pytest.raises(SystemExit):
        cli.build_parser().parse_args(["-qp", "14", "in.png"])                     # click.IntRange(0, 13)
    x = torch.arange(3 * 300 * 260, dtype=torch.float32).reshape(3, 300, 260)
    y = cli.aligned_crop(x)
    assert y.shape == (3, 256, 256) and torch.equal(y, x[:, 22:278, 2:258])
    assert cli.aligned_crop(x[:, :256, :128]).shape == (3, 256, 128)                # already aligned: untouched
    v = torch.tensor([-1.5, -1.0, -0.5, 0.0, 0.999, 1.0, 2.0])
    assert cli.de_transform(v).tolist() == [0, 0, 63, 127, 255, 255, 255]


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_image_helpers_match_the_reference_classes():
    ref_import.load()
    from mcquic.data.transforms import AlignedCrop
    from mcquic.utils.vision import DeTransform
    torch.manual_seed(0)
    for h, w in ((300, 260), (256, 256), (129, 511), (128, 640)):
        x = torch.rand(3, h, w)
        assert torch.equal(cli.aligned_crop(x), AlignedCrop()(x))
    x = torch.rand(2, 3, 64, 64) * 2.4 - 1.2
    assert torch.equal(cli.de_transform(x), DeTransform()(x))
