"""NEXT-2 end to end on the GPU: image -> `.mcq` file -> image through the CLI entry point (mcquic/demo.py:38-75 flow),
checked against the direct encode / rANS / decode calls and the oracle."""
import pathlib

import pytest
import torch

from mcquic_b200 import cli
from mcquic_b200.container import File, REFERENCE_VERSION
from mcquic_b200.utils.synthetic import uniform
from oracle import mcquic_oracle as O

pytestmark = pytest.mark.gpu


def _png(path: pathlib.Path, h: int, w: int, seed: int):
    from torchvision.io import write_png
    img = ((uniform((3, h, w), "cli.image", seed) * 0.5 + 0.5) * 255).clamp(0, 255).to(torch.uint8)
    write_png(img, str(path))
    return img


@pytest.mark.parametrize("h,w,crop", [(256, 256, False), (200, 328, False), (300, 260, True)])
def test_cli_round_trip(tmp_path, h, w, crop):
    src, mcq, out = tmp_path / "in.png", tmp_path / "in.mcq", tmp_path / "restored.png"
    img = _png(src, h, w, 3)
    argv = ["-q", "-qp", "1", "--synthetic"] + (["--crop"] if crop else [])
    assert cli.main(argv + [str(src), str(mcq)]) == 0
    f = File.deserialize(mcq.read_bytes())
    eh, ew = (h // 128 * 128, w // 128 * 128) if crop else (h, w)
    hd = f.FileHeader
    assert (hd.version, hd.qp) == (REFERENCE_VERSION, "qp_1_msssim")
    assert (hd.imageSize.height, hd.imageSize.width, hd.imageSize.channel) == (eh, ew, 3)
    ph, pw = -(-eh // 128) * 128, -(-ew // 128) * 128                               # AlignedPadding to multiples of 128
    assert hd.codeSize.m == [1, 1, 1] and hd.codeSize.k == [8192, 2048, 512]
    assert hd.codeSize.heights == [ph // 16, ph // 32, ph // 64] and hd.codeSize.widths == [pw // 16, pw // 32, pw // 64]
    assert len(f.Content) == 3 and f.BPP == sum(len(c) for c in f.Content) * 8 / (eh * ew)

    # the same through the direct calls: encode -> codes must be the ones inside the file, decode -> the same pixels
    import logging
    model = cli.load_model(1, None, torch.device("cuda"), False, logging.getLogger("t"), synthetic=True)
    x = img.float() / 255.0
    if crop:
        x = cli.aligned_crop(x)
    x = ((x - 0.5) * 2)[None].cuda()
    codes = model.encode(x)
    file_codes = model._quantizer._entropyCoder.decompress([f.Content], [hd.codeSize])
    assert all(torch.equal(a.cpu(), b.cpu()) for a, b in zip(codes, file_codes))
    sd_cpu = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref_codes = O.encode(sd_cpu, x.cpu())
    assert all(torch.equal(a.cpu(), b) for a, b in zip(codes, ref_codes))          # ... and the oracle's

    assert cli.main(["-q", "--synthetic", str(mcq), str(out)]) == 0
    from torchvision.io import read_image
    restored = read_image(str(out))
    assert restored.shape == (3, eh, ew) and restored.dtype == torch.uint8
    full = model.decode(codes)[0]
    top, left = (full.shape[-2] - eh) // 2, (full.shape[-1] - ew) // 2             # compressor.py:94-112 centre crop
    expect = cli.de_transform(full[:, top:top + eh, left:left + ew]).cpu()
    assert torch.equal(restored, expect)
    oracle_pixels = cli.de_transform(O.decode(sd_cpu, ref_codes)[0][:, top:top + eh, left:left + ew])
    assert int((restored.int() - oracle_pixels.int()).abs().max()) <= 1            # 1e-3 in [-1,1] is < 1 grey level


def test_cli_refuses_what_it_cannot_do(tmp_path):
    src = tmp_path / "in.png"
    _png(src, 128, 128, 1)
    with pytest.raises(RuntimeError, match="no network"):
        cli.main(["-q", str(src)])                                                 # pretrained weights need a download
    with pytest.raises(RuntimeError, match="no CPU path"):
        cli.main(["-q", "--disable-gpu", "--synthetic", str(src)])
    bad = tmp_path / "x.txt"
    bad.write_text("hello")
    with pytest.raises(ValueError, match="Invalid input file"):
        cli.main(["-q", "--synthetic", str(bad)])
    assert cli.main(["-q", "--synthetic", str(tmp_path / "missing.png")]) == 2


def test_speed_protocol_runs_and_round_trips(capsys):
    """`--speed` = the reference's Validator.speed (validator.py:60-97): same tensor shape, same call sequence, same
    formula; here with 2 repetitions.  The binaries it produced decode back to the codes it returned."""
    import logging
    model = cli.load_model(2, None, torch.device("cuda"), False, logging.getLogger("test"), synthetic=True)
    (enc, dec), summary = cli.speed(model, torch.device("cuda"), reps=2)
    assert enc > 0 and dec > 0 and summary.startswith("Coding throughput: encoder: ")
    x = torch.rand(2, 3, 768, 512, device="cuda")
    codes, binaries, headers = model.compress(x)
    assert [tuple(c.shape) for c in codes] == [(2, 2, 48, 32), (2, 2, 24, 16), (2, 2, 12, 8)]
    back = model._quantizer._entropyCoder.decompress(binaries, [h.CodeSize for h in headers])
    assert all(torch.equal(a.cpu(), b.cpu()) for a, b in zip(back, codes))
    assert cli.main(["--speed", "-q", "-qp", "1", "--synthetic"]) == 0
    assert "Coding throughput" in capsys.readouterr().out
