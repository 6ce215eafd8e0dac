"""`python -m mcquic_b200 [options] INPUT [OUTPUT]` -- the reference's default command (mcquic/cli.py:41-61,
mcquic/demo.py:38-163) on the B200 path: an image (.png/.jpg) is compressed to a `.mcq` file, a `.mcq` file is restored
to a `.png`.  Same options, same file format, same messages where they carry information; encode/decode run through
libmcquic_b200.so, the rANS coder on the host.

Differences that follow from the environment, not from the format: there is no network, so `-qp N` cannot download the
pretrained checkpoints (`--local model.mcquic` loads one; `--synthetic` builds seeded random weights for smoke tests),
and `--disable-gpu` is refused (this package has no CPU path)."""
import argparse
import logging
import pathlib
import sys
import warnings
from typing import Optional, Tuple

import torch

from . import Compressor
from .container import File, REFERENCE_VERSION, readable_size, version_check
from .modules.compressor import ALIGN_BASE


# ---- mcquic/data/transforms.py:60-80 (AlignedCrop) and mcquic/utils/vision.py:135-146 (DeTransform)
def aligned_crop(x: torch.Tensor, base: int = ALIGN_BASE) -> torch.Tensor:
    h, w = x.shape[-2], x.shape[-1]
    w_crop, h_crop = w - w // base * base, h - h // base * base
    left, top = w_crop // 2, h_crop // 2
    return x[..., top:h - (h_crop - top), left:w - (w_crop - left)]


def de_transform(x: torch.Tensor) -> torch.Tensor:
    """[-1, 1] float -> uint8 exactly as DeTransform: ((x + 1) / 2 * (256 - 1e-3)).clamp(0, 255).byte()."""
    x = (x - (-1.0)) / (1.0 - (-1.0))
    return (x * (255 + 1.0 - 1e-3)).clamp(0.0, 255.0).byte()


def parse_qp(qp: str) -> Optional[Tuple[int, bool]]:
    """demo.py:95-103: 'qp_2_msssim' -> (2, False)."""
    try:
        if not qp.startswith("qp_"):
            return None
        parsed = qp.split("_")
        return int(parsed[1]), parsed[2] == "mse"
    except Exception:
        return None


def model_params_of(config: dict) -> dict:
    """`Config.deserialize(ckpt["config"]).Model.Params` without marshmallow (mcquic/config.py:52-56): the exported
    config is a plain dict {"model": {"key": ..., "params": {channel, m, k}}, "train": {...}}."""
    try:
        params = dict(config["model"]["params"])
    except (KeyError, TypeError) as e:
        raise RuntimeError("checkpoint `config` has no model.params") from e
    return {k: params[k] for k in ("channel", "m", "k") if k in params} | (
        {"permutationRate": params["permutationRate"]} if "permutationRate" in params else {})


def _load_checkpoint(path: pathlib.Path, trusted: bool) -> dict:
    """A checkpoint is {"model": state_dict, "config": plain dict, "version": str}: tensors and plain containers, which
    the restricted unpickler (`weights_only=True`) reads.  Only a path the USER named on the command line (`--local`,
    upstream's "trusted source" warning) may fall back to the full unpickler; a path that comes out of a `.mcq` header
    (detect_model_from_file) never does -- a crafted file must not be able to make the CLI execute a pickle."""
    import pickle
    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except pickle.UnpicklingError as e:
        if not trusted:
            raise RuntimeError(f"checkpoint {path} named by the file header holds objects other than tensors and plain "
                               "containers; pass it with `--local` if you trust it") from e
        return torch.load(path, map_location="cpu", weights_only=False)


def load_model(qp: int, local: Optional[pathlib.Path], device, mse: bool, logger: logging.Logger,
               synthetic: bool = False, trusted: bool = True) -> Compressor:
    """demo.py:137-163.  Checkpoint = {"model": state_dict, "config": dict, "version": "0.1.x"}."""
    if local is not None:
        warnings.warn(f"By passing `--local`, `-qp` arg will be ignored. Checkpoint from {local} will be loaded. "
                      "Please ensure you obtain this local model from a trusted source.")
        ckpt = _load_checkpoint(local, trusted)
        if not isinstance(ckpt, dict) or "model" not in ckpt or "config" not in ckpt:
            raise RuntimeError(f"{local} is not a checkpoint {{model, config, version}}")
        logger.info("Use local model.")
        if "version" not in ckpt:
            raise RuntimeError("You are using a too old ckpt where `version` not in it.")
        version_check(ckpt["version"])
        model = Compressor(**model_params_of(ckpt["config"])).to(device).eval()
        model.QuantizationParameter = str(local)
        model.load_state_dict(ckpt["model"])
        logger.info("Model loaded, params: %s.", model_params_of(ckpt["config"]))
        return model
    key = f"qp_{qp}_{'mse' if mse else 'msssim'}"
    if not synthetic:
        raise RuntimeError(f"pretrained `{key}` has to be downloaded and this environment has no network: pass "
                           "`--local model.mcquic`, or `--synthetic` for seeded random weights")
    from .utils.synthetic import synthetic_state_dict
    params = {1: dict(channel=128, m=1, k=[8192, 2048, 512]), 2: dict(channel=128, m=2, k=[8192, 2048, 512])}.get(
        qp, dict(channel=192, m=max(qp, 1), k=[8192, 2048, 512]))
    model = Compressor(**params).eval()
    model.load_state_dict(synthetic_state_dict(params["channel"], params["m"], params["k"], seed=0))
    model = model.to(device)
    model.QuantizationParameter = key
    logger.info("Use SYNTHETIC (seeded random) weights for `%s`, params: %s.", key, params)
    return model


def compress_image(image: torch.Tensor, model: Compressor, crop: bool) -> File:
    """demo.py:105-121.  image: uint8 [3, h, w] on the model's device."""
    image = image.float() / 255.0                       # convert_image_dtype(uint8 -> float32)
    if crop:
        image = aligned_crop(image)
    image = (image - 0.5) * 2
    _, binaries, headers = model.compress(image[None, ...])
    return File(headers[0], binaries[0])


def decompress_image(source: File, model: Compressor) -> torch.Tensor:
    """demo.py:124-134: uint8 [3, h, w]."""
    restored = model.decompress([source.Content], [source.FileHeader])
    return de_transform(restored[0])


def detect_model_from_file(qp, local, mse, device, logger, source: File, synthetic: bool) -> Compressor:
    """demo.py:77-93: the header's qp field is a checkpoint path or `qp_N_target`."""
    path = pathlib.Path(source.FileHeader.qp)
    if path.exists() and path.is_file() and "mcquic" in path.suffix.lower():
        return load_model(-1, path, device, False, logger, trusted=False)      # the path is the FILE's claim, not the user's
    parsed = parse_qp(source.FileHeader.qp)
    if parsed is not None and local is None:
        return load_model(parsed[0], None, device, parsed[1], logger, synthetic)
    if parsed is None:
        warnings.warn("All qp detections failed. Fallback to use current args or you could try again after checks.")
    return load_model(qp, local, device, mse, logger, synthetic)


def run(debug: bool, quiet: bool, qp: int, local: Optional[pathlib.Path], disable_gpu: bool, mse: bool, crop: bool,
        input: pathlib.Path, output: Optional[pathlib.Path], synthetic: bool = False) -> Optional[File]:
    """demo.py:38-75."""
    from torchvision.io import ImageReadMode, read_image, write_png
    logging.basicConfig(level=logging.CRITICAL if quiet else logging.DEBUG if debug else logging.INFO,
                        format="%(message)s")
    logger = logging.getLogger("mcquic_b200")
    if disable_gpu or not torch.cuda.is_available():
        raise RuntimeError("mcquic_b200 needs a CUDA device (sm_100a); there is no CPU path -- use the reference for "
                           "`--disable-gpu`")
    device = torch.device("cuda")
    suffix = input.suffix.lower()
    with torch.inference_mode():
        if suffix in (".png", ".jpg", ".jpeg"):
            model = load_model(qp, local, device, mse, logger, synthetic)
            image = read_image(str(input), ImageReadMode.RGB).to(device)
            target = compress_image(image, model, crop)
            logger.info(target)
            raw = input.stat().st_size
            logger.info("%s => %s. Compression ratio: %.2f%%", readable_size(raw), target.size(True),
                        (raw - target.size(False)) / raw * 100)
            if output is not None:
                if output.is_dir():
                    output = output.joinpath(input.stem + ".mcq")
                with open(output, "wb") as fp:
                    fp.write(target.serialize())
                logger.info("Saved at %s", output)
            return target
        if suffix == ".mcq":
            with open(input, "rb") as fp:
                source = File.deserialize(fp.read())
            model = detect_model_from_file(qp, local, mse, device, logger, source, synthetic)
            restored = decompress_image(source, model)
            logger.info(source)
            if output is not None:
                if output.is_dir():
                    output = output.joinpath(input.stem + ".png")
                write_png(restored.cpu(), str(output))
            return source
    raise ValueError("Invalid input file.")


PUBLISHED_MPPS = {2: (25.45, 22.03), 12: (11.07, 10.21)}      # reference README.md:304-308, one RTX 3090


def speed(model: Compressor, device, reps: int = 50, batch: int = 10, height: int = 768, width: int = 512):
    """`Validator.speed` of the reference (mcquic/validate/validator.py:60-97), step for step: `torch.rand(10, 3, 768,
    512)` on the device, one warm-up `compress` + `decompress`, then 50 x `model.compress(tensor)` and 50 x
    `model.decompress(binaries, headers)` between CUDA events; Mpps = 50 * 10 * 768 * 512 / 1000 / ms.  Like upstream the
    figure INCLUDES the host rANS coder and the device-to-host reads of the codes, and excludes file I/O and model loading
    (README.md:308).  Returns ((encoder Mpps, decoder Mpps), summary string)."""
    with torch.inference_mode():
        tensor = torch.rand(batch, 3, height, width).to(device)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        codes, binaries, headers = model.compress(tensor)          # warm up
        restored = model.decompress(binaries, headers)
        start.record()
        for _ in range(reps):
            codes, binaries, headers = model.compress(tensor)
        end.record()
        torch.cuda.synchronize()
        encoder_ms = start.elapsed_time(end)
        start.record()
        for _ in range(reps):
            restored = model.decompress(binaries, headers)
        end.record()
        torch.cuda.synchronize()
        decoder_ms = start.elapsed_time(end)
    del restored, codes
    mp = reps * batch * height * width / 1000
    result = (mp / encoder_ms, mp / decoder_ms)
    return result, f"Coding throughput: encoder: {result[0]:.2f} Mpps, decoder: {result[1]:.2f} Mpps"


def run_speed(qp: int, local: Optional[pathlib.Path], mse: bool, synthetic: bool, quiet: bool = False):
    """`python -m mcquic_b200 --speed [-qp N] [--local ckpt | --synthetic]`"""
    logging.basicConfig(level=logging.CRITICAL if quiet else logging.INFO, format="%(message)s")
    logger = logging.getLogger("mcquic_b200")
    if not torch.cuda.is_available():
        raise RuntimeError("mcquic_b200 needs a CUDA device (sm_100a); there is no CPU path")
    device = torch.device("cuda")
    model = load_model(qp, local, device, mse, logger, synthetic)
    result, summary = speed(model, device)
    print(summary)
    if local is None and qp in PUBLISHED_MPPS:
        e, d = PUBLISHED_MPPS[qp]
        print(f"Published for qp={qp} (reference README, 1x RTX 3090, trained weights): encoder {e:.2f} Mpps, decoder "
              f"{d:.2f} Mpps -> x{result[0] / e:.1f} / x{result[1] / d:.1f}")
    return result


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="mcquic_b200", description="Compress/restore a file (B200 path of `mcquic`).")
    ap.add_argument("-v", "--version", action="version", version=f"mcquic_b200 (container of mcquic {REFERENCE_VERSION})")
    ap.add_argument("-D", "--debug", action="store_true", help="Set logging level to DEBUG to print verbose messages.")
    ap.add_argument("-q", "--quiet", action="store_true", help="Silence all messages, this option has higher priority to `-D/--debug`.")
    ap.add_argument("-qp", type=int, default=2, choices=range(0, 14), metavar="[0-13]",
                    help="Quantization parameter. Higher means better image quality and larger size.")
    ap.add_argument("--local", type=pathlib.Path, help="Use a local model path instead of download by `qp`.")
    ap.add_argument("--disable-gpu", action="store_true", help="(refused: this package has no CPU path)")
    ap.add_argument("--mse", action="store_true", help="Use model optimized for PSNR other than MsSSIM.")
    ap.add_argument("--crop", action="store_true", help="Crop the image to align feature patches.")
    ap.add_argument("--synthetic", action="store_true", help="Seeded random weights instead of a pretrained checkpoint (no network here).")
    ap.add_argument("--speed", action="store_true",
                    help="Run the reference's throughput protocol (Validator.speed: 10x3x768x512, 50x compress then 50x "
                         "decompress incl. rANS) instead of coding a file.")
    ap.add_argument("input", type=pathlib.Path, nargs="?", help="Image to compress, or `.mcq` file to restore.")
    ap.add_argument("output", type=pathlib.Path, nargs="?", help="Output file path or dir; omitted: only print the file information.")
    return ap


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    if args.speed:
        run_speed(args.qp, args.local, args.mse, args.synthetic, args.quiet)
        return 0
    if args.input is None:
        print("Error: Missing argument 'INPUT'.", file=sys.stderr)
        return 2
    if not args.input.is_file():
        print(f"Error: Invalid value for 'INPUT': File '{args.input}' does not exist.", file=sys.stderr)
        return 2
    if args.local is not None and not args.local.is_file():
        print(f"Error: Invalid value for '--local': File '{args.local}' does not exist.", file=sys.stderr)
        return 2
    run(args.debug, args.quiet, args.qp, args.local, args.disable_gpu, args.mse, args.crop, args.input.resolve(),
        None if args.output is None else args.output.resolve(), args.synthetic)
    return 0
