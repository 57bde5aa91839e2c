"""Model-card / checkpoint resolution and model assembly.

Mirrors videoseal/utils/cfg.py of the reference (setup_model_from_model_card :181-259, setup_model :88-154,
maybe_download_checkpoint :262-287) without OmegaConf: cards are read with PyYAML into plain dicts.  The card YAML
schema and the checkpoint format (`torch.load(..., weights_only=True)['model']` = Videoseal.state_dict()) are the
reference's, consumed as-is.
"""
from __future__ import annotations

import math
import os
from pathlib import Path
from urllib.parse import urlparse

import torch
import yaml

DEFAULT_CARD = "videoseal_1.0"
PKG_DIR = Path(__file__).resolve().parent
CARDS_DIR = PKG_DIR / "cards"


def resolve_config_path(cfg_path) -> Path:
    """working directory first, then the package directory (cfg.py:40-48 of the reference)"""
    p = Path(cfg_path)
    if p.is_file():
        return p
    q = PKG_DIR / cfg_path
    if q.is_file():
        return q
    raise FileNotFoundError(f"config file {cfg_path} not found (cwd or {PKG_DIR})")


def is_url(s) -> bool:
    try:
        r = urlparse(str(s))
        return all([r.scheme, r.netloc])
    except ValueError:
        return False


def maybe_download_checkpoint(url: str) -> str:
    """same cache naming as the reference: ckpts/<parent>_<file> (cfg.py:262-287)"""
    parts = urlparse(url).path.split("/")
    basename = f"{parts[-2]}_{parts[-1]}" if len(parts) >= 2 else parts[-1]
    os.makedirs("ckpts", exist_ok=True)
    filename = os.path.abspath(os.path.join("ckpts", basename))
    if os.path.exists(filename):
        return filename
    try:
        import requests
        resp = requests.get(url, timeout=30)
        resp.raise_for_status()
        with open(filename, "wb") as f:
            f.write(resp.content)
    except Exception as e:  # no network in most deployments of this package
        raise FileNotFoundError(
            f"checkpoint {url} is not cached at {filename} and could not be downloaded ({e}); "
            "put the file there or point the card's checkpoint_path at a local .pth") from e
    return filename


def spec_from_card(card: dict) -> dict:
    """Architecture hyper-parameters from a card (build_embedder embedder.py:243-282, build_extractor
    extractor.py:170-213, back-compat renames cfg.py:101-118)."""
    a = card["args"]
    nbits = int(a["nbits"])
    hsm = a.get("hidden_size_multiplier", 2)
    img_size = a["img_size_proc"] if "img_size_proc" in a else a["img_size_extractor"]
    emb_name = str(card["embedder"]["model"])
    ext_name = str(card["extractor"]["model"])
    if not emb_name.startswith("unet"):
        raise NotImplementedError(f"embedder '{emb_name}': only the unet* family is on the B200 hot path")
    if not ext_name.startswith("convnext"):
        raise NotImplementedError(f"extractor '{ext_name}': only the convnext* family is on the B200 hot path")
    u = card["embedder"]["params"]["unet"]
    mp = card["embedder"]["params"].get("msg_processor", {})
    if mp.get("msg_processor_type", "binary+concat") != "binary+concat":
        raise NotImplementedError("only 'binary+concat' message processors are implemented")
    if str(a.get("blending_method", "additive")) != "additive":
        raise NotImplementedError("only additive blending is implemented (all shipped cards use it)")
    ep = card["extractor"]["params"]
    pd = ep.get("pixel_decoder", {})
    if pd.get("pixelwise", False) or list(pd.get("upscale_stages", [1])) != [1] or pd.get("sigmoid_output", False):
        raise NotImplementedError("pixel decoder: only pixelwise=False, upscale_stages=[1], sigmoid_output=False")
    enc = ep["encoder"]
    dims = [int(x) for x in enc["dims"]]
    if ep.get("proportional_dim", False):
        mult = math.sqrt(nbits / 128)
        dims = [int(x * mult) for x in dims]
    if enc.get("temporal_convs", False) or enc.get("temporal_attention", False):
        raise NotImplementedError("temporal ConvNeXt variants are out of scope")
    att = str(a.get("attenuation", "None"))
    jnd = (0, 0)
    if att.lower().startswith("jnd"):
        att_cfg = yaml.safe_load(open(resolve_config_path(a.get("attenuation_config", "configs/attenuation.yaml"))))
        jnd = (int(att_cfg[att]["in_channels"]), int(att_cfg[att]["out_channels"]))
    return {
        "nbits": nbits, "hidden": int(nbits * hsm), "img_size": int(img_size), "yuv": "yuv" in emb_name,
        "scaling_w": float(a["scaling_w"]), "scaling_i": float(a["scaling_i"]),
        "chunk_size": int(a.get("videoseal_chunk_size", a.get("videowam_chunk_size", 8))),
        "step_size": int(a.get("videoseal_step_size", a.get("videowam_step_size", 4))),
        "unet": {
            "in_channels": int(u["in_channels"]), "out_channels": int(u["out_channels"]),
            "z": [int(u["z_channels"]) * int(m) for m in u["z_channels_mults"]], "num_blocks": int(u["num_blocks"]),
            "activation": str(u["activation"]), "normalization": str(u["normalization"]),
            "last_tanh": bool(u.get("last_tanh", True)),
        },
        "convnext": {"depths": [int(x) for x in enc["depths"]], "dims": dims, "stem_stride": int(enc.get("stem_stride", 4))},
        "jnd": jnd, "attenuation": att,
    }


def setup_model(card: dict, ckpt_path):
    from .model import Videoseal
    spec = spec_from_card(card)
    if not os.path.exists(ckpt_path):
        raise FileNotFoundError(f"Checkpoint path does not exist: {ckpt_path}")
    checkpoint = _safe_load(ckpt_path)
    model = Videoseal(spec)
    msg = model.load_state_dict(checkpoint["model"], strict=False)
    print(f"Model loaded successfully from {ckpt_path} with message: {msg}")
    return model


def _safe_load(ckpt_path):
    """torch.load with the safe unpickler only (weights_only=True, like utils/cfg.py:62,148 of the reference).  argparse.Namespace
    -- how training scripts save `args` next to `model` -- is allow-listed for the duration of the load: a plain attribute bag."""
    import argparse
    with torch.serialization.safe_globals([argparse.Namespace]):
        return torch.load(ckpt_path, map_location="cpu", weights_only=True)


def _plain(o):
    """checkpoint['args'] may be a dict, an argparse.Namespace or an OmegaConf container saved by the training script"""
    if hasattr(o, "items"):
        return {str(k): _plain(v) for k, v in o.items()}
    if hasattr(o, "__dict__") and not isinstance(o, (str, bytes)):
        return {str(k): _plain(v) for k, v in vars(o).items()}
    if isinstance(o, (list, tuple)):
        return [_plain(v) for v in o]
    return o


def get_config_from_checkpoint(ckpt_path) -> dict:
    """utils/cfg.py:50-85 of the reference: a raw training checkpoint carries its `args`; the embedder / extractor
    hyper-parameters are the presets named by args.embedder_model / args.extractor_model in the YAML files named by
    args.embedder_config / args.extractor_config (cwd first, then this package).  Returns a card-shaped dict."""
    # the safe unpickler only, like the reference (utils/cfg.py:62,148): a checkpoint is untrusted input.  Training scripts save
    # `args` as an argparse.Namespace, which is allow-listed for the duration of the load (a plain attribute bag, no code);
    # anything else the safe unpickler rejects (e.g. an OmegaConf container) is an error, never a silent full-pickle retry.
    checkpoint = _safe_load(ckpt_path)
    if "args" not in checkpoint:
        raise KeyError(f"{ckpt_path} holds no 'args': not a training checkpoint (use a model card for released weights)")
    args = _plain(checkpoint["args"])
    if not isinstance(args, dict):
        raise Exception("Expected logfile to contain params dictionary.")
    emb_cfg = yaml.safe_load(open(resolve_config_path(args.get("embedder_config", "configs/embedder.yaml"))))
    ext_cfg = yaml.safe_load(open(resolve_config_path(args.get("extractor_config", "configs/extractor.yaml"))))
    emb_model = args.get("embedder_model") or emb_cfg["model"]
    ext_model = args.get("extractor_model") or ext_cfg["model"]
    for name, cfg_, kind in ((emb_model, emb_cfg, "embedder"), (ext_model, ext_cfg, "extractor")):
        if name not in cfg_:
            raise NotImplementedError(f"{kind} preset '{name}' is not on the B200 hot path (known: "
                                      f"{', '.join(k for k in cfg_ if k != 'model')})")
    return {"checkpoint_path": str(ckpt_path), "args": args,
            "embedder": {"model": emb_model, "params": emb_cfg[emb_model]},
            "extractor": {"model": ext_model, "params": ext_cfg[ext_model]}}


def setup_model_from_checkpoint(ckpt_path) -> "Videoseal":
    """utils/cfg.py:156-178 of the reference: a model-card name, or the path of a raw training checkpoint (.pth)."""
    ckpt_path = str(ckpt_path)
    if "baseline" in ckpt_path:
        raise NotImplementedError("baseline watermarkers (videoseal/models/baselines.py) are outside the B200 hot path")
    if not ckpt_path.endswith(".pth") and "/" not in ckpt_path:
        return setup_model_from_model_card(ckpt_path)
    return setup_model(get_config_from_checkpoint(ckpt_path), ckpt_path)


def setup_model_from_model_card(model_card) -> "Videoseal":
    """`videoseal.load()`: card name (looked up in ./videoseal/cards, then in this package's cards/) or a Path to a YAML."""
    if model_card == "videoseal":
        model_card = DEFAULT_CARD
    if isinstance(model_card, str):
        candidates = [Path("videoseal/cards") / f"{model_card}.yaml", CARDS_DIR / f"{model_card}.yaml"]
        path = next((c for c in candidates if c.is_file()), None)
        if path is None:
            avail = sorted({c.stem for d in (Path("videoseal/cards"), CARDS_DIR) for c in d.glob("*.yaml")})
            print(f"Available model cards: {', '.join(avail)}")
            raise FileNotFoundError(f"Model card '{model_card}' not found")
    elif isinstance(model_card, Path):
        if not model_card.exists():
            raise FileNotFoundError(f"Model card file '{model_card}' not found")
        path = model_card
    else:
        raise TypeError("Model card must be a string or a Path object")
    with open(path, "r") as f:
        card = yaml.safe_load(f)
    cp = str(card["checkpoint_path"])
    if Path(cp).is_file():
        ckpt = Path(cp)
    elif (path.parent / cp).is_file():
        ckpt = path.parent / cp
    elif is_url(cp):
        ckpt = maybe_download_checkpoint(cp)
    else:
        raise RuntimeError(f"Path or uri {cp} is unknown or does not exist")
    return setup_model(card, ckpt)
