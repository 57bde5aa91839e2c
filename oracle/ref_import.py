"""TEST INFRASTRUCTURE ONLY — never imported by the product package.

Imports the UNMODIFIED reference modules from /root/reference (read-only mount that
exists only in the build container, not on the GPU box) with stub modules for the
optional third-party packages that are not installed here (timm, av, decord,
pycocotools, omegaconf).  None of the stubbed symbols execute on the embed/detect
path.  Used by oracle/make_golden.py to (a) validate the
restatement in oracle/restate.py and (b) generate the golden vectors under
tests/golden/.
"""
import os
import sys
import types

import torch.nn as nn
import yaml

REF_ROOT = os.environ.get("VIDEOSEAL_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "videoseal"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _DropPath(nn.Module):  # convnext.py:36 only instantiates it when drop_path > 0
    def __init__(self, p=0.0):
        super().__init__()

    def forward(self, x):
        return x


class AD(dict):
    """attribute-dict standing in for omegaconf.DictConfig in build_embedder/build_extractor"""

    def __getattr__(self, k):
        if k in self:
            return self[k]
        raise AttributeError(k)

    __setattr__ = dict.__setitem__


def wrap(o):
    return AD({k: wrap(v) for k, v in o.items()}) if isinstance(o, dict) else o


_installed = False


def install_stubs():
    global _installed
    if _installed:
        return
    for name in ("timm", "timm.models", "timm.optim", "timm.scheduler"):
        if name not in sys.modules:
            _mod(name)
    _mod("timm.models.layers", trunc_normal_=nn.init.trunc_normal_, DropPath=_DropPath)
    if "av" not in sys.modules:
        _mod("av")
    if "decord" not in sys.modules:
        _mod("decord", VideoReader=None, cpu=None)
    if "pycocotools" not in sys.modules:
        _mod("pycocotools").mask = _mod("pycocotools.mask")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def build_reference_model(card_name: str):
    """mirrors utils/cfg.py:88-144 (setup_model) without omegaconf; returns (model, cfg)"""
    install_stubs()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from videoseal.models.embedder import build_embedder
        from videoseal.models.extractor import build_extractor
        from videoseal.models.videoseal import Videoseal
        from videoseal.augmentation.augmenter import get_dummy_augmenter
        from videoseal.modules.jnd import JND
    cfg = wrap(yaml.safe_load(open(os.path.join(REF_ROOT, "videoseal/cards", card_name + ".yaml"))))
    a = cfg.args
    img_size = a.get("img_size_proc", a.get("img_size_extractor"))
    emb = build_embedder(cfg.embedder.model, cfg.embedder.params, a.nbits, a.get("hidden_size_multiplier", 2))
    ext = build_extractor(cfg.extractor.model, cfg.extractor.params, img_size, a.nbits)
    att = None
    if str(a.attenuation).lower().startswith("jnd"):
        att_cfg = yaml.safe_load(open(os.path.join(REF_ROOT, "configs/attenuation.yaml")))
        att = JND(**att_cfg[a.attenuation])
    model = Videoseal(
        emb, ext, get_dummy_augmenter(), attenuation=att, scaling_w=a.scaling_w, scaling_i=a.scaling_i,
        img_size=img_size,
        chunk_size=a.get("videoseal_chunk_size", a.get("videowam_chunk_size")),
        step_size=a.get("videoseal_step_size", a.get("videowam_step_size")),
    ).eval()
    return model, cfg
