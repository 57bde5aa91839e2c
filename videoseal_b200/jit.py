"""Tensor-returning facade with the surface of the reference's TorchScript artefact (docs/torchscript.md:25-175 of the
reference: `model(imgs, msgs) -> (imgs_w, preds)`, `embed`, `detect`, `detect_video_and_aggregate`, and the documented
attributes `scaling_w`, `img_size`, `clamp`, `do_attenuation`, `lowres_attenuation`, `chunk_size`, `step_size`,
`video_mode`) on top of this package's Videoseal.  Nothing is scripted: the work happens in libvsb200.so either way.

    m = videoseal_b200.jit.load("videoseal_1.0").to("cuda")
    img_w, preds = m(img, msg)                 # [B,3,H,W], [B,1+K]
    bits = m.detect_video_and_aggregate(video_w, aggregation="avg")    # float {0,1} [1,K]
"""
from __future__ import annotations

import torch


class TensorAPI:
    def __init__(self, model, lowres_attenuation: bool = True, chunk_size: int = 16):
        self.model = model
        self.lowres_attenuation = lowres_attenuation        # documented default of the artefact (docs/torchscript.md:176)
        self.model.chunk_size = chunk_size
        self._attenuation = model.attenuation

    # ---- documented attributes, forwarded to the model
    scaling_w = property(lambda s: s.model.blender.scaling_w, lambda s, v: setattr(s.model.blender, "scaling_w", float(v)))
    img_size = property(lambda s: s.model.img_size, lambda s, v: setattr(s.model, "img_size", int(v)))
    clamp = property(lambda s: s.model.clamp, lambda s, v: setattr(s.model, "clamp", bool(v)))
    chunk_size = property(lambda s: s.model.chunk_size, lambda s, v: setattr(s.model, "chunk_size", int(v)))
    step_size = property(lambda s: s.model.step_size, lambda s, v: setattr(s.model, "step_size", int(v)))
    video_mode = property(lambda s: s.model.video_mode, lambda s, v: setattr(s.model, "video_mode", str(v)))

    @property
    def do_attenuation(self) -> bool:
        return self.model.attenuation is not None

    @do_attenuation.setter
    def do_attenuation(self, on: bool):
        self.model.attenuation = self._attenuation if on else None

    def to(self, device):
        self.model = self.model.to(device)
        return self

    def eval(self):
        self.model.eval()
        return self

    # ---- functions
    @torch.no_grad()
    def embed(self, imgs: torch.Tensor, msgs: torch.Tensor, is_video: bool = False) -> torch.Tensor:
        return self.model.embed(imgs, msgs, is_video=is_video, lowres_attenuation=self.lowres_attenuation)["imgs_w"]

    @torch.no_grad()
    def detect(self, imgs: torch.Tensor, is_video: bool = False) -> torch.Tensor:
        return self.model.detect(imgs, is_video=is_video)["preds"]

    @torch.no_grad()
    def forward(self, imgs: torch.Tensor, msgs: torch.Tensor, is_video: bool = False):
        imgs_w = self.embed(imgs, msgs, is_video)
        return imgs_w, self.detect(imgs_w, is_video)

    __call__ = forward

    @torch.no_grad()
    def detect_video_and_aggregate(self, imgs: torch.Tensor, aggregation: str = "avg") -> torch.Tensor:
        return self.model.extract_message(imgs, aggregation).to(torch.float32)


def load(card_or_checkpoint, **kw) -> TensorAPI:
    from .cfg import setup_model_from_checkpoint
    return TensorAPI(setup_model_from_checkpoint(str(card_or_checkpoint)).eval(), **kw)
