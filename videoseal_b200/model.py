"""`Videoseal`: host-side mirror of the reference's inference API (models/videoseal.py:258-428, models/wam.py:134-234)
whose compute is entirely in libvsb200.so (sm_100a kernels).  PyTorch is used for tensor plumbing only: allocating
outputs, holding the checkpoint, device placement and streams.

Drop-in surface kept: `embed`, `detect`, `extract_message`, `get_random_msg`, mutable attributes
`blender.scaling_w/scaling_i`, `step_size`, `chunk_size`, `video_mode`, `img_size`, `clamp`, `attenuation` (may be set to
None), `embedder.yuv`, and the nn.Module verbs `.eval() .to(device) .compile()`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import nn

from . import _lib

_DEF_INTERP = {"mode": "bilinear", "align_corners": False, "antialias": True}


class Blender:
    """models/blender.py: only the additive method is on the hot path (all shipped cards)."""
    AVAILABLE_BLENDING_METHODS = ["additive"]

    def __init__(self, scaling_i: float, scaling_w: float, method: str = "additive"):
        if method != "additive":
            raise NotImplementedError(f"blending method '{method}' is not implemented")
        self.scaling_i, self.scaling_w, self.method = scaling_i, scaling_w, method


class JND:
    """modules/jnd.py stand-in: `heatmaps()` runs the fused CUDA kernel."""

    def __init__(self, owner: "Videoseal", in_channels: int, out_channels: int):
        self._owner, self.in_channels, self.out_channels = owner, in_channels, out_channels

    def to(self, *a, **k):
        return self

    def heatmaps(self, imgs: torch.Tensor) -> torch.Tensor:
        m = self._owner
        x = m._to_dev(imgs)
        out = torch.empty((x.shape[0], self.out_channels, x.shape[2], x.shape[3]), device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().vsb_jnd_heatmaps(m._handle(), x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[2], x.shape[3],
                                               m._stream()))
        return out.to(imgs.device)


class _Embedder:
    """models/embedder.py:130-165 UnetEmbedder seam: forward(imgs_res RGB [B,3,S,S] in [0,1], msgs [B,K]) -> delta"""

    def __init__(self, owner: "Videoseal"):
        self._owner = owner

    # the colour handling is baked into the native first layer (RGB -> Y fused): changing it rebuilds the native handle
    @property
    def yuv(self) -> bool:
        return bool(self._owner.spec["yuv"])

    @yuv.setter
    def yuv(self, v: bool):
        m = self._owner
        if bool(v) != bool(m.spec["yuv"]):
            want = 1 if v else 3
            if m.spec["unet"]["in_channels"] != want:
                raise ValueError(f"embedder.yuv={bool(v)} needs a U-Net with {want} input channel(s); this checkpoint has "
                                 f"{m.spec['unet']['in_channels']}")
            m._respec(yuv=bool(v))

    def get_random_msg(self, bsz: int = 1, nb_repetitions: int = 1) -> torch.Tensor:
        nbits = self._owner.spec["nbits"]
        if nb_repetitions != 1:   # modules/msg_processor.py:43-57
            assert nbits % nb_repetitions == 0
            aux = torch.randint(0, 2, (bsz, nbits // nb_repetitions))
            return aux.unsqueeze(1).repeat(1, nb_repetitions, 1).view(bsz, nbits)
        return torch.randint(0, 2, (bsz, nbits))

    def __call__(self, imgs_res: torch.Tensor, msgs: torch.Tensor) -> torch.Tensor:
        m = self._owner
        x = m._to_dev(imgs_res)
        S = m.spec["img_size"]
        assert x.shape[1:] == (3, S, S), "embedder seam takes RGB frames at processing size (Y extraction is fused)"
        mm = m._msgs_u8(msgs)
        out = torch.empty((x.shape[0], m.spec["unet"]["out_channels"], S, S), device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().vsb_embedder_forward(m._handle(), x.data_ptr(), mm.data_ptr(), mm.shape[0], out.data_ptr(),
                                                   x.shape[0], m._stream()))
        return out


class _Detector:
    """models/extractor.py:140-167 ConvnextExtractor seam: forward(imgs [B,3,S,S] in [0,1]) -> logits [B,1+K]"""

    def __init__(self, owner: "Videoseal"):
        self._owner = owner

    def __call__(self, imgs_res: torch.Tensor) -> torch.Tensor:
        return self._owner._detect_dev(self._owner._to_dev(imgs_res), True)


class Videoseal(nn.Module):
    def __init__(self, spec: dict):
        super().__init__()
        self.spec = spec
        # checkpoint tensors live in a plain dict (not nn.Parameters): the network is executed by the native library
        self._sd: dict = {}
        self._native = None          # (device_index, ctypes handle)
        self._dev = torch.device("cpu")
        self._anchor = nn.Parameter(torch.zeros(1), requires_grad=False)  # lets `.to()` / `.device` behave like a module
        self.blender = Blender(spec["scaling_i"], spec["scaling_w"])
        self._attenuation: Optional[JND] = JND(self, *spec["jnd"]) if spec["jnd"][0] else None
        self.clamp = True
        self.chunk_size, self.step_size = spec["chunk_size"], spec["step_size"]
        self.video_mode = "repeat"
        self.lowres_attenuation = False
        self.embedder = _Embedder(self)
        self.detector = _Detector(self)

    # ------------------------------------------------------------------ attributes the native handle bakes in
    # (ModelDesc of vsb_model_create): assigning them rebuilds the handle on the next call instead of being a silent no-op
    def _respec(self, **changes):
        self.spec = {**self.spec, **changes}
        self._release()

    @property
    def img_size(self) -> int:
        return self.spec["img_size"]

    @img_size.setter
    def img_size(self, v: int):
        v = int(v)
        if v != self.spec["img_size"]:
            if v <= 0 or v % 128 != 0:
                raise ValueError(f"img_size={v}: the native path needs a processing size that is a positive multiple of 128")
            self._respec(img_size=v)

    @property
    def attenuation(self) -> Optional[JND]:
        return self._attenuation

    @attenuation.setter
    def attenuation(self, jnd):
        """None switches the attenuation off (wam.py:196); a JND (ours or the reference's modules/jnd.py object: anything with
        in_channels / out_channels in {1, 3}) switches it on with that channel configuration"""
        if jnd is None:
            self._attenuation = None
            return
        cin, cout = int(getattr(jnd, "in_channels", 0)), int(getattr(jnd, "out_channels", 0))
        if cin not in (1, 3) or cout not in (1, 3):
            raise ValueError("attenuation must be None or a JND with in_channels / out_channels in {1, 3}")
        if (cin, cout) != tuple(self.spec["jnd"]):
            self._respec(jnd=(cin, cout))
        self._attenuation = jnd if isinstance(jnd, JND) and jnd._owner is self else JND(self, cin, cout)

    # ------------------------------------------------------------------ nn.Module plumbing
    @property
    def device(self) -> torch.device:
        return self._anchor.device

    def load_state_dict(self, state_dict, strict: bool = False):
        self._sd = {k: v.detach().to(torch.float32).cpu().contiguous() for k, v in state_dict.items()
                    if isinstance(v, torch.Tensor) and not k.endswith("num_batches_tracked")}
        self._release()
        return "<All keys matched successfully>" if self._sd else "<empty state dict>"

    def state_dict(self, *a, **k):
        return dict(self._sd)

    def compile(self, *a, **k):   # reference: nn.Module.compile only wraps forward(); embed/detect are unaffected
        return self

    def forward(self, *a, **k):
        raise NotImplementedError("training-time forward() is outside the inference hot path; use embed()/detect()")

    def _release(self):
        if self._native is not None:
            _lib.lib().vsb_model_destroy(self._native[1])
            self._native = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _handle(self):
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("videoseal_b200 runs on an NVIDIA B200 (sm_100a) only: call model.to('cuda') first; "
                               "there is no CPU fallback")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._native is not None and self._native[0] == idx:
            return self._native[1]
        self._release()
        L = _lib.lib()
        s = self.spec
        d = _lib.ModelDesc()
        d.nbits, d.hidden, d.img_size, d.yuv = s["nbits"], s["hidden"], s["img_size"], int(s["yuv"])
        u = s["unet"]
        d.unet_in_ch, d.unet_out_ch, d.unet_levels = u["in_channels"], u["out_channels"], len(u["z"])
        for i, z in enumerate(u["z"]):
            d.unet_z[i] = z
        d.unet_num_blocks = u["num_blocks"]
        acts, norms = {"relu": 0, "silu": 1}, {"batch": 0, "rms": 1}
        if u["activation"] not in acts or not any(u["normalization"].startswith(k) for k in norms):
            raise NotImplementedError(f"U-Net activation/normalization {u['activation']}/{u['normalization']} not implemented")
        d.unet_act = acts[u["activation"]]
        d.unet_norm = 0 if u["normalization"].startswith("batch") else 1
        d.unet_last_tanh = int(u["last_tanh"])
        cn = s["convnext"]
        for i in range(4):
            d.ext_depths[i], d.ext_dims[i] = cn["depths"][i], cn["dims"][i]
        d.ext_stem_stride = cn["stem_stride"]
        d.jnd_in_ch, d.jnd_out_ch = s["jnd"]
        h = C.c_void_p()
        _lib.check(L.vsb_model_create(C.byref(d), C.byref(h)))
        try:
            for k, v in self._sd.items():
                if k.startswith(("attenuation.", "rgb2yuv.")) or v.dim() > 4:
                    continue
                shape = (C.c_int64 * max(1, v.dim()))(*v.shape)
                _lib.check(L.vsb_model_set_tensor(h, k.encode(), v.data_ptr(), shape, v.dim()))
            _lib.check(L.vsb_model_finalize(h, idx))
        except Exception:
            L.vsb_model_destroy(h)
            raise
        self._native = (idx, h)
        return h

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _to_dev(self, x: torch.Tensor) -> torch.Tensor:
        return x.to(device=self.device, dtype=torch.float32).contiguous()

    def _msgs_u8(self, msgs: torch.Tensor) -> torch.Tensor:
        if msgs.dim() != 2 or msgs.shape[1] != self.spec["nbits"]:
            raise ValueError(f"msgs must be [B, {self.spec['nbits']}], got {tuple(msgs.shape)}")
        # thresholded on the side the tensor lives on: a host tensor (the usual case) costs one small H2D copy and no device kernel
        return (msgs > 0.5).to(torch.uint8).contiguous().to(self.device)

    def _flags(self, interpolation: dict, lowres: bool = False) -> int:
        if interpolation.get("mode", "bilinear") != "bilinear" or interpolation.get("align_corners", False):
            raise NotImplementedError("only bilinear, align_corners=False interpolation is implemented")
        f = 0
        if self.clamp:
            f |= _lib.FLAG_CLAMP
        if lowres:
            f |= _lib.FLAG_LOWRES_ATTN
        if self.attenuation is None:
            f |= _lib.FLAG_NO_ATTENUATION
        if not interpolation.get("antialias", False):
            f |= _lib.FLAG_RESIZE_NO_AA
        return f

    # ------------------------------------------------------------------ public API
    def get_random_msg(self, bsz: int = 1, nb_repetitions: int = 1) -> torch.Tensor:
        return self.embedder.get_random_msg(bsz, nb_repetitions)

    @torch.no_grad()
    def embed(self, imgs: torch.Tensor, msgs: torch.Tensor = None, is_video: bool = True, interpolation: dict = _DEF_INTERP,
              lowres_attenuation: bool = False) -> dict:
        """models/videoseal.py:258-350 (video) / models/wam.py:134-204 (images)."""
        if imgs.dim() != 4 or imgs.shape[1] != 3:
            raise ValueError("imgs must be [F, 3, H, W]")
        F_, H, W = imgs.shape[0], imgs.shape[2], imgs.shape[3]
        if msgs is None:
            msgs = self.get_random_msg(F_ if not is_video else 1)
        elif is_video:
            assert msgs.shape[0] == 1, "Message should be unique"
        x = self._to_dev(imgs)
        mm = self._msgs_u8(msgs)
        if not is_video and mm.shape[0] != F_:
            raise ValueError("image mode needs one message per image")
        step = self.step_size if is_video else 1
        imgs_w = torch.empty_like(x)
        preds_w = None
        if not is_video:
            pc = self.spec["unet"]["out_channels"]
            if self.attenuation is not None:       # hmaps * preds_w broadcasts to max(out_channels) channels (wam.py:189-193)
                pc = max(pc, self.attenuation.out_channels)
            preds_w = torch.empty((F_, pc, H, W), device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().vsb_embed(
            self._handle(), x.data_ptr(), mm.data_ptr(), mm.shape[0], imgs_w.data_ptr(),
            preds_w.data_ptr() if preds_w is not None else None, F_, H, W, step, _lib.VIDEO_MODES[self.video_mode],
            int(self.chunk_size), float(self.blender.scaling_i), float(self.blender.scaling_w), self._flags(interpolation, lowres_attenuation),
            self._stream()))
        if is_video:
            return {"imgs_w": imgs_w.to(imgs.device), "msgs": msgs[0:1].repeat(F_, 1)}
        return {"msgs": msgs, "preds_w": preds_w.to(imgs.device), "imgs_w": imgs_w.to(imgs.device)}

    def _detect_dev(self, x: torch.Tensor, antialias: bool) -> torch.Tensor:
        out = torch.empty((x.shape[0], 1 + self.spec["nbits"]), device=x.device, dtype=torch.float32)
        flags = 0 if antialias else _lib.FLAG_RESIZE_NO_AA
        _lib.check(_lib.lib().vsb_detect(self._handle(), x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[2], x.shape[3], flags,
                                         self._stream()))
        return out

    @torch.no_grad()
    def detect(self, imgs: torch.Tensor, is_video: bool = True, interpolation: dict = _DEF_INTERP) -> dict:
        """models/videoseal.py:352-388; frames are independent so the chunk loop is internal to the library."""
        if not is_video:
            interpolation = _DEF_INTERP       # the reference drops `interpolation` on this branch (videoseal.py:374)
        self._flags(interpolation)            # validates the interpolation dict
        preds = self._detect_dev(self._to_dev(imgs), bool(interpolation.get("antialias", False)))
        return {"preds": preds.to(imgs.device)}

    def extract_message(self, imgs: torch.Tensor, aggregation: str = "avg",
                        interpolation: dict = {"mode": "bilinear", "align_corners": False, "antialias": False}) -> torch.Tensor:
        """models/videoseal.py:390-428"""
        preds = self.detect(imgs, is_video=True, interpolation=interpolation)["preds"]
        bit_preds = preds[:, 1:]
        if aggregation is None:
            decoded = bit_preds
        elif aggregation == "avg":
            decoded = bit_preds.mean(dim=0)
        elif aggregation == "squared_avg":
            decoded = (bit_preds * bit_preds.abs()).mean(dim=0)
        elif aggregation == "l1norm_avg":
            decoded = (bit_preds * torch.norm(bit_preds, p=1, dim=1).unsqueeze(1)).mean(dim=0)
        elif aggregation == "l2norm_avg":
            decoded = (bit_preds * torch.norm(bit_preds, p=2, dim=1).unsqueeze(1)).mean(dim=0)
        else:
            raise ValueError(f"unknown aggregation {aggregation}")
        return (decoded > 0).squeeze().unsqueeze(0)

    # ------------------------------------------------------------------ test seam
    def debug_tensor(self, name: str) -> torch.Tensor:
        """Intermediate activation of the most recent sub-batch as an fp32 NHWC tensor (tests only)."""
        L = _lib.lib()
        shape = (C.c_int64 * 4)()
        n = L.vsb_debug_get_tensor(self._handle(), name.encode(), None, 0, shape)
        if n < 0:
            raise KeyError(L.vsb_last_error().decode())
        out = torch.empty(list(shape), dtype=torch.float32)
        n = L.vsb_debug_get_tensor(self._handle(), name.encode(), out.data_ptr(), out.numel(), shape)
        if n < 0:
            raise RuntimeError(L.vsb_last_error().decode())
        return out
