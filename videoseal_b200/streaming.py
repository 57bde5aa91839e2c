"""RGB24 clip I/O around embed/detect: the two inner functions of the reference's streaming CLI, same names and arguments
(inference_streaming.py:23-33 `embed_video_clip`, :116-124 `detect_video_clip`), running through `vsb_frames_host_u8`.

The reference converts uint8 HWC -> float32 CHW on the host, calls the model, and converts back; here the uint8 frames cross
PCIe as they are (4x fewer bytes each way) and both conversions run on the GPU inside the chunked, copy/compute-overlapped
host entry point.  Results are those of the reference's arithmetic: x = u8 / 255, u8 = (imgs_w * 255).byte().
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .model import _DEF_INTERP, Videoseal


def _as_u8_clip(clip) -> np.ndarray:
    a = clip.numpy() if isinstance(clip, torch.Tensor) else np.asarray(clip)
    if a.dtype != np.uint8 or a.ndim != 4 or a.shape[3] != 3:
        raise ValueError("clip must be uint8 [F, H, W, 3] (RGB24 frames)")
    return np.ascontiguousarray(a)


def _run(model: Videoseal, clip: np.ndarray, msgs, want_frames: bool, want_logits: bool, lowres_attenuation: bool, is_video: bool):
    F_, H, W = clip.shape[0], clip.shape[1], clip.shape[2]
    nbits = model.spec["nbits"]
    out = np.empty_like(clip) if want_frames else None
    logits = np.empty((F_, 1 + nbits), dtype=np.float32) if want_logits else None
    m8 = None
    n_msgs = 0
    if want_frames:
        if msgs is None:
            msgs = model.get_random_msg(1 if is_video else F_)
        m8 = (msgs.detach().cpu() > 0.5).to(torch.uint8).contiguous().numpy()
        if m8.ndim != 2 or m8.shape[1] != nbits or m8.shape[0] not in (1, F_) or (is_video and m8.shape[0] != 1):
            raise ValueError("msgs must be [1, nbits] for a video clip (or [F, nbits] with is_video=False)")
        n_msgs = m8.shape[0]
    interp = _DEF_INTERP
    flags = model._flags(interp, lowres_attenuation)
    _lib.check(_lib.lib().vsb_frames_host_u8(
        model._handle(), clip.ctypes.data_as(C.c_void_p), m8.ctypes.data_as(C.c_void_p) if m8 is not None else None, n_msgs,
        out.ctypes.data_as(C.c_void_p) if out is not None else None,
        logits.ctypes.data_as(C.c_void_p) if logits is not None else None,
        F_, H, W, model.step_size if is_video else 1, _lib.VIDEO_MODES[model.video_mode], int(model.chunk_size),
        float(model.blender.scaling_i), float(model.blender.scaling_w), flags))
    return out, (torch.from_numpy(logits) if logits is not None else None), msgs


def embed_video_clip(model: Videoseal, clip: np.ndarray, msgs: torch.Tensor) -> np.ndarray:
    """inference_streaming.py:23-33: uint8 [F,H,W,3] -> watermarked uint8 [F,H,W,3] (is_video=True, lowres_attenuation=True)."""
    out, _, _ = _run(model, _as_u8_clip(clip), msgs, True, False, True, True)
    return out


def detect_video_clip(model: Videoseal, clip: np.ndarray) -> torch.Tensor:
    """inference_streaming.py:116-124: uint8 [F,H,W,3] -> logits of the message bits [F, nbits] (column 0 dropped)."""
    _, logits, _ = _run(model, _as_u8_clip(clip), None, False, True, False, True)
    return logits[:, 1:]


def embed_detect_video_clip(model: Videoseal, clip: np.ndarray, msgs: torch.Tensor = None, lowres_attenuation: bool = True,
                            is_video: bool = True):
    """embed_video_clip followed by detect_video_clip on the watermarked uint8 frames, in one pass over the GPU.
    Returns (frames_w uint8 [F,H,W,3], preds [F, 1+nbits], msgs)."""
    return _run(model, _as_u8_clip(clip), msgs, True, True, lowres_attenuation, is_video)
