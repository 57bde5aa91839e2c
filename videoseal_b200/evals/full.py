"""Identity-augmentation slice of the reference's full evaluation (videoseal/evals/full.py:57-248 with only_identity=True,
skip_image_metrics for everything but PSNR): per item embed -> PSNR -> extract -> bit accuracy / p-value / capacity, one CSV
row per item with the reference's column names (`bit_acc_Identity_0`, `pvalue_Identity_0`, `log_pvalue_Identity_0`,
`capacity_Identity_0`, `embed_time`, `extract_time`, ...), so that its reports line up with the reference's own runs.
The augmentation bank (JPEG / H264 / crops ...), SSIM / LPIPS / VMAF and the image dumps need packages outside this path."""
from __future__ import annotations

import math
import os
import time
from typing import Iterable

import torch

from .metrics import bit_accuracy, capacity, psnr, pvalue

_DEF_INTERP = {"mode": "bilinear", "align_corners": False, "antialias": True}


def _sync(x: torch.Tensor):
    if x.is_cuda:
        torch.cuda.synchronize(x.device)


@torch.no_grad()
def evaluate(model, dataset: Iterable, is_video: bool, output_dir: str, num_frames: int = 24 * 3, video_aggregation: str = "avg",
             decoding: bool = True, interpolation: dict = _DEF_INTERP, lowres_attenuation: bool = False,
             skip_image_metrics: bool = False):
    os.makedirs(output_dir, exist_ok=True)
    all_metrics = []
    path = os.path.join(output_dir, "metrics.csv")
    with open(path, "w") as f:
        for it, item in enumerate(dataset):
            if item is None:
                continue
            imgs = item[0]
            if not is_video:
                imgs = imgs.unsqueeze(0)
            m = {"iteration": it, "t": imgs.shape[-4], "h": imgs.shape[-2], "w": imgs.shape[-1]}
            _sync(imgs)
            t0 = time.time()
            out = model.embed(imgs, is_video=is_video, interpolation=interpolation, lowres_attenuation=lowres_attenuation)
            _sync(out["imgs_w"])
            m["embed_time"] = time.time() - t0
            msgs, imgs_w = out["msgs"][:num_frames], out["imgs_w"][:num_frames]
            imgs = imgs[:num_frames]
            if not skip_image_metrics:
                m["psnr"] = psnr(imgs_w, imgs.to(imgs_w.device), is_video).mean().item()
            if decoding:
                t0 = time.time()
                if is_video:
                    bit_preds = model.extract_message(imgs_w, video_aggregation, interpolation).float() * 2 - 1   # bool [1,K] -> +-1
                    msgs = msgs[:1]
                else:
                    bit_preds = model.detect(imgs_w, is_video=False)["preds"][:, 1:]
                _sync(bit_preds)
                m["extract_time"] = time.time() - t0
                bit_preds, tgt = bit_preds.cpu(), msgs.cpu()
                pv = pvalue(bit_preds, tgt).nanmean().item()
                m["bit_acc_Identity_0"] = bit_accuracy(bit_preds, tgt).nanmean().item()
                m["pvalue_Identity_0"] = pv
                m["log_pvalue_Identity_0"] = -math.log10(pv) if pv > 0 else -100
                m["capacity_Identity_0"] = capacity(bit_preds, tgt).nanmean().item()
            all_metrics.append(m)
            if len(all_metrics) == 1:
                f.write(",".join(m.keys()) + "\n")
            f.write(",".join(map(str, m.values())) + "\n")
            f.flush()
    return all_metrics


def main(argv=None):
    import argparse
    from ..cfg import setup_model_from_checkpoint
    from .speed import _bool, synthetic_items
    ap = argparse.ArgumentParser(description="identity-augmentation evaluation of a card / checkpoint on this backend")
    ap.add_argument("--checkpoint", required=True)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--is_video", type=_bool, default=False)
    ap.add_argument("--num_frames", type=int, default=24 * 3)
    ap.add_argument("--num_samples", type=int, default=10)
    ap.add_argument("--size", type=int, nargs="+", default=[768])
    ap.add_argument("--video_aggregation", default="avg")
    ap.add_argument("--lowres_attenuation", type=_bool, default=False)
    ap.add_argument("--scaling_w", default=None)
    ap.add_argument("--output_dir", default="output/full")
    a = ap.parse_args(argv)
    model = setup_model_from_checkpoint(a.checkpoint).eval().to(a.device)
    if a.scaling_w is not None:
        model.blender.scaling_w = float(a.scaling_w)
    rows = evaluate(model, synthetic_items(a.num_samples, a.is_video, a.num_frames, a.size[0], a.size[-1], a.device), a.is_video,
                    a.output_dir, a.num_frames, a.video_aggregation, lowres_attenuation=a.lowres_attenuation)
    keys = [k for k in rows[0] if k not in ("iteration", "t", "h", "w")]
    print({k: sum(r[k] for r in rows) / len(rows) for k in keys})


if __name__ == "__main__":
    main()
