"""SURVEY 8(f)3 / 8(f)4 on the CUDA path: the raw-training-checkpoint loader (utils/cfg.py:50-85,156-178 of the reference), the
TorchScript artefact's tensor-returning surface (docs/torchscript.md:133-165), and the evaluation slice (evals/speed.py:36-148,
evals/full.py:57-248) all end in libvsb200.so; their outputs are compared with the CPU oracle on the same inputs."""
import argparse
import csv
import os

import pytest
import torch

from oracle import restate
from tests.util import load_card, tmpdir, SEED
from tests.test_e2e_gpu import logits_ok, PIX_TOL, LOGIT_RTOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def train_ckpt():
    """a raw TRAINING checkpoint ({'args': Namespace, 'model': state_dict}) of the videoseal_1.0 architecture, named presets only"""
    card = load_card("videoseal_1.0")
    spec = restate.spec_from_card(card)
    sd = restate.synth_state_dict(spec, seed=SEED)
    args = dict(card["args"])
    args.update({"embedder_config": "configs/embedder.yaml", "extractor_config": "configs/extractor.yaml",
                 "embedder_model": "unet_small2_yuv_quant", "extractor_model": "convnext_tiny", "lr": 1e-4, "epochs": 3})
    path = tmpdir() / "train_gpu_v1.pth"
    torch.save({"args": argparse.Namespace(**args), "model": sd, "epoch": 3}, path)
    return path, restate.OracleModel(spec, sd), spec


def test_setup_model_from_checkpoint_runs_on_the_gpu(train_ckpt):
    from videoseal_b200.utils.cfg import setup_model_from_checkpoint     # the reference's import path (evals/full.py:292)
    path, orc, spec = train_ckpt
    model = setup_model_from_checkpoint(str(path)).eval().to("cuda:0")
    g = torch.Generator().manual_seed(31)
    imgs = torch.rand(2, 3, 300, 340, generator=g)
    msgs = torch.randint(0, 2, (2, spec["nbits"]), generator=g)
    with torch.no_grad():
        ref = orc.embed(imgs, msgs, is_video=False)
        ref_det = orc.detect(ref["imgs_w"], is_video=False)["preds"]
    out = model.embed(imgs.cuda(), msgs, is_video=False)
    assert (out["imgs_w"].cpu() - ref["imgs_w"]).abs().max().item() <= PIX_TOL
    rel, flips, _ = logits_ok(model.detect(out["imgs_w"], is_video=False)["preds"].cpu(), ref_det)
    assert rel <= LOGIT_RTOL and flips == 0


def test_tensor_api_surface(train_ckpt):
    """docs/torchscript.md: model(imgs, msgs) -> (imgs_w, preds); detect_video_and_aggregate; documented attributes are live"""
    from videoseal_b200 import jit
    path, orc, spec = train_ckpt
    m = jit.load(str(path)).to("cuda:0").eval()
    assert m.lowres_attenuation is True and m.chunk_size == 16
    g = torch.Generator().manual_seed(32)
    imgs = torch.rand(2, 3, 288, 320, generator=g)
    msgs = torch.randint(0, 2, (2, spec["nbits"]), generator=g)
    imgs_w, preds = m(imgs.cuda(), msgs)
    assert imgs_w.shape == imgs.shape and preds.shape == (2, 1 + spec["nbits"])
    with torch.no_grad():
        ref = orc.embed(imgs, msgs, is_video=False, lowres_attenuation=True)
        ref_det = orc.detect(ref["imgs_w"], is_video=False)["preds"]
    assert (imgs_w.cpu() - ref["imgs_w"]).abs().max().item() <= PIX_TOL
    rel, flips, _ = logits_ok(preds.cpu(), ref_det)
    assert rel <= LOGIT_RTOL and flips == 0
    # attributes: scaling_w, do_attenuation, clamp, step_size / video_mode are forwarded to the native call
    m.scaling_w = 0.05
    m.do_attenuation = False
    old = (orc.scaling_w, orc.attenuation)
    try:
        orc.scaling_w, orc.attenuation = 0.05, None
        with torch.no_grad():
            ref2 = orc.embed(imgs, msgs, is_video=False, lowres_attenuation=True)
        assert (m.embed(imgs.cuda(), msgs).cpu() - ref2["imgs_w"]).abs().max().item() <= PIX_TOL
    finally:
        orc.scaling_w, orc.attenuation = old
        m.scaling_w = spec["scaling_w"]
        m.do_attenuation = True
    vid = torch.rand(9, 3, 272, 304, generator=g)
    vmsg = torch.randint(0, 2, (1, spec["nbits"]), generator=g)
    m.step_size = orc.step_size = 4
    try:
        vw = m.embed(vid.cuda(), vmsg, is_video=True)
        with torch.no_grad():
            rv = orc.embed(vid, vmsg, is_video=True, lowres_attenuation=True)
            ref_bits = orc.extract_message(rv["imgs_w"], "avg")
            agg = orc.detect(rv["imgs_w"], is_video=True)["preds"][:, 1:].mean(0)
        assert (vw.cpu() - rv["imgs_w"]).abs().max().item() <= PIX_TOL
        bits = m.detect_video_and_aggregate(vw, aggregation="avg")
        assert bits.dtype == torch.float32 and bits.shape == (1, spec["nbits"])
        sure = agg.abs() > 1e-2 * agg.abs().max()
        assert ((bits.cpu()[0] > 0.5) == ref_bits[0])[sure].all()
    finally:
        orc.step_size = spec["step_size"]


def test_speed_tester_and_full_eval_on_the_cuda_path(train_ckpt, tmp_path):
    from videoseal_b200.utils.cfg import setup_model_from_checkpoint
    from videoseal_b200.evals import full, speed
    from videoseal_b200 import _lib
    path, orc, spec = train_ckpt
    model = setup_model_from_checkpoint(str(path)).eval().to("cuda:0")
    L = _lib.lib()
    # evals/speed.py semantics: embed timed separately from extract, averages over the items
    L.vsb_launch_count(1)
    res = speed.SpeedTester("cuda").test_speed(model, speed.synthetic_items(2, False, 1, 320, 352, "cuda"), is_video=False, num_runs=2)
    assert res["avg_embedding_time"] > 0 and res["avg_extraction_time"] > 0 and len(res["image_shape"]) == 2
    assert L.vsb_launch_count(0) > 100, "the speed tester did not reach the native kernels"
    resv = speed.SpeedTester("cuda").test_speed(model, speed.synthetic_items(1, True, 12, 256, 256, "cuda"), is_video=True, num_frames=12, num_runs=2)
    assert resv["avg_embedding_ms_per_frame"] > 0 and resv["avg_extraction_ms_per_frame"] > 0
    # evals/full.py identity slice: metrics.csv with the reference's columns; PSNR / bit accuracy agree with the oracle's
    items = list(speed.synthetic_items(2, False, 1, 300, 280, "cuda", seed=5))
    torch.manual_seed(77)           # model.embed draws the random messages (wam.py:65)
    rows = full.evaluate(model, items, is_video=False, output_dir=str(tmp_path), lowres_attenuation=False)
    with open(os.path.join(tmp_path, "metrics.csv")) as f:
        got = list(csv.DictReader(f))
    assert len(got) == 2 and {"iteration", "t", "h", "w", "embed_time", "psnr", "extract_time", "bit_acc_Identity_0", "pvalue_Identity_0",
                              "log_pvalue_Identity_0", "capacity_Identity_0"} <= set(got[0])
    torch.manual_seed(77)
    for it, (x, _) in enumerate(items):
        msgs = model.get_random_msg(1)
        with torch.no_grad():
            ref = orc.embed(x.cpu().unsqueeze(0), msgs, is_video=False)
            ref_preds = orc.detect(ref["imgs_w"], is_video=False)["preds"]
        ref_bits = ref_preds[:, 1:]
        assert abs(rows[it]["psnr"] - restate.psnr(ref["imgs_w"], x.cpu().unsqueeze(0)).mean().item()) < 0.02
        unsure = int((ref_bits.abs() <= 1e-2 * ref_bits.abs().max()).sum())
        assert abs(rows[it]["bit_acc_Identity_0"] - restate.bit_accuracy(ref_preds, msgs).mean().item()) <= (unsure + 0.5) / spec["nbits"]
