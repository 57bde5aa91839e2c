"""A/B of the slice sizes of vsb_embed_detect_host (VSB_E2E_SL0 / VSB_E2E_SL1) in ONE process: pinned host frames in, watermarked
frames + logits out, wall clock around synchronous calls (the e2e leg of bench.py), batch 64 x 3x256x256, videoseal_1.0"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videoseal_b200 import _lib
from tests.util import make_model_pair
model, orc, spec = make_model_pair("videoseal_1.0")
L = _lib.lib()
B, S, K = 64, 256, spec["nbits"]
g = torch.Generator().manual_seed(0)
h_in = [torch.rand(B, 3, S, S, generator=g).pin_memory() for _ in range(2)]
h_msgs = torch.randint(0, 2, (B, K), generator=g).to(torch.uint8).pin_memory()
h_out = torch.empty(B, 3, S, S).pin_memory()
h_log = torch.empty(B, 1 + K).pin_memory()
h = model._handle()

def step(i):
    _lib.check(L.vsb_embed_detect_host(h, h_in[i % 2].data_ptr(), h_msgs.data_ptr(), B, h_out.data_ptr(), h_log.data_ptr(), B, S, S,
                                       1, 0, int(model.chunk_size), float(model.blender.scaling_i), float(model.blender.scaling_w), _lib.FLAG_CLAMP))

ref = None
for sl0, sl1 in [(32, 32), (16, 48), (8, 56), (16, 24), (24, 40), (32, 32)]:
    os.environ["VSB_E2E_SL0"], os.environ["VSB_E2E_SL1"] = str(sl0), str(sl1)
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for i in range(n):
        step(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    step(0)
    same = True if ref is None else bool(torch.equal(ref[0], h_out) and torch.equal(ref[1], h_log))
    if ref is None:
        ref = (h_out.clone(), h_log.clone())
    print(f"sl0={sl0} sl1={sl1}: {B * n / dt:.0f} frames/s ({1000 * dt / n:.3f} ms/step) identical_to_32_32={same}", flush=True)
