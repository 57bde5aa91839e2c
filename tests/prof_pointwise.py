"""embed + detect of n frames at 3x768x768 (for ncu captures of the full-resolution kernels: resize, JND blend).
usage: prof_pointwise.py [n=64] [video|image]   (image mode with n=32 = the workload of bench.py's roofline_hbm leg)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import make_model_pair
model, orc, spec = make_model_pair("videoseal_1.0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.rand(n, 3, 768, 768).cuda()
vid = not (len(sys.argv) > 2 and sys.argv[2] == "image")
msgs = torch.randint(0, 2, (1 if vid else n, spec["nbits"]))
for _ in range(2):
    o = model.embed(x, msgs, is_video=vid)
    model.detect(o["imgs_w"], is_video=vid)
torch.cuda.synchronize()
print("done")
