"""one detect() of 64 frames (for ncu captures of the extractor kernels)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import make_model_pair
model, orc, spec = make_model_pair("videoseal_1.0")
x = torch.rand(64, 3, 256, 256).cuda()
for _ in range(2):
    model.detect(x, is_video=False)
torch.cuda.synchronize()
print("done")
