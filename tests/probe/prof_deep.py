"""Profiling aid: the deepest 2-D down layer (conv5: 96 -> 192 channels, 4x4 -> 2x2) at N = 1024."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from lshm_b200._lib import lib
from lshm_b200.engine import conv_image
dev = torch.device("cuda:0"); L = lib(); st = torch.cuda.current_stream().cuda_stream
N, A, Bc, s = 1024, 192, 96, 2
big = torch.randn(N, Bc, 2 * s, 2 * s, device=dev); small = torch.empty(N, A, s, s, device=dev)
w = torch.randn(A, Bc, 4, 4, device=dev) * 0.1; bias = torch.randn(A, device=dev)
img = conv_image(w, 2, 0, st)
for _ in range(4):
    L.down2d(big.data_ptr(), big[0].numel(), img.data_ptr(), bias.data_ptr(), None, 0, small.data_ptr(), small[0].numel(), N, A, Bc, s, s, 1, st)
torch.cuda.synchronize(); print("ok")
