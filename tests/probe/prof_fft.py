import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from lshm_b200 import lofar_tools as T
x = torch.randn(1024, 8, 128, 128, device="cuda"); xh = torch.randn_like(x)
for _ in range(3):
    y = T.fft_features(x, xh)
torch.cuda.synchronize(); print("ok")
