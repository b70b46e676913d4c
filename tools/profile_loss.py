"""A few iterations of the fused image loss at 1080p (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from topo4d_b200 import losses  # noqa: E402

v = int(sys.argv[1]) if len(sys.argv) > 1 else 1
r = torch.rand(v, 3, 1080, 1920, device="cuda:0", requires_grad=True)
t = torch.rand(v, 3, 1080, 1920, device="cuda:0")
m = torch.zeros(v, 3, device="cuda:0", requires_grad=True)
c = torch.zeros(v, 3, device="cuda:0", requires_grad=True)
for _ in range(4):
    r.grad = None
    losses.image_loss(r, t, m, c).backward()
torch.cuda.synchronize()
