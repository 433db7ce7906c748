"""Single-config driver for ncu: BASELINE config 4 (parameter-shift gradient, grid_cluster(4,5)) on
B vectors.  `prof_grad.py [B]`"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mentpy_b200 as mb
from mentpy_b200.gradients import psr_gradient_batched

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
gs = mb.templates.grid_cluster(4, 5)
ps = mb.PatternSimulator(gs, backend="cuda-sv")
X = torch.rand((B, 16), dtype=torch.float64, device="cuda") * (2 * np.pi)
tgt = np.full(16, 0.25)
for _ in range(4):
    g = psr_gradient_batched(ps, X, tgt)
torch.cuda.synchronize()
print(float(g.abs().max()))
