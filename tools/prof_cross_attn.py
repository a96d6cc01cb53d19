#!/usr/bin/env python
"""One cross-attention launch (SD1.5 top level, UNet batch 16) for ncu captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cremage_b200 import ops  # noqa: E402

B, heads, nq, nk, d = 16, 8, 4096, 77, 40
inner = heads * d
q = (torch.randn(B * nq, inner, device="cuda") * 0.5).to(ops.ACT)
kv = (torch.randn(B * nk, 2 * inner, device="cuda") * 0.5).to(ops.ACT)
for _ in range(3):
    out = ops.attention(q, kv[:, :inner], kv[:, inner:], B, heads, nq, nk, d, d ** -0.5)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
