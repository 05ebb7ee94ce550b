#!/usr/bin/env python
"""Minimal driver for ncu: the fused compression-head pass (head_pool_fwd_kernel) on the BASELINE configs[4] shape,
hs = (64, 25, 1024, 199) fp32 = 1.304 GB read once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from wav2vec_contr_loss_b200.head import layer_time_pool
dev = torch.device("cuda:0")
hs = torch.randn(64, 25, 1024, 199, device=dev)
for _ in range(3):
    out = layer_time_pool(hs, 0.0, 0.01, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = layer_time_pool(hs, 0.0, 0.01, None)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"head_pool_fwd: {ms:.4f} ms  {hs.numel() * 4 / ms / 1e6:.0f} GB/s (algorithmic bytes = hs read once)")
