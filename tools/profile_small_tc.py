#!/usr/bin/env python
"""fwd+bwd at a mid-size batch on the tensor path (for an ncu launch list)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from wav2vec_contr_loss_b200 import functional as Fn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1337)
z = torch.nn.functional.normalize(torch.randn(n, 256, generator=g), dim=1).to(dev).to(torch.bfloat16)
y = (torch.rand(n, generator=g) < 0.5).to(torch.int32).to(dev)
prob = Fn.make_problem(n, 256, 1, tau=0.07, similarity=0, topk=15, alpha=alpha, flags=32)
for _ in range(3):
    stats, partials, loss = Fn.forward_rows(z, y, prob, want_loss=True)
    dz = Fn.backward_rows(z, y, stats, partials, None, prob, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
print("loss", float(loss))
