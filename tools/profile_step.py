#!/usr/bin/env python
"""Minimal fwd+bwd driver for ncu: N x 256 bf16 SupCon, a few steps.
    python tools/profile_step.py [N] [steps] [cosine|geodesic] [alpha] [topk]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from wav2vec_contr_loss_b200 import functional as Fn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sim = sys.argv[3] if len(sys.argv) > 3 else "cosine"
alpha = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
topk = int(sys.argv[5]) if len(sys.argv) > 5 else 15
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1337)
z = torch.nn.functional.normalize(torch.randn(n, 256, generator=g), dim=1).to(dev).to(torch.bfloat16)
y = (torch.rand(n, generator=g) < 0.5).to(torch.int32).to(dev)
prob = Fn.make_problem(n, 256, 1, tau=0.07, similarity=Fn.similarity_id(sim), topk=topk, alpha=alpha, flags=32)
for _ in range(steps):
    stats, partials, loss = Fn.forward_rows(z, y, prob, want_loss=True)
    dz = Fn.backward_rows(z, y, stats, partials, None, prob, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
print("loss", float(loss))
