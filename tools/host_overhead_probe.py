#!/usr/bin/env python
"""Rough host-side cost of one N = 64 fwd+bwd through the module: CPU tensors, every C call stubbed out
(runs without a GPU; the numbers say where the Python-side time of a call goes, not what the kernels cost)."""
import ctypes
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wav2vec_contr_loss_b200 import functional as Fn, _cabi
from wav2vec_contr_loss_b200.loss import SupConBinaryLoss

class FakeLib:
    def __getattr__(self, name):
        return lambda *a, **k: 0
fake = FakeLib()
_cabi.load = lambda: fake
Fn._require_cuda = lambda t, w: None
Fn._stream = lambda dev: ctypes.c_void_p(0)
Fn._on = lambda dev: Fn._NO_SWITCH
orig_ws = Fn.workspace_bytes
Fn.workspace_bytes = lambda prob, device: 4096
# labels: pretend cuda path
def canon(labels, n):
    lab = labels.reshape(-1)
    keys = torch.empty(n, dtype=torch.int32)
    fake.supcon_label_keys(Fn._p(lab), 1, n, Fn._p(keys), Fn._stream(None))
    return keys
Fn.canonical_labels = canon
import wav2vec_contr_loss_b200.loss as L
z0 = torch.nn.functional.normalize(torch.randn(64, 256), dim=1)
y = (torch.arange(64) % 2).long()
fn = SupConBinaryLoss(0.07, "cosine")
def one():
    zz = z0.detach().requires_grad_(True)
    ls = fn(zz, y, topk_neg=15, alpha=0.0)
    torch.autograd.grad(ls, zz)
for _ in range(200): one()
t0 = time.perf_counter()
for _ in range(2000): one()
print("per call us", (time.perf_counter() - t0) / 2000 * 1e6)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): one()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
