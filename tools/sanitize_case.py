#!/usr/bin/env python
"""Small forward+backward cases of every kernel family, for `compute-sanitizer --tool memcheck`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from wav2vec_contr_loss_b200 import functional as Fn
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
def run(n, d, dtype, sim, lam, k, alpha, row_offset=0, n_rows=None, two_phase=False):
    z = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1).to(dev).to(dtype)
    y = torch.randint(0, 3, (n,), generator=g).to(torch.int32).to(dev)
    prob = Fn.make_problem(n, d, Fn._dtype_id(z), tau=0.07, similarity=Fn.similarity_id(sim), lambda_uni=lam, topk=k,
                           alpha=alpha, row_offset=row_offset, n_rows=n_rows, flags=32)
    if two_phase:
        ws = Fn.forward_rows_local(z, y, prob)
        stats, partials = Fn.forward_rows_remote(z, y, prob, ws)
    else:
        stats, partials, _ = Fn.forward_rows(z, y, prob, want_loss=(n_rows is None))
    if n_rows is None:
        dz = Fn.backward_rows(z, y, stats, partials, None, prob, out_dtype=torch.float32)
        if n <= Fn.SMALL_BATCH_MAX:
            Fn.loss_and_grad(z, y, prob)
    torch.cuda.synchronize()
    print("ok", n, d, dtype, sim, lam, k, alpha, row_offset, n_rows, two_phase, flush=True)
run(64, 256, torch.float32, "geodesic", 0.05, 15, 0.5)            # cluster kernel
run(300, 100, torch.float32, "cosine", 0.1, 7, 1.0)               # tiled FFMA kernels (+ column splits)
run(1000, 256, torch.bfloat16, "geodesic", 0.05, 15, 0.5)         # tensor path, ragged N, mining + uniformity
run(512, 256, torch.bfloat16, "cosine", 0.0, 15, 0.0)             # tensor path, plain
run(1024, 256, torch.bfloat16, "cosine", 0.0, 15, 0.5, row_offset=256, n_rows=384, two_phase=True)
print("all done")
