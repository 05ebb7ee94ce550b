#!/usr/bin/env python
"""fwd+bwd device time of the small / mid-size batches (fp32, d = 256): one CUDA-graph replay of
supcon_loss_and_grad through the C-ABI (CUDA events), default dispatch vs the tiled exact kernels (flags = 4)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from wav2vec_contr_loss_b200 import functional as Fn
dev = torch.device("cuda:0")


def graph_us(fn, iters=50):
    gr = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=side):
            fn()
    torch.cuda.synchronize()
    for _ in range(5):
        gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    return round(1e3 * e0.elapsed_time(e1) / iters, 2)


for n in (64, 128, 160, 192, 256, 384, 512, 1024):
    g = torch.Generator().manual_seed(1337)
    z = torch.nn.functional.normalize(torch.randn(n, 256, generator=g), dim=1).to(dev)
    y = (torch.arange(n) % 2).to(torch.int32).to(dev)
    rec = {"N": n}
    for name, sim, lam, k, alpha in (("cosine", 0, 0.0, 15, 0.0), ("geodesic+uni", 1, 0.05, 15, 0.0), ("mined", 0, 0.0, 15, 0.5)):
        for tag, flags in (("default", 0), ("tiled", 4)):
            prob = Fn.make_problem(n, 256, 0, tau=0.07, similarity=sim, lambda_uni=lam, topk=k, alpha=alpha, flags=flags)
            rec[f"{name}/{tag}_us"] = graph_us(lambda: Fn.loss_and_grad(z, y, prob, want_grad=True))
    print(json.dumps(rec), flush=True)
