#!/usr/bin/env python
"""A/B timing of the tensor-path forward / backward of two builds of this repo in the same GPU call:
    python tools/ab_kernel_times.py [--root <tree>] [--n 65536] [--alpha 0.0]
CUDA-graph replays, CUDA events, L2 flushed between replays; alternates nothing -- run once per tree."""
import argparse, json, os, sys
ap = argparse.ArgumentParser()
ap.add_argument("--root", default=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap.add_argument("--n", type=int, default=65536)
ap.add_argument("--alpha", type=float, default=0.0)
ap.add_argument("--topk", type=int, default=15)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
sys.path.insert(0, os.path.abspath(args.root))
import torch
from wav2vec_contr_loss_b200 import functional as Fn
dev = torch.device("cuda:0")
n = args.n
g = torch.Generator().manual_seed(1337)
z = torch.nn.functional.normalize(torch.randn(n, 256, generator=g), dim=1).to(dev).to(torch.bfloat16)
y = (torch.rand(n, generator=g) < 0.5).to(torch.int32).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
try:
    prob = Fn.make_problem(n, 256, 1, tau=0.07, similarity=0, topk=args.topk, alpha=args.alpha, flags=32)
    Fn.forward_rows(z, y, prob, want_loss=True)
except Exception:   # a build that predates the unit-rows promise
    prob = Fn.make_problem(n, 256, 1, tau=0.07, similarity=0, topk=args.topk, alpha=args.alpha)
stats, partials, loss = Fn.forward_rows(z, y, prob, want_loss=True)


def graph_ms(fn):
    gr = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=side):
            fn()
    torch.cuda.synchronize()
    for _ in range(3):
        gr.replay()
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return round(sum(ts) / len(ts), 4), round(ts[len(ts) // 2], 4), round(ts[0], 4)


fwd = graph_ms(lambda: Fn.forward_rows(z, y, prob, want_loss=True))
bwd = graph_ms(lambda: Fn.backward_rows(z, y, stats, partials, None, prob, out_dtype=torch.bfloat16))
print(json.dumps(dict(root=os.path.abspath(args.root), n=n, alpha=args.alpha, topk=args.topk, loss=float(loss),
                      fwd_ms_mean_median_min=fwd, bwd_ms_mean_median_min=bwd,
                      bwd_tflops_median=round(4.0 * n * n * 256 / bwd[1] / 1e9, 1),
                      fwd_tflops_median=round(2.0 * n * n * 256 / fwd[1] / 1e9, 1))), flush=True)
