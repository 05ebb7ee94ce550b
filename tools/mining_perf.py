#!/usr/bin/env python
"""forward/backward time of the tensor path with and without hard-negative mining."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from wav2vec_contr_loss_b200 import functional as Fn
dev = torch.device("cuda:0")
for n in (1024, 4096, 16384, 65536):
    g = torch.Generator().manual_seed(1337)
    z = torch.nn.functional.normalize(torch.randn(n, 256, generator=g), dim=1).to(dev).to(torch.bfloat16)
    y = (torch.rand(n, generator=g) < 0.5).to(torch.int32).to(dev)
    for alpha, k in ((0.0, 15), (0.5, 15), (0.5, 32)):
        prob = Fn.make_problem(n, 256, 1, tau=0.07, similarity=0, topk=k, alpha=alpha, flags=32)
        def fwd(): return Fn.forward_rows(z, y, prob, want_loss=True)
        stats, partials, loss = fwd()
        def bwd(): return Fn.backward_rows(z, y, stats, partials, None, prob, out_dtype=torch.bfloat16)
        res = {}
        for name, fn in (("fwd", fwd), ("bwd", bwd)):
            gr = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn(); torch.cuda.synchronize()
                with torch.cuda.graph(gr, stream=side):
                    fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3): gr.replay()
            e0.record()
            for _ in range(10): gr.replay()
            e1.record(); torch.cuda.synchronize()
            res[name] = round(e0.elapsed_time(e1) / 10, 4)
        print(json.dumps(dict(n=n, alpha=alpha, k=k, **res)), flush=True)
