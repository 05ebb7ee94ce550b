"""Head step of BASELINE configs[4] (frozen encoder): eager launches vs one CUDA-graph replay.

    python tools/graphed_step_bench.py [--batch 64] [--iters 30] > gpurun_out/graphed_step.json

hs is the encoder output the reference hands the head: (B, 25, 1024, 199) fp32 (1.3 GB at B = 64), synthetic.
The head has the shape/ops of the reference's compression head (compression_module.py:35-67): layer mean,
Dropout(0.1), LeakyReLU, Linear(1024 -> 256) per frame.  Step = stage1_utils.py:121-130 without the encoder.
Times are CUDA events on the launching stream, median of --iters steps after warm-up.
"""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class LayerMeanHead(nn.Module):
    def __init__(self, feat=1024, dim=256, p=0.1):
        super().__init__()
        self.drop, self.act, self.proj = nn.Dropout(p), nn.LeakyReLU(), nn.Linear(feat, dim)

    def forward(self, hs):                      # (B, K, F, T) -> (B, dim, T)
        x = self.act(self.drop(hs.mean(dim=1)))
        return self.proj(x.transpose(1, 2)).transpose(1, 2)


def median_ms(fn, iters):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    torch.cuda.synchronize()
    for s, e in ev:
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    return statistics.median(s.elapsed_time(e) for s, e in ev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--frames", type=int, default=199)
    ap.add_argument("--head", choices=("torch", "fused"), default="torch",
                    help="torch: the reference's op chain; fused: FusedCompressionHead (csrc/supcon_head.cu)")
    args = ap.parse_args()
    from wav2vec_contr_loss_b200 import build
    build.build()
    from wav2vec_contr_loss_b200 import FusedCompressionHead, SupConBinaryLoss, layer_time_pool
    from wav2vec_contr_loss_b200 import stage1 as S

    dev = torch.device("cuda", 0)
    torch.manual_seed(1337)
    B = args.batch
    hs = torch.randn(B, 25, 1024, args.frames, device=dev)
    y = (torch.arange(B, device=dev) % 2).long()
    loss_fn = SupConBinaryLoss(temperature=0.07, similarity="cosine", uniformity_weight=0.0)
    out = {"batch": B, "hs_shape": list(hs.shape), "hs_GB": hs.numel() * 4 / 1e9, "iters": args.iters,
           "head": args.head}
    if args.head == "fused":
        rng = torch.tensor([1337, 0], dtype=torch.int64, device=dev)
        with torch.no_grad():
            t_eval = median_ms(lambda: layer_time_pool(hs, 0.0, 0.01, None), args.iters)
            t_drop = median_ms(lambda: layer_time_pool(hs, 0.1, 0.01, rng), args.iters)
        out["head_pool_kernel"] = {"eval_ms": t_eval, "dropout_ms": t_drop,
                                   "eval_GBps": hs.numel() * 4 / 1e6 / t_eval,
                                   "dropout_GBps": hs.numel() * 4 / 1e6 / t_drop}

    for alpha in (0.0, 0.5):
        head = (FusedCompressionHead() if args.head == "fused" else LayerMeanHead()).to(dev).train()
        opt = torch.optim.AdamW(head.parameters(), lr=5e-3, weight_decay=3e-3, capturable=True)

        def eager_step():
            z = S.embed(head, hs)
            loss = loss_fn(z, y, topk_neg=15, alpha=alpha)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(head.parameters(), 5.0)
            opt.step()
            return loss

        t_eager = median_ms(eager_step, args.iters)
        step = S.GraphedHeadStep(head, loss_fn, opt, hs, y, topk_neg=15)
        t_graph_copy = median_ms(lambda: step(hs, y, alpha), args.iters)             # + 1.3 GB input copy
        t_graph = median_ms(lambda: step(step.hs, step.labels, alpha), args.iters)  # producer writes in place

        z = S.embed(head, hs).detach()

        def loss_only():
            zz = z.clone().requires_grad_(True)
            loss_fn(zz, y, topk_neg=15, alpha=alpha).backward()

        t_loss = median_ms(loss_only, args.iters)
        with torch.no_grad():
            t_mean = median_ms(lambda: hs.mean(dim=1), args.iters)
        out[f"alpha_{alpha}"] = {"eager_step_ms": t_eager, "graphed_step_ms": t_graph,
                                 "graphed_step_with_input_copy_ms": t_graph_copy,
                                 "loss_fwd_bwd_eager_ms": t_loss, "layer_mean_only_ms": t_mean,
                                 "layer_mean_GBps": hs.numel() * 4 / 1e6 / t_mean}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
