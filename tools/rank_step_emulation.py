#!/usr/bin/env python
"""One rank's share of the row-sharded step on a SINGLE GPU, without the collectives: what each of R ranks
launches between its NCCL calls (rows [r N/R, (r+1) N/R) against all N columns).  Times, as CUDA-graph replays,
  single : supcon_forward_rows + supcon_backward_rows                       (no overlap structure)
  phased : forward_rows_local + _remote, finalize_sets, backward_rows_local + _remote  (what distributed.py issues)
so that the cost of the two-phase structure itself (extra launches, parked SMs, extra partial records) is known
apart from the communication.  The gap between `phased` and a rank's measured step is the exposed communication.

    python tools/rank_step_emulation.py [N] [R,R,...] [rank]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from wav2vec_contr_loss_b200 import functional as Fn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ranks = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]
which = int(sys.argv[3]) if len(sys.argv) > 3 else -1          # -1: a middle rank
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1337)
z = torch.nn.functional.normalize(torch.randn(n, 256, generator=g), dim=1).to(dev).to(torch.bfloat16)
y = (torch.rand(n, generator=g) < 0.5).to(torch.int32).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def graph_ms(fn, iters=20):
    gr = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=side):
            fn()
    torch.cuda.synchronize()
    for _ in range(3):
        gr.replay()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


whole = Fn.make_problem(n, 256, 1, tau=0.07, similarity=0, topk=15, alpha=0.0, flags=32)
stats_all, partials_all, _ = Fn.forward_rows(z, y, whole, want_loss=True)
t1 = None
for R in ranks:
    nl = n // R
    r = (R // 2) if which < 0 else which
    prob = Fn.make_problem(n, 256, 1, tau=0.07, similarity=0, topk=15, alpha=0.0, row_offset=r * nl, n_rows=nl, flags=32)
    sets = partials_all.repeat(R, 1) / R

    def single():
        st, pa, _ = Fn.forward_rows(z, y, prob, want_loss=False)
        return Fn.backward_rows(z, y, stats_all, partials_all, None, prob, out_dtype=torch.bfloat16)

    def phased():
        ws = Fn.forward_rows_local(z, y, prob)
        st, pa = Fn.forward_rows_remote(z, y, prob, ws)
        pg, loss = Fn.finalize_sets(prob, sets)
        ws2 = Fn.backward_rows_local(z, y, st, pa, prob)
        return Fn.backward_rows_remote(z, y, stats_all, pg, None, prob, ws2, out_dtype=torch.bfloat16)

    def fwd_phased():
        ws = Fn.forward_rows_local(z, y, prob)
        return Fn.forward_rows_remote(z, y, prob, ws)

    st_l, pa_l, _ = Fn.forward_rows(z, y, prob, want_loss=False)

    def bwd_phased():
        ws2 = Fn.backward_rows_local(z, y, st_l, pa_l, prob)
        return Fn.backward_rows_remote(z, y, stats_all, partials_all, None, prob, ws2, out_dtype=torch.bfloat16)

    rec = dict(N=n, R=R, rank=r, rows=nl, single_ms=round(graph_ms(single), 4))
    if R > 1:
        rec.update(phased_ms=round(graph_ms(phased), 4), fwd_phased_ms=round(graph_ms(fwd_phased), 4),
                   bwd_phased_ms=round(graph_ms(bwd_phased), 4))
    rec["fwd_single_ms"] = round(graph_ms(lambda: Fn.forward_rows(z, y, prob, want_loss=False)), 4)
    rec["bwd_single_ms"] = round(graph_ms(lambda: Fn.backward_rows(z, y, stats_all, partials_all, None, prob,
                                                                   out_dtype=torch.bfloat16)), 4)
    if R == 1:
        t1 = rec["single_ms"]
    if t1:
        rec["ideal_ms"] = round(t1 / R, 4)
        rec["compute_efficiency_single"] = round(t1 / R / rec["single_ms"], 4)
        if R > 1:
            rec["compute_efficiency_phased"] = round(t1 / R / rec["phased_ms"], 4)
    rec["tflops_single"] = round(6.0 * n * nl * 256 / rec["single_ms"] / 1e9, 1)
    print(json.dumps(rec), flush=True)
