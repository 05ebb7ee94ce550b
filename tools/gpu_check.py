#!/usr/bin/env python
"""One-shot GPU diagnostics: runs a grid of cases through the CUDA path, prints
(never asserts) the error of loss / dz / row stats against the CPU oracle, and
times fwd+bwd with CUDA events.  Output goes to stdout and gpurun_out/.

    python tools/gpu_check.py [--quick] [--flags F] [--perf]
"""
import argparse
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.nn.functional as F

from oracle import supcon_oracle as O
import gpu_util as G
from wav2vec_contr_loss_b200 import functional as Fn
from wav2vec_contr_loss_b200 import _cabi


def stats_errors(out, ref):
    st = out["stats"].cpu()
    sti = st.view(torch.int32)
    r = ref["stats"]
    res = {}
    for name, col in (("lse", 0), ("lse_m", 1), ("wsum", 6), ("pos_mean", 7)):
        a, b = st[:, col].double(), r[name].double()
        res[name] = float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))
    res["npos_eq"] = bool((sti[:, 2].long() == r["npos"]).all())
    res["nneg_eq"] = bool((sti[:, 3].long() == r["nneg"]).all())
    res["thr_idx_eq"] = float((sti[:, 5].long() == r["thr_idx"]).float().mean())
    return res


def one_case(n, d, kind, classes, sim, tau, lam, t, k, alpha, dtype, flags):
    x, y = O.make_inputs(n, d, kind, classes=max(classes, 2))
    z = F.normalize(x, dim=1).to(dtype)
    kw = dict(tau=tau, similarity=sim, lam=lam, t=t, topk=k, alpha=alpha)
    ref = G.oracle_for(z.float(), y, **kw)
    rec = dict(n=n, d=d, kind=kind, sim=sim, tau=tau, lam=lam, k=k, alpha=alpha, dtype=str(dtype)[6:], flags=flags)
    try:
        loss, dz = G.kernel_loss_and_grad(z, y, dtype=dtype, flags=flags, **kw)
        rec["loss"] = loss
        rec["loss_ref"] = ref["loss"]
        rec["loss_rel"] = abs(loss - ref["loss"]) / max(abs(ref["loss"]), 1e-30)
        rec["dz_rel"] = G.rel_err(dz, ref["dz"])
        rec["dz_maxabs"] = float((dz - ref["dz"]).abs().max())
        out = G.kernel_stats(z, y, dtype=dtype, flags=flags, **{**kw, "alpha": alpha if alpha != 0 else 0.0})
        rec["stats"] = stats_errors(out, ref)
    except Exception as e:  # noqa: BLE001
        rec["error"] = f"{type(e).__name__}: {e}"
        traceback.print_exc()
    return rec


def time_case(n, d, dtype, sim, tau, lam, k, alpha, flags, iters=20):
    dev = torch.device("cuda:0")
    x, y = O.make_inputs(n, d, "iso")
    z = F.normalize(x, dim=1).to(dev).to(dtype)
    yl = Fn.canonical_labels(y.to(dev), n)
    prob = Fn.make_problem(n, d, Fn._dtype_id(z), tau=tau, similarity=Fn.similarity_id(sim), lambda_uni=lam,
                           topk=k, alpha=alpha, flags=flags | 32)   # 32: rows are L2-normalised above

    def step():
        stats, partials, loss = Fn.forward_rows(z, yl, prob, want_loss=True)
        dz = Fn.backward_rows(z, yl, stats, partials, None, prob, out_dtype=dtype)
        return loss, stats, partials

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    fw = bw = 0.0
    for _ in range(iters):
        e0.record()
        stats, partials, loss = Fn.forward_rows(z, yl, prob, want_loss=True)
        e1.record()
        Fn.backward_rows(z, yl, stats, partials, None, prob, out_dtype=dtype)
        e2.record()
        torch.cuda.synchronize()
        fw += e0.elapsed_time(e1)
        bw += e1.elapsed_time(e2)
    fw, bw = fw / iters, bw / iters
    # CUDA-graph replay (device time without host launch overhead)
    graph_us = None
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            step()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                step()
        torch.cuda.synchronize()
        reps = 50
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph_us = 1e3 * e0.elapsed_time(e1) / reps
    except Exception as e:  # noqa: BLE001
        graph_us = f"graph failed: {e}"
    fused_us = None
    if n <= Fn.SMALL_BATCH_MAX:
        # the single-launch path (supcon_loss_and_grad), CUDA-graph replay = device time of fwd+bwd
        try:
            g2 = torch.cuda.CUDAGraph()
            s2 = torch.cuda.Stream()
            s2.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s2):
                Fn.loss_and_grad(z, yl, prob)
                torch.cuda.synchronize()
                with torch.cuda.graph(g2, stream=s2):
                    Fn.loss_and_grad(z, yl, prob)
            torch.cuda.synchronize()
            for _ in range(5):
                g2.replay()
            e0.record()
            for _ in range(200):
                g2.replay()
            e1.record()
            torch.cuda.synchronize()
            fused_us = 1e3 * e0.elapsed_time(e1) / 200
        except Exception as e:  # noqa: BLE001
            fused_us = f"failed: {e}"
    flops = (6 + (2 if lam > 0 else 0)) * n * n * d
    return dict(n=n, d=d, dtype=str(dtype)[6:], sim=sim, lam=lam, k=k, alpha=alpha, flags=flags, fwd_ms=fw,
                bwd_ms=bw, graph_us=graph_us, fused_graph_us=fused_us, pairs_per_s=n * n / ((fw + bw) * 1e-3),
                tflops=flops / ((fw + bw) * 1e-3) / 1e12)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--perf", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gpu_check.json"))
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    print("device:", torch.cuda.get_device_name(0), "lib:", _cabi.lib_path(), flush=True)
    f32, bf16 = torch.float32, torch.bfloat16
    cases = [
        (8, 4, "iso", 2, "cosine", 0.07, 0.0, 2.0, 15, 0.0, f32),
        (64, 256, "iso", 2, "cosine", 0.07, 0.0, 2.0, 15, 0.0, f32),
        (64, 256, "iso", 2, "geodesic", 0.07, 0.05, 2.0, 15, 0.0, f32),
        (64, 256, "clustered", 2, "cosine", 0.2, 0.2, 2.0, 15, 0.5, f32),
        (64, 256, "iso", 2, "geodesic", 0.03, 0.0, 2.0, 5, 1.0, f32),
        (130, 19, "iso", 7, "cosine", 0.07, 0.1, 2.0, 32, 0.37, f32),
        (256, 256, "iso", 2, "cosine", 0.07, 0.0, 2.0, 32, 0.37, f32),
        (1024, 256, "iso", 2, "cosine", 0.07, 0.0, 2.0, 15, 0.5, f32),
        (1024, 256, "iso", 2, "cosine", 0.07, 0.0, 2.0, 600, 1.0, f32),
        (1024, 256, "iso", 2, "cosine", 0.07, 0.0, 2.0, 200, 1.0, f32),
        (1000, 200, "iso", 5, "geodesic", 0.1, 0.2, 2.0, 15, 0.37, f32),
        (512, 256, "iso", 2, "cosine", 0.07, 0.05, 2.0, 15, 0.5, bf16),
        (512, 256, "iso", 2, "geodesic", 0.07, 0.05, 2.0, 15, 0.5, bf16),
    ]
    if not args.quick:
        cases += [(4096, 256, "iso", 2, "cosine", 0.07, 0.0, 2.0, 15, 0.5, f32),
                  (4096, 256, "iso", 2, "cosine", 0.07, 0.0, 2.0, 15, 0.0, bf16)]
    results = {"cases": [], "perf": []}
    for c in cases:
        rec = one_case(*c, flags=args.flags)
        results["cases"].append(rec)
        print(json.dumps(rec), flush=True)
    if args.perf:
        for (n, dtype, sim, lam, k, alpha) in [
                (64, f32, "cosine", 0.0, 15, 0.0), (64, f32, "geodesic", 0.05, 15, 0.0),
                (64, f32, "cosine", 0.0, 15, 0.5), (128, f32, "cosine", 0.0, 15, 0.0), (160, f32, "geodesic", 0.05, 15, 0.5),
                (256, f32, "cosine", 0.0, 15, 0.0), (1024, f32, "cosine", 0.0, 15, 0.0),
                (1024, f32, "cosine", 0.0, 15, 0.5), (1024, bf16, "cosine", 0.0, 15, 0.5),
                (4096, f32, "cosine", 0.0, 15, 0.0), (16384, bf16, "cosine", 0.0, 15, 0.0)]:
            try:
                rec = time_case(n, 256, dtype, sim, 0.07, lam, k, alpha, args.flags, iters=10 if n > 4096 else 20)
            except Exception as e:  # noqa: BLE001
                rec = dict(n=n, error=str(e))
            results["perf"].append(rec)
            print(json.dumps(rec), flush=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
