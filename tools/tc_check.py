#!/usr/bin/env python
"""Tensor-core path diagnostics: loss / dz / row-stat errors vs the CPU oracle for
bf16 inputs, with the tensor path on both directions (flags 0), forward only (8)
or backward only (16).  Prints, never asserts."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from oracle import supcon_oracle as O
import gpu_util as G
from wav2vec_contr_loss_b200 import functional as Fn

ap = argparse.ArgumentParser()
ap.add_argument("--big", action="store_true")
ap.add_argument("--perf", action="store_true")
args = ap.parse_args()
bf16 = torch.bfloat16
cases = [
    (512, "iso", 2, "cosine", 0.07, 0.0, 0.0, 15),
    (1024, "iso", 2, "cosine", 0.07, 0.0, 0.0, 15),
    (1000, "iso", 3, "cosine", 0.07, 0.0, 0.0, 15),
    (640, "clustered", 2, "cosine", 0.1, 0.05, 0.0, 15),
    (512, "iso", 2, "geodesic", 0.07, 0.0, 0.0, 15),
    (777, "iso", 5, "geodesic", 0.1, 0.2, 0.0, 15),
    (2048, "iso", 2, "cosine", 0.07, 0.0, 0.0, 0),
    (1024, "iso", 2, "cosine", 0.07, 0.0, 0.5, 15),      # hard-negative mining on the tensor path
    (1000, "iso", 3, "cosine", 0.07, 0.0, 1.0, 32),
    (640, "clustered", 2, "geodesic", 0.1, 0.05, 0.37, 5),
    (2048, "ties", 2, "cosine", 0.07, 0.0, 1.0, 15),
]
if args.big:
    cases += [(4096, "iso", 2, "cosine", 0.07, 0.0, 0.0, 15), (8192, "iso", 2, "cosine", 0.07, 0.05, 0.0, 15)]
for (n, kind, classes, sim, tau, lam, alpha, k) in cases:
    x, y = O.make_inputs(n, 256, kind, classes=classes)
    z = F.normalize(x, dim=1).to(bf16)
    kw = dict(tau=tau, similarity=sim, lam=lam, t=2.0, topk=k, alpha=alpha)
    ref = G.oracle_for(z.float(), y, **kw)
    for flags in (8, 16, 0):
        rec = dict(n=n, kind=kind, sim=sim, tau=tau, lam=lam, flags=flags)
        try:
            t0 = time.time()
            out = G.kernel_stats(z, y, dtype=bf16, flags=flags, **kw)
            torch.cuda.synchronize()
            st = out["stats"].cpu()
            rec["lse_err"] = float((st[:, 0].double() - ref["stats"]["lse"]).abs().max())
            rec["pos_mean_err"] = float((st[:, 7].double() - ref["stats"]["pos_mean"]).abs().max())
            rec["wsum_rel"] = float((st[:, 6].double() - ref["stats"]["wsum"]).abs().max() / max(float(ref["stats"]["wsum"].abs().max()), 1e-30))
            rec["npos_eq"] = bool((st.view(torch.int32)[:, 2].long() == ref["stats"]["npos"]).all())
            if alpha != 0:
                rec["lse_m_err"] = float((st[:, 1].double() - ref["stats"]["lse_m"]).abs().max())
                rec["thr_idx_eq"] = float((st.view(torch.int32)[:, 5].long() == ref["stats"]["thr_idx"]).float().mean())
                rec["thr_val_err"] = float((st[:, 4].double() - ref["stats"]["thr_val"]).abs().max())
            rec["loss"] = float(out["loss"]); rec["loss_ref"] = ref["loss"]
            rec["loss_rel"] = abs(rec["loss"] - ref["loss"]) / abs(ref["loss"])
            # backward through the C-ABI with fp32 dz
            prob = out["prob"]
            dz = Fn.backward_rows(out["z"], out["y"], out["stats"], out["partials"], None, prob, out_dtype=torch.float32)
            torch.cuda.synchronize()
            rec["dz_rel"] = G.rel_err(dz.cpu(), ref["dz"])
            rec["dz_maxabs"] = float((dz.cpu().double() - ref["dz"]).abs().max()); rec["dz_refmax"] = float(ref["dz"].abs().max())
            rec["sec"] = round(time.time() - t0, 3)
        except Exception as e:  # noqa: BLE001
            rec["error"] = f"{type(e).__name__}: {e}"
            print(json.dumps(rec), flush=True)
            raise SystemExit(1)
        print(json.dumps(rec), flush=True)
if args.perf:
    dev = torch.device("cuda:0")
    for n in (4096, 16384, 65536):
        x, y = O.make_inputs(n, 256, "iso")
        z = F.normalize(x, dim=1).to(dev).to(bf16)
        yl = Fn.canonical_labels(y.to(dev), n)
        prob = Fn.make_problem(n, 256, 1, tau=0.07, similarity=0, topk=15, alpha=0.0, flags=32)
        for _ in range(3):
            stats, partials, loss = Fn.forward_rows(z, yl, prob, want_loss=True)
            dz = Fn.backward_rows(z, yl, stats, partials, None, prob, out_dtype=bf16)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        fw = bw = 0.0
        iters = 10
        for _ in range(iters):
            e[0].record(); stats, partials, loss = Fn.forward_rows(z, yl, prob, want_loss=True)
            e[1].record(); dz = Fn.backward_rows(z, yl, stats, partials, None, prob, out_dtype=bf16)
            e[2].record(); torch.cuda.synchronize()
            fw += e[0].elapsed_time(e[1]); bw += e[1].elapsed_time(e[2])
        fw /= iters; bw /= iters
        print(json.dumps(dict(n=n, fwd_ms=fw, bwd_ms=bw, fwd_tflops=2 * n * n * 256 / fw / 1e9, bwd_tflops=4 * n * n * 256 / bw / 1e9,
                              total_tflops=6 * n * n * 256 / (fw + bw) / 1e9, loss=float(loss))), flush=True)
