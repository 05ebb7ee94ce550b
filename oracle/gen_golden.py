"""TEST INFRASTRUCTURE ONLY.  Regenerate tests/golden/*.npz from the REAL reference.

Run in the build container (where /root/reference exists):

    python -m oracle.gen_golden

Each fixture stores the exact inputs (z fp32, labels) plus what the reference's
``loss.py`` returns for them: the loss and autograd dz with the module run as
shipped (fp32) and run in fp64 on the same fp32 inputs (ground truth).  The GPU
box has no /root/reference, so these files are how its tests see the reference.
"""
import json
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import supcon_oracle as O
from oracle.ref_loader import load_reference_module

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, n, d, kind, classes, similarity, tau, lambda, t, K, alpha
CASES = [
    ("c1_cosine_n64",          64, 256, "iso",       2, "cosine",   0.07, 0.0,  2.0, 15, 0.0),
    ("c2_geodesic_uni_n64",    64, 256, "iso",       2, "geodesic", 0.07, 0.05, 2.0, 15, 0.0),
    ("mined_half_n64",         64, 256, "clustered", 2, "cosine",   0.2,  0.2,  2.0, 15, 0.5),
    ("geo_mined_full_n64",     64, 256, "iso",       2, "geodesic", 0.03, 0.0,  2.0, 5,  1.0),
    ("k32_n128_d64",           128, 64, "iso",       2, "cosine",   0.07, 0.0,  2.0, 32, 0.37),
    ("ties_n96_d32",           96,  32, "ties",      2, "cosine",   0.1,  0.0,  2.0, 7,  1.0),
    ("seven_class_n130",       130, 64, "iso",       7, "cosine",   0.07, 0.1,  2.0, 32, 0.37),
    ("k_zero_n64",             64,  16, "iso",       2, "cosine",   0.5,  0.0,  2.0, 0,  0.7),
    ("k_all_neg_n33",          33,  8,  "iso",       2, "cosine",   0.07, 0.05, 2.0, 100, 1.0),
    ("geo_clustered_n48_d20",  48,  20, "clustered", 3, "geodesic", 0.1,  0.2,  3.0, 3,  0.5),
    ("odd_d_n37_d19",          37,  19, "iso",       2, "geodesic", 0.2,  0.1,  2.0, 4,  0.25),
]

# label layouts with degenerate anchors (SURVEY 8a semantics)
LAYOUTS = [
    ("all_same_n16",     [1] * 16),
    ("all_distinct_n12", list(range(12))),
    ("singleton_n17",    [0] + [1] * 8 + [2] * 8),
    ("single_row_n1",    [1]),
    ("two_rows_n2",      [0, 0]),
]


def run_reference(ref, z32, y, sim, tau, lam, t, k, alpha, dtype):
    z = z32.to(dtype).clone().requires_grad_(True)
    mod = ref.SupConBinaryLoss(tau, sim, lam, t)
    loss = mod(z, y, topk_neg=k, alpha=alpha)
    if loss.requires_grad and z.requires_grad:
        try:
            (g,) = torch.autograd.grad(loss, z, allow_unused=True)
        except RuntimeError:
            g = None
    else:
        g = None
    if g is None:
        g = torch.zeros_like(z)
    return float(loss), g.detach()


def save(name, meta, z32, y, ref):
    l32, g32 = run_reference(ref, z32, y, meta["similarity"], meta["tau"], meta["lambda_uni"],
                             meta["uni_t"], meta["topk"], meta["alpha"], torch.float32)
    l64, g64 = run_reference(ref, z32, y, meta["similarity"], meta["tau"], meta["lambda_uni"],
                             meta["uni_t"], meta["topk"], meta["alpha"], torch.float64)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        meta=json.dumps(meta), z=z32.numpy(), labels=y.numpy(),
        loss32=np.float64(l32), loss64=np.float64(l64),
        dz32=g32.numpy(), dz64=g64.numpy())
    print(f"{name:28s} loss32={l32:.9f} loss64={l64:.12f} |dz|={g64.norm():.6e}")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    ref = load_reference_module("loss")
    for name, n, d, kind, classes, sim, tau, lam, t, k, alpha in CASES:
        x, y = O.make_inputs(n, d, kind, classes=classes)
        z32 = F.normalize(x, p=2, dim=1)
        meta = dict(n=n, d=d, kind=kind, classes=classes, similarity=sim, tau=tau,
                    lambda_uni=lam, uni_t=t, topk=k, alpha=alpha)
        save(name, meta, z32, y, ref)
    for name, labels in LAYOUTS:
        n = len(labels)
        x, _ = O.make_inputs(max(n, 2), 24, "iso", seed=7)
        z32 = F.normalize(x[:n], p=2, dim=1)
        y = torch.tensor(labels, dtype=torch.int64)
        for sim, lam, alpha in (("cosine", 0.0, 0.0), ("geodesic", 0.3, 0.5)):
            meta = dict(n=n, d=24, kind="layout", classes=len(set(labels)), similarity=sim, tau=0.1,
                        lambda_uni=lam, uni_t=2.0, topk=3, alpha=alpha)
            save(f"{name}_{sim}", meta, z32, y, ref)
    # un-normalised rows (the module takes z "as given", reference loss.py:96-100)
    x, y = O.make_inputs(40, 12, "iso", seed=11)
    meta = dict(n=40, d=12, kind="raw", classes=2, similarity="cosine", tau=0.5,
                lambda_uni=0.1, uni_t=0.5, topk=6, alpha=0.4)
    save("unnormalised_n40_d12", meta, (0.6 * x).contiguous(), y, ref)

    # multi-class class (reference loss.py:156-210)
    x, y = O.make_inputs(64, 32, "iso", classes=5)
    z32 = F.normalize(x, p=2, dim=1)
    out = {}
    for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
        z = z32.to(dt).clone().requires_grad_(True)
        loss = ref.SupConMultiClassLoss(0.1)(z, y)
        (g,) = torch.autograd.grad(loss, z)
        out["loss" + tag] = np.float64(float(loss))
        out["dz" + tag] = g.numpy()
    np.savez_compressed(os.path.join(OUT, "multiclass_n64_d32.npz"),
                        meta=json.dumps(dict(n=64, d=32, tau=0.1, classes=5)),
                        z=z32.numpy(), labels=y.numpy(), **out)
    print("multiclass_n64_d32", out["loss64"])

    # Appendix-B table (RNG-free inputs, gradient w.r.t. the un-normalised x, fp64)
    rows = []
    for b, d, sim, tau, lam, k, alpha in [
            (8, 4, "cosine", 0.07, 0.0, 15, 0.0), (64, 256, "cosine", 0.07, 0.0, 15, 0.0),
            (64, 256, "geodesic", 0.07, 0.05, 15, 0.0), (64, 256, "cosine", 0.2, 0.2, 15, 0.5),
            (64, 256, "geodesic", 0.03, 0.0, 5, 1.0), (256, 256, "cosine", 0.07, 0.0, 32, 0.37)]:
        x, y = O.appendix_b_inputs(b, d)
        x = x.requires_grad_(True)
        z = F.normalize(x, p=2, dim=1)
        loss = ref.SupConBinaryLoss(tau, sim, lam, 2.0)(z, y, topk_neg=k, alpha=alpha)
        (dx,) = torch.autograd.grad(loss, x)
        rows.append(dict(b=b, d=d, similarity=sim, tau=tau, lambda_uni=lam, topk=k, alpha=alpha,
                         loss=float(loss), dx_norm=float(dx.norm()), dx_sum=float(dx.sum()),
                         dx_first=float(dx[0, 0]), dx_last=float(dx[-1, -1])))
        print("appendixB", rows[-1])
    with open(os.path.join(OUT, "appendix_b.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
