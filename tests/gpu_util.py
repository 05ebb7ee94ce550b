"""Helpers shared by the GPU parity tests and tools/gpu_check.py."""
import torch

from oracle import supcon_oracle as O
from wav2vec_contr_loss_b200 import _cabi
from wav2vec_contr_loss_b200 import functional as Fn
from wav2vec_contr_loss_b200.loss import SupConBinaryLoss


def kernel_loss_and_grad(z_cpu, y_cpu, *, tau, similarity, lam=0.0, t=2.0, topk=32, alpha=0.0,
                         dtype=torch.float32, device="cuda:0", flags=0, grad_scale=1.0, unit_rows=None):
    """Through the public module + autograd. Returns (loss float, dz fp64 cpu)."""
    z = z_cpu.to(device=device, dtype=dtype).requires_grad_(True)
    y = y_cpu.to(device)
    mod = SupConBinaryLoss(tau, similarity, lam, t)
    mod.kernel_flags = flags
    mod.assume_unit_rows = unit_rows
    loss = mod(z, y, topk_neg=topk, alpha=alpha)
    if loss.requires_grad:
        (grad_scale * loss).backward()
        dz = z.grad if z.grad is not None else torch.zeros_like(z)
    else:
        dz = torch.zeros_like(z)
    return float(loss), (dz.detach().double().cpu() / grad_scale)


def oracle_for(z_cpu, y_cpu, *, tau, similarity, lam=0.0, t=2.0, topk=32, alpha=0.0, **kw):
    return O.closed_form(z_cpu, y_cpu, temperature=tau, similarity=similarity, uniformity_weight=lam,
                         uniformity_t=t, topk_neg=topk, alpha=alpha, **kw)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def kernel_stats(z_cpu, y_cpu, *, tau, similarity, lam=0.0, t=2.0, topk=32, alpha=1.0, dtype=torch.float32,
                 device="cuda:0", flags=0, row_offset=0, n_rows=None):
    """Row statistics + partials + top-k index sets straight from the C-ABI."""
    z = Fn.canonical_z(z_cpu.to(device=device, dtype=dtype))
    y = Fn.canonical_labels(y_cpu.to(device), z.size(0))
    prob = Fn.make_problem(z.size(0), z.size(1), Fn._dtype_id(z), tau=tau, similarity=Fn.similarity_id(similarity),
                           lambda_uni=lam, uni_t=t, topk=topk, alpha=alpha, flags=flags,
                           row_offset=row_offset, n_rows=n_rows)
    whole = row_offset == 0 and (n_rows is None or n_rows == z.size(0))
    stats, partials, loss = Fn.forward_rows(z, y, prob, want_loss=whole)
    # index sets are re-derived with the exact fp32 Gram: only meaningful for statistics of the exact paths
    uses_tc = (dtype == torch.bfloat16 and z.size(1) == 256 and z.size(0) >= 256 and tau >= 0.025
               and not (flags & 1) and (alpha == 0 or topk <= 32)
               and (similarity == "geodesic" or (flags & (2 | 32))))
    idx = Fn.topk_indices(z, y, stats, prob) if (topk >= 1 and not uses_tc) else None
    return dict(z=z, y=y, prob=prob, stats=stats, partials=partials, loss=loss, idx=idx)
