"""Row-sharded SupCon over several GPUs (one process per GPU, NCCL).

Rank r holds the embeddings/labels of its local batch = rows
[r*n_local, (r+1)*n_local) of the global N x N similarity matrix:

  forward   all_gather(z) + all_gather(labels)          (one coalesced NCCL launch over NVLink)
            row-block forward kernel  -> row stats + 8 partial sums
            all_gather(partial sums) + all_gather(row stats, N x 32 B)   (one coalesced launch;
            the partials are summed in rank order on every rank: deterministic)
            -> scalar loss (identical on every rank)
  backward  row-block backward kernel: dz_i = sum_j (G_ij + G_ji) z_j for the
            owned rows, recomputing the tiles.  Because the similarity matrix
            is symmetric the column-side term G_ji only needs the *statistics*
            of row j, so no N x d reduce-scatter of column partials is needed
            (deviation from the north_star plan, SURVEY H4/H6: the exchange is
            N x 32 B of statistics instead of N x d x 4 B of gradients).

The result equals the single-GPU loss on the concatenated batch (what the
reference's nn.DataParallel computes, train_stage1.py:82-84).  The returned
gradient is d(global loss)/d(z_local).

``kernels`` is the object providing forward_rows / finalize / backward_rows
(default: the CUDA C-ABI wrappers).  The CPU gloo tests inject an oracle-backed
stand-in to exercise this host logic without a GPU; the product default has no
CPU path.
"""
from typing import Optional

import torch
import torch.distributed as dist

from . import functional as Fn


class _CudaKernels:
    name = "cuda"

    @staticmethod
    def forward_rows(z_all, labels_all, prob):
        stats, partials, _ = Fn.forward_rows(z_all, labels_all, prob, want_loss=False)
        return stats, partials

    finalize = staticmethod(Fn.finalize)

    @staticmethod
    def backward_rows(z_all, labels_all, stats_all, partials, grad_out, prob, out_dtype):
        return Fn.backward_rows(z_all, labels_all, stats_all, partials, grad_out, prob, out_dtype=out_dtype)

    # two-phase forward: own columns while the all-gather is in flight, then the rest
    forward_rows_local = staticmethod(Fn.forward_rows_local)
    forward_rows_remote = staticmethod(Fn.forward_rows_remote)


_COMM_STREAMS = {}


def _comm_stream(device) -> "torch.cuda.Stream":
    key = (device.type, device.index)
    if key not in _COMM_STREAMS:
        _COMM_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COMM_STREAMS[key]


def gather_and_forward(z_local, labels_local, make_prob, group, kernels):
    """all-gather(z, labels) overlapped with the forward over this rank's own columns.

    The local block is copied into place first and the all-gather runs in place on a side stream, so the
    phase-1 kernels (which read only that block) race with nothing.  Returns (z_all, labels_all, prob,
    row_stats, partials)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local, d = z_local.shape
    dev = z_local.device
    prob = make_prob(n_local * world, d, rank * n_local, n_local)
    two_phase = dev.type == "cuda" and hasattr(kernels, "forward_rows_local")
    if not two_phase:
        z_all, labels_all = gather_inputs(z_local, labels_local, group)
        stats, partials = kernels.forward_rows(z_all, labels_all, prob)
        return z_all, labels_all, prob, stats, partials
    z_all = torch.empty((world * n_local, d), dtype=z_local.dtype, device=dev)
    y_all = torch.empty(world * n_local, dtype=labels_local.dtype, device=dev)
    zb, yb = z_all[rank * n_local:(rank + 1) * n_local], y_all[rank * n_local:(rank + 1) * n_local]
    zb.copy_(z_local)
    yb.copy_(labels_local)
    cur, comm = torch.cuda.current_stream(dev), _comm_stream(dev)
    comm.wait_stream(cur)
    with torch.cuda.stream(comm):
        with _coalesced(group, dev):
            dist.all_gather_into_tensor(z_all, zb, group=group)
            dist.all_gather_into_tensor(y_all, yb, group=group)
    z_all.record_stream(comm)
    y_all.record_stream(comm)
    ws = kernels.forward_rows_local(z_all, y_all, prob)      # runs while the all-gather is in flight
    cur.wait_stream(comm)
    stats, partials = kernels.forward_rows_remote(z_all, y_all, prob, ws)
    return z_all, y_all, prob, stats, partials


def _all_gather_rows(x: torch.Tensor, group) -> torch.Tensor:
    world = dist.get_world_size(group)
    out = torch.empty((world * x.size(0),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def _coalesced(group, device):
    """One NCCL group launch for the collectives issued inside the block (falls back to
    plain sequential collectives on backends without coalescing support, e.g. gloo)."""
    import contextlib
    mgr = getattr(dist, "_coalescing_manager", None)
    if mgr is None or device.type != "cuda":
        return contextlib.nullcontext()
    return mgr(group=group, device=device, async_ops=False)


def gather_inputs(z_local: torch.Tensor, labels_local: torch.Tensor, group=None):
    """all-gather of embeddings and labels in ONE collective launch. Returns (z_all, labels_all)."""
    world = dist.get_world_size(group)
    z_all = torch.empty((world * z_local.size(0), z_local.size(1)), dtype=z_local.dtype, device=z_local.device)
    y_all = torch.empty(world * labels_local.size(0), dtype=labels_local.dtype, device=labels_local.device)
    with _coalesced(group, z_local.device):
        dist.all_gather_into_tensor(z_all, z_local.contiguous(), group=group)
        dist.all_gather_into_tensor(y_all, labels_local.contiguous(), group=group)
    return z_all, y_all


def exchange_stats(partials: torch.Tensor, stats: torch.Tensor, group=None):
    """Global partial sums + row statistics of every rank in ONE collective launch: the 8 fp64 partial
    sums travel as 16 fp32 words beside the statistics (two all-gathers of the same dtype coalesce; an
    all-reduce would not) and are summed on every rank in rank order, so the result is deterministic.
    ``partials`` is overwritten with the global sums; returns stats_all."""
    world = dist.get_world_size(group)
    stats_all = torch.empty((world * stats.size(0),) + tuple(stats.shape[1:]), dtype=stats.dtype, device=stats.device)
    if partials.dtype == torch.float64 and stats.dtype == torch.float32:
        p32 = partials.view(torch.float32)
        p_all = torch.empty(world * p32.numel(), dtype=torch.float32, device=partials.device)
        with _coalesced(group, stats.device):
            dist.all_gather_into_tensor(p_all, p32, group=group)
            dist.all_gather_into_tensor(stats_all, stats.contiguous(), group=group)
        partials.copy_(p_all.view(torch.float64).view(world, -1).sum(dim=0))
    else:   # generic fallback (CPU stand-in kernels in the gloo tests)
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
        dist.all_gather_into_tensor(stats_all, stats.contiguous(), group=group)
    return stats_all


class _ShardedSupCon(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_local, labels_local, cfg, group, kernels):
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        zc = Fn.canonical_z(z_local.detach())

        def make_prob(n_total, d, row_offset, n_rows):
            return Fn.make_problem(n_total, d, Fn._dtype_id(zc), row_offset=row_offset, n_rows=n_rows, **cfg)

        z_all, labels_all, prob, stats, partials = gather_and_forward(zc, labels_local, make_prob, group, kernels)
        if ctx.needs_input_grad[0]:
            stats_all = exchange_stats(partials, stats, group)
        else:
            dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
        loss = kernels.finalize(prob, partials)
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(z_all, labels_all, stats_all, partials)
            ctx.prob, ctx.kernels, ctx.in_dtype, ctx.work_dtype = prob, kernels, z_local.dtype, zc.dtype
        return Fn._loss_dtype(loss, z_local.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        z_all, labels_all, stats_all, partials = ctx.saved_tensors
        dz = ctx.kernels.backward_rows(z_all, labels_all, stats_all, partials, grad_out, ctx.prob, ctx.work_dtype)
        return dz.to(ctx.in_dtype), None, None, None, None


class ShardedSupConLoss(torch.nn.Module):
    """SupConBinaryLoss over the global batch of all ranks (same constructor and
    call signature as reference loss.py:19-25,110-114, plus ``group``)."""

    def __init__(self, temperature: float = 0.2, similarity: str = "geodesic", uniformity_weight: float = 0.0,
                 uniformity_t: float = 2.0, group: Optional[dist.ProcessGroup] = None, kernels=None):
        super().__init__()
        self.tau = temperature
        self.similarity = similarity.lower()
        self.lambda_uni = float(uniformity_weight)
        self.uni_t = float(uniformity_t)
        Fn.similarity_id(similarity)
        self.group = group
        self.kernels = kernels if kernels is not None else _CudaKernels
        self.kernel_flags = 0

    def forward(self, z: torch.Tensor, labels: torch.Tensor, topk_neg: int = 32, alpha: float = 0.0):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("ShardedSupConLoss needs an initialised torch.distributed process group")
        if self.kernels is _CudaKernels:
            Fn._require_cuda(z, "z")
        cfg = dict(tau=self.tau, similarity=Fn.similarity_id(self.similarity), lambda_uni=self.lambda_uni,
                   uni_t=self.uni_t, topk=topk_neg, alpha=alpha, flags=self.kernel_flags)
        lab = Fn.canonical_labels(labels, z.size(0))
        return _ShardedSupCon.apply(z, lab, cfg, self.group, self.kernels)
